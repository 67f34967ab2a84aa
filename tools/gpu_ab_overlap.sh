#!/bin/bash
# Same-box A/B of the backward pass with the weight / bias gradients on a second stream (APH_BWD_OVERLAP).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_e2e_parity.py -x -q -m gpu > gpurun_out/ab_overlap_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_overlap_tests.log
for round in 1 2; do
  for v in 0 1; do
    APH_BWD_OVERLAP=$v timeout 600 python bench.py --workload train --skip-cpu-baseline > gpurun_out/ab_overlap_${v}_${round}.json 2> gpurun_out/ab_overlap_${v}_${round}.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_overlap_${v}_${round}.json").read().strip().splitlines()[-1])
print("overlap ${v} round ${round}: ms/step %.3f value %.0f e2e %.0f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
PY
  done
done
