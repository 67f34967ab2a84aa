#!/bin/bash
# Same-box A/B of the GEMM tail split: kernel + parity tests, then the predict bench (incl. the training sub-record) with the
# split off / on / off / on.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e_parity.py tests/test_gpu_training.py -x -q -m gpu > gpurun_out/ab_tail_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tail_tests.log
for round in 1 2; do
  for v in 0 1; do
    APH_GEMM_TAIL_SPLIT=$v timeout 600 python bench.py --skip-cpu-baseline --skip-membound --skip-ragged > gpurun_out/ab_tail_${v}_${round}.json 2> gpurun_out/ab_tail_${v}_${round}.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_tail_${v}_${round}.json").read().strip().splitlines()[-1])
t = d.get("train") or {}
k = d["roofline"].get("kernel_only") or {}
print("tail split ${v} round ${round}: ms/step %.3f value %.0f e2e %.0f gemm frac %.3f (cupti %.3f, %.1f us) | train ms %.3f frac %.3f" % (
    d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], k.get("frac", 0), 1e3 * k.get("avg_launch_ms", 0), t.get("ms_per_step", 0), (t.get("roofline") or {}).get("frac", 0)))
PY
  done
done
