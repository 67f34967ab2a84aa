#!/bin/bash
# A/B of programmatic dependent launch + packed GELU on one B200: parity tests first, then the bench line with PDL off and on.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_pdl_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/ab_pdl_tests.log
APH_PDL=0 timeout 600 python bench.py --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/ab_pdl_off.json 2> gpurun_out/ab_pdl_off.err; echo "off rc=$?"
APH_PDL=1 timeout 600 python bench.py --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/ab_pdl_on.json 2> gpurun_out/ab_pdl_on.err; echo "on rc=$?"
python - <<'PY'
import json
for name in ("off", "on"):
    try:
        d = json.loads(open(f"gpurun_out/ab_pdl_{name}.json").read().strip().splitlines()[-1])
        print(name, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm frac", round(d["roofline"]["frac"], 3),
              "kernel_only", round(d["roofline"].get("kernel_only", {}).get("frac", 0), 3))
    except Exception as e:
        print(name, "failed", e)
PY
