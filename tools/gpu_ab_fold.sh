#!/bin/bash
# LayerNorm folded into the GEMMs (APH_FOLD_LN): parity tests, then the bench line with the fold off and on.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e_parity.py tests/test_gpu_full_size.py tests/test_gpu_full_size_parity.py -m gpu -x -q > gpurun_out/ab_fold_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/ab_fold_tests.log
grep -E "^\[config" gpurun_out/ab_fold_tests.log | cut -c1-400
APH_FOLD_LN=0 timeout 600 python bench.py --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/ab_fold_off.json 2> gpurun_out/ab_fold_off.err; echo "off rc=$?"
APH_FOLD_LN=1 timeout 600 python bench.py --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/ab_fold_on.json 2> gpurun_out/ab_fold_on.err; echo "on rc=$?"
python - <<'PY'
import json
for name in ("off", "on"):
    try:
        d = json.loads(open(f"gpurun_out/ab_fold_{name}.json").read().strip().splitlines()[-1])
        print(name, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "gemm frac", round(d["roofline"]["frac"], 3),
              "avg gemm us", round(1000 * d["roofline"]["avg_launch_ms"], 1))
    except Exception as e:
        print(name, "failed", e)
        print(open(f"gpurun_out/ab_fold_{name}.err").read()[-1500:])
PY
