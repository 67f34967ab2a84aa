#!/bin/bash
# Round-end confirmation on one B200: GPU tests, smoke, both bench arms, launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench_predict.json 2> gpurun_out/final_bench_predict.err; tail -c 1800 gpurun_out/final_bench_predict.json
timeout 600 python bench.py --workload train --skip-cpu-baseline > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err; tail -c 1200 gpurun_out/final_bench_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches_predict.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/final_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 400 -c 8 -f -o gpurun_out/r01_gemm_final python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/final_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 80 -c 2 -f -o gpurun_out/r01_attention_final python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/final_ncu_att.log 2>&1; echo "ncu att rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 600 gpurun_out/final_bench_reference.json
timeout 300 python tools/bench_transformer.py 2>/dev/null | tail -1 > gpurun_out/final_bench_transformer.json; tail -c 300 gpurun_out/final_bench_transformer.json
