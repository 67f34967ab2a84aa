#!/bin/bash
# Round-2 confirmation on one B200: GPU tests, smoke, both bench arms, launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_gpu_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_predict.json 2> gpurun_out/r02_bench_predict.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_predict.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_predict_step.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 200 -c 4 -f -o gpurun_out/r02_gemm_final python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/r02_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pair_kernel -s 40 -c 1 -f -o gpurun_out/r02_attention_final python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/r02_ncu_att.log 2>&1; echo "ncu att rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 400 gpurun_out/r02_bench_reference.json
timeout 600 python tools/gpu_library_bar.py > gpurun_out/r02_library_bar.md 2> gpurun_out/r02_library_bar.err; echo "library bar rc=$?"; tail -12 gpurun_out/r02_library_bar.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1800 --csv --log-file gpurun_out/r02_launches_train_step.csv python bench.py --workload train --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_launches_train.log 2>&1; echo "train launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_pair -s 6 -c 3 -f -o gpurun_out/r02_ctc_pair_final python bench.py --workload train --steps 1 --warmup 2 --skip-cpu-baseline > gpurun_out/r02_ncu_ctc.log 2>&1; echo "ncu ctc rc=$?"
