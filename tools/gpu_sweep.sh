#!/bin/bash
# BASELINE configs[3] (hierarchical, full inventory) and configs[4] (30 s utterances, batch sweep) on one GPU.
# Usage (under gpurun): bash tools/gpu_sweep.sh  -> gpurun_out/sweep_*.json
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --hierarchical --batch 128 --seconds 10 --inventory 3183 --steps 5 --warmup 3 --skip-cpu-baseline \
  > gpurun_out/sweep_config4.json 2> gpurun_out/sweep_config4.err
echo "config4 rc=$?"
for b in 1 4 16 64 256; do
  timeout 400 python bench.py --batch $b --seconds 30 --steps 5 --warmup 3 --skip-cpu-baseline \
    > gpurun_out/sweep_30s_b$b.json 2> gpurun_out/sweep_30s_b$b.err
  echo "30s batch $b rc=$?"
done
