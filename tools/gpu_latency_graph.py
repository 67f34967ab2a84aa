"""Latency of small predict batches, eager vs CUDA graph (Estimator.predict(cuda_graph=True)), full XLS-R-300M shape."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from allophant_b200.dataset_processing import Batch

estimator, tfi = bench.build_estimator("cuda:0")
tfi = tfi.to("cuda:0")
for n_utt, seconds in [(1, 5), (1, 10), (4, 5), (8, 10), (32, 10)]:
    samples = seconds * 16000
    batch = Batch(0.1 * torch.randn(n_utt, samples, device="cuda:0"), torch.full((n_utt,), samples, device="cuda:0"), torch.zeros(n_utt, dtype=torch.long, device="cuda:0"))
    line = f"{n_utt} x {seconds} s:"
    for graphed in (False, True):
        for _ in range(5):
            estimator.predict(batch, tfi, cuda_graph=graphed)
        torch.cuda.synchronize()
        steps = 50
        start = time.perf_counter()
        for _ in range(steps):
            estimator.predict(batch, tfi, cuda_graph=graphed)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - start) / steps * 1e3
        line += f"  {'graph' if graphed else 'eager'} {ms:.3f} ms ({n_utt * seconds / ms * 1e3:,.0f} audio-s/s)"
    print(line, flush=True)
