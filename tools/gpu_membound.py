"""Times the HBM-bound kernels at BASELINE config sizes (the `membound` sub-record of bench.py) and prints a markdown table:
achieved GB/s from the ALGORITHMIC bytes against the measured copy bandwidth of MEASURED_PEAKS.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

records = bench.measure_membound("cuda")
print("| kernel | shape | us | algorithmic MB | GB/s | of measured HBM peak | note |")
print("|---|---|---:|---:|---:|---:|---|")
for r in records:
    print(f"| {r['kernel']} | {r['shape']} | {r['us']:.1f} | {r['algorithmic_bytes'] / 1e6:.1f} | {r['achieved_gbs']:.0f} | {100 * r['frac']:.1f} % | {r['note']} |")
print()
print(json.dumps(records))
