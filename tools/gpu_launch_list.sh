#!/bin/bash
# ncu launch list (durations only) of one predict step; $1 = output tag, rest = env assignments
tag=$1; shift
mkdir -p gpurun_out
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-train --skip-membound --skip-ragged > gpurun_out/launches_${tag}.log 2>&1; echo "launch list ${tag} rc=$?"
python - <<PY
import csv, re, collections
rows = list(csv.reader(open("gpurun_out/launches_${tag}.csv")))
hdr = next(r for r in rows if "Kernel Name" in r)
data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r != hdr]
starts = [i for i, d in enumerate(data) if "wave_stats" in d["Kernel Name"]]
step = data[starts[-2]:starts[-1]] if len(starts) >= 2 else data[starts[-1]:]
agg = collections.OrderedDict()
for d in step:
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void aph::", "").replace("aph::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(d["Metric Value"]) / 1000
total = sum(v[1] for v in agg.values())
print(f"one step: {len(step)} launches, {total / 1000:.3f} ms serialised")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:45s} {n:4d} {us:9.1f} us {100 * us / total:5.1f}%  avg {us / n:7.1f}")
layer = [d for d in step if "gemm" in d["Kernel Name"] or "attention" in d["Kernel Name"] or "layernorm" in d["Kernel Name"]]
print("first encoder layers:", [(re.sub(r"\(.*", "", d["Kernel Name"]).replace("void aph::", "")[:24], round(float(d["Metric Value"]) / 1000, 1)) for d in layer[16:34]])
PY
