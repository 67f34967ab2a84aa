"""Copies what `tools/gpu_round2_check.sh` left under gpurun_out/ into profiles/ (tracked) and writes the round-2 summaries:
launch-list tables, ncu --set full tables and the SASS instruction census of the built library.  Runs here (no GPU)."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROFILES = os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "allophant_b200", "liballophant_b200.so")


def run(*command):
    return subprocess.run(list(command), capture_output=True, text=True, cwd=ROOT).stdout


def copy(source, target):
    path = os.path.join(OUT, source)
    if os.path.exists(path) and os.path.getsize(path) > 0:
        shutil.copyfile(path, os.path.join(PROFILES, target))
        return True
    print("missing:", source)
    return False


copy("r02_bench_predict.json", "r02_bench_predict_n1_final.json")
copy("r02_bench_reference.json", "r02_bench_reference_arm.json")
copy("r02_launches_predict_step.csv", "r02_launches_predict_step.csv")
copy("r02_launches_train_step.csv", "r02_launches_train_step.csv")
copy("r02_gpu_tests.log", "r02_gpu_tests_final.log")
copy("r02_library_bar.md", "r02_library_bar_final.md")
copy("r02_full_size_parity.json", "r02_full_size_parity.json")

# ---- launch lists
bench = {}
path = os.path.join(OUT, "r02_bench_predict.json")
if os.path.exists(path):
    lines = [line for line in open(path).read().splitlines() if line.startswith("{")]
    if lines:
        bench = json.loads(lines[-1])
with open(os.path.join(PROFILES, "r02_launch_summary.md"), "w") as handle:
    handle.write("# Round 2 — launch lists of one predict step and one training step (B200, final tree)\n\n")
    handle.write(
        "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` (`tools/gpu_round2_check.sh`), summarised by\n"
        "`tools/launch_summary.py <csv> wave_stats_kernel` (one whole step = from one `wave_stats_kernel` to the next).  Per-launch times\n"
        "under ncu are cold-cache and serialised at ~1.6-1.75 GHz: read SHARES; the timed numbers are the bench line's.\n\n"
    )
    if bench:
        roof = bench.get("roofline", {})
        only = roof.get("kernel_only") or {}
        train = bench.get("train") or {}
        handle.write(
            f"Bench line of the same tree (`profiles/r02_bench_predict_n1_final.json`, not under ncu): {bench['value']:.0f} audio-s/s, "
            f"{bench['ms_per_step']:.2f} ms/step, e2e {bench['e2e']['value']:.0f} audio-s/s, {bench['gpu_launches']} launches per step; encoder GEMMs "
            f"{roof.get('achieved', 0):.0f} TFLOP/s = {roof.get('frac', 0):.3f} of the measured sustained bf16 peak (events around replays of the 96 launches), "
            f"{only.get('achieved', 0):.0f} TFLOP/s = {only.get('frac', 0):.3f} in the step itself (CUPTI); training step {train.get('ms_per_step', 0):.2f} ms "
            f"({train.get('value', 0):.0f} audio-s/s, {train.get('gpu_launches', 0)} launches).\n\n"
        )
    handle.write("## Predict step (BASELINE configs[1]: 32 x 10 s, 37 heads + greedy decode)\n\n")
    handle.write(run(sys.executable, "tools/launch_summary.py", "gpurun_out/r02_launches_predict_step.csv", "wave_stats_kernel"))
    handle.write("\n## Training step (BASELINE configs[2]: 8 utterances U[3, 15] s, train() mode, clip + Adam)\n\n")
    handle.write(run(sys.executable, "tools/launch_summary.py", "gpurun_out/r02_launches_train_step.csv", "wave_stats_kernel"))

# ---- ncu --set full
with open(os.path.join(PROFILES, "r02_ncu_summary.md"), "w") as handle:
    handle.write("# Round 2 — ncu --set full captures inside the bench step (B200, final tree)\n\n")
    handle.write(
        "`ncu --set full --clock-control none --import-source on -k regex:<kernel> ... python bench.py --steps 1 --warmup 3 ...`\n"
        "(`tools/gpu_round2_check.sh`), summarised with `python tools/ncu_summary.py <rep>`.  Launches under ncu are serialised, cold-cache and at\n"
        "~1.6-1.75 GHz.  `dram_rd_MB + dram_wr_MB` per launch is what `roofline.traffic` quotes; `xu_pct` is the MUFU / conversion pipe\n"
        "(`sm__inst_executed_pipe_xu`).\n\n"
    )
    for name, title in (
        ("r02_gemm_final.ncu-rep", "encoder GEMMs of one layer (FFN1 `<256,0>`, FFN2 `<256,2>`, QKV `<256,1>`, out-proj `<256,2>`)"),
        ("r02_attention_final.ncu-rep", "attention forward (query-tile-pair kernel), 32 x 16 heads x 499 frames x 64"),
        ("r02_ctc_pair_final.ncu-rep", "block-per-pair CTC recursions + gradient kernel of the training step (37 heads x 8 utterances)"),
        ("r02_ctc.ncu-rep", "CTC alpha / beta of the training step (before the staged log-sum-exp, see r02_ctc_recursion.md)"),
    ):
        rep = os.path.join(OUT, name)
        if os.path.exists(rep):
            handle.write(f"## {name}: {title}\n\n")
            handle.write(run(sys.executable, "tools/ncu_summary.py", rep))
            handle.write("\n")

# ---- SASS census
sass = run("cuobjdump", "-sass", LIB)
census = {}
kernels = {}
current = None
for line in sass.splitlines():
    line = line.strip()
    if line.startswith("Function :"):
        current = line.split(":", 1)[1].strip()
        continue
    for mnemonic in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS", "MUFU.EX2", "LDGSTS", "ACQBULK", "FFMA2", "FMUL2", "FADD2"):
        if f" {mnemonic}" in f" {line}" and (mnemonic + ".") in line + "." or f" {mnemonic} " in f" {line} ":
            census[mnemonic] = census.get(mnemonic, 0) + 1
            if current and mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG"):
                kernels.setdefault(current, {}).setdefault(mnemonic, 0)
                kernels[current][mnemonic] += 1
demangled = {}
if kernels:
    names = list(kernels)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(names, out))
with open(os.path.join(PROFILES, "r02_sass_census.md"), "w") as handle:
    handle.write("# Round 2 — SASS instruction census of `allophant_b200/liballophant_b200.so` (sm_100a)\n\n")
    handle.write("`cuobjdump -sass allophant_b200/liballophant_b200.so`, counted by `tools/collect_profiles.py` (static instruction counts).\n\n")
    handle.write("| mnemonic | what it proves | count |\n|---|---|---:|\n")
    meaning = {
        "UTCHMMA": "tcgen05.mma (bf16, TMEM accumulator)", "LDTM": "tcgen05.ld (TMEM -> registers)", "STTM": "tcgen05.st (registers -> TMEM: P of the attention)",
        "UTMALDG": "TMA tensor load", "UTMASTG": "TMA tensor store", "UTMAREDG": "TMA reduce-add store (split-K weight gradients)", "UTMAPF": "TMA L2 prefetch",
        "UTCBAR": "tcgen05.commit -> mbarrier", "SYNCS": "mbarrier operations", "MUFU.EX2": "exp2 on the MUFU pipe", "LDGSTS": "cp.async",
        "FFMA2": "packed fp32 FMA (f32x2)", "FMUL2": "packed fp32 multiply", "FADD2": "packed fp32 add",
    }  # fmt: skip
    for mnemonic, count in sorted(census.items(), key=lambda kv: -kv[1]):
        handle.write(f"| `{mnemonic}` | {meaning.get(mnemonic, '')} | {count} |\n")
    handle.write("\n## Kernels holding tensor-core / TMEM / TMA instructions\n\n| kernel | " + " | ".join(("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG")) + " |\n|---|" + "---:|" * 6 + "\n")
    for name, counts in sorted(kernels.items(), key=lambda kv: demangled.get(kv[0], kv[0])):
        short = demangled.get(name, name).split("(")[0].replace("void ", "").replace("aph::", "")
        handle.write(f"| `{short}` | " + " | ".join(str(counts.get(m, 0)) for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG")) + " |\n")
print("profiles written")
