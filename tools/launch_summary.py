"""Markdown table of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total, share and average per kernel.

usage: python tools/launch_summary.py <launches.csv> [first_kernel_substring]
With a second argument only the LAST whole step is summarised: the launches from the last-but-one occurrence of that kernel
(e.g. `wave_stats_kernel`, the first launch of a step) up to the last occurrence."""
import csv
import re
import sys


def read(path):
    rows = []
    with open(path, newline="") as handle:
        lines = [line for line in handle if not line.startswith("==")]
    reader = csv.DictReader(lines)
    for row in reader:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        value = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        us = {"ns": value / 1e3, "nsecond": value / 1e3, "us": value, "usecond": value, "ms": value * 1e3, "msecond": value * 1e3}.get(unit, value / 1e3)
        name = re.sub(r"^(void )?(aph::)?", "", row["Kernel Name"].split("(")[0])
        rows.append((name, us))
    return rows


def main():
    rows = read(sys.argv[1])
    if len(sys.argv) > 2:
        marks = [i for i, (name, _) in enumerate(rows) if sys.argv[2] in name]
        if len(marks) >= 2:
            rows = rows[marks[-2] : marks[-1]]
    total = sum(us for _, us in rows)
    table = {}
    for name, us in rows:
        entry = table.setdefault(name, [0, 0.0])
        entry[0] += 1
        entry[1] += us
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms serialised\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (count, us) in sorted(table.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {count} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / count:.1f} |")


if __name__ == "__main__":
    main()
