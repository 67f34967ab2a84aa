"""Summarises an .ncu-rep (read here, no GPU needed): per launch duration, DRAM traffic, tensor/DRAM utilisation."""
import csv
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
}


def to_unit(value: str, unit: str, key: str) -> float:
    v = float(value.replace(",", "")) if value else float("nan")
    if key.endswith("dur_us"):
        return v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    if key.endswith("_MB"):
        return {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)
    return v


def main(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]
    index = {name: i for i, name in enumerate(header)}
    print("| # | kernel | grid | " + " | ".join(KEYS.values()) + " |")
    print("|---|---|---|" + "---|" * len(KEYS))
    for n, row in enumerate(rows[2:]):
        name = row[index["Kernel Name"]].split("(")[0][-48:]
        grid = row[index["Grid Size"]] if "Grid Size" in index else ""
        cells = []
        for metric, short in KEYS.items():
            if metric in index:
                cells.append(f"{to_unit(row[index[metric]], units[index[metric]], short):.2f}")
            else:
                cells.append("n/a")
        print(f"| {n} | `{name}` | {grid} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
