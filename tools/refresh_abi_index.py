"""Regenerates the entry-point table of INTEGRATION.md from include/allophant_b200.h (tools/abi_index.py prints it)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "abi_index.py")], capture_output=True, text=True, check=True).stdout
path = os.path.join(ROOT, "INTEGRATION.md")
text = open(path).read()
start = text.index("| Entry point | Header section |")
lines = text[start:].split("\n")
count = 0
for line in lines:
    if not line.startswith("|"):
        break
    count += 1
end = start + len("\n".join(lines[:count]))
open(path, "w").write(text[:start] + table.rstrip("\n") + text[end:])
print("INTEGRATION.md: entry-point table refreshed,", table.count("\n") - 2, "entries")
