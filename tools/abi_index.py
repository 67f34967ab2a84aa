"""Entry-point index of include/allophant_b200.h as a markdown table (pasted into INTEGRATION.md §5).

For every `aph_*` function the header declares: the header section it sits in, the Python wrapper in
`allophant_b200/ops.py` / `phonemes.py` / ... that calls it, and the first sentence of its header comment (which
carries the reference file:line the entry point replaces).
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "allophant_b200.h")
PACKAGE = os.path.join(ROOT, "allophant_b200")


def header_entries(text):
    """-> [(section, symbol, first comment sentence)] in header order."""
    entries = []
    section = ""
    comment = ""
    position = 0
    token = re.compile(r"/\*(.*?)\*/|\b(aph_[a-z0-9_]+)\s*\(", re.S)
    for match in token.finditer(text):
        if match.group(1) is not None:
            body = " ".join(line.strip().lstrip("*").strip() for line in match.group(1).splitlines()).strip()
            banner = re.match(r"-{2,}\s*(.*?)\s*-*$", body)
            if banner:
                section = re.sub(r"\s*-{2,}.*?-{2,}\s*", " ", banner.group(1)).strip(" -")
                comment = ""
            else:
                comment = body
            position = match.end()
            continue
        symbol = match.group(2)
        # Only declarations: the text between the last comment and the symbol must look like a return type.
        between = text[position:match.start()]
        if "typedef" in between or ";" in between.split("\n")[-1]:
            pass
        if any(symbol == known for _, known, _ in entries):
            continue
        sentence = re.split(r"(?<=[^.\d])\.\s", comment, maxsplit=1)[0].strip()
        entries.append((section, symbol, sentence))
    return entries


def python_callers():
    """symbol -> sorted list of `module.function` names under allophant_b200/ that call `lib().symbol` / `check(...)`."""
    callers = {}
    for folder, _, files in os.walk(PACKAGE):
        for name in files:
            if not name.endswith(".py") or name == "_lib.py":
                continue
            path = os.path.join(folder, name)
            module = os.path.relpath(path, PACKAGE)[:-3].replace(os.sep, ".")
            current = None
            owner = None
            for line in open(path):
                klass = re.match(r"class\s+([A-Za-z0-9_]+)", line)
                if klass:
                    owner = klass.group(1)
                definition = re.match(r"(\s*)def\s+([A-Za-z0-9_]+)", line)
                if definition:
                    if not definition.group(1):
                        owner = None
                    current = f"{owner}.{definition.group(2)}" if owner else definition.group(2)
                for symbol in re.findall(r"\b(aph_[a-z0-9_]+)\b", line):
                    callers.setdefault(symbol, set()).add(f"{module}.{current}" if current else module)
    return {symbol: sorted(names) for symbol, names in callers.items()}


OVERRIDES = {
    "aph_debug_set_progress": ("library (debug aids)", "Progress markers of block (0,0) of the attention backward kernels go to a host-mapped int32[16] array (NULL = off, the default)"),
    "aph_debug_set_timeline": ("library (debug aids)", "clock64 stamps of CTA (0,0) of the attention forward kernel go to a device int64[32] array (NULL = off)"),
    "aph_abi_version": ("library", "APH_ABI_VERSION the library was built with (checked by `_lib.py` at load)"),
    "aph_last_error": ("library", "Text of the last error on the calling thread (every entry point returns 0 or a negative APH_ERR_* code)"),
    "aph_gemm_bf16": ("tensor-core GEMM (tcgen05 + TMEM accumulators, TMA-fed)",
                      "Every nn.Linear / Conv1d (as implicit GEMM) on the path and their data- / weight-gradient forms: "
                      "HF:275-323 (conv stack), 326-379 (positional conv), 422-434 (projection), 466-573 (attention and "
                      "feed-forward linears), acoustic_model.py:395-416 (classifier heads); one `aph_gemm_args` struct"),
}


def main():
    entries = [(OVERRIDES.get(symbol, (section, None))[0], symbol, OVERRIDES.get(symbol, (None, sentence))[1])
               for section, symbol, sentence in header_entries(open(HEADER).read())]
    callers = python_callers()
    print("| Entry point | Header section | Python caller | Replaces / does (header comment, first sentence) |")
    print("|---|---|---|---|")
    for section, symbol, sentence in entries:
        used = ", ".join(f"`{name}`" for name in callers.get(symbol, [])[:2]) or "—"
        sentence = sentence.replace("|", "\\|")
        if len(sentence) > 400:
            sentence = sentence[:397] + "…"
        print(f"| `{symbol}` | {section.replace('|', '/')} | {used} | {sentence} |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
