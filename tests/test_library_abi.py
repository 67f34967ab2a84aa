"""CPU: the C-ABI library loads without a GPU and exports every symbol include/allophant_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "allophant_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aph_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    symbols = declared_symbols()
    for required in ("aph_gemm_bf16", "aph_attention_bf16", "aph_ctc_forward", "aph_ctc_backward", "aph_ctc_greedy_collapse",
                     "aph_log_softmax_heads", "aph_conv0_ln_gelu", "aph_wave_stats", "aph_compose_embeddings"):  # fmt: skip
        assert required in symbols


def test_library_exports_every_declared_symbol():
    from allophant_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH)
    for symbol in declared_symbols():
        assert hasattr(_lib.lib, symbol), f"{symbol} is declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols(), "ctypes signatures and header disagree"
    assert _lib.lib.aph_abi_version() == 7


def test_integration_guide_indexes_every_entry_point():
    """INTEGRATION.md section 5 (tools/abi_index.py) names every declared symbol with the reference interface it replaces."""
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [symbol for symbol in declared_symbols() if f"| `{symbol}` |" not in guide]
    assert not missing, f"run tools/abi_index.py: INTEGRATION.md lacks {missing}"


def test_argument_validation_needs_no_gpu():
    """Launchers validate before touching CUDA and report through aph_last_error()."""
    from allophant_b200 import _lib

    args = _lib.GemmArgs()
    rc = _lib.lib.aph_gemm_bf16(ctypes.byref(args), None)
    assert rc == _lib.APH_ERR_INVALID
    assert b"null operand" in _lib.lib.aph_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "aph_gemm_bf16")
    assert _lib.lib.aph_ctc_states_pad(100) == 256
    assert _lib.lib.aph_ctc_states_pad(511) == 1024
    assert _lib.lib.aph_ctc_states_pad(512) == 1056  # block-per-pair path: 2 * 512 + 1 states rounded up to 32
    assert _lib.lib.aph_ctc_states_pad(8001) == _lib.APH_ERR_UNSUPPORTED


def test_no_cpu_fallback():
    """The product path fails loudly off-GPU instead of silently computing on the CPU."""
    import torch

    from allophant_b200 import ops
    from allophant_b200.predictions import GreedyCTCDecoder

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.log_softmax(torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="GPU"):
        GreedyCTCDecoder()(torch.zeros(1, 4, 4), torch.tensor([4]))


def test_product_never_imports_the_oracle():
    package = os.path.join(ROOT, "allophant_b200")
    for directory, _, files in os.walk(package):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(directory, name), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f"{name} imports the oracle"


def test_custom_ops_are_registered_with_fake_kernels():
    """The C-ABI entry points a user composes by hand are PyTorch custom ops (namespace allophant_b200) with fake kernels,
    and have no CPU implementation."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    from allophant_b200 import custom_ops

    for name in custom_ops.REGISTERED:
        assert hasattr(torch.ops.allophant_b200, name)
    with FakeTensorMode():
        x = torch.empty(7, 3, 40, device="cuda")
        assert torch.ops.allophant_b200.log_softmax(x).shape == (7, 3, 40)
        assert torch.ops.allophant_b200.linear_bf16(x, torch.empty(16, 40, device="cuda"), None, True).shape == (7, 3, 16)
        assert torch.ops.allophant_b200.ctc_nll(x, torch.empty(3, 5, dtype=torch.long, device="cuda"), torch.empty(3, dtype=torch.long, device="cuda"),
                                                 torch.empty(3, dtype=torch.long, device="cuda")).shape == (3,)  # fmt: skip
        q = torch.empty(6, 200, 64, device="cuda", dtype=torch.bfloat16)
        assert torch.ops.allophant_b200.attention(q, q, q, torch.empty(3, dtype=torch.int32, device="cuda"), 2).shape == (600, 128)
        loss, gradient = torch.ops.allophant_b200.ctc_loss_with_gradient(
            x, torch.empty(3, 5, dtype=torch.long, device="cuda"), torch.empty(3, dtype=torch.long, device="cuda"), torch.empty(3, dtype=torch.long, device="cuda"))  # fmt: skip
        assert loss.shape == (1,) and gradient.shape == x.shape
        tokens, counts, scores = torch.ops.allophant_b200.ctc_greedy_decode(x, torch.empty(3, dtype=torch.int32, device="cuda"))
        assert tokens.shape == (3, 7) and counts.shape == (3,) and scores.shape == (3,)
    with pytest.raises(NotImplementedError):
        torch.ops.allophant_b200.log_softmax(torch.zeros(2, 3))


def test_ctypes_mirrors_match_the_header_structs(tmp_path):
    """The by-value descriptor structs (`aph_ctc_head`, `aph_head_block`, `aph_gemm_args`) are mirrored as ctypes Structures in
    `_lib.py`: sizes and field offsets are compared with what gcc computes from `include/allophant_b200.h` itself."""
    import ctypes
    import re
    import subprocess

    from allophant_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mirrors = {"aph_ctc_head": _lib.CtcHead, "aph_head_block": _lib.HeadBlock, "aph_gemm_args": _lib.GemmArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "allophant_b200.h"', "int main(void) {"]
    for c_name, mirror in mirrors.items():
        lines.append(f'  printf("{c_name} size %zu\\n", sizeof({c_name}));')
        for field, _ in mirror._fields_:
            lines.append(f'  printf("{c_name} {field} %zu\\n", offsetof({c_name}, {field}));')
    lines += ["  return 0;", "}"]
    source = tmp_path / "layout.c"
    source.write_text("\n".join(lines))
    binary = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(source), "-o", str(binary)], check=True)
    output = subprocess.run([str(binary)], check=True, capture_output=True, text=True).stdout
    for line in output.splitlines():
        c_name, field, value = re.match(r"(\w+) (\w+) (\d+)", line).groups()
        mirror = mirrors[c_name]
        expected = ctypes.sizeof(mirror) if field == "size" else getattr(mirror, field).offset
        assert int(value) == expected, (c_name, field, value, expected)
