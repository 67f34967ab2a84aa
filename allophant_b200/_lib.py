"""ctypes binding of ``liballophant_b200.so`` (the C ABI in ``include/allophant_b200.h``).

The shared library is the product: there is no Python/torch fallback for any
entry point.  Importing this module on a machine without the built library
raises immediately; calling a compute entry point without a GPU fails inside
CUDA and is reported as ``RuntimeError`` with the library's own message.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liballophant_b200.so")

APH_OK = 0
APH_ERR_INVALID = -1
APH_ERR_CUDA = -2
APH_ERR_UNSUPPORTED = -3

APH_GEMM_ROWS = 0
APH_GEMM_TAPS = 1
APH_GEMM_DIAG_TAPS = 2
APH_EPI_STORE = 0
APH_EPI_QKV = 1


class GemmArgs(Structure):
    """Mirror of ``aph_gemm_args``."""

    _fields_ = [
        ("a", c_void_p),
        ("a_row_stride", c_int64),
        ("a_batch_stride", c_int64),
        ("a_rows", c_int32),
        ("a_inner", c_int32),
        ("batch", c_int32),
        ("mode", c_int32),
        ("tap_pad", c_int32),
        ("b", c_void_p),
        ("n", c_int32),
        ("k", c_int32),
        ("epilogue", c_int32),
        ("gelu", c_int32),
        ("scale", c_float),
        ("bias", c_void_p),
        ("resid", c_void_p),
        ("ld_resid", c_int64),
        ("out_f32", c_void_p),
        ("ld_f32", c_int64),
        ("out_bf16", c_void_p),
        ("ld_bf16", c_int64),
        ("out_batch_rows", c_int64),
        ("lengths", c_void_p),
        ("len_period", c_int32),
        ("q", c_void_p),
        ("kmat", c_void_p),
        ("vt", c_void_p),
        ("heads", c_int32),
        ("t_v", c_int32),
        ("q_scale", c_float),
        ("a_mn_major", c_int32),
        ("b_mn_major", c_int32),
        ("b_row_stride", c_int64),
        ("b_seg_stride", c_int64),
        ("k_seq", c_int32),
        ("k_batch", c_int32),
        ("b_k_shift", c_int32),
        ("n_taps", c_int32),
        ("aux_bf16", c_void_p),
        ("ld_aux", c_int64),
        ("gelu_bwd", c_void_p),
        ("ld_gelu_bwd", c_int64),
        ("vmat", c_void_p),
        ("drop_threshold", ctypes.c_uint32),
        ("drop_seed", ctypes.c_uint32),
        ("drop_scale", ctypes.c_float),
        ("act_bwd", c_int32),
        ("row_stats", c_void_p),
        ("row_stats_slots", c_int32),
        ("ln_stats", c_void_p),
        ("ln_slots", c_int32),
        ("ln_cols", c_int32),
        ("ln_colsum", c_void_p),
        ("ln_eps", c_float),
        ("taps_span", c_int32),
    ]


class HeadBlock(Structure):
    """Mirror of ``aph_head_block``."""

    _fields_ = [("ptr", c_void_p), ("ld", c_int64), ("width", c_int32), ("rows", c_int32)]


class CtcHead(Structure):
    """Mirror of ``aph_ctc_head``."""

    _fields_ = [
        ("log_probs", c_void_p),
        ("grad", c_void_p),
        ("stride_t", c_int64),
        ("stride_n", c_int64),
        ("n_classes", c_int32),
        ("s_pad", c_int32),
        ("labels", c_void_p),
        ("label_stride", c_int64),
        ("label_lengths", c_void_p),
        ("alpha_offset", c_int64),
    ]


_P = c_void_p
_I32 = c_int32
_I64 = c_int64
_F = c_float
_U32 = ctypes.c_uint32

# name -> argtypes; every function returns int (APH_OK or a negative APH_ERR_* code)
_SIGNATURES = {
    "aph_gemm_bf16": [POINTER(GemmArgs), _P],
    "aph_attention_bf16": [_P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "aph_attention_bf16_lse": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "aph_attention_backward_bf16": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P],
    "aph_attention_bf16_dropout": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _U32, _U32, c_float, _P],
    "aph_attention_backward_bf16_dropout": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _U32, _U32, c_float, _P],
    "aph_dropout_2d": [_P, _I64, _I64, _I32, _U32, _U32, c_float, _P, _P, _P, _I64, _P, _I64, _P],
    "aph_dropout_bf16_2d": [_P, _I64, _I64, _I32, _U32, _U32, c_float, _P],
    "aph_spec_augment_mask": [_P, _I32, _I32, c_float, _I32, _I32, _U32, _P, _P],
    "aph_mask_columns": [_P, _I64, _I32, _I32, _I32, _P, _P, _I64, _P],
    "aph_masked_rows_backward": [_P, _I64, _I64, _I32, _P, _P, _P],
    "aph_debug_set_progress": [_P],
    "aph_debug_set_timeline": [_P],
    "aph_wave_stats": [_P, _P, _I32, _I32, _P, _P, _P],
    "aph_wave_norm": [_P, _P, _P, _I32, _I32, _P, _P],
    "aph_frame_lengths": [_P, _I32, _P, _P, _I32, _P, _P, _P],
    "aph_conv0_ln_gelu": [_P, _P, _P, _I32, _I32, _P, _P, _P, _P, _F, _I32, _P, _P],
    "aph_conv0_gn_gelu": [_P, _P, _P, _I32, _I32, _P, _P, _P, _P, _F, _P, _P, _P, _P],
    "aph_layernorm_rows": [_P, _I32, _I64, _I64, _I32, _P, _P, _F, _I32, _P, _I64, _P, _I64, _P],
    "aph_compose_embeddings": [_P, _I32, _I32, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P],
    "aph_log_softmax_heads": [_P, _I64, _I64, _I32, _I32, _P, _P, _P, _I32, _P, _P, _P, _P],
    "aph_log_softmax_wide": [_P, _I64, _I64, _I32, _P, _I64, _P, _P, _P],
    "aph_dependency_softmax": [_P, _I64, _I64, _P, _P, _P, _I32, _I32, _P, _I64, _P],
    "aph_argmax_rows": [_P, _I64, _I64, _I32, _P, _P, _P],
    "aph_ctc_greedy_collapse": [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "aph_ctc_pack_hypotheses": [_P, _P, _P, _I32, _I32, _P, _P, _P, _P],
    "aph_cast_bf16": [_P, _P, _I64, _P],
    "aph_cast_bf16_2d": [_P, _I64, _P, _I64, _I64, _I32, _P],
    "aph_pack_conv_weight": [_P, _P, _I32, _I32, _I32, _P],
    "aph_pack_posconv_weight": [_P, _P, _P, _P, _I32, _I32, _I32, _P],
    "aph_allophone_forward": [_P, _I64, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P],
    "aph_allophone_backward": [_P, _P, _P, _I64, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "aph_transpose_cast_bf16": [_P, _I32, _I64, _I64, _I32, _P, _I64, _I64, _P],
    "aph_colsum_f32": [_P, _I64, _I64, _I32, _P, _P],
    "aph_colsum_bf16": [_P, _I64, _I64, _I32, _P, _P],
    "aph_layernorm_backward": [_P, _I32, _I64, _P, _I32, _I64, _I64, _I32, _P, _F, _P, _I64, _P, _I64, _P, _P, _P],
    "aph_mask_rows_f32": [_P, _I64, _I64, _I32, _P, _I32, _P],
    "aph_add_f32_2d": [_P, _I64, _P, _I64, _I64, _I32, _P],
    "aph_gelu_backward_bf16": [_P, _I64, _P, _I64, _I64, _I32, _P, _I64, _P],
    "aph_pack_posconv_weight_dgrad": [_P, _P, _P, _P, _I32, _I32, _I32, _P],
    "aph_posconv_weight_backward": [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P],
    "aph_posconv_weight_backward_blocks": [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P],
    "aph_embedding_bag_backward": [_P, _I64, _I32, _I32, _I32, _P, _P, _P, _P],
    "aph_softmax_backward_cols": [_P, _I64, _P, _I64, _I64, _P, _P, _P, _I32, _I32, _P, _I64, _P],
    "aph_multi_tensor_sumsq": [_P, _P, _I32, _P, _P],
    "aph_multi_tensor_scale": [_P, _P, _I32, _P, _F, _P],
    "aph_multi_tensor_sgd": [_P, _P, _P, _I32, _F, _F, _F, _I32, _P, _F, _P],
    "aph_multi_tensor_adam": [_P, _P, _P, _I32, _F, _F, _F, _F, _F, _I64, _P, _F, _P],
    "aph_conv0_raw_bf16": [_P, _P, _P, _I32, _I32, _P, _P, _P, _P],
    "aph_ln_gelu_backward_512": [_P, _P, _I64, _I64, _P, _P, c_float, _P, _P, _P, _P, _P],
    "aph_conv0_weight_backward": [_P, _P, _P, _P, _I32, _I32, _P, _P],
    "aph_layernorm_any": [_P, _I64, _I64, _I32, _P, _P, c_float, _P, _I64, _P, _I64, _P],
    "aph_layernorm_any_backward": [_P, _I64, _P, _I64, _I64, _I32, _P, c_float, _P, _I64, _P, _I64, _P, _P, _P],
    "aph_glu_backward_bf16": [_P, _I64, _P, _I64, _I64, _I32, _P, _I64, _P],
    "aph_conv_input_backward": [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I64, _P],
    "aph_activation_backward": [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _I64, _P],
    "aph_add_sinusoidal": [_P, _I64, _I32, _I32, _I32, _P, _P],
    "aph_transpose_nfl": [_P, _I32, _I32, _I32, _P, _I64, _P],
    "aph_reflect_pad_bf16": [_P, _I64, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P],
    "aph_glu_rows": [_P, _I64, _I64, _I32, _P, _I64, _P],
    "aph_ctc_beam_decode": [_P, _P, _I64, _I64, _I32, _I32, _I32, _I32, ctypes.c_double, _I32, _I32, _P, _P, _P, _P, _I32],
    "aph_edit_matrix": [_P, _I64, _P, _I64, _P],
    "aph_collate_pad_f32": [_P, _P, _I64, _I64, _P, _I32],
    "aph_edit_statistics_batch": [_P, _P, _P, _P, _I64, _P, _P, _I32],
    "aph_fold_layernorm_linear": [_P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P],
    "aph_attention_small": [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _I32, _P],
    "aph_copy_head_blocks": [POINTER(HeadBlock), POINTER(HeadBlock), _I32, _I64, _I32, _P],
    "aph_log_softmax_head_blocks": [POINTER(HeadBlock), POINTER(HeadBlock), _I32, _I64, _P],
    "aph_ctc_states_pad": [_I32],
    "aph_ctc_forward": [POINTER(CtcHead), _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "aph_ctc_backward": [POINTER(CtcHead), _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
}

EXPORTED_SYMBOLS = sorted(
    list(_SIGNATURES)
    + ["aph_abi_version", "aph_last_error", "aph_launch_count", "aph_reset_launch_count", "aph_set_pdl", "aph_set_gemm_tail_split", "aph_set_attention_kernel", "aph_word_error_rate"]
    + ["aph_edit_operations", "aph_edit_weighted", "aph_segmenter_create", "aph_segmenter_free", "aph_segmenter_find"]
)


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C allophant_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "allophant_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.aph_abi_version.restype = c_int
    lib.aph_last_error.restype = c_char_p
    lib.aph_launch_count.restype = c_int64
    lib.aph_reset_launch_count.restype = None
    lib.aph_set_pdl.argtypes = [c_int]
    lib.aph_set_pdl.restype = c_int
    lib.aph_set_gemm_tail_split.argtypes = [c_int]
    lib.aph_set_gemm_tail_split.restype = c_int
    lib.aph_set_attention_kernel.argtypes = [c_int]
    lib.aph_set_attention_kernel.restype = c_int
    lib.aph_word_error_rate.argtypes = [ctypes.c_uint64] * 4
    lib.aph_word_error_rate.restype = c_float
    lib.aph_edit_operations.argtypes = [_P, _I64, _P, _I64, _P, _P]
    lib.aph_edit_operations.restype = c_int64
    lib.aph_edit_weighted.argtypes = [_I64, _I64, _P, c_float, c_float, _I32, _P, _P, _P, _P]
    lib.aph_edit_weighted.restype = c_int64
    lib.aph_segmenter_create.argtypes = [c_char_p, _P, _I64]
    lib.aph_segmenter_create.restype = c_void_p
    lib.aph_segmenter_free.argtypes = [c_void_p]
    lib.aph_segmenter_free.restype = None
    lib.aph_segmenter_find.argtypes = [c_void_p, c_char_p, _I64, _P, _I64]
    lib.aph_segmenter_find.restype = c_int64
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    """Turns a negative APH_ERR_* code into the exception type the reference raises."""
    if rc == APH_OK:
        return
    message = f"{what}: {lib.aph_last_error().decode(errors='replace')} (code {rc})"
    if rc == APH_ERR_INVALID:
        raise ValueError(message)
    if rc == APH_ERR_UNSUPPORTED:
        raise NotImplementedError(message)
    raise RuntimeError(message)
