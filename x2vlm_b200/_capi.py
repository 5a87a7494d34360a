"""ctypes binding of libx2k.so (the C ABI declared in include/x2k.h).

No torch types cross the boundary: tensors are passed as raw device pointers (`tensor.data_ptr()`)
plus explicit shapes / leading dimensions, and the CUDA stream as a void*.
The library is REQUIRED: there is no CPU or eager fallback — a missing or unloadable libx2k.so
raises at first use.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("X2K_LIB") or os.path.join(_HERE, "lib", "libx2k.so")  # X2K_LIB: developer builds (tools/)

c_void_p = ctypes.c_void_p
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_uint64 = ctypes.c_uint64
c_float = ctypes.c_float


class X2kGemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", c_void_p), ("B", c_void_p),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("lda", c_int64), ("ldb", c_int64),
        ("a_mn_major", c_int32), ("b_mn_major", c_int32),
        ("bias", c_void_p),
        ("act", c_int32),
        ("aux", c_void_p), ("ld_aux", c_int64),
        ("preact_out", c_void_p), ("ld_preact", c_int64),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint64), ("dropout_offset", c_uint64),
        ("gamma", c_void_p),
        ("row_scale", c_void_p), ("rows_per_scale", c_int32),
        ("residual", c_void_p), ("ld_res", c_int64),
        ("accumulate", c_int32),
        ("out_bf16", c_void_p), ("ld_out_bf16", c_int64),
        ("out_f32", c_void_p), ("ld_out_f32", c_int64),
        ("tile_n", c_int32), ("max_ctas", c_int32), ("split_k", c_int32),
        ("dropout_offset_dev", c_void_p),
        ("ce_mode", c_int32), ("ce_labels", c_void_p), ("ce_partials", c_void_p), ("ce_target_logit", c_void_p),
        ("ce_lse", c_void_p), ("ce_row_grad", c_void_p),
    ]


class X2kAttnArgs(ctypes.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("ld_q", c_int64), ("ld_k", c_int64), ("ld_v", c_int64),
        ("B", c_int32), ("H", c_int32), ("Lq", c_int32), ("Lk", c_int32),
        ("kv_index", c_void_p), ("n_kv", c_int32),
        ("scale", c_float),
        ("bias", c_void_p), ("bias_h_stride", c_int64), ("bias_q_stride", c_int64),
        ("mask", c_void_p), ("mask_b_stride", c_int64), ("mask_q_stride", c_int64),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint64), ("dropout_offset", c_uint64),
        ("o", c_void_p), ("ld_o", c_int64),
        ("lse", c_void_p),
        ("d_o", c_void_p), ("ld_do", c_int64),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("ld_dq", c_int64), ("ld_dk", c_int64), ("ld_dv", c_int64),
        ("ds_out", c_void_p),
        ("ds_b_stride", c_int64), ("ds_h_stride", c_int64), ("ds_q_stride", c_int64),
        ("kv_groups", c_void_p),
        ("dropout_offset_dev", c_void_p),
        ("delta_ws", c_void_p),
        ("dq_ws", c_void_p),
    ]


ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX = 0, 1, 2, 3, 4

# every symbol include/x2k.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "x2k_version": (ctypes.c_int, []),
    "x2k_last_error": (ctypes.c_char_p, []),
    "x2k_launch_count": (c_int64, []),
    "x2k_gemm": (ctypes.c_int, [ctypes.POINTER(X2kGemmArgs), c_void_p]),
    "x2k_layernorm_fwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "x2k_layernorm_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p]),
    "x2k_layernorm_bwd_dropcast": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                                  c_float, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "x2k_scale_cast_colsum": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                             c_int32, c_float, c_uint64, c_uint64, c_void_p, c_void_p, c_int64,
                                             c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "x2k_colsum_bf16": (ctypes.c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "x2k_segment_sum_bf16": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "x2k_cast_f32_bf16": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "x2k_embed_ln_fwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_float, c_float, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "x2k_embed_ln_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_uint64, c_uint64, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "x2k_pool_tail_fwd": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "x2k_pool_tail_bwd": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "x2k_ce_finalize": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "x2k_attn_fwd": (ctypes.c_int, [ctypes.POINTER(X2kAttnArgs), c_void_p]),
    "x2k_attn_bwd": (ctypes.c_int, [ctypes.POINTER(X2kAttnArgs), c_void_p]),
    "x2k_attn_probs": (ctypes.c_int, [ctypes.POINTER(X2kAttnArgs), c_int32, c_void_p, c_void_p]),
    "x2k_attn_bwd_workspace_bytes": (c_int64, [ctypes.POINTER(X2kAttnArgs)]),
    "x2k_attn_group_slots": (c_int32, [c_int32]),
    "x2k_attn_group_table_ints": (c_int64, [c_int32, c_int32, c_int32]),
    "x2k_attn_group_build": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "x2k_relpos_bias_gather": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "x2k_relpos_bias_scatter": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int64, c_int64, c_int64,
                                               c_void_p, c_void_p, c_void_p]),
    "x2k_sumsq": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "x2k_adamw_flat": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_void_p, c_int32, c_float, c_float, c_float,
                                      c_int32, c_void_p, c_void_p, c_int32, c_void_p]),
}

_lib = None
_lock = threading.Lock()


class X2kError(RuntimeError):
    pass


def lib():
    """Load libx2k.so (once). Raises if it is missing — there is no fallback path."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise X2kError(
                        "libx2k.so not found at %s — build it with `python -m x2vlm_b200.build` "
                        "(or __graft_entry__.build()); the CUDA extension is mandatory" % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(handle, name)  # AttributeError if the symbol is not exported
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().x2k_last_error()
        raise X2kError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
