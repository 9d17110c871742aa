"""ctypes binding of libclover_b200.so (the C ABI declared in include/clover_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, a RuntimeError
is raised.  The library is built in-tree by ``__graft_entry__.build()`` / ``make -C clover_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclover_b200.so")

c_ll = C.c_longlong
c_vp = C.c_void_p


class WindowGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "D", "H", "W", "wd", "wh", "ww", "sd", "sh", "sw")]


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", c_vp), ("residual", c_vp), ("residual_is_bf16", C.c_int), ("ld_residual", c_ll),
        ("out", c_vp), ("out_is_bf16", C.c_int), ("ld_out", c_ll),
        ("out_pre", c_vp), ("ld_pre", c_ll), ("gelu_pre", c_vp), ("ld_gelu_pre", c_ll),
        ("act", C.c_int), ("scale_cols", C.c_int), ("scale", C.c_float),
        ("window", C.POINTER(WindowGeom)), ("k_splits", C.c_int), ("accumulate", C.c_int),
        ("row_scale", c_vp), ("row_scale_rows", c_ll), ("rowsum", c_vp),
    ]


class LnDesc(C.Structure):
    _fields_ = [
        ("x", c_vp), ("x_is_bf16", C.c_int), ("ld_x", c_ll),
        ("gamma", c_vp), ("beta", c_vp), ("eps", C.c_float),
        ("mean", c_vp), ("rstd", c_vp), ("rows", c_ll), ("C", C.c_int),
        ("window", C.POINTER(WindowGeom)),
        ("merge_B", C.c_int), ("merge_D", C.c_int), ("merge_H", C.c_int), ("merge_W", C.c_int), ("merge_C", C.c_int),
        ("add0", c_vp), ("add1", c_vp), ("div1", C.c_int), ("mod1", C.c_int),
        ("add2", c_vp), ("div2", C.c_int), ("mod2", C.c_int),
        ("group_rows", c_ll), ("group_stride", c_ll), ("row_offset", c_ll),
        ("blend_mask", c_vp), ("blend_token", c_vp),
        ("blend_D", C.c_int), ("blend_H", C.c_int), ("blend_W", C.c_int), ("blend_mh", C.c_int), ("blend_mw", C.c_int),
        ("row_index", c_vp),
    ]


class LnBwd(C.Structure):
    _fields_ = [
        ("dy", c_vp), ("dy_is_bf16", C.c_int), ("ld_dy", c_ll),
        ("dx", c_vp), ("ld_dx", c_ll), ("dres", c_vp), ("ld_dres", c_ll),
        ("dx_copy", c_vp), ("dx_copy_is_bf16", C.c_int), ("ld_copy", c_ll),
        ("copy_window", C.POINTER(WindowGeom)),
        ("dgamma", c_vp), ("dbeta", c_vp), ("dtoken", c_vp), ("dx_dense", C.c_int),
        ("dxsum", c_vp), ("copy_scale", c_vp), ("copy_scale_rows", c_ll),
    ]


class LnrDesc(C.Structure):
    _fields_ = [
        ("x", c_vp), ("x_is_bf16", C.c_int), ("gamma", c_vp), ("beta", c_vp), ("eps", C.c_float),
        ("mean", c_vp), ("rstd", c_vp), ("rows", c_ll), ("C", C.c_int), ("row_map", c_vp), ("map_period", C.c_int),
        ("row_blend", c_vp), ("blend_token", c_vp),
    ]


class LnrBwd(C.Structure):
    _fields_ = [
        ("dy", c_vp), ("dy_is_bf16", C.c_int), ("dy_mapped", C.c_int), ("dres", c_vp), ("dx", c_vp),
        ("dx_bf16", c_vp), ("dx_bf16_mapped", C.c_int), ("dgamma", c_vp), ("dbeta", c_vp), ("dxsum", c_vp),
        ("copy_scale", c_vp), ("copy_scale_rows", c_ll), ("dtoken", c_vp),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("seq", C.c_int), ("heads", C.c_int), ("head_dim", C.c_int),
        ("bias_table", c_vp), ("table_len", C.c_int), ("rel_code", c_vp), ("code_off", C.c_int),
        ("region", c_vp), ("nwin", C.c_int), ("key_mask", c_vp),
        ("drop_p", C.c_float), ("drop_seed", C.c_ulonglong), ("drop_offset", C.c_ulonglong),
    ]


class AttnW7Desc(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("heads", C.c_int), ("wd", C.c_int),
        ("bias_table", c_vp), ("table_len", C.c_int), ("cfg_wd", C.c_int),
        ("q_ext", c_vp), ("k_ext", c_vp), ("nwin", C.c_int),
    ]


class RowsAffine(C.Structure):
    _fields_ = [
        ("x", c_vp), ("x_is_bf16", C.c_int), ("ld_x", c_ll),
        ("in_group_rows", c_ll), ("in_group_stride", c_ll), ("in_offset", c_ll),
        ("y", c_vp), ("y_is_bf16", C.c_int), ("ld_y", c_ll),
        ("out_group_rows", c_ll), ("out_group_stride", c_ll), ("out_offset", c_ll),
        ("add0", c_vp), ("bvec", c_vp), ("bdiv", c_ll), ("bscale", C.c_float),
        ("rows", c_ll), ("C", C.c_int),
    ]


# name -> (restype, argtypes); every entry point declared in include/clover_b200.h
SIGNATURES = {
    "clv_last_error": (C.c_char_p, []),
    "clv_version": (C.c_int, []),
    "clv_launch_count": (c_ll, []),
    "clv_gemm_bf16": (C.c_int, [c_vp, c_ll, C.c_int, c_vp, c_ll, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(GemmEpilogue), c_vp]),
    "clv_layernorm_fwd": (C.c_int, [C.POINTER(LnDesc), c_vp, C.c_int, c_ll, c_vp]),
    "clv_layernorm_bwd": (C.c_int, [C.POINTER(LnDesc), C.POINTER(LnBwd), c_vp]),
    "clv_lnr_supported": (C.c_int, [C.c_int]),
    "clv_lnr_fwd": (C.c_int, [C.POINTER(LnrDesc), c_vp, C.c_int, C.c_int, c_vp]),
    "clv_lnr_bwd": (C.c_int, [C.POINTER(LnrDesc), C.POINTER(LnrBwd), c_vp]),
    "clv_attention_fwd": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp]),
    "clv_attention_fwd_tc": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp]),
    "clv_attention_bwd": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp, c_vp, C.c_float, c_vp, c_vp, c_vp]),
    "clv_attention_tc64_supported": (C.c_int, [C.c_int]),
    "clv_attention_fwd_tc64": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp]),
    "clv_attention_bwd_tc64": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp, c_vp, C.c_float, c_vp, c_vp]),
    "clv_attention_probs_mean": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp]),
    "clv_dropout": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_ll, C.c_float, C.c_ulonglong, C.c_ulonglong, c_vp]),
    "clv_keep_mask": (C.c_int, [c_vp, c_ll, C.c_float, C.c_ulonglong, C.c_ulonglong, c_vp]),
    "clv_dropout_threshold": (C.c_uint, [C.c_float]),
    "clv_rand_u32": (C.c_uint, [C.c_ulonglong, C.c_ulonglong]),
    "clv_rows_scale": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_ll, C.c_int, c_vp, c_ll, c_vp]),
    "clv_attention_bwd_tc_workspace_bytes": (c_ll, [C.POINTER(AttnDesc), C.c_int]),
    "clv_attention_bwd_tc": (C.c_int, [C.POINTER(AttnDesc), c_vp, c_vp, c_vp, c_vp, c_vp, C.c_float, c_vp, c_vp, c_vp]),
    "clv_attention_w7_fwd_workspace_bytes": (c_ll, [C.POINTER(AttnW7Desc)]),
    "clv_attention_w7_fwd": (C.c_int, [C.POINTER(AttnW7Desc), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "clv_attention_w7_bwd_workspace_bytes": (c_ll, [C.POINTER(AttnW7Desc), C.c_int]),
    "clv_attention_w7_bwd": (C.c_int, [C.POINTER(AttnW7Desc), c_vp, c_vp, c_vp, c_vp, c_vp, C.c_float, c_vp, c_vp, c_vp]),
    "clv_cast": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_ll, C.c_float, c_vp]),
    "clv_set_tunable": (C.c_int, [C.c_char_p, c_ll]),
    "clv_gelu": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_ll, c_vp]),
    "clv_tanh": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_ll, c_vp]),
    "clv_patchify": (C.c_int, [c_vp, c_vp] + [C.c_int] * 8 + [c_vp]),
    "clv_patchify_u8": (C.c_int, [c_vp, c_vp, c_vp, c_vp] + [C.c_int] * 8 + [c_vp]),
    "clv_grouped_colsum": (C.c_int, [c_vp, C.c_int, c_ll, c_ll, C.c_int, C.c_int, C.c_int, C.c_float, c_vp, C.c_int, c_vp]),
    "clv_rows_affine": (C.c_int, [C.POINTER(RowsAffine), c_vp]),
    "clv_cosine_scores": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, C.c_int, c_vp, c_ll, c_vp, c_vp]),
    "clv_retrieval_ranks": (C.c_int, [c_vp, c_ll, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "clv_adamw_step": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float,
                                 C.c_float, C.c_int, c_vp, c_vp, c_vp]),
    "clv_scatter_add_rows": (C.c_int, [c_vp, c_vp, c_vp, c_ll, C.c_int, c_vp]),
    "clv_nce_workspace_floats": (c_ll, [C.c_int, C.c_int, C.c_int]),
    "clv_nce_rank_fwd": (C.c_int, [C.POINTER(c_vp), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                   c_vp, c_vp, c_vp]),
    "clv_nce_rank_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, c_vp, c_vp, c_vp, C.POINTER(c_vp), c_vp]),
    "clv_softmax_focal_fwd": (C.c_int, [c_vp, c_ll, c_ll, C.c_int, c_vp, c_ll, C.c_float, c_vp, c_vp, c_vp, c_vp]),
    "clv_softmax_focal_bwd": (C.c_int, [c_vp, c_ll, c_ll, C.c_int, C.c_int, c_vp, C.c_float, c_vp, c_vp, c_vp, c_vp,
                                        C.c_int, c_ll, c_vp]),
}

_lib = None


def set_library(path):
    """Developer tooling (A/B runs of two builds on the same box: bench.py --lib): point the loader at another build of
    libclover_b200.so BEFORE the first call.  Not used by the product path."""
    global LIB_PATH, _lib
    if _lib is not None:
        raise RuntimeError("set_library must be called before the library is first loaded")
    LIB_PATH = os.path.abspath(path)


def load():
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C clover_b200/csrc`.  clover_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().clv_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def launch_count():
    return int(load().clv_launch_count())
