"""ctypes binding of the C-ABI in ``include/pd_b200.h`` (``libpd_b200.so``, built in-tree
by ``build.sh`` / ``__graft_entry__.build()``).

There is no CPU or PyTorch fallback: if the library is missing, importing this module
raises, and every op raises ``RuntimeError`` with ``pd_last_error()`` on a non-zero return.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PD_B200_LIB selects another build of the same library (kernel A/B experiments under scripts/); never a fallback
LIB_PATH = os.environ.get("PD_B200_LIB") or os.path.join(_HERE, "libpd_b200.so")

PD_F32, PD_BF16 = 0, 1
PD_ACT_NONE, PD_ACT_SILU, PD_ACT_GEGLU = 0, 1, 2
PD_ENGINE_AUTO, PD_ENGINE_SIMT, PD_ENGINE_TC = 0, 1, 2
PD_ATTN_AUTO, PD_ATTN_SIMT, PD_ATTN_MMA, PD_ATTN_TC, PD_ATTN_SHORT = 0, 1, 2, 3, 4


class ConvParams(C.Structure):
    """Mirror of ``pd_conv_params`` (include/pd_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("x2", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("rowvec", C.c_void_p), ("res", C.c_void_p), ("out", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("C2", C.c_int32), ("Cout", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("upsample", C.c_int32),
        ("ldx", C.c_int32), ("ldx2", C.c_int32), ("ldr", C.c_int32), ("ldo", C.c_int32),
        ("ldrv", C.c_int32),
        ("act", C.c_int32), ("dtype", C.c_int32), ("out_dtype", C.c_int32), ("engine", C.c_int32),
        ("alpha", C.c_float), ("w_blocked", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p),
        ("ln_parts", C.c_void_p), ("ln_parts_out", C.c_void_p), ("gn_stats_out", C.c_void_p),
        ("ln_rows", C.c_int64), ("out_sx", C.c_int64), ("out_sy", C.c_int64), ("out_sb", C.c_int64),
        ("ln_eps", C.c_float), ("gn_ld", C.c_int32), ("gn_rec_off", C.c_int32), ("gn_recs_per_image", C.c_int32),
        ("pad_y", C.c_int32), ("pad_x", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/pd_b200.h
SIGNATURES = {
    "pd_last_error": (C.c_char_p, []),
    "pd_abi_version": (C.c_int, []),
    "pd_launch_count": (C.c_uint64, []),
    "pd_device_is_sm100": (C.c_int, []),
    "pd_prof_enable": (C.c_int, [C.c_int]),
    "pd_prof_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "pd_prof_dump": (C.c_int, [C.c_char_p]),
    "pd_debug_tile_order": (C.c_int, [C.c_int32]),
    "pd_tune_dump": (C.c_int, [C.c_char_p]),
    "pd_debug_force_cta_group": (C.c_int, [C.c_int32]),
    "pd_debug_force_bn": (C.c_int, [C.c_int32]),
    "pd_debug_force_bres": (C.c_int, [C.c_int32]),
    "pd_debug_bres_launches": (C.c_uint64, []),
    "pd_debug_force_vh": (C.c_int, [C.c_int32]),
    "pd_debug_vh_launches": (C.c_uint64, []),
    "pd_debug_force_stream_k": (C.c_int, [C.c_int32]),
    "pd_layer_norm_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "pd_debug_group_norm_fused": (C.c_int, [C.c_int32]),
    "pd_debug_attention_timeline": (C.c_int, [C.c_void_p]),
    "pd_debug_attention_tc4": (C.c_int, [C.c_int32]),
    "pd_debug_attention_tc3": (C.c_int, [C.c_int32]),
    "pd_debug_attention_persistent": (C.c_int, [C.c_int32]),
    "pd_conv2d": (C.c_int, [C.POINTER(ConvParams), C.c_void_p]),
    "pd_conv2d_ln_parts_floats": (C.c_int64, [C.c_int64]),
    "pd_conv2d_gn_stats_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "pd_repack_conv_weight": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p]),
    "pd_group_norm_scratch_floats": (C.c_int64, [C.c_int32]),
    "pd_group_norm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_group_norm_apply": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_layer_norm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "pd_attention_causal": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                      C.c_void_p]),
    "pd_embedding_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_quick_gelu": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                C.c_void_p]),
    "pd_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float,
                                  C.c_int32, C.c_void_p]),
    "pd_geglu": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                           C.c_void_p]),
    "pd_attention": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                               C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "pd_attention_ex": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_timestep_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_float, C.c_int32, C.c_void_p]),
    "pd_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "pd_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "pd_nchw_to_nhwc_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p]),
    "pd_add2d": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                           C.c_void_p]),
    "pd_cast2d": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                            C.c_int32, C.c_void_p]),
    "pd_upsample2x": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pd_cfg_ddim_step": (C.c_int, [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]),
}


# exported by PD_DEBUG builds only (scripts/build_variant.sh, selected with PD_B200_LIB); bound when present
DEBUG_SIGNATURES = {
    "pd_debug_timeline": (C.c_int, [C.c_void_p]),
    "pd_debug_gemm_mode": (C.c_int, [C.c_int32]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with ./build.sh (or __graft_entry__.build()). "
            "prompt_diffusion_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in DEBUG_SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return lib.pd_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")


_graph_launches = 0


def note_graph_replay(n_kernels: int) -> None:
    """A replayed CUDA graph re-issues ``n_kernels`` of this library's kernels without passing through the
    C-ABI entry points; keep the launch counter honest."""
    global _graph_launches
    _graph_launches += int(n_kernels)


def launch_count() -> int:
    """Kernels of libpd_b200 launched so far: direct launches (counted inside the library) + kernels re-issued
    by CUDA-graph replays of the denoising step."""
    return int(lib.pd_launch_count()) + _graph_launches
