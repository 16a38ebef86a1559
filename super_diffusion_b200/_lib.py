"""ctypes loader for ``libsuperdiff_b200.so`` (the C ABI in include/superdiff_b200.h).

There is no fallback: if the shared library has not been built, or a call
fails, a RuntimeError is raised.  Nothing here imports ``oracle/``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsuperdiff_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_void_p = ctypes.c_void_p


class GemmSrc(ctypes.Structure):
    """struct sd_gemm_src (include/superdiff_b200.h)."""
    _fields_ = [("ptr", ctypes.c_void_p), ("C", ctypes.c_int), ("taps", ctypes.c_int), ("ld", ctypes.c_int)]


class ScoreNetDesc(ctypes.Structure):
    """struct sd_scorenet_desc (include/superdiff_b200.h)."""
    _fields_ = [("image_size", ctypes.c_int), ("channels", ctypes.c_int), ("nf", ctypes.c_int),
                ("num_res_blocks", ctypes.c_int), ("n_levels", ctypes.c_int), ("ch_mult", ctypes.c_int * 8),
                ("n_attn_res", ctypes.c_int), ("attn_resolutions", ctypes.c_int * 8), ("conditioned", ctypes.c_int),
                ("num_classes", ctypes.c_int), ("weights", ctypes.c_void_p), ("weights_bytes", ctypes.c_size_t),
                ("precision", ctypes.c_int)]


# name -> (restype, argtypes); must list every symbol include/superdiff_b200.h declares
_F, _I, _V, _U, _SZ = ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint, ctypes.c_size_t
_LL = ctypes.c_longlong
SIGNATURES = {
    "sd_step_vpsde": (_I, [_V, _V, ctypes.POINTER(_V), _I, _I, _I, _F, _F, _F, _F, _V, _V,
                           _I, _I, _F, _V, _F, _V, _V, _V, _V]),
    "sd_step_vpsde_ex": (_I, [_V, _V, ctypes.POINTER(_V), _I, _I, _I, _F, _F, _F, _F, _V, _V,
                              _I, _I, _F, _V, _F, _V, _V, _V, _V, _I, _I, _I]),
    "sd_step_vpsde_ode": (_I, [_V, ctypes.POINTER(_V), _I, _I, _I, _F, _F, _F, _F, _V, _V, _I, _I, _F, _V, _V, _V, _V, _V, _V]),
    "sd_rowdot": (_I, [_V, _V, _I, _I, _F, _V, _I, _V]),
    "sd_groupnorm_swish_jvp": (_I, [_V, _V, _I, _V, _V, _I, _I, _I, _V, _V, _F, _I, _V, _SZ, _V, _V, _V]),
    "sd_softmax_jvp": (_I, [_V, _V, _V, _F, _V, ctypes.c_long, _I, _V]),
    "sd_step_edm_cfg": (_I, [_V, _V, _V, _V, _V, _I, _I, _F, _F, _F, _F, _I, _F, _F, _F, _V, _V, _V, _V]),
    "sd_step_edm_ode": (_I, [_V, _V, _V, _V, _V, _I, _I, _F, _F, _F, _F, _V, _V, _V, _V]),
    "sd_counter_add": (_I, [_V, _I, _V]),
    "sd_counter_add_sat": (_I, [_V, _I, _I, _V]),
    "sd_conv_gemm": (_I, [ctypes.POINTER(GemmSrc), _I, _I, _I, _I, _V, _I, _V, _V, _I, _V, _U, _V, _I, _V, _V]),
    "sd_conv_gemm_gn": (_I, [ctypes.POINTER(GemmSrc), _I, _I, _I, _I, _V, _I, _V, _V, _I, _U, _V, _I, _V, _V, _V, _F, _I,
                             _V, ctypes.POINTER(_I), _V]),
    "sd_set_gn_fuse": (_I, [_I]),
    "sd_conv_gemm_s2": (_I, [_V, _I, _I, _I, _I, _V, _I, _V, _U, _V, _V, _V]),
    "sd_upconv_gemm": (_I, [_V, _I, _I, _I, _I, _V, _I, _V, _U, _V, _V, _V]),
    "sd_groupnorm_swish": (_I, [_V, _I, _V, _I, _I, _I, _V, _V, _F, _I, _V, _I, _V, _I, _V, _SZ, _V, _V]),
    "sd_groupnorm_swish_ex": (_I, [_V, _I, _V, _I, _I, _I, _V, _V, _F, _I, _V, _I, _V, _I, _V, _SZ, _V, _U, _V]),
    "sd_attention": (_I, [_V, _I, _I, _I, _V, _V]),
    "sd_upsample2x": (_I, [_V, _I, _I, _I, _I, _V, _V]),
    "sd_im2col_s2": (_I, [_V, _I, _I, _I, _I, _V, _V]),
    "sd_gather_row": (_I, [_V, _I, _I, _V, _V, _V]),
    "sd_im2col_in": (_I, [_V, _I, _I, _I, _I, _V, _V]),
    "sd_im2col_in_ex": (_I, [_V, _I, _I, _I, _I, _V, _U, _V]),
    "sd_softmax_rows_split": (_I, [_V, _V, ctypes.c_long, _I, _F, _I, _I, _V]),
    "sd_time_embedding_ex": (_I, [_V, _I, _V, _V, _I, _I, _V, _V, _V, _V, _V, _V, _V, _V, _U, _V]),
    "sd_conv_in": (_I, [_V, _I, _I, _I, _I, _V, _V, _I, _V, _V]),
    "sd_time_embedding": (_I, [_V, _I, _V, _V, _I, _I, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    "sd_batched_gemm": (_I, [_V, _I, _LL, _V, _I, _LL, _I, _I, _I, _I, _V, _V, _U, _V, _I, _LL, _V]),
    "sd_batched_gemm_stats": (_I, [_V, _I, _LL, _V, _I, _LL, _I, _I, _I, _I, _V, _V, _U, _V, _I, _LL, _V, _V]),
    "sd_attention_core": (_I, [_V, _I, _LL, _V, _I, _LL, _V, _I, _LL, _I, _I, _I, _F, _I, _V, _V, _V, _V, _V]),
    "sd_attention_probs": (_I, [_V, _I, _LL, _V, _I, _LL, _I, _I, _I, _F, _I, _V, _V]),
    "sd_softmax_rows": (_I, [_V, _V, ctypes.c_long, _I, _F, _V]),
    "sd_cast_f32_to_bf16": (_I, [_V, _V, _SZ, _V]),
    "sd_cast_bf16_to_f32": (_I, [_V, _V, _SZ, _V]),
    "sd_scorenet_weights_bytes": (_I, [ctypes.POINTER(ScoreNetDesc), ctypes.POINTER(_SZ)]),
    "sd_scorenet_workspace_bytes": (_I, [ctypes.POINTER(ScoreNetDesc), _I, _I, ctypes.POINTER(_SZ)]),
    "sd_scorenet_forward": (_I, [ctypes.POINTER(ScoreNetDesc), _V, _I, _V, _V, _I, _V, _V, _SZ, _I, _V]),
    "sd_scorenet_forward_sched": (_I, [ctypes.POINTER(ScoreNetDesc), _V, _V, _V, _V, _I, _V, _V, _SZ, _I, _V]),
    "sd_last_error": (ctypes.c_char_p, []),
    "sd_version": (_I, []),
    "sd_device_ok": (_I, []),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built. "
            "Run `python -m super_diffusion_b200.build` (or __graft_entry__.build()). "
            "There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sd_last_error()
        raise RuntimeError(f"superdiff_b200 {what} failed (code {rc}): {msg.decode() if msg else '?'}")


def require_device():
    """Raise unless the current CUDA device is an sm_100-family GPU."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("superdiff_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                           "there is no CPU fallback")
    if not load().sd_device_ok():
        raise RuntimeError("superdiff_b200 kernels are compiled for sm_100a only; the current device is not "
                           "compute capability 10.x")
