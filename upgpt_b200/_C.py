"""ctypes binding of libupgpt_b200.so (the C ABI declared in include/upgpt_b200.h).

The product path has NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_lib", "libupgpt_b200.so")
_lib = None


class UpgptError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("mode", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
        ("a_batch_stride", C.c_longlong), ("w_batch_stride", C.c_longlong),
        ("lda", C.c_int), ("ldw", C.c_int),
        ("n_imgs", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("block_n", C.c_int), ("splits", C.c_int),
        ("out32", C.c_void_p), ("ld32", C.c_int),
        ("out16", C.c_void_p), ("ld16", C.c_int),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("ld_rowvec", C.c_int), ("rows_per_group", C.c_int),
        ("res32", C.c_void_p), ("ldres", C.c_int), ("ldT", C.c_int),
        ("flags", C.c_uint), ("out_scale", C.c_float),
        ("rowstats_out", C.c_void_p), ("ln_stats", C.c_void_p), ("ln_slots", C.c_int), ("ln_eps", C.c_float), ("ln_colsum", C.c_void_p),
        ("gn_acc", C.c_void_p), ("gn_groups", C.c_int), ("gn_cpg", C.c_int), ("gn_choff", C.c_int),
        ("gn_acc2", C.c_void_p), ("gn_cpg2", C.c_int), ("gn_choff2", C.c_int),
        ("rowstats_slots", C.c_int),
    ]


GEMM_PLAIN, GEMM_CONV3X3, GEMM_CONV3X3_S2PHASE, GEMM_CONV1X1, GEMM_CONV3X3_S2PHASE_ASYM = 0, 1, 2, 3, 4
GEMM_CONV3X3_UP2 = 5
GEMM_F_GEGLU, GEMM_F_CHW, GEMM_F_SPLIT3OUT, GEMM_F_X3 = 1 << 1, 1 << 2, 1 << 4, 1 << 5
GEMM_F_W_STATIC = 1 << 6


def lib():
    """Loads (building first if sources changed and nvcc is present) and returns the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH) or os.environ.get("UPGPT_REBUILD"):
        from . import build as _build
        _build.build()
    if not os.path.exists(_LIB_PATH):
        raise UpgptError("libupgpt_b200.so missing at %s; run `python -m upgpt_b200.build`" % _LIB_PATH)
    l = C.CDLL(os.environ.get("UPGPT_LIB_PATH") or _LIB_PATH)     # UPGPT_LIB_PATH: bring-up (A/B of two builds inside one GPU call)
    l.upgpt_last_error.restype = C.c_char_p
    l.upgpt_launch_count.restype = C.c_longlong
    _bind(l)
    _lib = l
    return l


def check(rc, what):
    if rc != 0:
        raise UpgptError("%s failed (%d): %s" % (what, rc, lib().upgpt_last_error().decode()))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class PrepArgs(C.Structure):
    _fields_ = [
        ("x1", C.c_void_p), ("C1", C.c_int), ("x2", C.c_void_p), ("C2", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("groups", C.c_int),
        ("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("silu", C.c_int), ("layout", C.c_int), ("split3", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_int), ("raw", C.c_void_p), ("ldraw", C.c_int), ("scale_shift", C.c_void_p),
        ("gn_acc", C.c_void_p), ("raw_planes", C.c_int),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("k_batch_stride", C.c_longlong),
        ("vt", C.c_void_p), ("ldvt", C.c_int), ("out", C.c_void_p), ("ldo", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("dpad", C.c_int), ("scale", C.c_float), ("split3_out", C.c_int),
        ("v_rowmajor", C.c_int), ("v_batch_stride", C.c_longlong),
    ]


_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_PROTOS = {
    "upgpt_gemm": [C.POINTER(GemmArgs), _vp],
    "upgpt_gemm_plan": [C.POINTER(GemmArgs), C.POINTER(C.c_int * 8)],
    "upgpt_gemm_set_sm_weight": [C.c_double],
    "upgpt_set_pdl": [C.c_int],
    "upgpt_debug_set_gemm_timestamps": [_vp],
    "upgpt_trace_set": [_vp],
    "upgpt_groupnorm_stats": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp],
    "upgpt_groupnorm_affine": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp],
    "upgpt_prep_operand": [C.POINTER(PrepArgs), _vp],
    "upgpt_gn_accumulate": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "upgpt_zero": [_vp, _ll, _vp],
    "upgpt_groupnorm_prep": [C.POINTER(PrepArgs), _vp, _vp],
    "upgpt_layernorm": [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp],
    "upgpt_layernorm_split3": [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp],
    "upgpt_softmax_rows": [_vp, _i, _ll, _i, _f, _vp, _i, _vp],
    "upgpt_attention": [C.POINTER(AttnArgs), _vp],
    "upgpt_conv_small_cin": [_vp, _i, _vp, _i, _f, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp],
    "upgpt_timestep_embedding": [_vp, _i, _i, _f, _vp, _vp, _vp],
    "upgpt_linear_small_m": [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "upgpt_ddim_step": [_vp, _vp, _vp, _ll, _vp, _vp, _i, _vp, _vp, _ll, _vp],
    "upgpt_ddpm_step": [_vp, _vp, _vp, _ll, _vp, _vp, _i, _vp, _vp, _ll, _vp],
    "upgpt_step_state": [_vp, _i, _i, _vp, _i, _vp, _vp],
    "upgpt_axpby": [_vp, _f, _vp, _f, _vp, _ll, _vp],
    "upgpt_qsample_blend": [_vp, _vp, _ll, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "upgpt_gather_step_row": [_vp, _ll, _vp, _vp, _i, _i, _vp],
    "upgpt_to_uint8_nhwc": [_vp, _i, _i, _i, _vp, _vp],
    "upgpt_embed_tokens": [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp],
    "upgpt_patchify": [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "upgpt_vit_assemble": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "upgpt_layernorm_f32": [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp],
    "upgpt_quick_gelu_cast": [_vp, _ll, _i, _i, _vp, _i, _vp],
    "upgpt_attention_small": [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _i, _vp],
    "upgpt_gaussian_sample": [_vp, _vp, _f, _vp, _i, _i, _i, _vp],
    "upgpt_lincomb4": [_vp, _f, _vp, _f, _vp, _f, _vp, _f, _f, _vp, _ll, _vp],
    "upgpt_capture_begin": [_vp],
    "upgpt_capture_end": [_vp, C.POINTER(C.c_void_p)],
    "upgpt_graph_launch": [_vp, _vp],
    "upgpt_graph_destroy": [_vp],
    "upgpt_stream_fork": [_vp, _i],
    "upgpt_stream_join": [_vp, _i],
}
EXPORTS = sorted(list(_PROTOS) + ["upgpt_last_error", "upgpt_abi_version", "upgpt_launch_count", "upgpt_graph_kernel_count",
                                  "upgpt_aux_stream"])


def _bind(l):
    for name, argtypes in _PROTOS.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    l.upgpt_graph_kernel_count.argtypes = [_vp]
    l.upgpt_graph_kernel_count.restype = C.c_longlong
    l.upgpt_aux_stream.argtypes = [_i]
    l.upgpt_aux_stream.restype = C.c_void_p
