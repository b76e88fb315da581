"""Tensor-level wrappers over the C ABI (include/upgpt_b200.h).  Every function takes CUDA torch tensors, borrows their
device pointers for the duration of the call and enqueues kernels on torch's current stream.  No fallbacks."""
import ctypes as C

import torch

from . import _C


def _req(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise _C.UpgptError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise _C.UpgptError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _C.UpgptError(f"{name} must be contiguous")
    return t


def _p(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(**kw):
    """Raw access to upgpt_gemm: keyword = field of upgpt_gemm_args; tensors are replaced by their pointers."""
    a = _C.GemmArgs()
    for k, v in kw.items():
        setattr(a, k, _p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
    _C.check(_C.lib().upgpt_gemm(C.byref(a), stream()), "upgpt_gemm")


def groupnorm_stats(x1, x2, B, HW, stats, groups=32):
    _C.check(_C.lib().upgpt_groupnorm_stats(_p(x1), x1.shape[-1], _p(x2), 0 if x2 is None else x2.shape[-1], B, HW, groups,
                                            _p(stats), stream()), "upgpt_groupnorm_stats")


def prep(**kw):
    a = _C.PrepArgs()
    for k, v in kw.items():
        setattr(a, k, _p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
    _C.check(_C.lib().upgpt_prep_operand(C.byref(a), stream()), "upgpt_prep_operand")


def groupnorm_prep(stats, **kw):
    """Fused GroupNorm(+SiLU)+cast (upgpt_groupnorm_prep): keyword = field of upgpt_prep_args."""
    a = _C.PrepArgs()
    for k, v in kw.items():
        setattr(a, k, _p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
    _C.check(_C.lib().upgpt_groupnorm_prep(C.byref(a), C.c_void_p(stats.data_ptr()), stream()), "upgpt_groupnorm_prep")


def layernorm(x, gamma, beta, out16, eps=1e-5):
    rows, Cc = x.shape[0], x.shape[1]
    _C.check(_C.lib().upgpt_layernorm(_p(x), Cc, rows, Cc, _p(gamma), _p(beta), eps, _p(out16), out16.shape[1], stream()),
             "upgpt_layernorm")


def attention(**kw):
    a = _C.AttnArgs()
    for k, v in kw.items():
        setattr(a, k, _p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
    _C.check(_C.lib().upgpt_attention(C.byref(a), stream()), "upgpt_attention")


def timestep_freqs(dim, max_period=10000, device=None):
    """The reference's frequency table (util.py:163-165), computed with the same fp32 CPU torch ops."""
    import math
    half = dim // 2
    f = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    return f.to(device) if device is not None else f


_FREQS = {}


def timestep_embedding(t, dim, max_period=10000):
    t = _req(t.to(torch.int64), torch.int64, "timesteps")
    key = (dim, max_period, t.device)
    if key not in _FREQS:
        _FREQS[key] = timestep_freqs(dim, max_period, t.device)
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    _C.check(_C.lib().upgpt_timestep_embedding(_p(t), t.shape[0], dim, float(max_period), _p(_FREQS[key]), _p(out), stream()),
             "upgpt_timestep_embedding")
    return out


def linear_small_m(x, weight, bias=None, silu_in=False, silu_out=False, out=None):
    """out[..., n] = act(x[..., :] . W[n, :] + b[n]) for a handful of rows (fp32 weights, exact)."""
    lead = x.shape[:-1]
    x2 = _req(x.reshape(-1, x.shape[-1]).contiguous().float(), torch.float32, "x")
    w = _req(weight.detach().float().contiguous(), torch.float32, "weight")
    b = None if bias is None else _req(bias.detach().float().contiguous(), torch.float32, "bias")
    N, K = w.shape
    if out is None:
        out = torch.empty(x2.shape[0], N, device=x.device, dtype=torch.float32)
    _C.check(_C.lib().upgpt_linear_small_m(_p(x2), K, x2.shape[0], _p(w), _p(b), N, K, int(silu_in), int(silu_out),
                                           _p(out), out.shape[-1], stream()), "upgpt_linear_small_m")
    return out.reshape(*lead, N)


def conv_small_cin(x1, x2, wt_kmajor, bias, Cout, ksize, out, in_scale=1.0, out_nchw=False):
    B, C1, H, W = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    _C.check(_C.lib().upgpt_conv_small_cin(_p(x1), C1, _p(x2), C2, float(in_scale), B, H, W, ksize, _p(wt_kmajor), _p(bias),
                                           Cout, _p(out), int(out_nchw), stream()), "upgpt_conv_small_cin")


def ddim_step(x, eps, coef, x_prev, pred_x0=None, noise=None, noise_step_stride=0, step_ptr=None, step_imm=0):
    _C.check(_C.lib().upgpt_ddim_step(_p(x), _p(eps), _p(noise), noise_step_stride, _p(coef), _p(step_ptr), step_imm,
                                      _p(x_prev), _p(pred_x0), x.numel(), stream()), "upgpt_ddim_step")


def ddpm_step(x, eps, coef, x_prev, pred_x0=None, noise=None, noise_step_stride=0, step_ptr=None, step_imm=0):
    _C.check(_C.lib().upgpt_ddpm_step(_p(x), _p(eps), _p(noise), noise_step_stride, _p(coef), _p(step_ptr), step_imm,
                                      _p(x_prev), _p(pred_x0), x.numel(), stream()), "upgpt_ddpm_step")


def step_state(step_ptr, op, value, t_buf=None, t_table=None):
    B = 0 if t_buf is None else t_buf.numel()
    _C.check(_C.lib().upgpt_step_state(_p(step_ptr), op, value, _p(t_buf), B, _p(t_table), stream()), "upgpt_step_state")


def axpby(a, sa, b, sb, out):
    _C.check(_C.lib().upgpt_axpby(_p(a), float(sa), _p(b), float(sb), _p(out), a.numel(), stream()), "upgpt_axpby")


def qsample_blend(x0, noise, sqrt_acp, sqrt_1m_acp, t=None, t_imm=0, mask=None, img=None, out=None, t_table=None, step_ptr=None,
                  noise_step_stride=0):
    """out = q (* mask + (1 - mask) * img) with q = sqrt_acp[t] x0 + sqrt_1m_acp[t] noise (upgpt_qsample_blend). t: (B,) int64 tensor,
    or (t_table, step_ptr) for the sampler graphs, or the immediate t_imm."""
    x0 = _req(x0.contiguous().float(), torch.float32, "x0")
    B, Cc = x0.shape[0], x0.shape[1]
    HW = x0.numel() // (B * Cc)
    noise = _req(noise.contiguous().float(), torch.float32, "noise")
    if out is None:
        out = torch.empty_like(x0)
    mask_c = 0
    if mask is not None:
        mask = _req(mask.contiguous().float(), torch.float32, "mask")
        mask_c = mask.shape[1]
        if mask.shape[0] != B:
            mask = mask.expand(B, *mask.shape[1:]).contiguous()
        img = _req(img.contiguous().float(), torch.float32, "img")
    if t is not None:
        t = _req(t.to(device=x0.device, dtype=torch.int64).contiguous(), torch.int64, "t")
    _C.check(_C.lib().upgpt_qsample_blend(_p(x0), _p(noise), int(noise_step_stride), _p(mask), mask_c, _p(img), _p(out), _p(sqrt_acp),
                                          _p(sqrt_1m_acp), _p(t), _p(t_table), _p(step_ptr), int(t_imm), B, Cc, HW, stream()),
             "upgpt_qsample_blend")
    return out


def softmax_rows(x, n, scale, out16):
    rows = x.numel() // x.shape[-1]
    _C.check(_C.lib().upgpt_softmax_rows(_p(x), x.shape[-1], rows, n, float(scale), _p(out16), out16.shape[-1], stream()),
             "upgpt_softmax_rows")


def to_uint8_nhwc(x):
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.uint8)
    _C.check(_C.lib().upgpt_to_uint8_nhwc(_p(x), B, Cc, H * W, _p(out), stream()), "upgpt_to_uint8_nhwc")
    return out


class Graph:
    """A captured sequence of upgpt_* calls (CUDA graph)."""

    def __init__(self):
        self.handle = C.c_void_p(0)

    def capture(self, fn):
        """Captures the upgpt_* calls made by fn() (on torch's current stream) into an executable graph.
        The legacy default stream cannot be captured, so the capture happens on a side stream."""
        torch.cuda.synchronize()
        # lane 0: any side stream (scratch slot 0, like the caller's own stream). Lane i > 0: the lane's own library stream, so the
        # captured nodes bake that lane's scratch slot (common.cuh: stream_slot) -- graphs of different lanes replay side by side
        from . import lanes
        side = torch.cuda.Stream() if lanes.current() == 0 else lanes.stream(lanes.current())
        with torch.cuda.stream(side):
            _C.check(_C.lib().upgpt_capture_begin(C.c_void_p(side.cuda_stream)), "capture_begin")
            try:
                fn()
            finally:
                rc = _C.lib().upgpt_capture_end(C.c_void_p(side.cuda_stream), C.byref(self.handle))
            _C.check(rc, "capture_end")
        torch.cuda.synchronize()
        return self

    def launch(self):
        _C.check(_C.lib().upgpt_graph_launch(self.handle, stream()), "graph_launch")

    @property
    def kernels(self):
        return int(_C.lib().upgpt_graph_kernel_count(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _C.lib().upgpt_graph_destroy(self.handle)
        except Exception:
            pass
