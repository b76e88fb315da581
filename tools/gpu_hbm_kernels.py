"""HBM-side kernels at the VAE decoder's largest level (B=8, 256x256x128 fp32 = 268 MB): prep (GroupNorm apply + swish + cast) and
gn_stats, timed alone with CUDA events (operands exceed L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import ops
dev = torch.device("cuda:0")
B, H, W, C = 8, 256, 256, 128
x = torch.randn(B, H * W, C, device=dev); ss = torch.randn(B, 2, C, device=dev)
out = torch.empty(B * H * W * C, device=dev, dtype=torch.half); st = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64)
def t(fn, n=7):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]
ms = t(lambda: ops.prep(x1=x, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, stats=None, gamma=None, beta=None, eps=1e-6, silu=1, layout=0, split3=0, out=out, raw=None, scale_shift=ss))
print("prep   chunk_max=%s: %.1f us  %.0f GB/s (read 4B + write 2B per element)" % (os.environ.get("UPGPT_PREP_CHUNK_MAX", "64"), ms * 1e3, B * H * W * C * 6 / ms / 1e6))
ms = t(lambda: ops.groupnorm_stats(x, None, B, H * W, st))
print("gn_stats: %.1f us  %.0f GB/s (read 4B per element)" % (ms * 1e3, B * H * W * C * 4 / ms / 1e6))
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x))
print("torch copy_ fp32 (read 4B + write 4B): %.1f us  %.0f GB/s" % (ms * 1e3, B * H * W * C * 8 / ms / 1e6))
