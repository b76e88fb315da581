"""A few launches of one non-GEMM kernel at the BASELINE configs[1] shape for `ncu --set full` (argv: gn_fused | attn_self | attn_cross)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import _C, ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "gn_fused"
if which == "gn_fused":
    B, H, W, C = 8, 32, 32, 224
    x = torch.randn(B, H * W, C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    st = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64); out = torch.zeros(B * H * W * 2 * C, device=dev, dtype=torch.half)
    fn = lambda: ops.groupnorm_prep(st, x1=x, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, gamma=gamma, beta=beta, eps=1e-5, silu=1, layout=0,
                                    split3=1, out=out, raw=None)
else:
    B, Hh, Nq, d, dpad = 8, 8, 1024, 28, 64
    Nk = 1024 if which == "attn_self" else 87
    Nkp = (Nk + 7) // 8 * 8
    q = (torch.randn(B, Nq, Hh * dpad, device=dev) * 0.5).half(); k = (torch.randn(B, Nk, Hh * dpad, device=dev) * 0.5).half()
    vt = (torch.randn(B, Hh * dpad, Nkp, device=dev) * 0.5).half(); o = torch.zeros(B, Nq, 2 * Hh * dpad, device=dev, dtype=torch.half)
    a = _C.AttnArgs()
    a.q, a.ldq, a.k, a.ldk, a.k_batch_stride, a.vt, a.ldvt, a.out, a.ldo = q.data_ptr(), Hh * dpad, k.data_ptr(), Hh * dpad, 0, vt.data_ptr(), Nkp, o.data_ptr(), 2 * Hh * dpad
    a.B, a.H, a.Nq, a.Nk, a.dpad, a.scale, a.split3_out = B, Hh, Nq, Nk, dpad, d ** -0.5, 1
    import ctypes as C
    fn = lambda: _C.check(_C.lib().upgpt_attention(C.byref(a), ops.stream()), "attn")
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("ok")
