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
elif which == "gn_group":
    # the one-CTA-per-(image, 4 groups) form at the 8x8 level (896 channels, B = 8)
    B, H, W, C = 8, 8, 8, 896
    x = torch.randn(B, H * W, C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    st = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64); out = torch.zeros(B * H * W * C, device=dev, dtype=torch.half)
    fn = lambda: ops.groupnorm_prep(st, x1=x, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, gamma=gamma, beta=beta, eps=1e-5, silu=1, layout=0,
                                    split3=0, out=out, raw=None)
elif which == "prep":
    # GroupNorm apply + swish + fp16 cast at the VAE decoder's 256x256x128 level, B = 8 (bench.py's roofline_hbm kernel)
    B, H, W, C = 8, 256, 256, 128
    x = torch.randn(B, H * W, C, device=dev); ss = torch.randn(B, 2, C, device=dev)
    out = torch.empty(B * H * W * C, device=dev, dtype=torch.half)
    fn = lambda: ops.prep(x1=x, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, stats=None, gamma=None, beta=None, eps=1e-6, silu=1,
                          layout=0, split3=0, out=out, raw=None, scale_shift=ss)
else:
    # level-0 attention cores as the engine runs them: 8 heads of 28 padded to 32 (head pairs in 64-wide rows), V row-major
    B, Hh, Nq, d, dpad = 8, 8, 1024, 28, 32
    HD = Hh * dpad
    if which == "attn_self":
        qkv = (torch.randn(B * Nq, 3 * HD, device=dev) * 0.5).half()
        kw = dict(q=qkv, ldq=3 * HD, k=qkv.reshape(-1)[HD:], ldk=3 * HD, k_batch_stride=Nq * 3 * HD, vt=qkv.reshape(-1)[2 * HD:], ldvt=3 * HD,
                  v_rowmajor=1, v_batch_stride=Nq * 3 * HD, Nk=Nq)
    else:
        Nk = 87
        qq = (torch.randn(B * Nq, HD, device=dev) * 0.5).half(); kv = (torch.randn(B * Nk, 2 * HD, device=dev) * 0.5).half()
        kw = dict(q=qq, ldq=HD, k=kv, ldk=2 * HD, k_batch_stride=Nk * 2 * HD, vt=kv.reshape(-1)[HD:], ldvt=2 * HD, v_rowmajor=1,
                  v_batch_stride=Nk * 2 * HD, Nk=Nk)
    o = torch.zeros(B * Nq, 2 * HD, device=dev, dtype=torch.half)
    fn = lambda: ops.attention(out=o, ldo=2 * HD, B=B, H=Hh, Nq=Nq, dpad=dpad, scale=float(d) ** -0.5, split3_out=1, **kw)
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("ok")
