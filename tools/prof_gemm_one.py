"""A few tc_gemm launches for `ncu --set full` (shape chosen by argv: conv224 | gemm_small | gemm_k224)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import _C, ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "conv224"
if which == "conv224":
    B, H, W, C = 8, 32, 32, 224
    x3 = int(os.environ.get("X3", 0))    # 1 = the error-compensated fp16x3 mode the parity gate runs ([hi | lo] operand planes)
    kx = 2 if x3 else 1
    x = (torch.randn(B, H, W, C * kx, device=dev) * 0.5).half(); w = (torch.randn(C, 9, C * kx, device=dev) * 0.02).half()
    out = torch.empty(B * H * W, C, device=dev); bias = torch.randn(C, device=dev)
    fn = lambda: ops.gemm(a=x, w=w, mode=_C.GEMM_CONV3X3, N=C, K=C, n_imgs=B, H=H, W=W, out32=out, bias=bias, flags=_C.GEMM_F_X3 if x3 else 0)
elif which == "conv_deep":
    # the weight-bandwidth-bound ResBlock conv of the 4x4 level (896 -> 896, M = 8 x 16 = 128 rows, K = 9 x 896): 14.5 MB of fp16 weights
    # (single plane in the calibrated plan) against 0.23 MB of activations. 8 weight sets (116 MB) rotate so the weights come from HBM.
    B, H, W, C = 8, 4, 4, 896
    x = (torch.randn(B, H, W, C, device=dev) * 0.5).half()
    ws = [(torch.randn(C, 9, C, device=dev) * 0.02).half() for _ in range(8)]
    out = torch.empty(B * H * W, C, device=dev); bias = torch.randn(C, device=dev); e = torch.randn(B, C, device=dev)
    it = [0]
    def fn():
        it[0] += 1
        ops.gemm(a=x, w=ws[it[0] % 8], mode=_C.GEMM_CONV3X3, N=C, K=C, n_imgs=B, H=H, W=W, out32=out, bias=bias, rowvec=e)
elif which == "up2":
    # Upsample of the 16x16 -> 32x32 level (448 -> 448, fp16x3) as four parity-wise 2x2 convolutions over the low-resolution operand
    from upgpt_b200.unet_engine import split3_w, up2_conv_w
    B, H, W, C = 8, 16, 16, 448
    x = split3_w(torch.randn(B, H, W, C, device=dev) * 0.5); w = split3_w(up2_conv_w(torch.randn(C, C, 3, 3, device=dev) * 0.02))
    out = torch.empty(B * 4 * H * W, C, device=dev); bias = torch.randn(C, device=dev)
    fn = lambda: ops.gemm(a=x, w=w, mode=_C.GEMM_CONV3X3_UP2, N=C, K=C, n_imgs=B, H=H, W=W, out32=out, bias=bias, flags=_C.GEMM_F_X3 | _C.GEMM_F_W_STATIC)
elif which == "geglu":
    # GEGLU feed-forward projection of the 32x32 level (attention.py:37-44): M 8192, N 2 x 896, K 224, single-plane operands, folded LayerNorm
    M3, K3, inner = 8192, 224, 896
    a3 = (torch.randn(M3, K3, device=dev) * 0.5).half(); w3 = (torch.randn(2 * inner, K3, device=dev) * 0.05).half()
    b3 = torch.randn(2 * inner, device=dev); o3 = torch.empty(M3, inner, device=dev, dtype=torch.half); cs = torch.randn(2 * inner, device=dev)
    st = torch.rand(M3 * 2 * 2, device=dev) + 1.0
    fn = lambda: ops.gemm(a=a3, w=w3, mode=0, M=M3, N=2 * inner, K=K3, out16=o3, bias=b3, ln_stats=st, ln_slots=2, ln_eps=1e-5, ln_colsum=cs,
                          flags=_C.GEMM_F_GEGLU | _C.GEMM_F_W_STATIC)
elif which == "gemm_small":
    M, N, K = 128, 896, 896
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.02).half(); out = torch.empty(M, N, device=dev)
    fn = lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, block_n=64, splits=1, out32=out)
else:
    M, N, K = 8192, 224, 224
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.02).half(); out = torch.empty(M, N, device=dev)
    r = torch.randn(M, N, device=dev)
    fn = lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out32=out, res32=r)
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("ok")
