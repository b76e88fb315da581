"""Tile-width / split-K sweep over the shapes that carry the B=8 step (round 2), with the real epilogue extras (residual, fp16x3
operands). Same timing method as gpu_sweep_gemm.py (graph of 16 launches rotating over > L2 of operands)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from upgpt_b200 import _C, ops
from upgpt_b200.unet_engine import split3_w
from gpu_sweep_gemm import sweep, dev


def conv(B, H, W, Cin, Cout, x3, res):
    wbytes = (Cout * 9 * Cin * 2 + B * H * W * Cin * 2) * (2 if x3 else 1)
    xs, ws = {}, {}
    out = torch.empty(B * H * W, Cout, device=dev); bias = torch.randn(Cout, device=dev); r = torch.randn(B * H * W, Cout, device=dev)
    cv = (lambda t: split3_w(t)) if x3 else (lambda t: t.half())

    def mk(ncopies, bn, sp):
        for i in range(ncopies):
            if i not in xs:
                xs[i] = cv(torch.randn(B, H, W, Cin, device=dev) * 0.5); ws[i] = cv(torch.randn(Cout, 9, Cin, device=dev) * 0.02)
        return lambda i: ops.gemm(a=xs[i], w=ws[i], mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W, block_n=bn, splits=sp, out32=out,
                                  bias=bias, res32=r if res else None, flags=(_C.GEMM_F_X3 if x3 else 0) | _C.GEMM_F_W_STATIC)
    cfgs = [(bn, sp) for bn in (224, 128, 112, 64) if Cout % bn == 0 for sp in (1, 2, 3, 4, 6, 8)]
    sweep(f"conv B{B} {H}x{W} {Cin}->{Cout}{' x3' if x3 else ''}{' +res' if res else ''}", 2 * B * H * W * 9 * Cin * Cout * (3 if x3 else 1), wbytes, cfgs, mk)


def gemm(M, N, K, x3, res, h16):
    wbytes = (N * K * 2 + M * K * 2) * (2 if x3 else 1)
    xs, ws = {}, {}
    out = torch.empty(M, N, device=dev); bias = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev)
    o16 = torch.empty(M, 2 * N, device=dev, dtype=torch.half)
    cv = (lambda t: split3_w(t)) if x3 else (lambda t: t.half())

    def mk(ncopies, bn, sp):
        for i in range(ncopies):
            if i not in xs:
                xs[i] = cv(torch.randn(M, K, device=dev) * 0.5); ws[i] = cv(torch.randn(N, K, device=dev) * 0.02)
        return lambda i: ops.gemm(a=xs[i], w=ws[i], mode=0, M=M, N=N, K=K, block_n=bn, splits=sp, out32=out, bias=bias, res32=r if res else None,
                                  out16=o16 if h16 else None, ld16=2 * N if h16 else 0,
                                  flags=(_C.GEMM_F_X3 if x3 else 0) | _C.GEMM_F_W_STATIC | (_C.GEMM_F_SPLIT3OUT if h16 else 0))
    cfgs = [(bn, sp) for bn in (256, 224, 128, 112, 96, 64) if N % bn == 0 for sp in (1, 2, 3, 4, 7, 8)]
    sweep(f"gemm M{M} N{N} K{K}{' x3' if x3 else ''}{' +res' if res else ''}{' h16' if h16 else ''}", 2 * M * N * K * (3 if x3 else 1), wbytes, cfgs, mk)


if __name__ == "__main__":
    conv(8, 32, 32, 224, 224, 1, 1)
    conv(8, 16, 16, 448, 448, 1, 1)
    conv(8, 8, 8, 896, 896, 0, 1)
    conv(8, 4, 4, 896, 896, 0, 1)
    gemm(8192, 224, 256, 0, 1, 1)
    gemm(8192, 224, 224, 1, 1, 0)
    gemm(8192, 224, 896, 0, 1, 1)
    gemm(2048, 448, 512, 0, 1, 1)
    gemm(2048, 448, 448, 1, 1, 0)
    gemm(512, 896, 1024, 0, 1, 1)
    gemm(512, 896, 896, 1, 1, 0)
    gemm(512, 896, 3584, 0, 1, 1)
    gemm(128, 896, 1792, 0, 0, 0)
