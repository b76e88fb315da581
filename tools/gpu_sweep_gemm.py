"""Tile-shape / split-K sweep of tc_gemm_kernel over the U-Net's layer shapes (B=8).
Each config is timed as a CUDA graph of REP back-to-back launches rotating over enough weight/activation copies to
exceed the 126 MB L2 (weights are cold in the real step: 850 MB per U-Net pass), so host launch latency is excluded."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import _C, ops

dev = torch.device("cuda:0")
REP = 16


def time_graph(make_call, ncopies):
    """make_call(i) enqueues one launch using copy i % ncopies."""
    for i in range(min(ncopies, 2)): make_call(i)
    torch.cuda.synchronize()
    g = ops.Graph().capture(lambda: [make_call(i % ncopies) for i in range(REP)])
    g.launch(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.launch(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / REP)
    return best


def sweep(name, fl, bytes_per_copy, cfgs, mk):
    ncopies = max(2, min(16, int(200e6 / max(bytes_per_copy, 1)) + 1))
    res = []
    for cfg in cfgs:
        try:
            res.append((time_graph(mk(ncopies, *cfg), ncopies), cfg))
        except Exception as e:
            pass
    auto = time_graph(mk(ncopies, 0, 0), ncopies)
    res.sort()
    best = ", ".join("bn%d sp%d %.1fus" % (c[0], c[1], u) for u, c in res[:5])
    print(f"{name}: auto {auto:.1f}us ({fl / auto / 1e6:.0f} TF/s) | best {best} ({fl / res[0][0] / 1e6:.0f} TF/s)", flush=True)


def conv(B, H, W, Cin, Cout):
    wbytes = Cout * 9 * Cin * 2 + B * H * W * Cin * 2
    xs, ws = {}, {}
    out = torch.empty(B * H * W, Cout, device=dev); bias = torch.randn(Cout, device=dev)

    def mk(ncopies, bn, sp):
        for i in range(ncopies):
            if i not in xs:
                xs[i] = (torch.randn(B, H, W, Cin, device=dev) * 0.5).half(); ws[i] = (torch.randn(Cout, 9, Cin, device=dev) * 0.02).half()
        return lambda i: ops.gemm(a=xs[i], w=ws[i], mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W, block_n=bn, splits=sp, out32=out, bias=bias)
    cfgs = [(bn, sp) for bn in (256, 224, 128, 112, 64) if Cout % bn == 0 for sp in (1, 2, 3, 4, 6, 9, 12, 18)]
    sweep(f"conv B{B} {H}x{W} {Cin}->{Cout}", 2 * B * H * W * 9 * Cin * Cout, wbytes, cfgs, mk)


def gemm(M, N, K):
    wbytes = N * K * 2 + M * K * 2
    xs, ws = {}, {}
    out = torch.empty(M, N, device=dev); bias = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev)

    def mk(ncopies, bn, sp):
        for i in range(ncopies):
            if i not in xs:
                xs[i] = (torch.randn(M, K, device=dev) * 0.5).half(); ws[i] = (torch.randn(N, K, device=dev) * 0.02).half()
        return lambda i: ops.gemm(a=xs[i], w=ws[i], mode=0, M=M, N=N, K=K, block_n=bn, splits=sp, out32=out, bias=bias, res32=r)
    cfgs = [(bn, sp) for bn in (256, 224, 128, 112, 64) if N % bn == 0 for sp in (1, 2, 4, 7, 8)]
    sweep(f"gemm M{M} N{N} K{K}", 2 * M * N * K, wbytes, cfgs, mk)


if __name__ == "__main__":
    conv(8, 32, 32, 224, 224)
    conv(8, 32, 32, 448, 224)
    conv(8, 32, 32, 672, 224)
    conv(8, 16, 16, 448, 448)
    conv(8, 16, 16, 1344, 448)
    conv(8, 8, 8, 896, 896)
    conv(8, 8, 8, 1792, 896)
    conv(8, 4, 4, 896, 896)
    conv(8, 4, 4, 1792, 896)
    gemm(8192, 224, 224)
    gemm(8192, 1024, 224)
    gemm(8192, 512, 224)
    gemm(8192, 224, 896)
    gemm(8192, 224, 512)
    gemm(2048, 448, 448)
    gemm(2048, 448, 1792)
    gemm(512, 896, 896)
    gemm(512, 896, 3584)
    gemm(512, 2048, 896)
    gemm(128, 896, 896)
    gemm(696, 512, 768)
