"""Round-2 in-kernel %globaltimer timeline of tc_gemm_kernel for the hot level-0 / level-2 transformer shapes (fp16x3): plain vs LayerNorm
producer (raw planes + row stats) vs LayerNorm consumer epilogues."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import _C, ops
dev = torch.device("cuda:0")
L = _C.lib()
names = ["start", "setup done", "prod before first TMA", "prod last issue", "mma first full", "mma second full", "mma last full", "mma tile committed",
         "epi tfull", "epi stores issued", "all joined", "dealloc done", "c0 tmem loaded | splitK partials fenced", "c0 staged | siblings arrived", "c0 barrier | my slice reduced", "c0 flushed | slice barrier"]
ts = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


def run(name, fn):
    fn(); fn(); torch.cuda.synchronize()
    ts.zero_(); torch.cuda.synchronize()
    L.upgpt_debug_set_gemm_timestamps(ts.data_ptr())
    fn(); torch.cuda.synchronize()
    L.upgpt_debug_set_gemm_timestamps(None)
    t = ts.cpu().reshape(148, 16)
    act = t[:, 0] > 0
    t0 = t[act, 0].min()
    end = t[act, 11]
    slow = int(torch.nonzero(act)[end.argmax()][0]) if act.any() else 0
    print(f"--- {name}: {int(act.sum())} CTAs, kernel span {(t[act, 11].max() - t0).item() / 1e3:.2f} us")
    for c in (0, slow):
        row = t[c]
        print(f"  CTA {c}: " + ", ".join(f"{n}={(row[i] - t0).item() / 1e3:.2f}" for i, n in enumerate(names) if row[i] > 0))


def mk(M, N, K, x3=True):
    kx = 2 if x3 else 1
    a = (torch.randn(M, kx * K, device=dev) * 0.5).half(); w = (torch.randn(N, kx * K, device=dev) * 0.05).half()
    return a, w


X3 = _C.GEMM_F_X3
for (M, N, K, tag) in ((8192, 224, 256, "attn.out L0"), (8192, 224, 896, "ff2 L0"), (2048, 448, 512, "attn.out L1"), (512, 896, 1024, "attn.out L2 (x1)")):
    x3 = M > 512
    a, w = mk(M, N, K, x3)
    out = torch.empty(M, N, device=dev); res = torch.randn(M, N, device=dev); b = torch.randn(N, device=dev)
    raw = torch.empty(M, 2 * N, device=dev, dtype=torch.half); st = torch.empty(M * 16 * 2, device=dev)
    fl = X3 if x3 else 0
    run(f"{tag}: M{M} N{N} K{K} +res", lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out32=out, res32=res, bias=b, flags=fl))
    run(f"{tag}: M{M} N{N} K{K} +res + raw planes + rowstats (LN producer)",
        lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out32=out, res32=res, bias=b, out16=raw, rowstats_out=st, flags=fl | _C.GEMM_F_SPLIT3OUT))
for (M, N, K, tag) in ((8192, 768, 224, "qkv L0"), (8192, 256, 224, "q L0"), (2048, 1536, 448, "qkv L1")):
    a, w = mk(M, N, K)
    o16 = torch.empty(M, N, device=dev, dtype=torch.half); b = torch.randn(N, device=dev); cs = torch.randn(N, device=dev)
    st = torch.rand(M * 2 * 2, device=dev) + 1.0
    run(f"{tag}: M{M} N{N} K{K} f16 out", lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out16=o16, flags=X3))
    run(f"{tag}: M{M} N{N} K{K} f16 out, folded LN", lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out16=o16, bias=b, ln_stats=st, ln_slots=2, ln_eps=1e-5,
                                                                      ln_colsum=cs, flags=X3))
M3, K3, inner = 8192, 224, 896
a3, w3 = mk(M3, 2 * inner, K3)
b3 = torch.randn(2 * inner, device=dev); o3 = torch.empty(M3, 2 * inner, device=dev, dtype=torch.half); cs = torch.randn(2 * inner, device=dev)
st = torch.rand(M3 * 2 * 2, device=dev) + 1.0
run("GEGLU M8192 N1792 K224 bn224", lambda: ops.gemm(a=a3, w=w3, mode=0, M=M3, N=2 * inner, K=K3, block_n=224, out16=o3, bias=b3, flags=_C.GEMM_F_GEGLU | X3 | _C.GEMM_F_SPLIT3OUT))
run("GEGLU M8192 N1792 K224 bn224, folded LN", lambda: ops.gemm(a=a3, w=w3, mode=0, M=M3, N=2 * inner, K=K3, block_n=224, out16=o3, bias=b3, ln_stats=st, ln_slots=2,
                                                                 ln_eps=1e-5, ln_colsum=cs, flags=_C.GEMM_F_GEGLU | X3 | _C.GEMM_F_SPLIT3OUT))
B, H, W, C = 8, 32, 32, 224
x = (torch.randn(B, H, W, 2 * C, device=dev) * 0.5).half(); wc = (torch.randn(C, 9, 2 * C, device=dev) * 0.02).half()
outc = torch.empty(B * H * W, C, device=dev); bias = torch.randn(C, device=dev); e4 = torch.randn(B, C, device=dev); r4 = torch.randn(B * H * W, C, device=dev)
run("X3 conv 224->224 @32 +rowvec (conv1)", lambda: ops.gemm(a=x, w=wc, mode=_C.GEMM_CONV3X3, N=C, K=C, n_imgs=B, H=H, W=W, out32=outc, bias=bias, rowvec=e4, flags=X3))
run("X3 conv 224->224 @32 +res (conv2)", lambda: ops.gemm(a=x, w=wc, mode=_C.GEMM_CONV3X3, N=C, K=C, n_imgs=B, H=H, W=W, out32=outc, bias=bias, res32=r4, flags=X3))
