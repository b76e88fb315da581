"""In-kernel %globaltimer timeline of tc_gemm_kernel (CTA 0 and the slowest CTA): where does the time go?"""
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
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

def run(name, fn, cold):
    fn(); fn(); torch.cuda.synchronize()
    if cold: flush.zero_()
    ts.zero_(); torch.cuda.synchronize()
    L.upgpt_debug_set_gemm_timestamps(ts.data_ptr())
    fn(); torch.cuda.synchronize()
    L.upgpt_debug_set_gemm_timestamps(None)
    t = ts.cpu().reshape(148, 16)
    act = t[:, 0] > 0
    t0 = t[act, 0].min()
    end = t[act, 11]
    slow = int(torch.nonzero(act)[end.argmax()][0]) if act.any() else 0
    print(f"--- {name} [{'cold' if cold else 'warm'}]: {int(act.sum())} CTAs, kernel span {(t[act, 11].max() - t0).item() / 1e3:.2f} us")
    for c in (0, slow):
        row = t[c]
        print(f"  CTA {c}: " + ", ".join(f"{n}={(row[i] - t0).item() / 1e3:.2f}" for i, n in enumerate(names) if row[i] > 0))

M, N, K = 128, 896, 896
a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.02).half(); out = torch.empty(M, N, device=dev)
for cold in (False, True):
    run("gemm M128 N896 K896 bn64", lambda: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, block_n=64, splits=1, out32=out), cold)
B, H, W, C = 8, 32, 32, 224
x = (torch.randn(B, H, W, C, device=dev) * 0.5).half(); wc = (torch.randn(C, 9, C, device=dev) * 0.02).half()
outc = torch.empty(B * H * W, C, device=dev); bias = torch.randn(C, device=dev)
for cold in (False, True):
    run("conv 224->224 @32 B8", lambda: ops.gemm(a=x, w=wc, mode=_C.GEMM_CONV3X3, N=C, K=C, n_imgs=B, H=H, W=W, out32=outc, bias=bias), cold)
M2 = 8192
a2 = (torch.randn(M2, 224, device=dev) * 0.5).half(); w2 = (torch.randn(224, 224, device=dev) * 0.02).half(); out2 = torch.empty(M2, 224, device=dev); r2 = torch.randn(M2, 224, device=dev)
for cold in (False, True):
    run("gemm M8192 N224 K224 +res", lambda: ops.gemm(a=a2, w=w2, mode=0, M=M2, N=224, K=224, out32=out2, res32=r2), cold)

for (B_, H_, C_) in ((8, 4, 896), (8, 8, 896), (8, 16, 448)):
    x4 = (torch.randn(B_, H_, H_, C_, device=dev) * 0.5).half(); w4 = (torch.randn(C_, 9, C_, device=dev) * 0.02).half()
    o4 = torch.empty(B_ * H_ * H_, C_, device=dev); b4 = torch.randn(C_, device=dev); e4 = torch.randn(B_, C_, device=dev)
    for cold in (False, True):
        run(f"conv {C_}->{C_} @{H_}x{H_} B8 auto", lambda: ops.gemm(a=x4, w=w4, mode=_C.GEMM_CONV3X3, N=C_, K=C_, n_imgs=B_, H=H_, W=H_, out32=o4, bias=b4, rowvec=e4), cold)

# error-compensated mode (UPGPT_GEMM_F_X3): [hi | lo] operand planes, 3 MMAs per k-step on 2 loaded plane pairs
for (B_, H_, C_) in ((8, 32, 224), (8, 16, 448), (8, 8, 896), (8, 4, 896)):
    x4 = (torch.randn(B_, H_, H_, 2 * C_, device=dev) * 0.5).half(); w4 = (torch.randn(C_, 9, 2 * C_, device=dev) * 0.02).half()
    o4 = torch.empty(B_ * H_ * H_, C_, device=dev); b4 = torch.randn(C_, device=dev); e4 = torch.randn(B_, C_, device=dev)
    for cold in (False, True):
        run(f"X3 conv {C_}->{C_} @{H_}x{H_} B8 auto", lambda: ops.gemm(a=x4, w=w4, mode=_C.GEMM_CONV3X3, N=C_, K=C_, n_imgs=B_, H=H_, W=H_, out32=o4, bias=b4, rowvec=e4, flags=_C.GEMM_F_X3), cold)

# GEGLU feed-forward projection at level 0 (M = 8192, N = 2*896 packed [x | gate] per 224-wide tile, K = 224), fp16x3
M3, K3, inner = 8192, 224, 896
a3 = (torch.randn(M3, 2 * K3, device=dev) * 0.5).half(); w3 = (torch.randn(2 * inner, 2 * K3, device=dev) * 0.05).half()
b3 = torch.randn(2 * inner, device=dev); o3 = torch.empty(M3, 2 * inner, device=dev, dtype=torch.half)
for bn in (128, 256):
    for cold in (False, True):
        run(f"X3 GEGLU gemm M8192 N1792 K224 bn{bn}", lambda: ops.gemm(a=a3, w=w3, mode=0, M=M3, N=2 * inner, K=K3, block_n=bn, out16=o3, bias=b3,
                                                                 flags=_C.GEMM_F_GEGLU | _C.GEMM_F_X3 | _C.GEMM_F_SPLIT3OUT), cold)
# the same product without the GEGLU epilogue (plain fp16 hi/lo store of all 1792 columns)
o4 = torch.empty(M3, 4 * inner, device=dev, dtype=torch.half)
for cold in (False, True):
    run("X3 gemm M8192 N1792 K224 (no GEGLU, fp16 [hi|lo] out)", lambda: ops.gemm(a=a3, w=w3, mode=0, M=M3, N=2 * inner, K=K3, out16=o4, bias=b3,
                                                                             flags=_C.GEMM_F_X3 | _C.GEMM_F_SPLIT3OUT), cold)
