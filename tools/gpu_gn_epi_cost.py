"""Cost of the GroupNorm-moment epilogue (upgpt_gemm gn_acc) on the producer GEMMs of the B=8 step: each shape timed as a graph of 16
back-to-back launches with and without gn_acc (UPGPT_GN_DBG=1: without the atomics, 2: without the commit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from upgpt_b200 import _C, ops
from upgpt_b200.unet_engine import split3_w
from gpu_sweep_gemm import time_graph, dev

def conv(B, H, W, Cin, Cout, x3, res, two):
    cv = (lambda t: split3_w(t)) if x3 else (lambda t: t.half())
    n = 6
    xs = [cv(torch.randn(B, H, W, Cin, device=dev) * 0.5) for _ in range(n)]; ws = [cv(torch.randn(Cout, 9, Cin, device=dev) * 0.02) for _ in range(n)]
    out = torch.empty(B * H * W, Cout, device=dev); bias = torch.randn(Cout, device=dev); r = torch.randn(B * H * W, Cout, device=dev)
    acc = torch.zeros(B, 32, 2, device=dev, dtype=torch.int64); acc2 = torch.zeros(B, 32, 2, device=dev, dtype=torch.int64)
    res_ = {}
    for gn in (0, 1):
        kw = dict(mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W, out32=out, bias=bias, res32=r if res else None,
                  flags=(_C.GEMM_F_X3 if x3 else 0) | _C.GEMM_F_W_STATIC)
        if gn:
            kw.update(gn_acc=acc, gn_groups=32, gn_cpg=Cout // 32, gn_choff=0)
            if two: kw.update(gn_acc2=acc2, gn_cpg2=Cout // 16, gn_choff2=0)
            # pin a power-of-two split like the engine does
            import ctypes as C
            a = _C.GemmArgs()
            for k, v in dict(kw, a=xs[0], w=ws[0]).items():
                setattr(a, k, v.data_ptr() if isinstance(v, torch.Tensor) else (0 if v is None else v))
            a.gn_acc = 0; a.gn_acc2 = 0
            plan = (C.c_int * 8)(); _C.check(_C.lib().upgpt_gemm_plan(C.byref(a), C.byref(plan)), "plan")
            sp = 1
            while sp * 2 <= plan[2]: sp *= 2
            kw["splits"] = sp; kw["block_n"] = int(plan[0])
            res_["plan"] = (plan[0], plan[2], sp)
        res_[gn] = time_graph(lambda i: ops.gemm(a=xs[i], w=ws[i], **kw), n)
    print(f"conv B{B} {H}x{W} {Cin}->{Cout}{' x3' if x3 else ''}{' +res' if res else ''}{' 2 consumers' if two else ''}: "
          f"plain {res_[0]:.1f} us, gn_acc {res_[1]:.1f} us (bn, splits, pinned = {res_.get('plan')})", flush=True)

print("UPGPT_GN_DBG =", os.environ.get("UPGPT_GN_DBG", "0"))
conv(8, 32, 32, 224, 224, 1, 1, 0)
conv(8, 32, 32, 224, 224, 1, 1, 1)
conv(8, 16, 16, 448, 448, 1, 1, 0)
conv(8, 8, 8, 896, 896, 0, 1, 0)
conv(8, 4, 4, 896, 896, 0, 1, 0)
