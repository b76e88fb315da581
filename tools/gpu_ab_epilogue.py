"""Epilogue-bound GEMMs of the step timed alone (graph of 16 launches): GEGLU projections of the three transformer levels and the
folded-LayerNorm fp16-output projections. Run under UPGPT_LIB_PATH=<other build> for an A/B inside one GPU call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from upgpt_b200 import _C, ops
from gpu_sweep_gemm import time_graph, dev
print("lib:", os.environ.get("UPGPT_LIB_PATH", "default"))
for (M, inner, K) in ((8192, 896, 224), (2048, 1792, 448), (512, 3584, 896)):
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(2 * inner, K, device=dev) * 0.05).half()
    b = torch.randn(2 * inner, device=dev); o = torch.empty(M, inner, device=dev, dtype=torch.half); cs = torch.randn(2 * inner, device=dev)
    st = torch.rand(M * 2 * 2, device=dev) + 1.0
    t = time_graph(lambda i: ops.gemm(a=a, w=w, mode=0, M=M, N=2 * inner, K=K, out16=o, bias=b, ln_stats=st, ln_slots=2, ln_eps=1e-5, ln_colsum=cs,
                                      flags=_C.GEMM_F_GEGLU | _C.GEMM_F_W_STATIC), 2)
    print(f"GEGLU M{M} N{2 * inner} K{K} folded LN: {t:.2f} us", flush=True)
for (M, N, K) in ((8192, 768, 224), (8192, 256, 224), (2048, 1536, 448), (512, 3072, 896)):
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev); o = torch.empty(M, N, device=dev, dtype=torch.half); cs = torch.randn(N, device=dev)
    st = torch.rand(M * 2 * 2, device=dev) + 1.0
    t = time_graph(lambda i: ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out16=o, bias=b, ln_stats=st, ln_slots=2, ln_eps=1e-5, ln_colsum=cs,
                                      flags=_C.GEMM_F_W_STATIC), 2)
    print(f"gemm M{M} N{N} K{K} f16 out, folded LN: {t:.2f} us", flush=True)
