"""GPU bring-up battery for the tcgen05 GEMM / implicit-conv kernel. Prints one line per case; never stops early."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from upgpt_b200 import _C

torch.manual_seed(0)
dev = torch.device("cuda:0")
L = _C.lib()
results = []


def report(name, got, ref, tol=2e-3):
    got = got.float().cpu(); ref = ref.float().cpu()
    bad = not torch.isfinite(got).all()
    err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    ok = (not bad) and err < tol
    results.append(dict(name=name, err=err, ok=bool(ok)))
    print(("PASS " if ok else "FAIL ") + name + "  max-rel-err=%.3e" % err, flush=True)
    return ok


def run_gemm(**kw):
    a = _C.GemmArgs()
    for k, v in kw.items():
        setattr(a, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    rc = L.upgpt_gemm(a, _C.stream_ptr())
    if rc != 0:
        print("   rc=%d %s" % (rc, L.upgpt_last_error().decode()))
    torch.cuda.synchronize()
    return rc


def case_plain(M, N, K, block_n=0, splits=0, bias=True, res=True, out16=False, batch=1):
    A = (torch.randn(batch, M, K, device=dev) * 0.5).half()
    Wt = (torch.randn(batch, N, K, device=dev) * 0.1).half()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(batch * M, N, device=dev) if res else None
    out = torch.full((batch * M, N), float("nan"), device=dev)
    o16 = torch.zeros(batch * M, N, device=dev, dtype=torch.half) if out16 else None
    ref = torch.einsum("bmk,bnk->bmn", A.float(), Wt.float()).reshape(batch * M, N)
    if bias: ref = ref + b
    if res: ref = ref + r
    name = "plain M%d N%d K%d bn%d sp%d b%d" % (M, N, K, block_n, splits, batch)
    try:
        rc = run_gemm(a=A, w=Wt, mode=0, M=M, N=N, K=K, batch=batch, block_n=block_n, splits=splits, out32=out,
                      out16=o16 if out16 else 0, bias=b if bias else 0, res32=r if res else 0)
        ok = rc == 0 and report(name, out, ref)
        if out16 and rc == 0: report(name + " [fp16 copy]", o16, ref, 4e-3)
    except Exception as e:  # noqa
        print("FAIL", name, "EXC", e); results.append(dict(name=name, err=None, ok=False))


def case_conv(B, H, W, Cin, Cout, block_n=0, splits=0, emb=True, res=False, chw=False):
    x = (torch.randn(B, H, W, Cin, device=dev) * 0.5).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) * (1.0 / (9 * Cin)) ** 0.5).half()
    wp = w.permute(0, 2, 3, 1).contiguous()   # [Cout][3][3][Cin]
    b = torch.randn(Cout, device=dev)
    e = torch.randn(B, Cout, device=dev) if emb else None
    r = torch.randn(B * H * W, Cout, device=dev) if res else None
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, padding=1)
    if emb: ref = ref + e[:, :, None, None]
    if chw:
        out = torch.full((B, Cout, H, W), float("nan"), device=dev)
        refc = ref
    else:
        out = torch.full((B * H * W, Cout), float("nan"), device=dev)
        refc = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
        if res: refc = refc + r
    name = "conv3x3 B%d %dx%d %d->%d bn%d sp%d%s" % (B, H, W, Cin, Cout, block_n, splits, " chw" if chw else "")
    try:
        rc = run_gemm(a=x, w=wp, mode=1, N=Cout, K=Cin, n_imgs=B, H=H, W=W, block_n=block_n, splits=splits, out32=out,
                      bias=b, rowvec=e if emb else 0, res32=r if res else 0, flags=(_C.GEMM_F_CHW if chw else 0))
        if rc == 0: report(name, out, refc)
        else: results.append(dict(name=name, err=None, ok=False))
    except Exception as ex:  # noqa
        print("FAIL", name, "EXC", ex); results.append(dict(name=name, err=None, ok=False))


def case_conv_s2(B, H, W, C, Cout):
    """stride-2 conv through the 4-phase layout. H, W = INPUT size."""
    x = (torch.randn(B, H, W, C, device=dev) * 0.5).half()
    w = (torch.randn(Cout, C, 3, 3, device=dev) * (1.0 / (9 * C)) ** 0.5).half()
    wp = w.permute(0, 2, 3, 1).contiguous()
    b = torch.randn(Cout, device=dev)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, stride=2, padding=1)
    Ho, Wo = H // 2, W // 2
    ph = torch.stack([x[:, p::2, q::2, :] for p in (0, 1) for q in (0, 1)], 0).contiguous()  # [4][B][Ho][Wo][C]
    out = torch.full((B * Ho * Wo, Cout), float("nan"), device=dev)
    name = "conv3x3-s2 B%d %dx%d %d->%d" % (B, H, W, C, Cout)
    rc = run_gemm(a=ph, w=wp, mode=2, N=Cout, K=C, n_imgs=B, H=Ho, W=Wo, out32=out, bias=b)
    if rc == 0: report(name, out, ref.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Cout))
    else: results.append(dict(name=name, err=None, ok=False))


def case_geglu(M, C, block_n=224):
    inner = 4 * C
    A = (torch.randn(M, C, device=dev) * 0.5).half()
    Wt = (torch.randn(2 * inner, C, device=dev) * (1.0 / C) ** 0.5).half()
    b = torch.randn(2 * inner, device=dev) * 0.1
    h = block_n // 2
    # pack rows per tile: [x(j0..j0+h) | gate(j0..j0+h)]
    idx = []
    for t in range(inner // h):
        idx += list(range(t * h, (t + 1) * h)) + list(range(inner + t * h, inner + (t + 1) * h))
    idx = torch.tensor(idx, device=dev)
    Wp, bp = Wt[idx].contiguous(), b[idx].contiguous()
    y = A.float() @ Wt.float().t() + b
    ref = y[:, :inner] * F.gelu(y[:, inner:])
    o16 = torch.zeros(M, inner, device=dev, dtype=torch.half)
    name = "geglu M%d C%d bn%d" % (M, C, block_n)
    rc = run_gemm(a=A, w=Wp, mode=0, M=M, N=2 * inner, K=C, block_n=block_n, out16=o16, bias=bp, flags=_C.GEMM_F_GEGLU)
    if rc == 0: report(name, o16, ref, 4e-3)
    else: results.append(dict(name=name, err=None, ok=False))


def case_chw_plain(B, T, N, K, ldT):
    """V^T style store: out16[(b*N + n)*ldT + t]"""
    A = (torch.randn(B * T, K, device=dev) * 0.5).half()
    Wt = (torch.randn(N, K, device=dev) * 0.1).half()
    o16 = torch.zeros(B, N, ldT, device=dev, dtype=torch.half)
    ref = (A.float() @ Wt.float().t()).reshape(B, T, N).permute(0, 2, 1)
    name = "plain-CHW B%d T%d N%d K%d ldT%d" % (B, T, N, K, ldT)
    rc = run_gemm(a=A, w=Wt, mode=0, M=B * T, N=N, K=K, out16=o16, rows_per_group=T, ldT=ldT, flags=_C.GEMM_F_CHW)
    if rc == 0: report(name, o16[:, :, :T], ref, 4e-3)
    else: results.append(dict(name=name, err=None, ok=False))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    case_plain(128, 64, 64, res=False, bias=False)
    case_plain(128, 128, 256)
    case_plain(1024, 224, 224)
    case_plain(8192, 224, 224, out16=True)
    case_plain(8192, 672, 224)
    case_plain(300, 448, 448)
    case_plain(696, 256, 768, res=False)          # context K/V projection rows (8*87)
    case_plain(128, 896, 896, splits=4, res=False)
    case_plain(128, 896, 8064, splits=0, res=True)
    case_plain(1024, 1024, 512, batch=2, res=False, bias=False)
    case_conv(1, 32, 32, 64, 64)
    case_conv(2, 32, 32, 224, 224)
    case_conv(8, 32, 32, 224, 224, res=True)
    case_conv(8, 16, 16, 448, 448)
    case_conv(8, 8, 8, 896, 896)
    case_conv(8, 4, 4, 896, 896)
    case_conv(8, 4, 4, 1792, 896)
    case_conv(2, 32, 24, 224, 224)
    case_conv(3, 4, 3, 896, 896)
    case_conv(2, 64, 64, 224, 224)
    case_conv(8, 32, 32, 224, 4, emb=False, chw=True)
    case_conv(1, 128, 128, 128, 128, emb=False)
    case_conv_s2(8, 32, 32, 224, 224)
    case_conv_s2(2, 8, 8, 896, 896)
    case_geglu(1024, 224)
    case_geglu(256, 448)
    case_chw_plain(8, 87, 256, 768, 96)
    case_chw_plain(2, 1024, 256, 224, 1024)
    # quick timing of the big conv
    x = (torch.randn(8, 32, 32, 224, device=dev) * 0.5).half(); w = (torch.randn(224, 3, 3, 224, device=dev) * 0.02).half()
    out = torch.empty(8 * 1024, 224, device=dev)
    a = _C.GemmArgs(); a.a = x.data_ptr(); a.w = w.data_ptr(); a.mode = 1; a.N = 224; a.K = 224; a.n_imgs = 8; a.H = 32; a.W = 32; a.out32 = out.data_ptr()
    for _ in range(3): L.upgpt_gemm(a, _C.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.upgpt_gemm(a, _C.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 2 * 8192 * 224 * 224 * 9
    print("conv 224->224@32x32 B8: %.3f us  %.1f TFLOP/s" % (ms * 1e3, fl / ms / 1e9))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/gemm_check.json", "w"), indent=1)
    nfail = sum(1 for r in results if not r["ok"])
    print("TOTAL %d cases, %d failed" % (len(results), nfail))
