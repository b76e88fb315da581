"""Per-kernel parity through the C ABI (ctypes) against fp32 PyTorch-on-CPU restatements of the reference ops.
Tolerances: fp32-accumulated results of fp16 operands are compared against the same fp16-rounded operands in fp32
(tol 1e-5..1e-4); outputs stored as fp16 carry one extra rounding (tol 2e-3 = 4 fp16 ulps of the max)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from upgpt_b200 import _C
    _C.lib()
    return torch.device("cuda:0")


def relerr(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9))


@pytest.mark.parametrize("M,N,K,splits,batch", [(128, 64, 64, 0, 1), (8192, 224, 224, 0, 1), (300, 448, 448, 0, 1), (696, 256, 768, 0, 1),
                                                (128, 896, 896, 4, 1), (128, 896, 8064, 0, 1), (1024, 1024, 512, 0, 2), (1, 16, 8, 0, 1)])
def test_gemm_plain(dev, M, N, K, splits, batch):
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(batch, M, K, generator=g) * 0.5).half(); W = (torch.randn(batch, N, K, generator=g) * 0.1).half()
    b = torch.randn(N, generator=g); r = torch.randn(batch * M, N, generator=g)
    ref = torch.einsum("bmk,bnk->bmn", A.float(), W.float()).reshape(batch * M, N) + b + r
    out = torch.full((batch * M, N), float("nan"), device=dev)
    o16 = torch.zeros(batch * M, N, device=dev, dtype=torch.half) if N % 8 == 0 else None
    ops.gemm(a=A.to(dev), w=W.to(dev), mode=0, M=M, N=N, K=K, batch=batch, splits=splits, out32=out, out16=o16, bias=b.to(dev), res32=r.to(dev))
    torch.cuda.synchronize()
    assert relerr(out, ref) < 2e-5
    if o16 is not None:
        assert relerr(o16, ref) < 2e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 32, 32, 64, 64), (8, 32, 32, 224, 224), (8, 16, 16, 448, 448), (8, 8, 8, 896, 896), (8, 4, 4, 1792, 896),
                                            (2, 32, 24, 224, 224), (3, 4, 3, 896, 896), (2, 64, 64, 224, 224), (1, 256, 256, 128, 128), (1, 64, 192, 128, 128), (1, 8, 8, 32, 16)])
def test_conv3x3_implicit_gemm(dev, B, H, W, Cin, Cout):
    """Zero padding comes from TMA out-of-bounds fill; edge cases: ragged 32x24, 4x3 (multi-image tiles), W=256 (column tiles)."""
    from upgpt_b200 import ops, _C
    g = torch.Generator().manual_seed(B * H + Cin)
    x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).half(); w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).half()
    b = torch.randn(Cout, generator=g); e = torch.randn(B, Cout, generator=g)
    ref = (F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, padding=1) + e[:, :, None, None]).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    out = torch.full((B * H * W, Cout), float("nan"), device=dev)
    ops.gemm(a=x.to(dev), w=w.permute(0, 2, 3, 1).contiguous().to(dev), mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W, out32=out,
             bias=b.to(dev), rowvec=e.to(dev))
    torch.cuda.synchronize()
    assert relerr(out, ref) < 2e-5


def test_conv3x3_stride2_phases_and_nchw_out(dev):
    from upgpt_b200 import ops, _C
    g = torch.Generator().manual_seed(5)
    B, H, W, C, Co = 4, 16, 16, 224, 224
    x = (torch.randn(B, H, W, C, generator=g) * 0.5); w = (torch.randn(Co, C, 3, 3, generator=g) * (9 * C) ** -0.5).half(); b = torch.randn(Co, generator=g)
    xd = x.to(dev).reshape(B, H * W, C).contiguous()
    ph = torch.zeros(4 * B * (H // 2) * (W // 2) * C, device=dev, dtype=torch.half)
    ops.prep(x1=xd, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, stats=None, gamma=None, beta=None, eps=0.0, silu=0, layout=2, split3=0, out=ph, raw=None)
    out = torch.full((B * 64, Co), float("nan"), device=dev)
    ops.gemm(a=ph, w=w.permute(0, 2, 3, 1).contiguous().to(dev), mode=_C.GEMM_CONV3X3_S2PHASE, N=Co, K=C, n_imgs=B, H=H // 2, W=W // 2, out32=out, bias=b.to(dev))
    ref = F.conv2d(x.half().float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=1)
    torch.cuda.synchronize()
    assert relerr(out, ref.permute(0, 2, 3, 1).reshape(B * 64, Co)) < 2e-5
    # N=4 output conv written straight to NCHW (eps layout at the boundary)
    w4 = (torch.randn(4, C, 3, 3, generator=g) * 0.02).half()
    o = torch.full((B, 4, H, W), float("nan"), device=dev)
    ops.gemm(a=x.half().to(dev), w=w4.permute(0, 2, 3, 1).contiguous().to(dev), mode=_C.GEMM_CONV3X3, N=4, K=C, n_imgs=B, H=H, W=W, block_n=16, out32=o,
             flags=_C.GEMM_F_CHW)
    torch.cuda.synchronize()
    assert relerr(o, F.conv2d(x.half().float().permute(0, 3, 1, 2), w4.float(), padding=1)) < 2e-5


@pytest.mark.parametrize("M,C,N,x3,prod_splits,cons", [(8192, 224, 768, 1, 0, "f16"), (300, 224, 1792, 0, 0, "geglu"), (128, 896, 1024, 0, 4, "f16"),
                                                       (512, 896, 3072, 1, 2, "f16"), (128, 896, 896, 0, 0, "split"), (2048, 448, 512, 1, 0, "f32")])
def test_layernorm_folded_into_gemms(dev, M, C, N, x3, prod_splits, cons):
    """x = A0 W0^T + b0 + r (the GEMM that writes the residual stream) emits fp16 planes of the raw x and per-row {sum, sumsq} partials
    per N tile (TMA-store epilogue or the cluster split-K reduction); the consumer GEMM runs on those planes with gamma-scaled weights
    and applies mean / rstd in its epilogue (fp16 TMA-store, GEGLU, fp32, split-K reduction): result == LayerNorm(x) W^T + b."""
    import ctypes as C_
    from upgpt_b200 import ops, _C
    from upgpt_b200.unet_engine import pack_geglu, geglu_half, split3_w
    g = torch.Generator().manual_seed(M + C + N)
    K0 = 256
    A0 = (torch.randn(M, K0, generator=g) * 0.5).half(); W0 = (torch.randn(C, K0, generator=g) * 0.1).half()
    b0 = torch.randn(C, generator=g) + 0.7; r = torch.randn(M, C, generator=g)          # rows with a non-zero mean
    x_ref = A0.float() @ W0.float().t() + b0 + r
    gamma = 1 + 0.2 * torch.randn(C, generator=g); beta = 0.3 * torch.randn(C, generator=g)
    W = torch.randn(N, C, generator=g) * C ** -0.5; b = 0.1 * torch.randn(N, generator=g)
    y_ref = F.layer_norm(x_ref, (C,), gamma, beta, 1e-5) @ W.t() + b
    # ---- producer ----
    x32 = torch.full((M, C), float("nan"), device=dev)
    planes = 2 if x3 else 1
    raw16 = torch.zeros(M, planes * C, device=dev, dtype=torch.half)
    stats = torch.full((M * 16 * 2,), float("nan"), device=dev)     # dense [M][slots][2], slots known after planning
    pa = _C.GemmArgs()
    kw = dict(a=A0.to(dev), w=W0.to(dev), mode=0, M=M, N=C, K=K0, splits=prod_splits, out32=x32, out16=raw16, bias=b0.to(dev), res32=r.to(dev),
              rowstats_out=stats, flags=_C.GEMM_F_SPLIT3OUT if x3 else 0)
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v); v = v.data_ptr()
        setattr(pa, k, v)
    plan = (C_.c_int * 8)()
    _C.check(_C.lib().upgpt_gemm_plan(C_.byref(pa), C_.byref(plan)), "plan")
    slots = int(plan[1])
    assert 1 <= slots <= 16 and (prod_splits == 0 or int(plan[2]) == prod_splits)
    _C.check(_C.lib().upgpt_gemm(C_.byref(pa), ops.stream()), "producer")
    torch.cuda.synchronize()
    assert relerr(x32, x_ref) < 2e-5
    st = stats[:M * slots * 2].reshape(M, slots, 2).double().sum(1).cpu()
    assert relerr(st[:, 0], x_ref.double().sum(1)) < 1e-5 and relerr(st[:, 1], (x_ref.double() ** 2).sum(1)) < 1e-5
    # ---- consumer ----
    wg = W * gamma[None, :]
    bias_f = (b.double() + W.double() @ beta.double()).float()
    if cons == "geglu":
        inner = N // 2; half = geglu_half(inner, bool(x3))
        wg, bias_f = pack_geglu(wg, bias_f, inner, half)
        yr = y_ref[:, :inner] * F.gelu(y_ref[:, inner:])
    wp = split3_w(wg) if x3 else wg.half()
    colsum = wp.double().sum(-1).float()
    ck = dict(a=raw16, w=wp.to(dev), mode=0, M=M, N=N, K=C, bias=bias_f.to(dev), ln_stats=stats, ln_slots=slots, ln_eps=1e-5, ln_colsum=colsum.to(dev),
              flags=_C.GEMM_F_X3 if x3 else 0)
    if cons == "geglu":
        o16 = torch.zeros(M, inner, device=dev, dtype=torch.half)
        ops.gemm(out16=o16, block_n=2 * half, **dict(ck, flags=ck["flags"] | _C.GEMM_F_GEGLU))
        torch.cuda.synchronize()
        assert relerr(o16, yr) < 3e-3
    elif cons == "f16":
        o16 = torch.zeros(M, N, device=dev, dtype=torch.half)
        ops.gemm(out16=o16, **ck)
        torch.cuda.synchronize()
        assert relerr(o16, y_ref) < (2e-3 if x3 else 3e-3)
    else:
        o32 = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(out32=o32, splits=4 if cons == "split" else 0, **ck)
        torch.cuda.synchronize()
        assert relerr(o32, y_ref) < (1e-4 if x3 else 2e-3)     # single-plane operands: x rounded to fp16 (relative to |x|, not |x - mean|)


@pytest.mark.parametrize("B,H,W,Cin,Cout,x3,splits", [(8, 32, 32, 224, 224, 1, 0), (8, 32, 32, 224, 224, 0, 1), (8, 16, 16, 448, 448, 1, 0), (8, 8, 8, 896, 896, 0, 0),
                                                      (8, 4, 4, 896, 896, 0, 0), (8, 4, 4, 1792, 896, 0, 4), (2, 32, 24, 224, 224, 0, 0), (3, 8, 8, 448, 896, 0, 2)])
def test_groupnorm_moments_from_conv_epilogue(dev, B, H, W, Cin, Cout, x3, splits):
    """upgpt_gemm(gn_acc=...): the conv that produces a tensor adds its per-(image, group) moments to int64 fixed-point accumulators
    (TMA-store epilogue and cluster split-K reduction; 1 .. 8 images per tile; two consumers with different groupings: the next block's
    GroupNorm over the tensor alone and a decoder GroupNorm over a concatenation it is the first part of). upgpt_prep_operand(gn_acc=...)
    then applies GroupNorm + SiLU without a statistics pass: same operand as the fused one-launch GroupNorm kernel; accumulators are
    bit-identical across runs and equal to upgpt_gn_accumulate on the stored tensor up to fixed-point rounding."""
    from upgpt_b200 import ops, _C
    from upgpt_b200.unet_engine import split3_w
    g = torch.Generator().manual_seed(B * H + Cin + Cout)
    HW = H * W
    x = torch.randn(B, H, W, Cin, generator=g) * 0.5
    w = torch.randn(Cout, 9, Cin, generator=g) * (9 * Cin) ** -0.5
    b = torch.randn(Cout, generator=g); e = torch.randn(B, Cout, generator=g)
    xa = (split3_w(x) if x3 else x.half()).to(dev); wa = (split3_w(w) if x3 else w.half()).to(dev)
    out = torch.full((B * HW, Cout), float("nan"), device=dev)
    groups = 32
    cpg1 = Cout // groups
    C2 = Cout // 2                                   # second consumer: GroupNorm over [this tensor | C2 more channels]
    cpg2 = (Cout + C2) // groups
    acc1 = torch.zeros(B, groups, 2, device=dev, dtype=torch.int64)
    acc2 = torch.zeros(B, groups, 2, device=dev, dtype=torch.int64)
    kw = dict(a=xa, w=wa, mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W, out32=out, bias=b.to(dev), rowvec=e.to(dev), splits=splits,
              flags=_C.GEMM_F_X3 if x3 else 0, gn_acc=acc1, gn_groups=groups, gn_cpg=cpg1, gn_choff=0, gn_acc2=acc2, gn_cpg2=cpg2, gn_choff2=0)
    try:
        ops.gemm(**kw)
    except _C.UpgptError as ex:
        # the tiling picked a split factor of 3 / 5 / 6 / 7: the moments need power-of-two row slices -> the caller pins one (the engine
        # does the same through upgpt_gemm_plan, unet_engine.py:_gn_from_epilogues)
        assert "power-of-two" in str(ex) and splits == 0
        kw["splits"] = 4
        ops.gemm(**kw)
    torch.cuda.synchronize()
    y = out.reshape(B, HW, Cout).double().cpu()
    ref1 = torch.stack([y.reshape(B, HW, groups, cpg1).sum((1, 3)), (y ** 2).reshape(B, HW, groups, cpg1).sum((1, 3))], -1)
    got1 = torch.stack([acc1[..., 0].double() / 2 ** 24, acc1[..., 1].double() / 2 ** 20], -1).cpu()
    assert relerr(got1[..., 1], ref1[..., 1]) < 1e-5 and float((got1[..., 0] - ref1[..., 0]).abs().max()) < 1e-3 * float(ref1[..., 1].max()) ** 0.5
    # second consumer: only the groups the tensor's channels fall into carry anything
    ch = torch.arange(Cout) // cpg2
    ref2 = torch.zeros(B, groups, 2, dtype=torch.float64)
    ref2[..., 0].index_add_(1, ch, y.sum(1)); ref2[..., 1].index_add_(1, ch, (y ** 2).sum(1))
    got2 = torch.stack([acc2[..., 0].double() / 2 ** 24, acc2[..., 1].double() / 2 ** 20], -1).cpu()
    assert relerr(got2[..., 1], ref2[..., 1]) < 1e-5
    # bit-reproducible, and equal to the stand-alone accumulation kernel on the stored tensor up to fixed-point rounding of the partials
    a1b = torch.zeros_like(acc1)
    ops.gemm(**dict(kw, gn_acc=a1b, gn_acc2=None))
    torch.cuda.synchronize()
    assert torch.equal(a1b, acc1)
    a1c = torch.zeros_like(acc1)
    _C.check(_C.lib().upgpt_gn_accumulate(out.data_ptr(), Cout, B, HW, groups, cpg1, 0, a1c.data_ptr(), ops.stream()), "gn_accumulate")
    torch.cuda.synchronize()
    assert float((a1c - acc1).abs().max()) <= 2e-7 * float(acc1.abs().max())     # fp32 partial sums taken in a different (fixed) order
    # apply from the accumulators == the fused one-launch GroupNorm(+SiLU) operand
    gamma = (1 + 0.1 * torch.randn(Cout, generator=g)).to(dev); beta = (0.1 * torch.randn(Cout, generator=g)).to(dev)
    o_acc = torch.zeros(B * HW, 2 * Cout, device=dev, dtype=torch.half); o_fused = torch.zeros_like(o_acc)
    ops.prep(x1=out, C1=Cout, x2=None, C2=0, B=B, H=H, W=W, groups=groups, stats=None, gamma=gamma, beta=beta, eps=1e-5, silu=1, layout=0, split3=1,
             out=o_acc, raw=None, gn_acc=acc1)
    st = torch.zeros(B, groups, 2, device=dev, dtype=torch.float64)
    ops.groupnorm_prep(st, x1=out, C1=Cout, x2=None, C2=0, B=B, H=H, W=W, groups=groups, gamma=gamma, beta=beta, eps=1e-5, silu=1, layout=0, split3=1,
                       out=o_fused, raw=None)
    torch.cuda.synchronize()
    full = lambda t: t[:, :Cout].float() + t[:, Cout:].float()
    assert relerr(full(o_acc), full(o_fused)) < 2e-6


@pytest.mark.parametrize("B,H,W,Cin,Cout,x3,splits", [(8, 4, 4, 896, 896, 0, 0), (8, 8, 8, 896, 448, 1, 0), (8, 16, 16, 448, 448, 1, 0), (2, 16, 16, 448, 224, 0, 1),
                                                      (3, 6, 10, 64, 96, 1, 0), (1, 64, 64, 128, 128, 0, 0), (2, 5, 3, 32, 48, 0, 2)])
def test_upsample_fold_conv(dev, B, H, W, Cin, Cout, x3, splits):
    """UPGPT_GEMM_CONV3X3_UP2 == conv3x3(nearest x2 upsample) (openaimodel.py:116-118, model.py:49-52) from the LOW-resolution operand:
    several images per tile (4x4, 8x8), one image over several tiles, ragged tiles, split-K through the cluster and the flat path."""
    from upgpt_b200 import ops, _C
    from upgpt_b200.unet_engine import split3_w, up2_conv_w
    g = torch.Generator().manual_seed(B * H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g) * 0.5
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5
    b = torch.randn(Cout, generator=g)
    wp = up2_conv_w(w)
    xa = (split3_w(x) if x3 else x.half()).to(dev); wa = (split3_w(wp) if x3 else wp.half()).to(dev)
    out = torch.full((B * 4 * H * W, Cout), float("nan"), device=dev)
    ops.gemm(a=xa, w=wa, mode=_C.GEMM_CONV3X3_UP2, N=Cout, K=Cin, n_imgs=B, H=H, W=W, out32=out, bias=b.to(dev), splits=splits,
             flags=(_C.GEMM_F_X3 if x3 else 0) | _C.GEMM_F_W_STATIC)
    torch.cuda.synchronize()
    xr = x if x3 else x.half().float()
    ref = F.conv2d(F.interpolate(xr.permute(0, 3, 1, 2).double(), scale_factor=2, mode="nearest"), w.double(), b.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(B * 4 * H * W, Cout).float()
    assert not torch.isnan(out).any()
    assert relerr(out.cpu(), ref) < (2e-5 if x3 else 2e-3)       # single plane: the folded weights are rounded to fp16


def test_geglu_epilogue(dev):
    from upgpt_b200 import ops, _C
    from upgpt_b200.unet_engine import pack_geglu, geglu_half
    g = torch.Generator().manual_seed(3)
    M, C = 300, 224
    inner = 4 * C; half = geglu_half(inner)
    A = (torch.randn(M, C, generator=g) * 0.5).half(); W = (torch.randn(2 * inner, C, generator=g) * C ** -0.5).half(); b = torch.randn(2 * inner, generator=g) * 0.1
    y = A.float() @ W.float().t() + b
    ref = y[:, :inner] * F.gelu(y[:, inner:])     # exact erf GELU (attention.py:43)
    wp, bp = pack_geglu(W, b, inner, half)
    o16 = torch.zeros(M, inner, device=dev, dtype=torch.half)
    ops.gemm(a=A.to(dev), w=wp.to(dev), mode=0, M=M, N=2 * inner, K=C, block_n=2 * half, out16=o16, bias=bp.to(dev), flags=_C.GEMM_F_GEGLU)
    torch.cuda.synchronize()
    assert relerr(o16, ref) < 2e-3


@pytest.mark.parametrize("B,H,W,C1,C2,silu,eps", [(2, 16, 16, 224, 0, True, 1e-5), (2, 8, 8, 448, 224, True, 1e-5), (1, 32, 32, 128, 0, False, 1e-6), (3, 4, 3, 896, 896, True, 1e-5),
                                                  (1, 128, 128, 128, 0, True, 1e-6)])
def test_groupnorm_silu_prep(dev, B, H, W, C1, C2, silu, eps):
    """GroupNorm32 over the channel concat of two tensors (group boundaries straddle the sources), fp32 statistics."""
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(C1 + C2)
    x1 = torch.randn(B, H * W, C1, generator=g) * 2 + 0.5
    x2 = torch.randn(B, H * W, C2, generator=g) - 1.0 if C2 else None
    Cc = C1 + C2
    gamma = 1 + 0.1 * torch.randn(Cc, generator=g); beta = 0.1 * torch.randn(Cc, generator=g)
    xc = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.group_norm(xc.reshape(B, H, W, Cc).permute(0, 3, 1, 2), 32, gamma, beta, eps)
    ref = F.silu(ref) if silu else ref
    stats = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64)
    out = torch.zeros(B * H * W * Cc, device=dev, dtype=torch.half)
    x1d, x2d = x1.to(dev), (x2.to(dev) if C2 else None)
    ops.groupnorm_stats(x1d, x2d, B, H * W, stats)
    ops.prep(x1=x1d, C1=C1, x2=x2d, C2=C2, B=B, H=H, W=W, groups=32, stats=stats, gamma=gamma.to(dev), beta=beta.to(dev), eps=eps, silu=int(silu), layout=0,
             split3=0, out=out, raw=None)
    torch.cuda.synchronize()
    n = (Cc // 32) * H * W
    mean_ref = xc.reshape(B, H * W, 32, Cc // 32).double().mean(dim=(1, 3))
    assert float(((stats[:, :, 0].cpu() / n) - mean_ref).abs().max()) < 1e-6      # statistics themselves are fp64-accurate
    assert relerr(out.reshape(B, H, W, Cc).permute(0, 3, 1, 2), ref) < 2e-3


@pytest.mark.parametrize("B,H,W,C1,C2,silu,eps,layout,split3", [
    (8, 32, 32, 224, 0, True, 1e-5, 0, 1),      # U-Net level 0: 8-CTA cluster per image
    (8, 32, 32, 448, 224, True, 1e-5, 0, 1),    # level-0 skip concat (2.75 MB per image): 16-CTA cluster or the two-launch fallback
    (2, 16, 16, 896, 448, True, 1e-5, 0, 0),    # concat, group boundaries straddle the two sources
    (3, 4, 3, 896, 896, True, 1e-5, 0, 1),      # 12 pixels: ragged split over the cluster (empty trailing CTAs)
    (2, 8, 8, 448, 0, False, 1e-6, 1, 1),       # SpatialTransformer / Upsample flavour: eps 1e-6, no SiLU, nearest-x2 layout
    (1, 24, 32, 224, 0, True, 1e-5, 2, 0),      # stride-2 phase layout on the real bbox.yaml latent aspect (32x24)
    (1, 128, 128, 128, 0, True, 1e-6, 0, 0),    # VAE-sized: does not fit a cluster -> internal gn_stats + prep
])
def test_groupnorm_prep_fused_cluster(dev, B, H, W, C1, C2, silu, eps, layout, split3):
    """upgpt_groupnorm_prep (one cluster per image, moments over DSMEM) == the two-launch path bit for bit, and == F.group_norm."""
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(C1 + C2 + H)
    x1 = torch.randn(B, H * W, C1, generator=g) * 2 + 0.5
    x2 = torch.randn(B, H * W, C2, generator=g) - 1.0 if C2 else None
    Cc = C1 + C2
    gamma = 1 + 0.1 * torch.randn(Cc, generator=g); beta = 0.1 * torch.randn(Cc, generator=g)
    x1d, x2d, gd, bd = x1.to(dev), (x2.to(dev) if C2 else None), gamma.to(dev), beta.to(dev)
    kx = 2 if split3 else 1
    mult = 4 if layout == 1 else 1
    outs, raws, stats = [], [], []
    for fused in (True, False):
        st = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64)
        out = torch.zeros(B * H * W * mult * Cc * kx, device=dev, dtype=torch.half)
        raw = torch.zeros(B * H * W * Cc * kx, device=dev, dtype=torch.half) if layout == 0 else None
        kw = dict(x1=x1d, C1=C1, x2=x2d, C2=C2, B=B, H=H, W=W, groups=32, gamma=gd, beta=bd, eps=eps, silu=int(silu), layout=layout,
                  split3=split3, out=out, raw=raw)
        if fused:
            ops.groupnorm_prep(st, **kw)
        else:
            ops.groupnorm_stats(x1d, x2d, B, H * W, st)
            ops.prep(stats=st, **kw)
        torch.cuda.synchronize()
        outs.append(out); raws.append(raw); stats.append(st)
    assert float((stats[0] - stats[1]).abs().max() / stats[1].abs().max()) < 1e-6
    # same arithmetic after the moments; the moments differ only in the fp32 partial-sum tree -> a few fp16 ulps at most
    d = (outs[0].float() - outs[1].float()).abs().max()
    assert float(d) <= 2e-3 * float(outs[1].float().abs().max()), float(d)
    if raws[0] is not None:
        assert torch.equal(raws[0], raws[1])
    xc = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.group_norm(xc.reshape(B, H, W, Cc).permute(0, 3, 1, 2), 32, gamma, beta, eps)
    ref = F.silu(ref) if silu else ref
    if layout == 1:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
        got = outs[0].reshape(B, 2 * H, 2 * W, kx, Cc).float().cpu()
    elif layout == 2:
        got = outs[0].reshape(2, 2, B, H // 2, W // 2, kx, Cc).float().cpu()
        full = torch.zeros(B, H, W, kx, Cc)
        for i in range(2):
            for j in range(2):
                full[:, i::2, j::2] = got[i, j]
        got = full
    else:
        got = outs[0].reshape(B, H, W, kx, Cc).float().cpu()
    val = got.sum(3)      # hi (+ lo)
    assert relerr(val.permute(0, 3, 1, 2), ref) < (2e-5 if split3 else 2e-3)


def test_prep_layouts_and_split3(dev):
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, H, W, C = 2, 8, 8, 64
    x = torch.randn(B, H * W, C, generator=g)
    xd = x.to(dev)
    up = torch.zeros(B * 4 * H * W * C, device=dev, dtype=torch.half)
    ops.prep(x1=xd, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, stats=None, gamma=None, beta=None, eps=0.0, silu=0, layout=1, split3=0, out=up, raw=None)
    ref = F.interpolate(x.reshape(B, H, W, C).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert relerr(up.reshape(B, 2 * H, 2 * W, C).permute(0, 3, 1, 2), ref) < 1e-3
    s3 = torch.zeros(B * H * W * 2 * C, device=dev, dtype=torch.half)
    ops.prep(x1=xd, C1=C, x2=None, C2=0, B=B, H=H, W=W, groups=32, stats=None, gamma=None, beta=None, eps=0.0, silu=0, layout=0, split3=1, out=s3, raw=None)
    o = s3.reshape(B * H * W, 2, C).float().cpu()      # planes [hi | lo]: hi = fp16(x), lo = fp16(x - hi)
    assert float((o[:, 0] + o[:, 1] - x.reshape(-1, C)).abs().max()) < 1e-6 and torch.equal(o[:, 0], x.reshape(-1, C).half().float())


def _planes(t):
    """fp32 [..., K] -> fp16 [..., 2K] = [hi | lo]."""
    hi = t.half()
    return torch.cat([hi, (t - hi.float()).half()], -1)


@pytest.mark.parametrize("M,N,K,splits", [(128, 64, 64, 0), (8192, 224, 224, 0), (300, 448, 448, 0), (128, 896, 896, 4), (128, 896, 8064, 0),
                                          (696, 256, 768, 0), (512, 2048, 224, 0)])
def test_gemm_x3_error_compensated(dev, M, N, K, splits):
    """UPGPT_GEMM_F_X3: operands are [hi | lo] planes, product Ah*Wh + Al*Wh + Ah*Wl -> fp32-grade result (the parity mode).
    K = 224 / 448 exercise the zero-filled k-block tail per plane (planes are a tensor-map dimension, not a K offset)."""
    from upgpt_b200 import ops, _C
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g) * 0.5; W = torch.randn(N, K, generator=g) * 0.1
    b = torch.randn(N, generator=g); r = torch.randn(M, N, generator=g)
    ref = (A.double() @ W.double().t() + b + r).float()
    out = torch.full((M, N), float("nan"), device=dev)
    o16 = torch.zeros(M, 2 * N, device=dev, dtype=torch.half)
    ops.gemm(a=_planes(A).to(dev), w=_planes(W).to(dev), mode=0, M=M, N=N, K=K, splits=splits, out32=out, out16=o16, bias=b.to(dev), res32=r.to(dev),
             flags=_C.GEMM_F_X3 | _C.GEMM_F_SPLIT3OUT)
    torch.cuda.synchronize()
    err = relerr(out, ref)
    assert err < 1e-5, err                  # single-plane fp16 operands give ~3e-4 here
    o = o16.float().cpu().reshape(M, 2, N)
    assert relerr(o[:, 0] + o[:, 1], ref) < 1e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(8, 32, 32, 224, 224), (8, 16, 16, 672, 448), (8, 4, 4, 1792, 896), (2, 32, 24, 224, 224)])
def test_conv3x3_x3_error_compensated(dev, B, H, W, Cin, Cout):
    from upgpt_b200 import ops, _C
    g = torch.Generator().manual_seed(B * H + Cin)
    x = torch.randn(B, H, W, Cin, generator=g) * 0.5
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (Cin * 9) ** -0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1).float()
    out = torch.full((B * H * W, Cout), float("nan"), device=dev)
    ops.gemm(a=_planes(x).to(dev), w=_planes(w.permute(0, 2, 3, 1).contiguous()).to(dev), mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W,
             out32=out, bias=b.to(dev), flags=_C.GEMM_F_X3)
    torch.cuda.synchronize()
    err = relerr(out.reshape(B, H, W, Cout).permute(0, 3, 1, 2), ref)
    assert err < 1e-5, err                  # fp32 accumulation over K = 9*Cin; single-plane fp16 operands give ~3e-4


@pytest.mark.parametrize("rows,C", [(300, 224), (64, 448), (17, 896), (5, 64)])
def test_layernorm(dev, rows, C):
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, C, generator=g) * 3 + 1; gamma = 1 + 0.1 * torch.randn(C, generator=g); beta = 0.1 * torch.randn(C, generator=g)
    out = torch.zeros(rows, C, device=dev, dtype=torch.half)
    ops.layernorm(x.to(dev), gamma.to(dev), beta.to(dev), out)
    torch.cuda.synchronize()
    assert relerr(out, F.layer_norm(x, (C,), gamma, beta, 1e-5)) < 2e-3


@pytest.mark.parametrize("B,Hh,Nq,Nk,d", [(1, 1, 128, 128, 64), (2, 8, 1024, 1024, 28), (2, 8, 256, 256, 56), (2, 8, 64, 64, 112), (2, 8, 16, 16, 112),
                                          (2, 8, 1024, 87, 28), (2, 8, 64, 87, 112), (1, 4, 384, 300, 16), (1, 8, 4096, 4096, 28), (1, 2, 1, 1, 8)])
def test_flash_attention(dev, B, Hh, Nq, Nk, d):
    """softmax(q k^T d^-1/2) v: self (Nk = Nq) and cross (87 context keys), ragged key tails masked, head dims 28/56/112 padded."""
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(Nq + Nk + d)
    dpad = 64 if d <= 64 else 128
    q = torch.randn(B, Hh, Nq, d, generator=g).half(); k = torch.randn(B, Hh, Nk, d, generator=g).half(); v = torch.randn(B, Hh, Nk, d, generator=g).half()
    ref = torch.softmax(torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * d ** -0.5, -1) @ v.float()
    HD = Hh * dpad
    Q = torch.zeros(B, Nq, Hh, dpad, dtype=torch.half); Q[..., :d] = q.permute(0, 2, 1, 3)
    K = torch.zeros(B, Nk, Hh, dpad, dtype=torch.half); K[..., :d] = k.permute(0, 2, 1, 3)
    ldvt = (Nk + 7) // 8 * 8
    Vt = torch.zeros(B, Hh, dpad, ldvt, dtype=torch.half); Vt[:, :, :d, :Nk] = v.permute(0, 1, 3, 2)
    out = torch.full((B, Nq, Hh, dpad), float("nan"), device=dev, dtype=torch.half)
    ops.attention(q=Q.to(dev), ldq=HD, k=K.to(dev), ldk=HD, k_batch_stride=0, vt=Vt.to(dev), ldvt=ldvt, out=out, ldo=HD, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=dpad,
                  scale=d ** -0.5)
    torch.cuda.synchronize()
    assert relerr(out[..., :d].permute(0, 2, 1, 3), ref) < 3e-3
    if d < dpad:
        assert float(out[..., d:].float().abs().max()) == 0.0      # padded head columns stay exactly zero
    # row-major V (a slice of a fused q|k|v projection output): the P V product reads it as an MN-major operand, no V^T anywhere
    QKV = torch.zeros(B, max(Nq, Nk), 3, Hh, dpad, dtype=torch.half)
    QKV[:, :Nq, 0, :, :d] = q.permute(0, 2, 1, 3); QKV[:, :Nk, 1, :, :d] = k.permute(0, 2, 1, 3); QKV[:, :Nk, 2, :, :d] = v.permute(0, 2, 1, 3)
    qkv = QKV.to(dev)
    flat = qkv.reshape(-1)
    n_rows = max(Nq, Nk)
    out2 = torch.full((B, Nq, Hh, dpad), float("nan"), device=dev, dtype=torch.half)
    ops.attention(q=flat, ldq=3 * HD, k=flat[HD:], ldk=3 * HD, k_batch_stride=n_rows * 3 * HD, vt=flat[2 * HD:], ldvt=3 * HD,
                  v_rowmajor=1, v_batch_stride=n_rows * 3 * HD, out=out2, ldo=HD, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=dpad, scale=d ** -0.5) if Nq == n_rows else None
    if Nq == n_rows:
        torch.cuda.synchronize()
        assert relerr(out2[..., :d].permute(0, 2, 1, 3), ref) < 3e-3
        assert torch.equal(out2, out), "same arithmetic whichever way V is laid out"


@pytest.mark.parametrize("B,Hh,Nq,Nk,d,ramp,split3", [(2, 8, 1024, 1024, 28, 1.0, 1), (1, 8, 1024, 87, 28, 1.0, 1), (1, 2, 256, 300, 32, 20.0, 0),
                                                      (2, 4, 100, 100, 16, 1.0, 0), (1, 8, 4096, 4096, 28, 1.0, 0)])
def test_flash_attention_head_pairs_d32(dev, B, Hh, Nq, Nk, d, ramp, split3):
    """dpad = 32: q / k / v rows hold head PAIRS in 64-wide (128-byte swizzled) rows, head h at columns 32 h. The S MMA of head h covers
    the two 16-wide k-steps of its half-row, P V runs over the pair's 64 V columns and the head keeps its own 32 (the level-0 shape of
    bbox.yaml: 8 heads of 28). Self (K ring, one V slot, one S buffer, two CTAs per SM) and cross (87 keys, one tile); the lo plane of
    the [hi | lo] output; rows whose max keeps growing (rescale path)."""
    from upgpt_b200 import ops
    g = torch.Generator().manual_seed(Nq + Nk + d)
    dpad = 32
    q = torch.randn(B, Hh, Nq, d, generator=g).half()
    k = (torch.randn(B, Hh, Nk, d, generator=g) * torch.linspace(1.0, ramp, Nk)[None, None, :, None]).half()
    v = torch.randn(B, Hh, Nk, d, generator=g).half()
    ref = torch.softmax(torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * d ** -0.5, -1) @ v.float()
    HD = Hh * dpad
    Q = torch.zeros(B, Nq, Hh, dpad, dtype=torch.half); Q[..., :d] = q.permute(0, 2, 1, 3)
    KV = torch.zeros(B, Nk, 2, Hh, dpad, dtype=torch.half); KV[:, :, 0, :, :d] = k.permute(0, 2, 1, 3); KV[:, :, 1, :, :d] = v.permute(0, 2, 1, 3)
    Qd, KVd = Q.to(dev), KV.to(dev)
    planes = 2 if split3 else 1
    out = torch.full((B, Nq, planes, Hh, dpad), float("nan"), device=dev, dtype=torch.half)
    ops.attention(q=Qd, ldq=HD, k=KVd, ldk=2 * HD, k_batch_stride=Nk * 2 * HD, vt=KVd.reshape(-1)[HD:], ldvt=2 * HD, v_rowmajor=1,
                  v_batch_stride=Nk * 2 * HD, out=out, ldo=planes * HD, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=dpad, scale=d ** -0.5, split3_out=split3)
    torch.cuda.synchronize()
    hi = out[:, :, 0]
    assert relerr(hi[..., :d].permute(0, 2, 1, 3), ref) < 3e-3
    if d < dpad:
        assert float(hi[..., d:].float().abs().max()) == 0.0      # padded head columns stay exactly zero
    if split3:
        full = hi.float() + out[:, :, 1].float()
        assert relerr(full[..., :d].permute(0, 2, 1, 3), ref) < 1.5e-3     # the lo plane refines the fp16 rounding of O
    # the same problem through the 64-wide single-head path must agree to fp16 rounding of P / O
    Q64 = torch.zeros(B, Nq, Hh, 64, dtype=torch.half); Q64[..., :d] = q.permute(0, 2, 1, 3)
    KV64 = torch.zeros(B, Nk, 2, Hh, 64, dtype=torch.half); KV64[:, :, 0, :, :d] = k.permute(0, 2, 1, 3); KV64[:, :, 1, :, :d] = v.permute(0, 2, 1, 3)
    out64 = torch.full((B, Nq, Hh, 64), float("nan"), device=dev, dtype=torch.half)
    KVd64 = KV64.to(dev)
    ops.attention(q=Q64.to(dev), ldq=Hh * 64, k=KVd64, ldk=2 * Hh * 64, k_batch_stride=Nk * 2 * Hh * 64, vt=KVd64.reshape(-1)[Hh * 64:], ldvt=2 * Hh * 64,
                  v_rowmajor=1, v_batch_stride=Nk * 2 * Hh * 64, out=out64, ldo=Hh * 64, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=64, scale=d ** -0.5)
    torch.cuda.synchronize()
    assert relerr(hi[..., :d], out64[..., :d]) < 2e-3


@pytest.mark.parametrize("Nk,ramp", [(1024, 12.0), (1024, 1.0), (300, 30.0)])
def test_flash_attention_growing_row_max(dev, Nk, ramp):
    """The online softmax keeps its reference max until a row's max has grown by more than 2^8 and only then rescales O in TMEM:
    keys whose magnitude ramps up along the sequence push every row through several such rescales (ramp = 1: through none after the
    first tile); both must agree with the fp32 softmax."""
    from upgpt_b200 import ops
    B, Hh, Nq, d, dpad = 1, 2, 256, 28, 64
    g = torch.Generator().manual_seed(int(Nk + ramp))
    q = torch.randn(B, Hh, Nq, d, generator=g).half()
    k = (torch.randn(B, Hh, Nk, d, generator=g) * torch.linspace(1.0, ramp, Nk)[None, None, :, None]).half()
    v = torch.randn(B, Hh, Nk, d, generator=g).half()
    ref = torch.softmax(torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * d ** -0.5, -1) @ v.float()
    HD = Hh * dpad
    Q = torch.zeros(B, Nq, Hh, dpad, dtype=torch.half); Q[..., :d] = q.permute(0, 2, 1, 3)
    KV = torch.zeros(B, Nk, 2, Hh, dpad, dtype=torch.half); KV[:, :, 0, :, :d] = k.permute(0, 2, 1, 3); KV[:, :, 1, :, :d] = v.permute(0, 2, 1, 3)
    Qd, KVd = Q.to(dev), KV.to(dev)
    out = torch.full((B, Nq, Hh, dpad), float("nan"), device=dev, dtype=torch.half)
    ops.attention(q=Qd, ldq=HD, k=KVd, ldk=2 * HD, k_batch_stride=Nk * 2 * HD, vt=KVd.reshape(-1)[HD:], ldvt=2 * HD, v_rowmajor=1,
                  v_batch_stride=Nk * 2 * HD, out=out, ldo=HD, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=dpad, scale=d ** -0.5)
    torch.cuda.synchronize()
    assert relerr(out[..., :d].permute(0, 2, 1, 3), ref) < 3e-3


def test_small_kernels(dev):
    from upgpt_b200 import ops
    from oracle import ldm_oracle as O
    g = torch.Generator().manual_seed(2)
    t = torch.tensor([981, 481, 1, 0])
    assert relerr(ops.timestep_embedding(t.to(dev), 224), O.timestep_embedding(t, 224)) < 2e-6
    x = torch.randn(4, 896, generator=g); w = torch.randn(300, 896, generator=g) * 0.03; b = torch.randn(300, generator=g)
    got = ops.linear_small_m(x.to(dev), w.to(dev), b.to(dev), silu_in=True)
    assert relerr(got, F.linear(F.silu(x), w, b)) < 1e-5
    # boundary conv: latent (4ch) ++ mask (1ch) -> 224, NCHW in, NHWC out
    xi = torch.randn(2, 4, 16, 24, generator=g); m = torch.randn(2, 1, 16, 24, generator=g); wc = torch.randn(224, 5, 3, 3, generator=g) * 0.1; bc = torch.randn(224, generator=g)
    out = torch.zeros(2, 16 * 24, 224, device=dev)
    ops.conv_small_cin(xi.to(dev), m.to(dev), wc.permute(1, 2, 3, 0).reshape(45, 224).contiguous().to(dev), bc.to(dev), 224, 3, out)
    ref = F.conv2d(torch.cat([xi, m], 1), wc, bc, padding=1).permute(0, 2, 3, 1).reshape(2, 384, 224)
    assert relerr(out, ref) < 1e-5


@pytest.mark.parametrize("eta", [0.0, 1.0])
def test_ddim_and_ddpm_update_kernels(dev, eta):
    from upgpt_b200 import ops
    from oracle import ldm_oracle as O
    g = torch.Generator().manual_seed(4)
    sched = O.register_schedule(1000, 0.00085, 0.012)
    ts, a, ap, sg, s1m = O.ddim_schedule(sched["alphas_cumprod"], 50, eta)
    x = torch.randn(2, 4, 32, 32, generator=g); e = torch.randn(2, 4, 32, 32, generator=g); nz = torch.randn(2, 4, 32, 32, generator=g)
    import numpy as np
    rows = np.stack([np.asarray(a, np.float32), np.asarray(ap, np.float32), np.asarray(sg, np.float32), np.asarray(s1m, np.float32), np.ones(50, np.float32)], 1)
    coef = torch.from_numpy(rows).to(dev)
    for index in (49, 17, 0):
        ref_prev, ref_x0 = O.ddim_step(x, e, a[index], ap[index], sg[index], s1m[index], nz if eta > 0 else None)
        xp, p0 = torch.empty(2, 4, 32, 32, device=dev), torch.empty(2, 4, 32, 32, device=dev)
        ops.ddim_step(x.to(dev), e.to(dev), coef, xp, p0, noise=nz.to(dev) if eta > 0 else None, step_imm=index)
        assert relerr(xp, ref_prev) < 2e-6 and relerr(p0, ref_x0) < 2e-6
    model_rows = torch.stack([sched["sqrt_recip_alphas_cumprod"], sched["sqrt_recipm1_alphas_cumprod"], sched["posterior_mean_coef1"], sched["posterior_mean_coef2"],
                              sched["posterior_log_variance_clipped"], (torch.arange(1000) != 0).float()], 1).contiguous().to(dev)
    for tt in (999, 500, 0):
        ref_prev, ref_x0 = O.ddpm_step(x, e, torch.full((2,), tt), sched, nz)
        xp, p0 = torch.empty(2, 4, 32, 32, device=dev), torch.empty(2, 4, 32, 32, device=dev)
        ops.ddpm_step(x.to(dev), e.to(dev), model_rows, xp, p0, noise=nz.to(dev), step_imm=tt)
        assert relerr(xp, ref_prev) < 5e-6 and relerr(p0, ref_x0) < 5e-6


def test_launch_trace_stamps_every_kernel_in_order(dev):
    """upgpt_trace_set: block 0 of every kernel of the library stamps %globaltimer on entry (kind 0) and when its dependency wait returns
    (kind 1): a chain of n launches leaves n stamps of each kind, the kind-1 stamps are non-decreasing (completion order of a dependent
    chain), nothing is written when the trace is off, and the capacity word bounds the writes."""
    from upgpt_b200 import ops, _C
    L = _C.lib()
    M, N, K = 256, 128, 128
    a = (torch.randn(M, K) * 0.5).half().to(dev); w = (torch.randn(N, K) * 0.05).half().to(dev)
    out = torch.empty(M, N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.half)

    def chain():
        ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out32=out)
        _C.check(L.upgpt_axpby(out.data_ptr(), 2.0, out.data_ptr(), 0.5, out.data_ptr(), out.numel(), ops.stream()), "axpby")
        ops.gemm(a=a, w=w, mode=0, M=M, N=N, K=K, out16=o16)
    chain(); torch.cuda.synchronize()
    cap = 64
    buf = torch.zeros(cap + 2, dtype=torch.int64, device=dev); buf[1] = cap
    torch.cuda.synchronize()
    _C.check(L.upgpt_trace_set(buf.data_ptr()), "trace_set")
    chain(); chain(); torch.cuda.synchronize()
    _C.check(L.upgpt_trace_set(None), "trace_set")
    chain(); torch.cuda.synchronize()                      # trace off: no further stamps
    h = buf.cpu()
    n = int(h[0])
    assert n == 12, "6 launches x {entry, dependency wait returned}"
    st = h[2:2 + n]
    kind, t = st & 3, st >> 2
    assert int((kind == 0).sum()) == 6 and int((kind == 1).sum()) == 6
    tw = t[kind == 1]
    assert bool((tw[1:] >= tw[:-1]).all()) and int(tw[-1] - tw[0]) < 10_000_000      # ns; a handful of small launches
    assert int(h[2 + n:].abs().sum()) == 0
    small = torch.zeros(2 + 2, dtype=torch.int64, device=dev); small[1] = 2       # capacity 2: the counter runs on, the writes stop
    torch.cuda.synchronize()
    _C.check(L.upgpt_trace_set(small.data_ptr()), "trace_set")
    chain(); torch.cuda.synchronize()
    _C.check(L.upgpt_trace_set(None), "trace_set")
    hs = small.cpu()
    assert int(hs[0]) == 6 and int((hs[2:] != 0).sum()) == 2
