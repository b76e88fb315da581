"""GPU bring-up: per-op checks (norm / layernorm / attention) and whole-U-Net eps parity vs the CPU oracle + goldens."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
from upgpt_b200 import _C, ops, synth
from oracle import ldm_oracle as O

dev = torch.device("cuda:0")
res = []


def report(name, got, ref, tol):
    got = got.float().cpu(); ref = ref.float().cpu()
    finite = bool(torch.isfinite(got).all())
    err = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9)).item() if finite else float("nan")
    l2 = ((got - ref).norm() / ref.norm()).item() if finite else float("nan")
    ok = finite and err < tol
    res.append(dict(name=name, err=err, l2=l2, ok=ok))
    print(("PASS " if ok else "FAIL ") + f"{name}: max-rel {err:.3e} l2-rel {l2:.3e} (tol {tol:g})", flush=True)


def check_norms():
    g = torch.Generator().manual_seed(0)
    for (B, H, W, C1, C2, silu, layout) in [(2, 16, 16, 224, 0, True, 0), (2, 8, 8, 448, 224, True, 0), (1, 32, 32, 128, 0, False, 0),
                                            (2, 8, 8, 64, 0, False, 1), (2, 8, 8, 64, 0, False, 2), (3, 4, 3, 896, 896, True, 0)]:
        x1 = (torch.randn(B, H * W, C1, generator=g) * 2 + 0.5).to(dev)
        x2 = (torch.randn(B, H * W, C2, generator=g) - 1.0).to(dev) if C2 else None
        Cc = C1 + C2
        gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev); beta = (0.1 * torch.randn(Cc, generator=g)).to(dev)
        xc = torch.cat([x1, x2], -1) if C2 else x1
        xn = xc.reshape(B, H, W, Cc).permute(0, 3, 1, 2)
        use_norm = layout == 0
        if use_norm:
            ref = F.group_norm(xn, 32, gamma, beta, 1e-5)
            if silu: ref = F.silu(ref)
        else:
            ref = xn
        stats = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64)
        mult = 4 if layout == 1 else 1
        out = torch.zeros(B * H * W * mult * Cc, device=dev, dtype=torch.half)
        raw = torch.zeros(B * H * W * Cc, device=dev, dtype=torch.half) if layout == 0 else None
        if use_norm:
            ops.groupnorm_stats(x1, x2, B, H * W, stats)
        ops.prep(x1=x1, C1=C1, x2=x2, C2=C2, B=B, H=H, W=W, groups=32, stats=stats if use_norm else None, gamma=gamma if use_norm else None,
                 beta=beta if use_norm else None, eps=1e-5, silu=int(silu), layout=layout, split3=0, out=out, raw=raw)
        torch.cuda.synchronize()
        if layout == 0:
            got = out.reshape(B, H, W, Cc).permute(0, 3, 1, 2)
            report(f"gn+prep B{B} {H}x{W} C{C1}+{C2} silu{int(silu)}", got, ref, 2e-3)
            report("  raw16 copy", raw.reshape(B, H, W, Cc).permute(0, 3, 1, 2), xn, 1e-3)
        elif layout == 1:
            got = out.reshape(B, 2 * H, 2 * W, Cc).permute(0, 3, 1, 2)
            report(f"prep up2 B{B} {H}x{W} C{C1}", got, F.interpolate(xn, scale_factor=2, mode="nearest"), 1e-3)
        else:
            got = out.reshape(4, B, H // 2, W // 2, Cc)
            refp = torch.stack([xc.reshape(B, H, W, Cc)[:, p::2, q::2] for p in (0, 1) for q in (0, 1)], 0)
            report(f"prep s2phase B{B} {H}x{W} C{C1}", got, refp, 1e-3)
    # split3 planes
    x1 = torch.randn(2, 64, 64, generator=g).to(dev)
    out = torch.zeros(2 * 64 * 192, device=dev, dtype=torch.half)
    ops.prep(x1=x1, C1=64, x2=None, C2=0, B=2, H=8, W=8, groups=32, stats=None, gamma=None, beta=None, eps=0.0, silu=0, layout=0, split3=1, out=out, raw=None)
    o = out.reshape(128, 3, 64).float()
    report("prep split3 hi+lo", o[:, 0] + o[:, 1], x1.reshape(128, 64), 1e-6)
    report("prep split3 hi==hi", o[:, 2], o[:, 0], 1e-9)
    for rows, Cc in [(300, 224), (64, 448), (17, 896)]:
        x = (torch.randn(rows, Cc, generator=g) * 3 + 1).to(dev)
        gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev); beta = (0.1 * torch.randn(Cc, generator=g)).to(dev)
        out = torch.zeros(rows, Cc, device=dev, dtype=torch.half)
        ops.layernorm(x, gamma, beta, out)
        report(f"layernorm {rows}x{Cc}", out, F.layer_norm(x, (Cc,), gamma, beta, 1e-5), 2e-3)


def check_attention():
    g = torch.Generator().manual_seed(1)
    for (B, Hh, Nq, Nk, d, dpad) in [(1, 1, 128, 128, 64, 64), (2, 8, 1024, 1024, 28, 64), (2, 8, 256, 256, 56, 64), (2, 8, 64, 64, 112, 128),
                                     (2, 8, 16, 16, 112, 128), (2, 8, 1024, 87, 28, 64), (2, 8, 64, 87, 112, 128), (1, 4, 384, 300, 16, 64),
                                     (1, 8, 4096, 4096, 28, 64)]:
        q = torch.randn(B, Hh, Nq, d, generator=g).to(dev); k = torch.randn(B, Hh, Nk, d, generator=g).to(dev)
        v = torch.randn(B, Hh, Nk, d, generator=g).to(dev)
        scale = d ** -0.5
        qh, kh, vh = q.half().float(), k.half().float(), v.half().float()
        ref = torch.softmax(torch.einsum("bhid,bhjd->bhij", qh, kh) * scale, -1) @ vh
        HD = Hh * dpad
        Q = torch.zeros(B, Nq, Hh, dpad, device=dev, dtype=torch.half); Q[..., :d] = q.permute(0, 2, 1, 3).half()
        K = torch.zeros(B, Nk, Hh, dpad, device=dev, dtype=torch.half); K[..., :d] = k.permute(0, 2, 1, 3).half()
        ldvt = (Nk + 7) // 8 * 8
        Vt = torch.zeros(B, Hh, dpad, ldvt, device=dev, dtype=torch.half); Vt[:, :, :d, :Nk] = v.permute(0, 1, 3, 2).half()
        out = torch.full((B, Nq, Hh, dpad), float("nan"), device=dev, dtype=torch.half)
        try:
            ops.attention(q=Q, ldq=HD, k=K, ldk=HD, k_batch_stride=0, vt=Vt, ldvt=ldvt, out=out, ldo=HD, B=B, H=Hh, Nq=Nq, Nk=Nk, dpad=dpad, scale=scale)
            torch.cuda.synchronize()
            report(f"attention B{B} H{Hh} Nq{Nq} Nk{Nk} d{d}", out[..., :d].permute(0, 2, 1, 3), ref, 3e-3)
        except Exception as e:
            print("FAIL attention", (B, Hh, Nq, Nk, d), e); res.append(dict(name="attention", ok=False, err=None))


def check_unet(tag, kw, B, H, W, ctx_len, t_values, seed, golden):
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    m = UNetModel(**kw)
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    x, mask, ctx = synth.synth_inputs(B, H, W, ctx_len, kw["context_dim"], seed)
    xc = torch.cat([x, mask], 1)
    for prec in ("fp16", "fp16x3"):
        for t in t_values:
            tt = torch.full((B,), t, dtype=torch.long)
            with torch.no_grad():
                t0 = time.time()
                eng = m.engine(B, H, W, ctx_len, precision=prec)
                eng.set_context(ctx.to(dev)); eng.stage_inputs(xc.to(dev), tt.to(dev))
                y = eng.run(use_graph=False).clone()
                torch.cuda.synchronize()
                dt = time.time() - t0
            gold = torch.from_numpy(golden[f"{tag}_eps_t{t}"])
            report(f"unet[{tag}] {prec} t={t} eager vs golden(reference) ({dt:.2f}s, {eng.launches_per_step} launches)", y, gold, 2e-3 if prec == "fp16" else 5e-4)
            with torch.no_grad():
                y2 = eng.run(use_graph=True).clone(); y3 = eng.run(use_graph=True).clone()
                torch.cuda.synchronize()
            report(f"unet[{tag}] {prec} t={t} graph replay == eager", y3, y, 1e-6)
    return m


if __name__ == "__main__":
    which = sys.argv[1:] or ["norms", "attention", "unet"]
    print(torch.cuda.get_device_name(0))
    if "norms" in which: check_norms()
    if "attention" in which: check_attention()
    if "unet" in which:
        golden = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from oracle.make_golden import TINY_UNET_KW
        from oracle.ref_loader import BBOX_UNET_KW
        try:
            check_unet("tiny", TINY_UNET_KW, 2, 16, 16, 87, [981, 1], 0, golden)
            check_unet("tinyrect", TINY_UNET_KW, 3, 16, 24, 20, [500], 1, golden)
            m = check_unet("bbox", BBOX_UNET_KW, 1, 32, 32, 87, [981, 481], 0, golden)
            # timing at B=8
            B = 8
            x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
            eng = m.engine(B, 32, 32, 87, precision="fp16")
            eng.set_context(ctx.to(dev)); eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long).to(dev))
            for _ in range(3): eng.run(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): eng.run(True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"U-Net bbox B=8 32x32 graph step: {ms:.3f} ms  ({8 * 91.03 / ms:.1f} TFLOP/s algorithmic), {eng.launches_per_step} launches/step")
            e0.record()
            for _ in range(3): eng.run(False)
            e1.record(); torch.cuda.synchronize()
            print(f"U-Net bbox B=8 eager step: {e0.elapsed_time(e1) / 3:.3f} ms")
        except Exception as e:
            import traceback; traceback.print_exc(); res.append(dict(name="unet-exception", ok=False, err=None))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "unet_check.json"), "w"), indent=1)
    print("TOTAL %d checks, %d failed" % (len(res), sum(1 for r in res if not r["ok"])))
