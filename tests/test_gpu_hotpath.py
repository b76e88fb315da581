"""Whole-path parity on the GPU: U-Net eps vs the reference's golden vectors and the CPU oracle, the fused DDIM / DDPM
samplers vs the oracle's restatement of the reference loops, the KL-f8 VAE decode, and size-independent properties
at BASELINE's full size (B=8, 32x32, bbox.yaml U-Net).

Tolerances (metric: max|a-b| / max|b| as defined in SURVEY.md 7.2):
  fp16 fast mode (opt-in): 2.5e-3 on eps (measured 1.3e-3..1.7e-3; single fp16 operand plane, fp32 everywhere else)
  fp16x3 (every GEMM/conv operand error-compensated with hi/lo fp16 planes): 5e-4 asserted, 1.3e-4..1.9e-4 measured --
                       this is the mode that meets BASELINE.json's 1e-3 tolerance.
"""
import numpy as np
import pytest
import torch

from conftest import tiny_ldm_config
from oracle import ldm_oracle as O
from oracle.make_golden import TINY_UNET_KW, TINY_VAE_KW
from oracle.ref_loader import BBOX_UNET_KW, BBOX_VAE_KW, UPSCALE_UNET_KW, UPSCALE_VAE_KW
from upgpt_b200 import synth

pytestmark = pytest.mark.gpu
EPS_TOL = 2.5e-3


def relerr(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _unet(kw, seed, dev):
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    m = UNetModel(**kw)
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m.to(dev).eval(), sd


@pytest.mark.parametrize("tag,kw,B,H,W,L,ts,seed", [
    ("tiny", TINY_UNET_KW, 2, 16, 16, 87, [981, 1], 0),
    ("tinyrect", TINY_UNET_KW, 3, 16, 24, 20, [500], 1),
    ("bbox", BBOX_UNET_KW, 1, 32, 32, 87, [981, 481], 0),
    ("upscale", UPSCALE_UNET_KW, 1, 32, 24, 86, [481], 2),      # the 4x super-resolution U-Net family
])
def test_unet_eps_vs_reference_golden(dev, golden, tag, kw, B, H, W, L, ts, seed):
    """Public UNetModel.forward (default precision: error-compensated fp16x3 operands, in "mixed" relaxed to single-plane fp16 in the
    deep levels of a >= 4-level U-Net) against the reference's own outputs: tolerance 1e-3 (BASELINE.json north_star), 5e-4 asserted;
    then the opt-in fp16 fast mode at its own bound."""
    from upgpt_b200.unet_engine import default_precision
    m, _ = _unet(kw, seed, dev)
    x, mask, ctx = synth.synth_inputs(B, H, W, L, kw["context_dim"], seed, concat_channels=kw["in_channels"] - kw["out_channels"])
    xc = torch.cat([x[:, :kw["out_channels"]], mask], 1).to(dev)
    for t in ts:
        tt = torch.full((B,), t, dtype=torch.long, device=dev)
        ref = torch.from_numpy(golden[f"{tag}_eps_t{t}"])
        with torch.no_grad():
            y = m(xc, tt, ctx.to(dev))                     # public UNetModel.forward (graph replay)
            eng = m.engine(B, H, W, L)
            assert eng.precision == default_precision() and eng.precision in ("fp16x3", "mixed")
            eng.stage_inputs(xc, tt); y_eager = eng.run(use_graph=False).clone()
        assert relerr(y, ref) < 5e-4
        assert torch.equal(y, y_eager), "graph replay must be bit-identical to the eager program (deterministic reductions)"
        fast = m.engine(B, H, W, L, precision="fp16")
        fast.set_context(ctx.to(dev)); fast.stage_inputs(xc, tt)
        yf = fast.run(use_graph=True).clone()
        assert relerr(yf, ref) < EPS_TOL
        assert torch.equal(yf, fast.run(use_graph=False))


@pytest.mark.parametrize("tag,kw,B,H,W,L,ts,seed", [
    ("tiny", TINY_UNET_KW, 2, 16, 16, 87, [981], 0),
    ("bbox", BBOX_UNET_KW, 1, 32, 32, 87, [981, 481], 0),
])
def test_unet_parity_mode_meets_1e3(dev, golden, tag, kw, B, H, W, L, ts, seed):
    """BASELINE.json's tolerance: eps within 1e-3 rel of the reference U-Net.  The error-compensated operand mode
    (precision="fp16x3": every GEMM/conv operand split into hi/lo fp16 planes, fp32 accumulate) meets it with margin
    (measured 1.3e-4 .. 1.9e-4); the default single-plane fp16 mode sits at 1.3e-3 .. 1.7e-3 (EPS_TOL above)."""
    m, _ = _unet(kw, seed, dev)
    x, mask, ctx = synth.synth_inputs(B, H, W, L, kw["context_dim"], seed)
    eng = m.engine(B, H, W, L, precision="fp16x3")
    eng.set_context(ctx.to(dev))
    for t in ts:
        eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), t, dtype=torch.long, device=dev))
        y = eng.run(use_graph=True).clone()
        assert relerr(y, torch.from_numpy(golden[f"{tag}_eps_t{t}"])) < 5e-4      # target 1e-3 (north_star), measured <= 1.9e-4
        assert torch.equal(y, eng.run(use_graph=False))


def test_unet_weight_repack_after_change(dev):
    """load_state_dict / ema_scope style in-place weight changes must invalidate the packed fp16 shadow copy."""
    m, sd = _unet(TINY_UNET_KW, 0, dev)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    xc, tt, c = torch.cat([x, mask], 1).to(dev), torch.full((2,), 500, dtype=torch.long, device=dev), ctx.to(dev)
    y0 = m(xc, tt, c)
    sd2 = synth.synth_state_dict(m.state_dict(), 5)
    m.load_state_dict(sd2)
    y1 = m(xc, tt, c)
    with torch.no_grad():
        ref = O.unet_forward(sd2, TINY_UNET_KW, xc.cpu(), tt.cpu(), ctx)
    assert relerr(y1, ref) < EPS_TOL and relerr(y1, y0) > 0.1


def _tiny_ldm(dev):
    from ldm.util import instantiate_from_config
    model = instantiate_from_config(tiny_ldm_config())
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("model.", "first_stage_model.", "extra_cond_models."))}
    sdn = synth.synth_state_dict(sd, 0)
    model.load_state_dict(sdn, strict=False)
    return model.to(dev).eval(), sdn


@pytest.mark.parametrize("S,eta", [(10, 0.0), (10, 1.0), (50, 0.0)])
def test_ddim_sampler_fused_vs_oracle(dev, S, eta):
    """DDIMSampler.sample (one CUDA graph per step) vs the oracle's restatement of ddim.py:114-204, same x_T / noise."""
    from ldm.models.diffusion.ddim import DDIMSampler
    model, sd = _tiny_ldm(dev)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    noises = torch.randn(S, *x.shape, generator=torch.Generator().manual_seed(123))
    sched = O.register_schedule(1000, 0.00085, 0.012)
    with torch.no_grad():
        ref, traj = O.ddim_sample(lambda xx, tt: O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx), x, S, eta, sched,
                                  noises if eta > 0 else None, return_all=True)
    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    kw = dict(conditioning=cond, eta=eta, x_T=x.to(dev), verbose=False, log_every_t=1, x_noise=noises.to(dev) if eta > 0 else None)
    out, inter = DDIMSampler(model).sample(S, 2, (4, 16, 16), **kw)
    assert len(inter["x_inter"]) == S + 1
    assert relerr(inter["x_inter"][1], traj[0]) < 2e-3          # first step
    tol = 2e-2 if S == 50 else 1e-2                                # fp16 operand noise accumulated over the chain
    assert relerr(out, ref) < tol
    out2, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), fused=False, **kw)     # general (python-loop) path
    assert torch.equal(out, out2), "fused graph loop and general loop must agree bit-for-bit"
    out3, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), **kw)
    assert torch.equal(out, out3), "sampling must be deterministic"


def test_ddim_mask_x0_blend_stochastic_encode_decode_vs_oracle(dev):
    """the known-region branch of ddim_sampling (ddim.py:144-147: img <- q_sample(x0, t) * mask + (1 - mask)
    * img before every step), DDPM.q_sample with per-sample t (ddpm.py:281-284), stochastic_encode and decode (ddim.py:207-240), all on
    upgpt_qsample_blend + the step graphs, against the oracle (pinned to the reference DDIMSampler's own outputs by
    tests/test_oracle_golden.py::test_ddim_mask_blend_and_encode_decode_match_reference_golden)."""
    from ldm.models.diffusion.ddim import DDIMSampler
    model, sd = _tiny_ldm(dev)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    g = torch.Generator().manual_seed(321)
    S = 10
    x0 = torch.randn(*x.shape, generator=g) * 0.8
    keep = (torch.rand(2, 1, 16, 16, generator=g) > 0.5).float()
    q_noises = torch.randn(S, *x.shape, generator=g)
    enc_noise = torch.randn(*x.shape, generator=g)
    sched = O.register_schedule(1000, 0.00085, 0.012)
    apply = lambda xx, tt: O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    # q_sample, per-sample t
    t = torch.tensor([981, 17])
    assert relerr(model.q_sample(x0.to(dev), t.to(dev), noise=enc_noise.to(dev)), O.q_sample(x0, t, sched, enc_noise)) < 1e-6
    # mask / x0 blend: fused step graphs, then the general loop; a (B, C, H, W) mask as well as the broadcast (B, 1, H, W) one
    with torch.no_grad():
        ref = O.ddim_sample(apply, x, S, 0.0, sched, mask=keep, x0=x0, q_noises=q_noises)
    kw = dict(conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False, mask=keep.to(dev), x0=x0.to(dev), x0_noise=q_noises.to(dev))
    out, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), **kw)
    assert relerr(out, ref) < 1e-2
    out_g, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), fused=False, **kw)
    assert torch.equal(out, out_g), "fused graph loop and general loop must agree bit-for-bit"
    out_c, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), **dict(kw, mask=keep.expand(2, 4, 16, 16).contiguous().to(dev)))
    assert torch.equal(out, out_c)
    kept = keep.expand_as(x0).bool()
    free, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False)
    assert relerr(out.cpu()[kept], ref[kept]) < 1e-2 and not torch.allclose(out.cpu()[kept], free.cpu()[kept], atol=1e-2)
    # stochastic_encode on the DDIM grid (per-sample index), then decode = the last t_start steps
    sampler = DDIMSampler(model)
    sampler.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
    t_idx = torch.tensor([6, 6])
    z = sampler.stochastic_encode(x0.to(dev), t_idx.to(dev), noise=enc_noise.to(dev))
    z_ref = O.stochastic_encode(x0, t_idx, S, sched, enc_noise)
    assert relerr(z, z_ref) < 1e-6
    t_mix = torch.tensor([6, 2])
    assert relerr(sampler.stochastic_encode(x0.to(dev), t_mix.to(dev), noise=enc_noise.to(dev)), O.stochastic_encode(x0, t_mix, S, sched, enc_noise)) < 1e-6
    dec = sampler.decode(z, cond, 6)
    with torch.no_grad():
        dec_ref = O.ddim_sample(apply, z_ref, S, 0.0, sched, t_start=6)
    assert relerr(dec, dec_ref) < 1e-2


def test_ddpm_ancestral_sampler_vs_oracle(dev):
    model, sd = _tiny_ldm(dev)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    T = 6
    noises = torch.randn(T, *x.shape, generator=torch.Generator().manual_seed(7))
    sched = O.register_schedule(1000, 0.00085, 0.012)
    img = x
    with torch.no_grad():
        for i, t in enumerate(reversed(range(T))):
            tt = torch.full((2,), t, dtype=torch.long)
            e = O.unet_forward(usd, TINY_UNET_KW, torch.cat([img, mask], 1), tt, ctx)
            img, _ = O.ddpm_step(img, e, tt, sched, noises[i])
    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    out = model.p_sample_loop(cond, tuple(x.shape), x_T=x.to(dev), timesteps=T, x_noise=noises.to(dev))
    assert relerr(out, img) < 5e-3


@pytest.mark.parametrize("tag,kw,hw", [("vaetiny", TINY_VAE_KW, 16), ("vaebbox", BBOX_VAE_KW, 32)])
def test_vae_decode_vs_reference_golden(dev, golden, tag, kw, hw):
    from ldm.models.autoencoder import AutoencoderKL
    ae = AutoencoderKL(kw, embed_dim=4)
    sd = synth.synth_state_dict(ae.state_dict(), 0)
    ae.load_state_dict(sd); ae = ae.to(dev).eval()
    z = synth.synth_inputs(1, hw, hw, 1, 8, 7)[0]
    y = ae.decode(z.to(dev), in_scale=1. / 0.18215)
    ref = torch.from_numpy(golden[f"{tag}_img_sub"])
    got = y[:, :, ::8, ::8] if hw == 32 else y
    assert relerr(got, ref) < 5e-3
    with torch.no_grad():
        full = O.decode_first_stage(sd, kw, z, 0.18215)
    assert relerr(y, full) < 5e-3
    z2 = torch.cat([z, z.flip(0) * 0.5 + 0.1], 0)                 # batch > 1 through a second engine
    y2 = ae.decode(z2.to(dev), in_scale=1. / 0.18215)
    assert relerr(y2[:1], y) < 5e-3, "decode of a sample must not depend on its batch neighbours (tiling differs with B)"


def test_latent_diffusion_end_to_end_bbox_yaml(dev):
    """configs/deepfashion/bbox.yaml unchanged -> LatentDiffusion -> log_images path (cond assembly, DDIM, decode)."""
    import os
    from conftest import ROOT
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    cfg.model.params["use_ema"] = False
    model = instantiate_from_config(cfg.model)
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("model.", "first_stage_model.", "extra_cond_models."))}
    model.load_state_dict(synth.synth_state_dict(sd, 0), strict=False)
    model = model.to(dev).eval()
    from ldm.modules.poses.poses import DummyModel
    model.extra_cond_models[0] = DummyModel()               # as InferenceModel does (generate_utils.py:142)
    g = torch.Generator().manual_seed(0)
    B = 2
    batch = {"txt": torch.randn(B, 77, 768, generator=g).to(dev), "styles": torch.randn(B, 9, 768, generator=g).to(dev),
             "smpl": torch.randn(B, 1, 85, generator=g).to(dev) * 0.5, "person_mask": torch.full((B, 1, 32, 24), -1.0).to(dev)}
    out = model.log_images(batch, N=B, ddim_steps=4, ddim_eta=1.0, seed=1, use_ema_scope=False)
    img = out["samples"]
    assert tuple(img.shape) == (B, 3, 256, 192) and torch.isfinite(img).all()
    z, c = model.get_input(batch, "image", bs=B)
    assert tuple(c["c_crossattn"].shape) == (B, 87, 768)
    ref_tok = torch.nn.functional.linear(batch["smpl"].cpu(), sd_cpu(model, "extra_cond_models.1.model.weight"), sd_cpu(model, "extra_cond_models.1.model.bias"))
    assert relerr(c["c_crossattn"][:, 86:], ref_tok) < 1e-5      # LinearProject SMPL token (poses.py:3-9)


def sd_cpu(model, k):
    return model.state_dict()[k].detach().float().cpu()


def test_full_size_properties_b8(dev):
    """BASELINE config-2 size (B=8, 32x32, 87 tokens): properties that do not need the CPU oracle at full size."""
    m, _ = _unet(BBOX_UNET_KW, 0, dev)
    B = 8
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
    xc, c = torch.cat([x, mask], 1).to(dev), ctx.to(dev)
    tt = torch.full((B,), 501, dtype=torch.long, device=dev)
    y = m(xc, tt, c)
    assert torch.isfinite(y).all() and float(y.abs().max()) > 0.1
    assert torch.equal(y, m(xc, tt, c)), "determinism"
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=dev)
    yp = m(xc[perm].contiguous(), tt, c[perm].contiguous())
    assert relerr(yp, y[perm]) < 1e-6, "samples are independent chains: permuting the batch permutes the output"
    y1 = m.engine(1, 32, 32, 87).forward(xc[:1].contiguous(), tt[:1], c[:1].contiguous())
    assert relerr(y1, y[:1]) < 5e-3, "batch-of-1 engine (different tiling / split-K) agrees with row 0 of the batch-of-8 engine"


def test_unet_eps_b8_benchmarked_config_vs_oracle(dev, monkeypatch):
    """The configuration bench.py times (BASELINE configs[1]: bbox.yaml U-Net, B=8, 32x32x4 latent, 87x768 context) against the CPU
    oracle AT B=8 (the B=8 engine tiles / splits differently from the B=1 engine the golden vectors pin), in the default precision
    plan -- CALIBRATED on these weights exactly as bench.py gets it (upgpt_b200/precision.py) -- and in uniform fp16x3, at the tolerance
    BASELINE.json states: eps max-rel < 1e-3."""
    import json, os
    from conftest import ROOT
    monkeypatch.setenv("UPGPT_CALIBRATE", "1")
    m, sd = _unet(BBOX_UNET_KW, 0, dev)
    B = 8
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
    xc = torch.cat([x, mask], 1)
    rec = {}
    for t in (981, 481):
        tt = torch.full((B,), t, dtype=torch.long)
        with torch.no_grad():
            ref = O.unet_forward(sd, BBOX_UNET_KW, xc, tt, ctx)
        for prec in (None, "fp16x3"):
            eng = m.engine(B, 32, 32, 87, precision=prec)
            eng.set_context(ctx.to(dev)); eng.stage_inputs(xc.to(dev), tt.to(dev))
            y = eng.run(use_graph=True).clone()
            e = relerr(y, ref)
            per_sample = max(relerr(y[i:i + 1], ref[i:i + 1]) for i in range(B))
            rec[f"{eng.precision}_t{t}"] = {"batch_max_rel": e, "worst_sample_max_rel": per_sample, "plan": eng.plan_name}
            assert e < 1e-3 and per_sample < 1e-3, (eng.precision, t, e, per_sample)
            assert torch.equal(y, eng.run(use_graph=False)), "graph replay == eager program"
    rec["calibration"] = m._plans[("raw", 32, 32, 87)][2]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "parity_b8.json"), "w"), indent=1)
    print("eps parity at B=8:", rec)


def test_precision_calibration_falls_back_when_the_limit_is_tight(dev, monkeypatch):
    """upgpt_b200/precision.py: the "mixed" plan is measured per checkpoint against the fp16x3 eps of the same weights; a tighter limit (a
    checkpoint that is more sensitive) must walk down the candidate list to a safer plan, a limit of 0 must end at uniform fp16x3, and
    new weights must trigger a new calibration."""
    from upgpt_b200 import precision as P
    monkeypatch.setenv("UPGPT_CALIBRATE", "1")
    m, sd = _unet(TINY_UNET_KW, 0, dev)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    xc, tt, c = torch.cat([x, mask], 1).to(dev), torch.full((2,), 481, dtype=torch.long, device=dev), ctx.to(dev)
    with torch.no_grad():
        ref = O.unet_forward(sd, TINY_UNET_KW, xc.cpu(), tt.cpu(), ctx)
    y = m(xc, tt, c)
    ver, plan, rep = m._plans[("raw", 16, 16, 87)]
    assert rep["chosen"] == plan["name"] and all(v >= 0 for v in rep["deviation_vs_fp16x3"].values())
    assert relerr(y, ref) < 1e-3
    assert len(m._engines) == 1, "calibration engines are throw-away: one engine (and one packed weight set) remains"
    plan0, rep0 = P.calibrate(m, 16, 16, 87, limit=-1.0)     # nothing passes (a plan that changes no layer of this U-Net deviates by exactly 0)
    assert plan0["name"] == "fp16x3" and len(rep0["deviation_vs_fp16x3"]) == len(P.candidates(16, 16, len(m.channel_mult))) - 1
    devs = rep0["deviation_vs_fp16x3"]
    first = next(iter(devs))
    plan1, _ = P.calibrate(m, 16, 16, 87, limit=devs[first] * 0.999)      # just too tight for the fastest candidate
    assert plan1["name"] != first
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), 7))
    m(xc, tt, c)
    assert m._plans[("raw", 16, 16, 87)][0] == m._weights_version != ver, "new weights -> new calibration"


def test_config4_smpl_interpolation_sequence_cond_cache(dev):
    """BASELINE configs[3] in miniature: keyframes alpha in linspace(1, 0, K) lerp the SMPL vector and the person mask between two
    poses (app.py:298-301) with text / style tokens fixed, one DDIM sample per keyframe through ONE sampler / engine. Each keyframe's
    context differs from the previous one only in its SMPL token: the K/V cond-cache must be rebuilt per keyframe (a stale cache
    would reproduce keyframe 0) and every keyframe must match the CPU oracle run on that keyframe's conditioning."""
    from ldm.models.diffusion.ddim import DDIMSampler
    model, sd = _tiny_ldm(dev)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    B, K, S = 2, 4, 4          # S must divide 1000 as in the reference (util.py:46-60 indexes alphas_cumprod[t+1])
    x, mask_a, ctx = synth.synth_inputs(B, 16, 16, 87, 128, 0)
    _, mask_b, _ = synth.synth_inputs(B, 16, 16, 87, 128, 5)
    g = torch.Generator().manual_seed(17)
    smpl_a, smpl_b = torch.randn(B, 1, 85, generator=g) * 0.5, torch.randn(B, 1, 85, generator=g) * 0.5
    Wp, bp = sd["extra_cond_models.1.model.weight"], sd["extra_cond_models.1.model.bias"]
    sched = O.register_schedule(1000, 0.00085, 0.012)
    sampler = DDIMSampler(model)
    outs = []
    for alpha in torch.linspace(1, 0, K).tolist():
        smpl = alpha * smpl_a + (1 - alpha) * smpl_b
        mask = alpha * mask_a + (1 - alpha) * mask_b
        tok_dev = model.extra_cond_models[1](smpl.to(dev))                    # LinearProject on the device (poses.py:3-9)
        tok_ref = torch.nn.functional.linear(smpl, Wp, bp)
        assert relerr(tok_dev, tok_ref) < 1e-5
        c_ref = torch.cat([ctx[:, :86], tok_ref], 1)
        cond = {"c_crossattn": torch.cat([ctx[:, :86].to(dev), tok_dev], 1), "c_concat": [mask.to(dev)]}
        z, _ = sampler.sample(S, B, (4, 16, 16), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False)
        with torch.no_grad():
            z_ref = O.ddim_sample(lambda xx, tt: O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, c_ref), x, S, 0.0, sched)
        assert relerr(z, z_ref) < 5e-3, "keyframe alpha=%g" % alpha
        outs.append(z)
        # the cond-cache itself: staged context == this keyframe's, and the cached K of the first cross-attention (context @ Wk^T,
        # head-padded) follows the SMPL token (row 86) -- a stale cache (e.g. keyed on a recycled device address) fails here
        eng = next(iter(model.model.diffusion_model._engines.values()))
        assert torch.equal(eng.bufs["ctx32"].cpu(), c_ref.to(eng.bufs["ctx32"].dtype)) or relerr(eng.bufs["ctx32"], c_ref) < 1e-6
        qn = next(k for k in eng.bufs if k.endswith(".ctx_kv"))
        wk = usd[qn[:-len(".ctx_kv")] + ".attn2.to_k.weight"]
        k_ref = (c_ref.reshape(-1, c_ref.shape[-1]) @ wk.t())                  # [B*87, heads*d]
        kc = eng.bufs[qn].float().cpu()
        kc = kc[:, :kc.shape[1] // 2]                                          # cond-cache rows are [K | V]
        heads = TINY_UNET_KW["num_heads"] if "num_heads" in TINY_UNET_KW else kc.shape[1] // 64
        dpad = kc.shape[1] // heads
        d = k_ref.shape[1] // heads
        kc = kc.reshape(-1, heads, dpad)[:, :, :d].reshape(-1, heads * d)
        assert relerr(kc[86::87], k_ref[86::87]) < 2e-3, "cached K row of the SMPL token, keyframe alpha=%g" % alpha
    assert not torch.equal(outs[1], outs[0]) and not torch.equal(outs[-1], outs[0])
    assert len(model.model.diffusion_model._engines) == 1, "all keyframes ran through one engine (one packed weight set, one step graph)"
    st = eng.cond.stats
    assert st["full_rebuilds"] == 1 and st["row_updates"] == K - 1, "keyframes after the first refresh the SMPL row (86) only: %r" % st
    assert eng.cond.row_version[86] == K and eng.cond.row_version[0] == 1


def test_cond_cache_row_refresh_and_style_slots(dev):
    """CondCache (upgpt_b200/cond_cache.py): a row refresh must leave the cache equal to a full rebuild on the same context, for the
    SMPL row (interpolation, app.py:296-301) and for replaced style slots (mix_style, generate_utils.py:172-190); many changed rows fall
    back to the full rebuild; eps through the refreshed cache equals eps through a rebuilt one."""
    m, _ = _unet(TINY_UNET_KW, 0, dev)
    B, L = 3, 87
    x, mask, ctx = synth.synth_inputs(B, 16, 16, L, 128, 0)
    eng = m.engine(B, 16, 16, L)
    xc, tt = torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev)
    c0 = ctx.to(dev)
    eng.set_context(c0)
    snap = lambda: {k: v.clone() for k, v in eng.bufs.items() if k.endswith(".ctx_kv")}
    c1 = c0.clone(); c1[:, 86] = torch.randn(B, 128, device=dev)            # new SMPL token
    assert eng.cond._changed_rows(c1) == [86]
    eng.set_context(c1)
    assert eng.cond.stats["row_updates"] == 1 and eng.cond.stats["full_rebuilds"] == 1
    part = snap()
    eng.stage_inputs(xc, tt); y_part = eng.run(use_graph=False).clone()
    eng.set_context(c1.clone(), force=True)                                    # rebuild every row on the same values
    full = snap()
    for k in part:
        assert relerr(part[k], full[k]) < 2e-3, k                              # fp16 cache rows; M = B tiles vs M = B*L tiles
        assert torch.equal(part[k].reshape(B, L, -1)[:, :86], full[k].reshape(B, L, -1)[:, :86])
    y_full = eng.run(use_graph=False).clone()
    assert relerr(y_part, y_full) < 1e-3
    # style slots 2 and 7 replaced (rows 79, 84), as mix_style does for masked / text-overridden styles
    emb = torch.randn(128, device=dev)
    eng.cond.set_style_slot(2, emb); eng.cond.set_style_slot(7, emb)
    c2 = c1.clone(); c2[:, 79] = emb; c2[:, 84] = emb
    assert torch.equal(eng.bufs["ctx32"], c2)
    part = snap()
    eng.set_context(c2, force=True)
    for k, v in snap().items():
        assert relerr(part[k], v) < 2e-3, k
    n_full = eng.cond.stats["full_rebuilds"]
    eng.set_context(torch.randn_like(c2))                                      # everything changed -> one k | v GEMM per layer
    assert eng.cond.stats["full_rebuilds"] == n_full + 1


def test_bbox_unet_ddim_trajectory_s10_vs_oracle(dev):
    """A sampler trajectory on the REAL bbox.yaml U-Net (the chain tests above run the tiny U-Net): DDIM S = 10, eta = 0, B = 2, fused
    step graphs vs the oracle's restatement of the reference loop. Bound: the first step within 1e-3 (one eps evaluation), the final
    latent within 1e-2 (fp16-operand eps errors of <= 4e-4 per step accumulated over the chain); the measured numbers are recorded."""
    import json, os
    from conftest import ROOT
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    cfg.model.params["use_ema"] = False
    model = instantiate_from_config(cfg.model)
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("model.diffusion_model.", "first_stage_model.", "extra_cond_models."))}
    sdn = synth.synth_state_dict(sd, 0)
    model.load_state_dict(sdn, strict=False)
    model = model.to(dev).eval()
    usd = {k[len("model.diffusion_model."):]: v for k, v in sdn.items() if k.startswith("model.diffusion_model.")}
    B, S = 2, 10
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 31)
    sched = O.register_schedule(1000, 0.00085, 0.012)
    with torch.no_grad():
        ref, traj = O.ddim_sample(lambda xx, tt: O.unet_forward(usd, BBOX_UNET_KW, torch.cat([xx, mask], 1), tt, ctx), x, S, 0.0, sched,
                                  return_all=True)
    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    out, inter = DDIMSampler(model).sample(S, B, (4, 32, 32), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False, log_every_t=1)
    errs = [relerr(inter["x_inter"][i + 1], traj[i]) for i in range(S)]
    rec = {"first_step": errs[0], "final": relerr(out, ref), "per_step": errs}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "parity_bbox_ddim_s10.json"), "w"), indent=1)
    print("bbox DDIM S=10 trajectory vs oracle:", rec)
    assert errs[0] < 1e-3 and rec["final"] < 1e-2


def test_config4_at_size_keyframes_bbox_unet(dev):
    """BASELINE configs[3] at its real size: the bbox.yaml U-Net, bs = 4, SMPL / mask lerp keyframes alpha in linspace(1, 0, K) with text
    and style tokens fixed (app.py:298-301), S = 10 DDIM steps per keyframe through ONE sampler / engine / step graph. Every keyframe's
    final latent is checked against the oracle trajectory (samples 0 and 1 of the batch -- samples are independent chains), and the
    cond-cache is checked per keyframe: only the SMPL row (86) of the 16 K | V caches is refreshed after the first keyframe."""
    import os
    from conftest import ROOT
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    cfg.model.params["use_ema"] = False
    model = instantiate_from_config(cfg.model)
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("model.diffusion_model.", "first_stage_model.", "extra_cond_models."))}
    sdn = synth.synth_state_dict(sd, 0)
    model.load_state_dict(sdn, strict=False)
    model = model.to(dev).eval()
    usd = {k[len("model.diffusion_model."):]: v for k, v in sdn.items() if k.startswith("model.diffusion_model.")}
    B, K, S, nref = 4, 3, 10, 2
    x, mask_a, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 40)
    _, mask_b, _ = synth.synth_inputs(B, 32, 32, 87, 768, 41)
    g = torch.Generator().manual_seed(17)
    smpl_a, smpl_b = torch.randn(B, 1, 85, generator=g) * 0.5, torch.randn(B, 1, 85, generator=g) * 0.5
    Wp, bp = sdn["extra_cond_models.1.model.weight"], sdn["extra_cond_models.1.model.bias"]
    sched = O.register_schedule(1000, 0.00085, 0.012)
    sampler = DDIMSampler(model)
    unet = model.model.diffusion_model
    for kf, alpha in enumerate(torch.linspace(1, 0, K).tolist()):
        smpl = alpha * smpl_a + (1 - alpha) * smpl_b
        mask = alpha * mask_a + (1 - alpha) * mask_b
        tok_dev = model.extra_cond_models[1](smpl.to(dev))
        tok_ref = torch.nn.functional.linear(smpl, Wp, bp)
        c_ref = torch.cat([ctx[:, :86], tok_ref], 1)
        cond = {"c_crossattn": torch.cat([ctx[:, :86].to(dev), tok_dev], 1), "c_concat": [mask.to(dev)]}
        z, _ = sampler.sample(S, B, (4, 32, 32), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False)
        with torch.no_grad():
            z_ref = O.ddim_sample(lambda xx, tt: O.unet_forward(usd, BBOX_UNET_KW, torch.cat([xx, mask[:nref]], 1), tt, c_ref[:nref]), x[:nref], S, 0.0, sched)
        e = relerr(z[:nref], z_ref)
        print("config-4 keyframe %d (alpha %.2f): final latent max-rel %.2e" % (kf, alpha, e))
        assert e < 1e-2, "keyframe alpha=%g" % alpha
        eng = next(iter(unet._engines.values()))
        st = eng.cond.stats
        assert len(unet._engines) == 1 and st["full_rebuilds"] == 1 and st["row_updates"] == kf, st
        qn = next(k for k in eng.bufs if k.endswith(".ctx_kv"))
        wk = usd[qn[:-len(".ctx_kv")] + ".attn2.to_k.weight"]
        k_ref = (c_ref.reshape(-1, 768) @ wk.t())
        kc = eng.bufs[qn].float().cpu()
        kc = kc[:, :kc.shape[1] // 2]
        heads, d = 8, k_ref.shape[1] // 8
        dpad = kc.shape[1] // heads
        kc = kc.reshape(-1, heads, dpad)[:, :, :d].reshape(-1, heads * d)
        assert relerr(kc, k_ref) < 2e-3, "cached K of every context row, keyframe %d" % kf


def test_config5_64x64_latent_b4(dev, golden):
    """BASELINE configs[4] shape: 64x64x4 latent (512x512 image), B=4, bbox.yaml U-Net. eps of sample 0 against the CPU oracle at
    B=1 (the oracle needs ~20 s for one 64x64 forward), plus size-independent properties at B=4 (N = 4096 self-attention keys,
    16-CTA GroupNorm clusters / two-launch fallback, column-tiled convs)."""
    m, sd = _unet(BBOX_UNET_KW, 0, dev)
    B = 4
    x, mask, ctx = synth.synth_inputs(B, 64, 64, 87, 768, 21)
    xc, c = torch.cat([x, mask], 1), ctx
    tt = torch.full((B,), 481, dtype=torch.long)
    y = m(xc.to(dev), tt.to(dev), c.to(dev))
    assert tuple(y.shape) == (B, 4, 64, 64) and torch.isfinite(y).all()
    assert torch.equal(y, m(xc.to(dev), tt.to(dev), c.to(dev))), "determinism"
    with torch.no_grad():
        ref0 = O.unet_forward(sd, BBOX_UNET_KW, xc[:1], tt[:1], c[:1])
    assert relerr(y[:1], ref0) < 1e-3, "eps tolerance of BASELINE.json at the 64x64 shape (fp16x3 parity mode)"
    perm = torch.tensor([2, 0, 3, 1])
    yp = m(xc[perm].contiguous().to(dev), tt.to(dev), c[perm].contiguous().to(dev))
    assert relerr(yp, y[perm.to(dev)]) < 1e-6


@pytest.mark.parametrize("tag,kw,hw", [("vaetiny", TINY_VAE_KW, 32), ("vaebbox", BBOX_VAE_KW, 64)])
def test_vae_encode_vs_reference_golden(dev, golden, tag, kw, hw):
    """AutoencoderKL.encode (Encoder with (0,1,0,1)-padded stride-2 convs as TMA phase planes, quant_conv,
    posterior sampling kernel) against the reference's own moments (golden) on a rectangular image."""
    from ldm.models.autoencoder import AutoencoderKL
    ae = AutoencoderKL(kw, embed_dim=4)
    sd = synth.synth_state_dict(ae.state_dict(), 0)
    ae.load_state_dict(sd); ae = ae.to(dev).eval()
    x = torch.tanh(torch.randn(1, 3, hw, hw + 16, generator=torch.Generator().manual_seed(11)))
    post = ae.encode(x.to(dev))
    ref = torch.from_numpy(golden[f"{tag}_enc_moments"])
    assert tuple(post.parameters.shape) == tuple(ref.shape)
    assert relerr(post.parameters, ref) < 5e-3
    noise = torch.randn(1, 4, *ref.shape[2:], generator=torch.Generator().manual_seed(3))
    z = post.sample(noise=noise.to(dev), scale=0.18215)
    assert relerr(z, O.gaussian_sample(post.parameters.cpu(), noise, 0.18215)) < 1e-5
    assert relerr(post.mode(), post.parameters[:, :4]) < 1e-7
    x2 = torch.cat([x, x.flip(0).flip(3) * 0.5], 0)                      # batch > 1 through a second engine
    assert relerr(ae.encode(x2.to(dev)).parameters[:1], post.parameters) < 5e-3


def test_vae_encode_full_size_and_reconstruction_path(dev):
    """Real KL-f8 encoder at the bbox.yaml image size (256x192, B=2) vs the CPU oracle, then LatentDiffusion.get_input /
    log_images' reconstruction branch (encode -> scale_factor * sample -> decode, ddpm.py:689-692,761)."""
    import os
    from conftest import ROOT
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    cfg.model.params["use_ema"] = False
    model = instantiate_from_config(cfg.model)
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("model.", "first_stage_model.", "extra_cond_models."))}
    sdn = synth.synth_state_dict(sd, 0)
    model.load_state_dict(sdn, strict=False)
    model = model.to(dev).eval()
    B = 2
    g = torch.Generator().manual_seed(5)
    img = torch.tanh(torch.randn(B, 256, 192, 3, generator=g))            # b h w c, as the data loaders emit
    vsd = {k[len("first_stage_model."):]: v for k, v in sdn.items() if k.startswith("first_stage_model.")}
    post = model.encode_first_stage(img.permute(0, 3, 1, 2).contiguous().to(dev))
    with torch.no_grad():
        m_ref = O.encode_first_stage_moments(vsd, BBOX_VAE_KW, img.permute(0, 3, 1, 2))
    assert tuple(post.parameters.shape) == (B, 8, 32, 24)
    assert relerr(post.parameters, m_ref) < 5e-3
    from ldm.modules.poses.poses import DummyModel
    model.extra_cond_models[0] = DummyModel()
    batch = {"image": img.to(dev), "txt": torch.randn(B, 77, 768, generator=g).to(dev), "styles": torch.randn(B, 9, 768, generator=g).to(dev),
             "smpl": torch.randn(B, 1, 85, generator=g).to(dev) * 0.5, "person_mask": torch.full((B, 1, 32, 24), -1.0).to(dev)}
    z, c, x, xrec = model.get_input(batch, "image", return_first_stage_outputs=True, bs=B)
    assert tuple(z.shape) == (B, 4, 32, 24) and tuple(xrec.shape) == (B, 3, 256, 192) and torch.isfinite(xrec).all()
    mean_z = 0.18215 * m_ref[:, :4]
    std_z = 0.18215 * torch.exp(0.5 * m_ref[:, 4:].clamp(-30, 20))
    assert float(((z.cpu() - mean_z).abs() / std_z).max()) < 6.0, "z is a draw from N(mean, std) * scale_factor"
    out = model.log_images(batch, N=B, ddim_steps=2, ddim_eta=0.0, seed=1, use_ema_scope=False)
    assert "reconstruction" in out and tuple(out["reconstruction"].shape) == (B, 3, 256, 192)


def test_plms_sampler_vs_oracle(dev):
    """PLMSSampler.sample (eps history combined by upgpt_lincomb4, DDIM update kernel with sigma = 0)
    against the reference PLMSSampler's own output (golden) and the oracle."""
    from ldm.models.diffusion.plms import PLMSSampler
    model, sd = _tiny_ldm(dev)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    out, inter = PLMSSampler(model).sample(10, 2, (4, 16, 16), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False, log_every_t=1)
    assert len(inter["x_inter"]) == 11
    # the oracle's PLMS is pinned against the reference PLMSSampler's own output by tests/test_oracle_golden.py (golden plms_S10_x0);
    # here it runs on this model's weights (name-seeded under the LatentDiffusion prefixes, so not the golden's tensors)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    sched = O.register_schedule(1000, 0.00085, 0.012)
    with torch.no_grad():
        ref, traj = O.plms_sample(lambda xx, tt: O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx), x, 10, sched, return_all=True)
    assert relerr(inter["x_inter"][1], traj[0]) < 2e-3          # pseudo improved Euler step
    assert relerr(out, ref) < 1e-2
    with pytest.raises(ValueError):
        PLMSSampler(model).sample(10, 2, (4, 16, 16), conditioning=cond, eta=1.0, x_T=x.to(dev), verbose=False)
    out2, _ = PLMSSampler(model).sample(10, 2, (4, 16, 16), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False)
    assert torch.equal(out, out2)


def test_classifier_free_guidance_vs_oracle(dev):
    """True classifier-free guidance with dict conditioning (ddim.py:171-178; the reference's torch.cat of dicts cannot run):
    e = e_u + s (e_c - e_u) with an unconditional context, through DDIM (general loop) and PLMS, vs the oracle."""
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.plms import PLMSSampler
    model, sd = _tiny_ldm(dev)
    usd = {k[len("model.diffusion_model."):]: v for k, v in sd.items() if k.startswith("model.diffusion_model.")}
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    uctx = torch.zeros_like(ctx)
    scale, S = 3.0, 5
    sched = O.register_schedule(1000, 0.00085, 0.012)

    def guided(xx, tt):
        e_c = O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
        e_u = O.unet_forward(usd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, uctx)
        return e_u + scale * (e_c - e_u)

    cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
    ucond = {"c_crossattn": uctx.to(dev), "c_concat": [mask.to(dev)]}
    kw = dict(conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False, unconditional_guidance_scale=scale, unconditional_conditioning=ucond)
    with torch.no_grad():
        ref_ddim = O.ddim_sample(guided, x, S, 0.0, sched)
        ref_plms = O.plms_sample(guided, x, S, sched)
    out, inter = DDIMSampler(model).sample(S, 2, (4, 16, 16), log_every_t=1, **kw)     # fused: one 2B-batch graph per step
    assert relerr(out, ref_ddim) < 1e-2
    assert tuple(out.shape) == (2, 4, 16, 16) and all(tuple(t.shape) == (2, 4, 16, 16) for t in inter["x_inter"] + inter["pred_x0"])
    out_g, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), fused=False, **kw)          # general loop: two passes per step
    assert relerr(out_g, ref_ddim) < 1e-2
    assert relerr(out, out_g) < 2e-3, "the 2B-batch graph and the two-pass loop agree"
    kw1 = dict(kw, eta=1.0, x_noise=torch.randn(S, 2, 4, 16, 16, generator=torch.Generator().manual_seed(3)).to(dev))
    o1, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), **kw1)
    o2, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), fused=False, **kw1)
    assert relerr(o1, o2) < 2e-3
    unguided, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), conditioning=cond, eta=0.0, x_T=x.to(dev), verbose=False)
    assert relerr(unguided, ref_ddim) > relerr(out, ref_ddim), "guidance changes the sample"
    outp, _ = PLMSSampler(model).sample(S, 2, (4, 16, 16), **kw)
    assert relerr(outp, ref_plms) < 1e-2


def test_upscale_model_full_size_and_kl_f4(dev):
    """the upscale configuration at its real shapes -- U-Net on a 128x96x3 latent + 3 low-res concat channels
    (12288 pixels at level 0, 3072-token self-attention at ds 2, GroupNorm on 12.6 MB images = the two-launch path), B=1 vs the
    CPU oracle; KL-f4 autoencoder (embed_dim 3, ch_mult [1,2,4]) encode + decode vs the oracle."""
    from ldm.models.autoencoder import AutoencoderKL
    m, sd = _unet(UPSCALE_UNET_KW, 2, dev)
    g = torch.Generator().manual_seed(9)
    xc = torch.randn(1, 6, 128, 96, generator=g); ctx = torch.randn(1, 86, 768, generator=g)
    tt = torch.full((1,), 481, dtype=torch.long)
    y = m(xc.to(dev), tt.to(dev), ctx.to(dev))
    with torch.no_grad():
        ref = O.unet_forward(sd, UPSCALE_UNET_KW, xc, tt, ctx)
    assert tuple(y.shape) == (1, 3, 128, 96)
    assert relerr(y, ref) < 1e-3
    ae = AutoencoderKL(UPSCALE_VAE_KW, embed_dim=3)
    vsd = synth.synth_state_dict(ae.state_dict(), 0)
    ae.load_state_dict(vsd); ae = ae.to(dev).eval()
    img = torch.tanh(torch.randn(1, 3, 128, 96, generator=g))
    post = ae.encode(img.to(dev))
    with torch.no_grad():
        m_ref = O.encode_first_stage_moments(vsd, UPSCALE_VAE_KW, img)
    assert tuple(post.parameters.shape) == (1, 6, 32, 24) and relerr(post.parameters, m_ref) < 5e-3
    z = post.mode()
    rec = ae.decode(z)
    with torch.no_grad():
        rec_ref = O.decode_first_stage(vsd, UPSCALE_VAE_KW, m_ref[:, :3], 1.0)
    assert tuple(rec.shape) == (1, 3, 128, 96) and relerr(rec, rec_ref) < 1e-2


def test_lanes_two_batches_in_flight_are_bit_identical(dev):
    """upgpt_b200/lanes.py: two independent batches sampled + decoded side by side (own stream, engines, step graphs and library scratch
    slot per lane; shared packed weights) give exactly the results of running them one after the other, on the tiny model and -- the
    cluster split-K / fork-join paths of the benchmarked configuration -- on the bbox U-Net at B = 8."""
    from ldm.models.diffusion.ddim import DDIMSampler
    from upgpt_b200 import lanes
    model, sd = _tiny_ldm(dev)
    S = 10
    ins = []
    for seed in (0, 1):
        x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, seed)
        noises = torch.randn(S, *x.shape, generator=torch.Generator().manual_seed(50 + seed))
        ins.append((x.to(dev), {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}, noises.to(dev)))

    def run(i):
        x, cond, noises = ins[i]
        z, _ = DDIMSampler(model).sample(S, 2, (4, 16, 16), conditioning=cond, eta=1.0, x_T=x, verbose=False, log_every_t=1000, x_noise=noises)
        return z, model.decode_first_stage(z)
    seq = [run(0), run(1)]                       # one after the other, lane 0
    torch.cuda.synchronize()
    outs = {}
    for rep in range(2):                         # first round builds lane 1's engines / graphs, second round is pure replay
        with lanes.lane(0):
            outs[0] = run(0)
        with lanes.lane(1):
            outs[1] = run(1)                     # enqueued while lane 0's batch is still running
        torch.cuda.synchronize()
        for i in (0, 1):
            assert torch.equal(outs[i][0], seq[i][0]) and torch.equal(outs[i][1], seq[i][1]), "lane %d, round %d" % (i, rep)
    unet = model.model.diffusion_model
    assert {k[-1] for k in unet._engines} == {0, 1} and len({id(e.w) for e in unet._engines.values()}) >= 1
    assert lanes.current() == 0
    # bbox U-Net at B = 8: one step per lane, concurrently, vs sequentially
    m, _ = _unet(BBOX_UNET_KW, 0, dev)
    xs = []
    for seed in (3, 5, 7):
        x, mask, ctx = synth.synth_inputs(8, 32, 32, 87, 768, seed)
        xs.append((torch.cat([x, mask], 1).to(dev), torch.full((8,), 481, dtype=torch.long, device=dev), ctx.to(dev)))
    ref = [m(*xs[i]).clone() for i in range(3)]
    torch.cuda.synchronize()
    for rep in range(2):
        ys = []
        for i in range(3):                       # three batches in flight (bench.py's default schedule)
            with lanes.lane(i):
                ys.append(m(*xs[i]))
        torch.cuda.synchronize()
        assert all(torch.equal(ys[i], ref[i]) for i in range(3)), "bbox U-Net, round %d" % rep


def test_throughput_mode_tiling_keeps_eps_parity(dev, golden):
    """lanes.set_throughput_mode(): the GEMM tiler counts SM time (upgpt_gemm_set_sm_weight), so small layers take fewer CTAs / split-K
    slices -- a different summation split, the same result within the tolerance; the tiling really changes; the mode is process-global and
    is switched back."""
    import ctypes as C_
    from upgpt_b200 import lanes, _C
    L = _C.lib()

    def plan_of(M, N, K):
        a = _C.GemmArgs(); a.mode, a.M, a.N, a.K = _C.GEMM_PLAIN, M, N, K
        a.out32 = 16       # (any non-null value: the plan only looks at shapes and flags)
        p = (C_.c_int * 8)()
        _C.check(L.upgpt_gemm_plan(C_.byref(a), C_.byref(p)), "plan")
        return int(p[0]), int(p[2])
    lat = plan_of(128, 896, 1792)
    try:
        assert lanes.set_throughput_mode(True) == lanes.THROUGHPUT_SM_WEIGHT
        thr = plan_of(128, 896, 1792)
        assert thr != lat and thr[1] <= lat[1], (lat, thr)
        m, _ = _unet(BBOX_UNET_KW, 0, dev)
        x, mask, ctx = synth.synth_inputs(1, 32, 32, 87, 768, 0)
        for t in (981, 481):
            y = m(torch.cat([x, mask], 1).to(dev), torch.full((1,), t, dtype=torch.long, device=dev), ctx.to(dev))
            assert relerr(y, torch.from_numpy(golden[f"bbox_eps_t{t}"])) < 1e-3
    finally:
        lanes.set_throughput_mode(False)
    assert plan_of(128, 896, 1792) == lat
    # the engine above was recorded under the throughput tiling. An eager replay re-plans every GEMM under the settings of the moment:
    # either the LayerNorm producers keep their N tiling (then the result is still right) or the launch fails loudly -- never garbage
    eng = m.engine(1, 32, 32, 87)
    eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((1,), 481, dtype=torch.long, device=dev))
    try:
        y = eng.run(use_graph=False)
        assert relerr(y, torch.from_numpy(golden["bbox_eps_t481"])) < 1e-3
    except _C.UpgptError as ex:
        assert "rebuild the engine" in str(ex)
