"""The CPU oracle (oracle/ldm_oracle.py) against the golden vectors produced by the REAL reference modules
(oracle/make_golden.py, run in the build container).  Pins the oracle on every box, no GPU needed."""
import numpy as np
import pytest
import torch

from oracle import ldm_oracle as O
from oracle.make_golden import TINY_UNET_KW, TINY_VAE_KW
from oracle.ref_loader import BBOX_UNET_KW, BBOX_VAE_KW, UPSCALE_UNET_KW, UPSCALE_VAE_KW
from upgpt_b200 import synth

TOL = 2e-5   # fp32 reassociation between the reference's nn.Modules and the functional restatement


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _unet_sd(kw, seed):
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    return synth.synth_state_dict(UNetModel(**kw).state_dict(), seed)


@pytest.mark.parametrize("tag,kw,B,H,W,L,ts,seed", [
    ("tiny", TINY_UNET_KW, 2, 16, 16, 87, [981, 1], 0),
    ("tinyrect", TINY_UNET_KW, 3, 16, 24, 20, [500], 1),
    ("bbox", BBOX_UNET_KW, 1, 32, 32, 87, [981, 481], 0),
    ("upscale", UPSCALE_UNET_KW, 1, 32, 24, 86, [481], 2),      # models/upgpt/upscale/config.yaml: 6 -> 3 channels, 86 tokens
])
def test_unet_eps_matches_reference_golden(golden, tag, kw, B, H, W, L, ts, seed):
    sd = _unet_sd(kw, seed)
    x, mask, ctx = synth.synth_inputs(B, H, W, L, kw["context_dim"], seed, concat_channels=kw["in_channels"] - kw["out_channels"])
    x = x[:, :kw["out_channels"]]
    for t in ts:
        with torch.no_grad():
            y = O.diffusion_wrapper_hybrid(sd, kw, x, torch.full((B,), t, dtype=torch.long), [mask], [ctx])
        ref = torch.from_numpy(golden[f"{tag}_eps_t{t}"])
        assert y.shape == ref.shape
        assert ref.abs().max() > 0.1, "golden eps must not be the all-zero output of a zero_module-initialised U-Net"
        assert relerr(y, ref) < TOL


@pytest.mark.parametrize("S,eta", [(50, 0.0), (10, 1.0)])
def test_ddim_sampler_matches_reference_golden(golden, S, eta):
    sd = _unet_sd(TINY_UNET_KW, 0)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    sched = O.register_schedule(1000, 0.00085, 0.012)
    noises = torch.randn(S, *x.shape, generator=torch.Generator().manual_seed(123))
    apply = lambda xx, tt: O.unet_forward(sd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
    with torch.no_grad():
        x0 = O.ddim_sample(apply, x, S, eta, sched, noises if eta > 0 else None)
    assert relerr(x0, torch.from_numpy(golden[f"ddim_S{S}_eta{int(eta)}_x0"])) < 1e-4
    ts, alphas, alphas_prev, sigmas, _ = O.ddim_schedule(sched["alphas_cumprod"], S, eta)
    np.testing.assert_array_equal(ts, golden[f"ddim_S{S}_eta{int(eta)}_timesteps"])
    np.testing.assert_allclose(np.asarray(alphas, dtype=np.float64), golden[f"ddim_S{S}_eta{int(eta)}_alphas"], rtol=0, atol=0)
    np.testing.assert_allclose(alphas_prev, golden[f"ddim_S{S}_eta{int(eta)}_alphas_prev"], rtol=0, atol=0)
    np.testing.assert_allclose(np.asarray(sigmas, dtype=np.float64), golden[f"ddim_S{S}_eta{int(eta)}_sigmas"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("tag,kw,hw", [("vaetiny", TINY_VAE_KW, 16), ("vaebbox", BBOX_VAE_KW, 32)])
def test_vae_decode_matches_reference_golden(golden, tag, kw, hw):
    from ldm.models.autoencoder import AutoencoderKL
    sd = synth.synth_state_dict(AutoencoderKL(kw, embed_dim=4).state_dict(), 0)
    z = synth.synth_inputs(1, hw, hw, 1, 8, 7)[0]
    with torch.no_grad():
        y = O.decode_first_stage(sd, kw, z, 0.18215)
    ref = torch.from_numpy(golden[f"{tag}_img_sub"])
    got = y[:, :, ::8, ::8] if hw == 32 else y
    assert relerr(got, ref) < TOL
    mean, std, amax = golden[f"{tag}_stats"]
    assert abs(float(y.mean()) - mean) < 1e-4 * max(1.0, abs(amax)) and abs(float(y.abs().max()) - amax) < 1e-4 * amax


@pytest.mark.parametrize("tag,kw,hw", [("vaetiny", TINY_VAE_KW, 32), ("vaebbox", BBOX_VAE_KW, 64)])
def test_vae_encode_matches_reference_golden(golden, tag, kw, hw):
    """Encoder + quant_conv restatement (asymmetric (0,1,0,1) pad before the stride-2 convs) vs the reference's own moments."""
    from ldm.models.autoencoder import AutoencoderKL
    sd = synth.synth_state_dict(AutoencoderKL(kw, embed_dim=4).state_dict(), 0)
    x = torch.tanh(torch.randn(1, 3, hw, hw + 16, generator=torch.Generator().manual_seed(11)))
    with torch.no_grad():
        m = O.encode_first_stage_moments(sd, kw, x)
    assert relerr(m, torch.from_numpy(golden[f"{tag}_enc_moments"])) < TOL
    # posterior sampling identities (distributions.py:24-37)
    noise = torch.randn(1, 4, *m.shape[2:], generator=torch.Generator().manual_seed(3))
    z = O.gaussian_sample(m, noise, 0.18215)
    assert relerr(O.gaussian_sample(m, None, 0.18215), 0.18215 * m[:, :4]) < 1e-7
    assert float((z / 0.18215 - m[:, :4] - torch.exp(0.5 * m[:, 4:].clamp(-30, 20)) * noise).abs().max()) < 1e-5


def test_plms_sampler_matches_reference_golden(golden):
    """plms.py restatement (pseudo improved Euler + Adams-Bashforth orders 2..4) vs the reference PLMSSampler's own output."""
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    sd = synth.synth_state_dict(UNetModel(**TINY_UNET_KW).state_dict(), 0)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    sched = O.register_schedule(1000, 0.00085, 0.012)
    with torch.no_grad():
        got = O.plms_sample(lambda xx, tt: O.unet_forward(sd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx), x, 10, sched)
    assert relerr(got, torch.from_numpy(golden["plms_S10_x0"])) < 1e-4


def test_ddpm_step_closed_form():
    """q_posterior / predict_start_from_noise identities (ddpm.py:224-237): with eps = true noise, x0 is recovered."""
    sched = O.register_schedule(1000, 0.00085, 0.012)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 4, 8, 8, generator=g); eps = torch.randn(2, 4, 8, 8, generator=g)
    t = torch.tensor([500, 500])
    xt = sched["sqrt_alphas_cumprod"][t].reshape(-1, 1, 1, 1) * x0 + sched["sqrt_one_minus_alphas_cumprod"][t].reshape(-1, 1, 1, 1) * eps
    x_prev, x0_hat = O.ddpm_step(xt, eps, t, sched, torch.zeros_like(xt))
    assert relerr(x0_hat, x0) < 1e-4


def _mask_case():
    """The inputs oracle/make_golden.py drew for the mask / x0 and stochastic_encode -> decode cases (generator seed 321)."""
    g = torch.Generator().manual_seed(321)
    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    x0 = torch.randn(*x.shape, generator=g) * 0.8
    keep = (torch.rand(2, 1, 16, 16, generator=g) > 0.5).float()
    q_noises = torch.randn(10, *x.shape, generator=g)
    enc_noise = torch.randn(*x.shape, generator=g)
    return x, mask, ctx, x0, keep, q_noises, enc_noise


def test_ddim_mask_blend_and_encode_decode_match_reference_golden(golden):
    """ddim.py:144-147 (img <- q_sample(x0, t) * mask + (1 - mask) * img before every step) and ddim.py:207-240 (stochastic_encode on
    the DDIM grid, decode = the last t_start steps) vs the reference DDIMSampler's own outputs."""
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    sd = synth.synth_state_dict(UNetModel(**TINY_UNET_KW).state_dict(), 0)
    x, mask, ctx, x0, keep, q_noises, enc_noise = _mask_case()
    sched = O.register_schedule(1000, 0.00085, 0.012)
    apply = lambda xx, tt: O.unet_forward(sd, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
    with torch.no_grad():
        got = O.ddim_sample(apply, x, 10, 0.0, sched, mask=keep, x0=x0, q_noises=q_noises)
        enc = O.stochastic_encode(x0, torch.tensor([6, 6]), 10, sched, enc_noise)
        dec = O.ddim_sample(apply, enc, 10, 0.0, sched, t_start=6)
    assert relerr(got, torch.from_numpy(golden["ddim_mask_S10_x0"])) < 1e-4
    assert relerr(enc, torch.from_numpy(golden["ddim_encode_S10_t6"])) < 1e-6
    assert relerr(dec, torch.from_numpy(golden["ddim_decode_S10_t6"])) < 1e-4
