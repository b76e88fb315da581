"""TEST INFRASTRUCTURE ONLY -- pins the oracle and writes tests/golden/*.npz.

Runs ONLY in the build container (needs /root/reference).  For each case it
  1. builds the REAL reference module (UNetModel / Decoder / DDIMSampler imported from /root/reference),
  2. loads the deterministic synthetic weights of upgpt_b200.synth (name-seeded, reproducible on the GPU box),
  3. runs the reference on seeded inputs,
  4. checks oracle/ldm_oracle.py reproduces it (max-abs-diff printed, asserted tight),
  5. stores inputs' seeds + the reference outputs as small fixtures.

    python -m oracle.make_golden
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ldm_oracle as O            # noqa: E402
from oracle import ref_loader                 # noqa: E402
from upgpt_b200 import synth                  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_UNET_KW = dict(image_size=16, in_channels=5, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
                    num_res_blocks=1, channel_mult=[1, 2, 2], num_heads=4, use_spatial_transformer=True,
                    transformer_depth=1, context_dim=128, use_checkpoint=False, legacy=False)
TINY_VAE_KW = dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2],
                   num_res_blocks=1, attn_resolutions=[], dropout=0.0)


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def unet_case(ref, kw, B, H, W, ctx_len, t_values, seed, tag, out):
    torch.manual_seed(0)
    m = ref.UNetModel(**kw).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    x, mask, ctx = synth.synth_inputs(B, H, W, ctx_len, kw["context_dim"], seed, concat_channels=kw["in_channels"] - kw["out_channels"])
    xc = torch.cat([x[:, :kw["out_channels"]], mask], 1)
    for t in t_values:
        tt = torch.full((B,), t, dtype=torch.long)
        with torch.no_grad():
            t0 = time.time(); y_ref = m(xc, tt, ctx); dt = time.time() - t0
            y_or = O.unet_forward(sd, kw, xc, tt, ctx)
        e = relerr(y_or, y_ref)
        print(f"[{tag}] t={t}: reference fwd {dt:.2f}s |eps|max={y_ref.abs().max():.4f} oracle-vs-reference max-rel={e:.2e}")
        assert e < 2e-5, "oracle restatement deviates from the reference"
        out[f"{tag}_eps_t{t}"] = y_ref.numpy()
    out[f"{tag}_meta"] = np.array([B, H, W, ctx_len, seed], dtype=np.int64)
    return m, sd


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    out = {}
    # ---- U-Net: tiny config (fast everywhere) and the real bbox.yaml config (C1: B=1, 32x32, t in {981, 481}) ----
    m_tiny, sd_tiny = unet_case(ref, TINY_UNET_KW, 2, 16, 16, 87, [981, 1], 0, "tiny", out)
    unet_case(ref, TINY_UNET_KW, 3, 16, 24, 20, [500], 1, "tinyrect", out)
    unet_case(ref, ref_loader.BBOX_UNET_KW, 1, 32, 32, 87, [981, 481], 0, "bbox", out)
    # the second U-Net family (models/upgpt/upscale/config.yaml): 6 -> 3 channels, 256 model channels, attention at ds 2/4/8, 86 tokens
    unet_case(ref, ref_loader.UPSCALE_UNET_KW, 1, 32, 24, 86, [481], 2, "upscale", out)

    # ---- DDIM: reference DDIMSampler over the tiny U-Net (eta = 0 and eta = 1 with injected noise) ----
    sched = O.register_schedule(1000, 0.00085, 0.012)

    class Shim:  # the duck-typed `model` DDIMSampler needs
        num_timesteps = 1000
        betas, alphas_cumprod, alphas_cumprod_prev = sched["betas"], sched["alphas_cumprod"], sched["alphas_cumprod_prev"]
        device = torch.device("cpu")
        parameterization = "eps"

        def __init__(self, mask, ctx):
            self.mask, self.ctx = mask, ctx

        def apply_model(self, x, t, c):
            with torch.no_grad():
                return m_tiny(torch.cat([x, self.mask], 1), t, self.ctx)

        def q_sample(self, x_start, t, noise=None):      # DDPM.q_sample (ddpm.py:281-284) with injected noise (mask / x0 path)
            noise = next(self.q_noise) if noise is None else noise
            e = lambda a: a[t].reshape(-1, 1, 1, 1)
            return e(sched["sqrt_alphas_cumprod"]) * x_start + e(sched["sqrt_one_minus_alphas_cumprod"]) * noise

    x, mask, ctx = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    shim = Shim(mask, ctx)
    for S, eta in ((50, 0.0), (10, 1.0)):
        sampler = ref.DDIMSampler(shim)
        g = torch.Generator().manual_seed(123)
        noises = torch.randn(S, *x.shape, generator=g)
        if eta > 0:
            it = iter(noises)
            # p_sample_ddim calls noise_like(x.shape, device, repeat_noise) (ddim.py:199): inject our draws
            ref.DDIMSampler.make_schedule.__globals__["noise_like"] = lambda shape, device, repeat=False: next(it)
        with torch.no_grad():
            samples, inter = sampler.sample(S, 2, (4, 16, 16), conditioning=None, eta=eta, x_T=x, verbose=False,
                                            log_every_t=1)
        apply = lambda xx, tt: O.unet_forward(sd_tiny, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
        with torch.no_grad():
            mine = O.ddim_sample(apply, x, S, eta, sched, noises if eta > 0 else None)
        e = relerr(mine, samples)
        print(f"[ddim S={S} eta={eta}] oracle-vs-reference sampler max-rel={e:.2e}")
        assert e < 1e-4
        out[f"ddim_S{S}_eta{int(eta)}_x0"] = samples.numpy()
        out[f"ddim_S{S}_eta{int(eta)}_timesteps"] = np.asarray(sampler.ddim_timesteps)
        out[f"ddim_S{S}_eta{int(eta)}_alphas"] = np.asarray(sampler.ddim_alphas, dtype=np.float64)
        out[f"ddim_S{S}_eta{int(eta)}_alphas_prev"] = np.asarray(sampler.ddim_alphas_prev, dtype=np.float64)
        out[f"ddim_S{S}_eta{int(eta)}_sigmas"] = np.asarray(sampler.ddim_sigmas, dtype=np.float64)

    # ---- DDIM mask / x0 blending (ddim.py:144-147), stochastic_encode -> decode (ddim.py:207-240): reference DDIMSampler, S = 10 ----
    g = torch.Generator().manual_seed(321)
    S = 10
    x0 = torch.randn(*x.shape, generator=g) * 0.8
    keep = (torch.rand(2, 1, 16, 16, generator=g) > 0.5).float()          # 1 = keep the known latent x0, 0 = generate
    q_noises = torch.randn(S, *x.shape, generator=g)
    shim.q_noise = iter(q_noises)
    # p_sample_ddim always calls noise_like (times sigma = 0 here): the iterator injected above for the eta = 1 case is exhausted
    ref.DDIMSampler.make_schedule.__globals__["noise_like"] = lambda shape, device, repeat=False: torch.zeros(shape)
    sampler = ref.DDIMSampler(shim)
    with torch.no_grad():
        samples, _ = sampler.sample(S, 2, (4, 16, 16), conditioning=None, eta=0.0, x_T=x, mask=keep, x0=x0, verbose=False)
        apply = lambda xx, tt: O.unet_forward(sd_tiny, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
        mine = O.ddim_sample(apply, x, S, 0.0, sched, mask=keep, x0=x0, q_noises=q_noises)
    e = relerr(mine, samples)
    print(f"[ddim mask/x0 S={S}] oracle-vs-reference max-rel={e:.2e}")
    assert e < 1e-4
    out["ddim_mask_S10_x0"] = samples.numpy()
    enc_noise = torch.randn(*x.shape, generator=g)
    t_enc = 6
    with torch.no_grad():
        z_enc = sampler.stochastic_encode(x0, torch.tensor([t_enc] * 2), noise=enc_noise)
        z_dec = sampler.decode(z_enc, None, t_enc)
        mine_enc = O.stochastic_encode(x0, torch.tensor([t_enc] * 2), S, sched, enc_noise)
        mine_dec = O.ddim_sample(apply, mine_enc, S, 0.0, sched, t_start=t_enc)
    e1, e2 = relerr(mine_enc, z_enc), relerr(mine_dec, z_dec)
    print(f"[ddim stochastic_encode t={t_enc} -> decode] oracle-vs-reference max-rel={e1:.2e} / {e2:.2e}")
    assert e1 < 1e-5 and e2 < 1e-4
    out["ddim_encode_S10_t6"] = z_enc.numpy()
    out["ddim_decode_S10_t6"] = z_dec.numpy()

    # ---- PLMS ('next' row 8(f)-3): reference PLMSSampler over the tiny U-Net, S = 10 (exercises all four multistep orders) ----
    sampler = ref.PLMSSampler(shim)
    with torch.no_grad():
        samples, _ = sampler.sample(10, 2, (4, 16, 16), conditioning=None, eta=0.0, x_T=x, verbose=False, log_every_t=1)
        apply = lambda xx, tt: O.unet_forward(sd_tiny, TINY_UNET_KW, torch.cat([xx, mask], 1), tt, ctx)
        mine = O.plms_sample(apply, x, 10, sched)
    e = relerr(mine, samples)
    print(f"[plms S=10] oracle-vs-reference sampler max-rel={e:.2e}")
    assert e < 1e-4
    out["plms_S10_x0"] = samples.numpy()

    # ---- VAE decoder: tiny and the real KL-f8 decoder (bbox.yaml ddconfig), B=1 ----
    for tag, kw, hw in (("vaetiny", TINY_VAE_KW, 16), ("vaebbox", ref_loader.BBOX_VAE_KW, 32)):
        torch.manual_seed(0)
        dec = ref.Decoder(**kw).eval()
        pq = torch.nn.Conv2d(4, 4, 1)
        full_sd = {("decoder." + k): v for k, v in dec.state_dict().items()}
        full_sd.update({("post_quant_conv." + k): v for k, v in pq.state_dict().items()})
        sd = synth.synth_state_dict(full_sd, 0)
        dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
        pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")})
        z = synth.synth_inputs(1, hw, hw, 1, 8, 7)[0]
        with torch.no_grad():
            t0 = time.time(); y_ref = dec(pq(z / 0.18215)); dt = time.time() - t0   # autoencoder.py:330-333, ddpm.py:779
            y_or = O.decode_first_stage(sd, kw, z, 0.18215)
        e = relerr(y_or, y_ref)
        print(f"[{tag}] reference decode {dt:.2f}s out {tuple(y_ref.shape)} oracle-vs-reference max-rel={e:.2e}")
        assert e < 2e-5
        out[f"{tag}_img_sub"] = y_ref[:, :, ::8, ::8].numpy() if hw == 32 else y_ref.numpy()
        out[f"{tag}_stats"] = np.array([y_ref.mean().item(), y_ref.std().item(), y_ref.abs().max().item()])

    # ---- VAE encoder ('next' row 8(f)-2): Encoder + quant_conv, tiny and the real KL-f8 encoder, B=1 ----
    for tag, kw, hw in (("vaetiny", TINY_VAE_KW, 32), ("vaebbox", ref_loader.BBOX_VAE_KW, 64)):
        torch.manual_seed(0)
        enc = ref.Encoder(**kw).eval()
        qc = torch.nn.Conv2d(2 * kw["z_channels"], 2 * 4, 1)
        full_sd = {("encoder." + k): v for k, v in enc.state_dict().items()}
        full_sd.update({("quant_conv." + k): v for k, v in qc.state_dict().items()})
        sd = synth.synth_state_dict(full_sd, 0)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
        qc.load_state_dict({k[len("quant_conv."):]: v for k, v in sd.items() if k.startswith("quant_conv.")})
        x = torch.tanh(torch.randn(1, 3, hw, hw + 16, generator=torch.Generator().manual_seed(11)))     # rectangular, like 256x192
        with torch.no_grad():
            t0 = time.time(); m_ref = qc(enc(x)); dt = time.time() - t0        # autoencoder.py:324-327
            m_or = O.encode_first_stage_moments(sd, kw, x)
        e = relerr(m_or, m_ref)
        print(f"[{tag}-enc] reference encode {dt:.2f}s moments {tuple(m_ref.shape)} oracle-vs-reference max-rel={e:.2e}")
        assert e < 2e-5
        out[f"{tag}_enc_moments"] = m_ref.numpy()

    np.savez_compressed(os.path.join(OUT, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "hotpath_golden.npz"), os.path.getsize(os.path.join(OUT, "hotpath_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
