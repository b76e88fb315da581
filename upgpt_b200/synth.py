"""Deterministic synthetic weights and inputs (BASELINE.json: random-init U-Net, synthetic latents / conditioning).

Every tensor is drawn from a CPU torch.Generator seeded by crc32(name) ^ seed, so the values depend only on the
parameter name and shape -- not on module construction order -- and are identical in the build container (where the
golden vectors are made with the real reference) and on the GPU box.

A freshly constructed reference U-Net outputs exactly 0 (zero_module on 39 layers, openaimodel.py:229-231,685;
attention.py:244-248), so all parameters are re-drawn: weights ~ N(0, 1/fan_in), biases ~ 0.05 N(0,1),
norm scales ~ 1 + 0.1 N(0,1).
"""
import zlib

import torch


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(name, shape, seed=0):
    g = _gen(name, seed)
    shape = tuple(shape)
    is_norm = (".norm" in name or "in_layers.0." in name or "out_layers.0." in name or name.startswith("out.0.")
               or "norm_out" in name or "layer_norm" in name or ".ln_" in name or name.startswith("ln_"))
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
    if is_norm and name.endswith("weight"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm:
        return 0.1 * torch.randn(shape, generator=g)
    return 0.05 * torch.randn(shape, generator=g)


def synth_state_dict(shapes, seed=0):
    """shapes: {name: shape} (e.g. from module.state_dict()) -> {name: fp32 CPU tensor}."""
    return {k: synth_tensor(k, tuple(v.shape) if hasattr(v, "shape") else tuple(v), seed) for k, v in shapes.items()}


def synth_inputs(B, H, W, ctx_len=87, ctx_dim=768, seed=0, concat_channels=1):
    """Latent x_T, bbox person_mask (-1 outside / -0.99215686 inside, deepfashion_inshop.py:234-239), context."""
    g = _gen("inputs", seed)
    x = torch.randn(B, 4, H, W, generator=g)
    mask = torch.full((B, concat_channels, H, W), -1.0)
    for b in range(B):
        y0 = int(torch.randint(0, H // 2, (1,), generator=g)); x0 = int(torch.randint(0, W // 2, (1,), generator=g))
        mask[b, :, y0:y0 + H // 2, x0:x0 + W // 2] = -0.99215686
    ctx = torch.randn(B, ctx_len, ctx_dim, generator=g)
    return x, mask, ctx
