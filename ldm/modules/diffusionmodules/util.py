"""Schedule tables and small helpers of the diffusion hot path (reference ldm/modules/diffusionmodules/util.py).

Host-side (numpy, float64 -> float32) pieces only: the per-step arithmetic runs in the CUDA extension.
"""
import math

import numpy as np
import torch
from torch import nn


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """beta_t table (util.py:21-43). 'linear' = squared linspace of the square roots, computed in float64."""
    if schedule == "linear":
        betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    elif schedule == "cosine":
        ts = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + cosine_s
        alphas = np.cos(ts / (1 + cosine_s) * np.pi / 2) ** 2
        alphas = alphas / alphas[0]
        betas = np.clip(1 - alphas[1:] / alphas[:-1], 0, 0.999)
    elif schedule == "sqrt_linear":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    elif schedule == "sqrt":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """Sub-sequence of DDPM steps visited by DDIM, shifted by +1 (util.py:46-60)."""
    if ddim_discr_method == "uniform":
        stride = num_ddpm_timesteps // num_ddim_timesteps
        ts = np.arange(0, num_ddpm_timesteps, stride)
    elif ddim_discr_method == "quad":
        ts = (np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ts + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigma_t, a_t, a_prev) per DDIM step (util.py:63-74); alphacums may be a tensor or array."""
    # Mixed torch-fp32 / numpy-float64 arithmetic exactly as the reference performs it: `alphas` stays a torch fp32
    # tensor, `alphas_prev` is a float64 numpy array of fp32 values, and the sigma expression promotes term by term.
    ac = alphacums.detach().float().cpu() if isinstance(alphacums, torch.Tensor) else torch.as_tensor(np.asarray(alphacums))
    alphas = ac[ddim_timesteps]
    alphas_prev = np.asarray([ac[0]] + ac[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule "
              f"for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """[cos | sin] sinusoidal embedding (util.py:151-171); on CUDA it is the extension's kernel."""
    if timesteps.is_cuda and not repeat_only:
        from upgpt_b200 import ops
        return ops.timestep_embedding(timesteps, dim, max_period)
    raise RuntimeError("upgpt_b200: timestep_embedding runs on the CUDA extension only (no CPU fallback)")


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    """32-group GroupNorm whose statistics are always fp32 (util.py:214-216). Parameter container here."""


def normalization(channels):
    return GroupNorm32(32, channels)


def conv_nd(dims, *args, **kwargs):
    if dims == 2:
        return nn.Conv2d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims} (the B200 path implements 2-D U-Nets)")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def checkpoint(func, inputs, params, flag):
    """Gradient checkpointing is a no-op for inference (util.py:102-148 under no_grad)."""
    return func(*inputs)
