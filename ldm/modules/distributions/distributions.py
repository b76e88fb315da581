"""Posterior of the KL autoencoder (reference ldm/modules/distributions/distributions.py:24-60).

`parameters` are the NCHW moments {mean | logvar} produced by AutoencoderKL.encode; sampling, the mode and the scale_factor
multiply of `get_first_stage_encoding` run in one kernel (upgpt_gaussian_sample)."""
import ctypes as C

import torch


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.deterministic = deterministic

    @property
    def mean(self):
        return torch.chunk(self.parameters, 2, dim=1)[0]

    @property
    def logvar(self):
        return torch.clamp(torch.chunk(self.parameters, 2, dim=1)[1], -30.0, 20.0)

    def _draw(self, noise, scale):
        from upgpt_b200 import _C, ops
        p = self.parameters
        if not p.is_cuda:
            raise RuntimeError("upgpt_b200: the VAE posterior lives on the CUDA device (no CPU fallback)")
        p = p.contiguous()
        B, C2, H, W = p.shape
        out = torch.empty(B, C2 // 2, H, W, device=p.device, dtype=torch.float32)
        _C.check(_C.lib().upgpt_gaussian_sample(p.data_ptr(), 0 if noise is None else noise.data_ptr(), float(scale), out.data_ptr(),
                                                B, C2 // 2, H * W, ops.stream()), "upgpt_gaussian_sample")
        return out

    def sample(self, noise=None, scale=1.0):
        """mean + std * N(0,1) (distributions.py:35-37); `noise` may be given for reproducibility, `scale` folds scale_factor in."""
        if self.deterministic:
            return self._draw(None, scale)
        if noise is None:
            noise = torch.randn(self.mean.shape, device=self.parameters.device)
        return self._draw(noise.contiguous().float(), scale)

    def mode(self, scale=1.0):
        return self._draw(None, scale)
