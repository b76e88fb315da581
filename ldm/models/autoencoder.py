"""AutoencoderKL first stage (reference ldm/models/autoencoder.py:285-423) on the B200 engines.

`decode` (post_quant_conv -> Decoder, autoencoder.py:330-333) is on the denoising hot path.  `encode` (Encoder -> quant_conv ->
DiagonalGaussianDistribution, autoencoder.py:324-328) is the first 'next' row of SURVEY.md section 8(f): log_images'
reconstruction, img2img / mask-inpaint starts (`stochastic_encode`) and `get_input` need it.
"""
import torch
from torch import nn

from ldm.modules.diffusionmodules.model import Decoder, Encoder
from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
from upgpt_b200.host import EngineHostMixin


class AutoencoderKL(nn.Module, EngineHostMixin):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None):
        super().__init__()
        ddconfig = dict(ddconfig)
        assert ddconfig["double_z"]
        self.image_key, self.embed_dim, self.ddconfig = image_key, embed_dim, ddconfig
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.monitor = monitor
        self._host_init()
        if ckpt_path is not None:
            import os
            if os.path.exists(ckpt_path):
                self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)
            else:
                print(f"AutoencoderKL: checkpoint {ckpt_path} not found; keeping random-init weights")

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        sd = {k: v for k, v in sd.items() if not any(k.startswith(ik) for ik in ignore_keys)}
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.mark_weights_changed()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._host_reset()
        return out

    def engine(self, B, H, W, precision=None):
        from upgpt_b200.vae_engine import VAEDecoderEngine
        from upgpt_b200.unet_engine import default_precision
        precision = precision or default_precision()
        return self._engine_get((B, H, W, precision), lambda: VAEDecoderEngine(self, B, H, W, precision=precision))

    def decode(self, z, in_scale=1.0):
        """z (B, embed_dim, h, w) fp32 -> image (B, out_ch, 8h, 8w) fp32.  in_scale multiplies z first (1/scale_factor)."""
        if not z.is_cuda:
            raise RuntimeError("upgpt_b200: AutoencoderKL.decode requires CUDA tensors (sm_100a engine; no CPU fallback)")
        B, _, H, W = z.shape
        return self.engine(B, H, W).decode(z, in_scale)

    def encoder_engine(self, B, H, W):
        from upgpt_b200.vae_engine import VAEEncoderEngine
        return self._engine_get(("enc", B, H, W), lambda: VAEEncoderEngine(self, B, H, W))

    def encode(self, x):
        """x (B, 3, H, W) fp32 NCHW image in [-1, 1] -> DiagonalGaussianDistribution over (B, embed_dim, H/8, W/8) (autoencoder.py:324-328)."""
        if not x.is_cuda:
            raise RuntimeError("upgpt_b200: AutoencoderKL.encode requires CUDA tensors (sm_100a engine; no CPU fallback)")
        B, _, H, W = x.shape
        return DiagonalGaussianDistribution(self.encoder_engine(B, H, W).encode(x))

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("autoencoder training/reconstruction is outside the B200 hot path")


class IdentityFirstStage(nn.Module):
    def __init__(self, *args, vq_interface=False, **kwargs):
        super().__init__()

    def encode(self, x, *args, **kwargs):
        return x

    def decode(self, x, *args, **kwargs):
        return x
