"""UNetModel of the latent-diffusion hot path (reference ldm/modules/diffusionmodules/openaimodel.py:413-742).

Constructor signature, attribute names and the parameter tree (=> state_dict keys such as
`input_blocks.1.0.in_layers.2.weight`) follow the reference so configs and checkpoints load unchanged.  `forward`
hands the whole network to the B200 engine: NCHW fp32 in/out at the boundary, NHWC fp16 tensor-core operands inside.
"""
import torch as th
from torch import nn

from ldm.modules.attention import SpatialTransformer
from ldm.modules.diffusionmodules.util import conv_nd, linear, normalization, zero_module
from upgpt_b200.host import EngineHostMixin


class TimestepBlock(nn.Module):
    """Marker: blocks that consume the timestep embedding."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None):
        raise RuntimeError("upgpt_b200: blocks execute inside UNetModel.forward on the CUDA engine")


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if not use_conv:
            raise NotImplementedError("avg-pool downsampling is not used by any UPGPT config")
        self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=padding)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv:
            raise NotImplementedError("scale-shift norm / resblock_updown / 3x3 skip are not used by any UPGPT config")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_checkpoint, self.use_scale_shift_norm, self.updown = use_checkpoint, False, False
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        self.skip_connection = (nn.Identity() if self.out_channels == channels
                                else conv_nd(dims, channels, self.out_channels, 1))


class UNetModel(nn.Module, EngineHostMixin):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        assert use_spatial_transformer and context_dim is not None, \
            "the B200 engine implements the SpatialTransformer U-Net (use_spatial_transformer=True with context_dim)"
        assert num_classes is None and n_embed is None and not resblock_updown and dims == 2
        if not isinstance(context_dim, int):
            context_dim = list(context_dim)
            assert len(context_dim) == 1
            context_dim = context_dim[0]
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        assert (num_heads != -1) or (num_head_channels != -1), "Either num_heads or num_head_channels has to be set"

        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions = list(attention_resolutions)
        self.dropout, self.channel_mult, self.conv_resample = dropout, list(channel_mult), conv_resample
        self.num_classes, self.use_checkpoint = num_classes, use_checkpoint
        self.dtype = th.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.predict_codebook_ids = False
        self.context_dim, self.transformer_depth = context_dim, transformer_depth

        ted = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, ted), nn.SiLU(), linear(ted, ted))

        def heads_for(ch):
            if num_head_channels == -1:
                return num_heads, ch // num_heads
            return ch // num_head_channels, num_head_channels

        def transformer(ch):
            nh, dh = heads_for(ch)
            if legacy:
                dh = ch // nh
            return SpatialTransformer(ch, nh, dh, depth=transformer_depth, context_dim=context_dim)

        def res(cin, cout):
            return ResBlock(cin, ted, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        skip_chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(transformer(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), transformer(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in self.attention_resolutions:
                    layers.append(transformer(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self._host_init()

    # -- weight-change tracking: ONE packed fp16 shadow copy per module (and weights tag), shared by its engines (SURVEY.md 8b;
    #    upgpt_b200/host.py): mark_weights_changed / use_weights_tag / the LRU-bounded engine cache come from EngineHostMixin --
    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.mark_weights_changed()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._host_reset()
        return out

    def precision_plan(self, H, W, ctx_len):
        """The calibrated "mixed" plan of the CURRENT weights at this latent size (upgpt_b200/precision.py): measured on the device at the
        first request after a weight change, cached per (weights tag, version, size). None: calibration disabled -> static profile."""
        from upgpt_b200 import precision as P
        if not P.enabled() or not next(self.parameters()).is_cuda:
            return None
        plans = self.__dict__.setdefault("_plans", {})
        key = (self._weights_tag, H, W, ctx_len)
        hit = plans.get(key)
        if hit is None or hit[0] != self._weights_version:
            plan, report = P.calibrate(self, H, W, ctx_len)
            hit = plans[key] = (self._weights_version, plan, report)
        return hit[1]

    def engine(self, B, H, W, ctx_len, precision=None):
        from upgpt_b200.unet_engine import UNetEngine, default_precision
        precision = precision or default_precision()
        plan = self.precision_plan(H, W, ctx_len) if precision == "mixed" else None
        pname = None if plan is None else plan["name"]
        return self._engine_get((B, H, W, ctx_len, precision, pname),
                                lambda: UNetEngine(self, B, H, W, ctx_len, precision=precision, plan=plan))

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """x (B, in_channels, H, W) fp32 NCHW, timesteps (B,) int64, context (B, L, context_dim) -> eps (B, out, H, W)."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("upgpt_b200: UNetModel.forward requires CUDA tensors (sm_100a engine; no CPU fallback)")
        B, _, H, W = x.shape
        eng = self.engine(B, H, W, context.shape[1])
        return eng.forward(x, timesteps, context)
