"""Conditioning encoders named by configs/deepfashion/bbox.yaml:81-87, with the reference's dotted paths and call signatures
(ldm/modules/encoders/modules.py:137-162 FrozenCLIPEmbedder, :234-256 FrozenClipImageEmbedder2).

The reference runs both CLIP ViT-L/14 towers once per request, before the denoising loop.  Here the classes hold the towers'
parameters in plain nn.Modules under the reference's state-dict names (`transformer.text_model...` as in transformers.CLIPTextModel,
`model.visual...` as in OpenAI clip's VisionTransformer) and NO arithmetic: forward hands over to the sm_100a engines in
upgpt_b200/clip_engine.py (tcgen05 GEMMs + attention kernels).  The pretrained weights cannot be fetched in this environment, so the
parameter trees are materialised lazily: `materialize()` (random init) or `load_state_dict` of a checkpoint that carries them.
Until then:
  * inputs that are already embeddings -- (B, 77, 768) text / (B, n, 768) style, what InferenceModel feeds -- pass through,
  * raw inputs raise (no CPU / library fallback).
"""
from collections import OrderedDict

import torch
from torch import nn

from upgpt_b200.host import EngineHostMixin


class AbstractEncoder(nn.Module):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


# ---------------------------------------------------------------------------------------------------- parameter containers
def _ns(**children):
    m = nn.Module()
    for k, v in children.items():
        setattr(m, k, v)
    return m


def clip_text_params(vocab=49408, width=768, layers=12, mlp=3072, positions=77):
    """Parameter tree with the key names of transformers.CLIPTextModel(...).state_dict() (openai/clip-vit-large-patch14)."""
    def layer():
        return _ns(self_attn=_ns(k_proj=nn.Linear(width, width), v_proj=nn.Linear(width, width), q_proj=nn.Linear(width, width),
                                 out_proj=nn.Linear(width, width)),
                   layer_norm1=nn.LayerNorm(width), mlp=_ns(fc1=nn.Linear(width, mlp), fc2=nn.Linear(mlp, width)),
                   layer_norm2=nn.LayerNorm(width))
    return _ns(text_model=_ns(embeddings=_ns(token_embedding=nn.Embedding(vocab, width), position_embedding=nn.Embedding(positions, width)),
                              encoder=_ns(layers=nn.ModuleList([layer() for _ in range(layers)])),
                              final_layer_norm=nn.LayerNorm(width)))


def clip_visual_params(width=1024, layers=24, heads=16, patch=14, resolution=224, output_dim=768):
    """Parameter tree with the key names of OpenAI clip's `model.visual` (clip/model.py VisionTransformer, ViT-L/14)."""
    def block():
        return _ns(attn=nn.MultiheadAttention(width, heads), ln_1=nn.LayerNorm(width),
                   mlp=nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, width * 4)), ("gelu", nn.Identity()), ("c_proj", nn.Linear(width * 4, width))])),
                   ln_2=nn.LayerNorm(width))
    scale = width ** -0.5
    v = _ns(conv1=nn.Conv2d(3, width, patch, patch, bias=False), ln_pre=nn.LayerNorm(width),
            transformer=_ns(resblocks=nn.Sequential(*[block() for _ in range(layers)])), ln_post=nn.LayerNorm(width))
    v.class_embedding = nn.Parameter(scale * torch.randn(width))
    v.positional_embedding = nn.Parameter(scale * torch.randn((resolution // patch) ** 2 + 1, width))
    v.proj = nn.Parameter(scale * torch.randn(width, output_dim))
    return v


class _EngineHost(nn.Module, EngineHostMixin):
    """Weight-change tracking shared by the two towers (the engines keep packed fp16 shadow copies)."""

    def __init__(self):
        super().__init__()
        self._host_init()      # LRU-bounded engine cache + one shared packed weight store (upgpt_b200/host.py)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.mark_weights_changed()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._host_reset()
        return out

    _PARAM_ROOT = None   # name of the lazily created parameter tree ("transformer" / "model")

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """A checkpoint that carries the tower's weights materialises the parameter tree before the keys are matched, so
        `LatentDiffusion.load_state_dict(ckpt, strict=False)` works as in the reference (generate_utils.py:33-48)."""
        root = prefix + self._PARAM_ROOT + "."
        if getattr(self, self._PARAM_ROOT) is None and any(k.startswith(root) for k in state_dict):
            self.materialize()
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class FrozenCLIPEmbedder(AbstractEncoder, _EngineHost):
    """CLIP text tower -> (B, 77, 768) last hidden state (modules.py:137-162)."""
    ARCH = dict(vocab=49408, width=768, layers=12, heads=12, mlp=3072, positions=77)
    _PARAM_ROOT = "transformer"

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, arch=None):
        _EngineHost.__init__(self)
        self.version, self.device, self.max_length = version, device, max_length
        self.arch = dict(self.ARCH, **(arch or {}))
        self.embed_dim = self.arch["width"]
        self.tokenizer = None
        self.transformer = None

    def materialize(self):
        """Creates the parameter tree (random init; load_state_dict afterwards for real weights)."""
        if self.transformer is None:
            a = self.arch
            self.transformer = clip_text_params(a["vocab"], a["width"], a["layers"], a["mlp"], a["positions"])
            for p in self.transformer.parameters():
                p.requires_grad = False
            self.mark_weights_changed()
        return self

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False
        return self.eval()

    def tokenize(self, text):
        if self.tokenizer is None:
            from transformers import CLIPTokenizer
            self.tokenizer = CLIPTokenizer.from_pretrained(self.version, local_files_only=True)   # raises offline without a cache
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return enc["input_ids"]

    def engine(self, B, L, precision=None):
        from upgpt_b200.clip_engine import ClipTextEngine, clip_precision
        precision = precision or clip_precision()
        return self._engine_get((B, L, precision), lambda: ClipTextEngine(self, B, L, precision=precision))

    @torch.no_grad()
    def forward(self, text):
        if isinstance(text, torch.Tensor) and text.is_floating_point() and text.dim() == 3 and text.shape[-1] == self.embed_dim:
            return text  # pre-computed (B, 77, 768) embeddings
        if self.transformer is None:
            raise RuntimeError("FrozenCLIPEmbedder: CLIP text weights are not loaded (call materialize() / load_state_dict), "
                               "or pass pre-computed (B, 77, 768) embeddings")
        ids = text if isinstance(text, torch.Tensor) else self.tokenize(text)
        dev = next(self.transformer.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("upgpt_b200: FrozenCLIPEmbedder runs on the sm_100a engine only (module is on %s; no CPU fallback)" % dev)
        ids = ids.to(dev, torch.long)
        return self.engine(ids.shape[0], ids.shape[1]).forward(ids)

    def encode(self, text):
        return self(text)


class FrozenClipImageEmbedder2(_EngineHost):
    """CLIP image tower over the style crops: (B, n, 3, 224, 224) -> (B, n, 768) (modules.py:234-256)."""
    ARCH = dict(width=1024, layers=24, heads=16, patch=14, resolution=224, output_dim=768)
    _PARAM_ROOT = "model"

    def __init__(self, model="ViT-L/14", jit=False, device="cuda", antialias=False, arch=None):
        super().__init__()
        self.arch = dict(self.ARCH, **(arch or {}))
        self.embed_dim = self.arch["output_dim"]
        self.model_name = model
        self.model = None

    def materialize(self):
        if self.model is None:
            a = self.arch
            self.model = _ns(visual=clip_visual_params(a["width"], a["layers"], a["heads"], a["patch"], a["resolution"], a["output_dim"]))
            for p in self.model.parameters():
                p.requires_grad = False
            self.mark_weights_changed()
        return self

    def engine(self, n, precision=None):
        from upgpt_b200.clip_engine import ClipVisionEngine, clip_precision
        precision = precision or clip_precision()
        return self._engine_get((n, precision), lambda: ClipVisionEngine(self, n, precision=precision))

    @torch.no_grad()
    def forward(self, x):
        if isinstance(x, torch.Tensor) and x.dim() == 3 and x.shape[-1] == self.embed_dim:
            return x  # pre-computed style embeddings (what InferenceModel feeds through DummyModel)
        if self.model is None:
            raise RuntimeError("FrozenClipImageEmbedder2: CLIP ViT-L/14 image weights are not loaded (call materialize() / "
                               "load_state_dict), or pass pre-computed (B, n, 768) style embeddings")
        if not x.is_cuda:
            raise RuntimeError("upgpt_b200: FrozenClipImageEmbedder2 runs on the sm_100a engine only (no CPU fallback)")
        b, n = x.shape[:2]
        ret = self.engine(b * n).forward(x.reshape((b * n,) + tuple(x.shape[2:])).float())
        return ret.reshape(b, n, -1)

    def encode(self, x):
        return self(x)
