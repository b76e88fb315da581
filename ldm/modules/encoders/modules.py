"""Conditioning encoders named by configs/deepfashion/bbox.yaml:81-87.

The real CLIP ViT-L/14 towers (reference ldm/modules/encoders/modules.py:137-162,234-256) run once per request, before
the denoising loop; they are outside the B200 hot path and their weights cannot be fetched here.
These classes keep the dotted paths and call signatures so the config instantiates unchanged:
  * if `transformers` can build the text tower from a local cache it is used,
  * otherwise inputs that are already embeddings pass through, and raw inputs raise.
"""
import torch
from torch import nn


class AbstractEncoder(nn.Module):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


class FrozenCLIPEmbedder(AbstractEncoder):
    """CLIP text tower -> (B, 77, 768) last hidden state."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77):
        super().__init__()
        self.version, self.device, self.max_length = version, device, max_length
        self.tokenizer = None
        self.transformer = None
        self.embed_dim = 768

    def _lazy_load(self):
        if self.transformer is None:
            from transformers import CLIPTokenizer, CLIPTextModel
            self.tokenizer = CLIPTokenizer.from_pretrained(self.version, local_files_only=True)
            self.transformer = CLIPTextModel.from_pretrained(self.version, local_files_only=True).eval().to(self.device)
            for p in self.transformer.parameters():
                p.requires_grad = False

    def forward(self, text):
        if isinstance(text, torch.Tensor) and text.dim() == 3 and text.shape[-1] == self.embed_dim:
            return text  # pre-computed (B, 77, 768) embeddings
        self._lazy_load()
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return self.transformer(input_ids=enc["input_ids"].to(self.device)).last_hidden_state

    def encode(self, text):
        return self(text)


class FrozenClipImageEmbedder2(nn.Module):
    """CLIP image tower over the 9 style crops: (B, n, 3, 224, 224) -> (B, n, 768)."""

    def __init__(self, model="ViT-L/14", jit=False, device="cuda", antialias=False):
        super().__init__()
        self.embed_dim = 768
        self.model_name = model

    def forward(self, x):
        if isinstance(x, torch.Tensor) and x.dim() == 3 and x.shape[-1] == self.embed_dim:
            return x  # pre-computed style embeddings (what InferenceModel feeds through DummyModel)
        raise RuntimeError("FrozenClipImageEmbedder2: CLIP ViT-L/14 image weights are not available offline; "
                           "pass pre-computed (B, n, 768) style embeddings")

    def encode(self, x):
        return self(x)
