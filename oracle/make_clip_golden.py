"""Pins oracle/clip_oracle.py against the `transformers` CLIP implementations (build container only) and writes
tests/golden/clip_golden.npz = outputs of transformers.CLIPTextModel / CLIPVisionModelWithProjection on name-seeded synthetic weights.

    python oracle/make_clip_golden.py

The reference calls transformers.CLIPTextModel directly (ldm/modules/encoders/modules.py:140-159); its image tower is OpenAI clip's
VisionTransformer (absent here), whose HF port CLIPVisionModelWithProjection is the executable stand-in (key map in clip_oracle)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import clip_oracle as CO                       # noqa: E402
from upgpt_b200 import synth                               # noqa: E402
from ldm.modules.encoders.modules import FrozenCLIPEmbedder, FrozenClipImageEmbedder2   # noqa: E402  (parameter containers only)

TINY_TEXT = dict(vocab=1000, width=128, layers=2, heads=2, mlp=512, positions=77)
TINY_VIS = dict(width=128, layers=2, heads=2, patch=14, resolution=56, output_dim=96)


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())


def text_case(arch, B, seed):
    from transformers import CLIPTextConfig, CLIPTextModel
    host = FrozenCLIPEmbedder(arch=arch).materialize()
    sd = synth.synth_state_dict(host.transformer.state_dict(), seed)
    cfg = CLIPTextConfig(vocab_size=arch["vocab"], hidden_size=arch["width"], intermediate_size=arch["mlp"], num_hidden_layers=arch["layers"],
                         num_attention_heads=arch["heads"], max_position_embeddings=arch["positions"], hidden_act="quick_gelu")
    hf = CLIPTextModel(cfg).eval()
    missing = hf.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    ids = torch.randint(0, arch["vocab"], (B, 77), generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        ref = hf(input_ids=ids).last_hidden_state
        got = CO.clip_text_forward(sd, arch["heads"], ids)
    return ids, ref, relerr(got, ref)


def vision_case(arch, n, seed):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    host = FrozenClipImageEmbedder2(arch=arch).materialize()
    sd = synth.synth_state_dict(host.model.state_dict(), seed)
    cfg = CLIPVisionConfig(hidden_size=arch["width"], intermediate_size=4 * arch["width"], num_hidden_layers=arch["layers"],
                           num_attention_heads=arch["heads"], image_size=arch["resolution"], patch_size=arch["patch"],
                           projection_dim=arch["output_dim"], hidden_act="quick_gelu")
    hf = CLIPVisionModelWithProjection(cfg).eval()
    missing = hf.load_state_dict(CO.openai_to_hf_vision(sd), strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    img = torch.randn(n, 3, arch["resolution"], arch["resolution"], generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        ref = hf(pixel_values=img).image_embeds
        got = CO.clip_vision_forward(sd, arch["heads"], img)
    return img, ref, relerr(got, ref)


def main():
    out = {}
    ids, ref, e = text_case(TINY_TEXT, 2, 0)
    print(f"text  tiny : oracle vs transformers.CLIPTextModel max-rel {e:.2e}"); assert e < 1e-5
    out["text_tiny_ids"], out["text_tiny_out"] = ids.numpy(), ref.numpy()
    img, ref, e = vision_case(TINY_VIS, 3, 1)
    print(f"image tiny : oracle vs transformers.CLIPVisionModelWithProjection max-rel {e:.2e}"); assert e < 1e-5
    out["vis_tiny_out"] = ref.numpy()
    if "--full" in sys.argv:   # ViT-L/14 shapes (not stored: 0.4 G parameters; the GPU tests recompute the oracle at these shapes)
        _, ref, e = text_case(dict(FrozenCLIPEmbedder.ARCH), 1, 2)
        print(f"text  L/14 : oracle vs transformers max-rel {e:.2e}"); assert e < 2e-5
        out["text_full_probe"] = ref[0, ::19, ::97].numpy()
        _, ref, e = vision_case(dict(FrozenClipImageEmbedder2.ARCH), 1, 3)
        print(f"image L/14 : oracle vs transformers max-rel {e:.2e}"); assert e < 2e-5
        out["vis_full_probe"] = ref[0, ::37].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clip_golden.npz"), **out)
    print("wrote tests/golden/clip_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
