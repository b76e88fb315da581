"""CPU checks for the CLIP conditioning towers: the oracle restatement against the golden outputs of the
`transformers` implementations (oracle/make_clip_golden.py), the parameter containers' state-dict names, and the engines' host logic
(program recording without a device; no CPU execution path)."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as CO
from oracle.make_clip_golden import TINY_TEXT, TINY_VIS
from upgpt_b200 import synth
from ldm.modules.encoders.modules import FrozenCLIPEmbedder, FrozenClipImageEmbedder2

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip_golden.npz"))


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_text_oracle_matches_transformers_golden():
    host = FrozenCLIPEmbedder(arch=TINY_TEXT).materialize()
    sd = synth.synth_state_dict(host.transformer.state_dict(), 0)
    with torch.no_grad():
        y = CO.clip_text_forward(sd, TINY_TEXT["heads"], torch.from_numpy(GOLD["text_tiny_ids"]))
    assert relerr(y, torch.from_numpy(GOLD["text_tiny_out"])) < 1e-5
    # causal: the first token's state does not depend on later tokens
    ids2 = torch.from_numpy(GOLD["text_tiny_ids"]).clone(); ids2[:, 5:] = 7
    with torch.no_grad():
        y2 = CO.clip_text_forward(sd, TINY_TEXT["heads"], ids2)
    assert torch.allclose(y[:, :5], y2[:, :5], atol=1e-6) and not torch.allclose(y[:, 5:], y2[:, 5:], atol=1e-3)


def test_vision_oracle_matches_transformers_golden():
    host = FrozenClipImageEmbedder2(arch=TINY_VIS).materialize()
    sd = synth.synth_state_dict(host.model.state_dict(), 1)
    img = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = CO.clip_vision_forward(sd, TINY_VIS["heads"], img)
        ys = CO.style_embed(sd, TINY_VIS["heads"], img[None])
    assert relerr(y, torch.from_numpy(GOLD["vis_tiny_out"])) < 1e-5
    assert tuple(ys.shape) == (1, 3, TINY_VIS["output_dim"]) and torch.equal(ys[0], y)


def test_vit_l14_oracle_probe():
    """ViT-L/14 shapes (the bbox.yaml towers): a strided probe of the transformers outputs recorded by make_clip_golden.py --full."""
    host = FrozenClipImageEmbedder2().materialize()
    sd = synth.synth_state_dict(host.model.state_dict(), 3)
    img = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y = CO.clip_vision_forward(sd, 16, img)
    assert tuple(y.shape) == (1, 768)
    assert relerr(y[0, ::37], torch.from_numpy(GOLD["vis_full_probe"])) < 2e-5


def test_state_dict_names_follow_the_reference_dependencies():
    t = FrozenCLIPEmbedder().materialize()
    keys = set(t.state_dict())
    assert len(keys) == 2 + 12 * 16 + 2
    for k in ("transformer.text_model.embeddings.token_embedding.weight", "transformer.text_model.embeddings.position_embedding.weight",
              "transformer.text_model.encoder.layers.11.self_attn.q_proj.bias", "transformer.text_model.encoder.layers.0.mlp.fc1.weight",
              "transformer.text_model.encoder.layers.3.layer_norm2.weight", "transformer.text_model.final_layer_norm.bias"):
        assert k in keys, k
    assert tuple(t.state_dict()["transformer.text_model.embeddings.token_embedding.weight"].shape) == (49408, 768)
    v = FrozenClipImageEmbedder2().materialize()
    sd = v.state_dict()
    assert len(sd) == 4 + 2 + 24 * 12 + 2
    shapes = {"model.visual.conv1.weight": (1024, 3, 14, 14), "model.visual.class_embedding": (1024,), "model.visual.positional_embedding": (257, 1024),
              "model.visual.proj": (1024, 768), "model.visual.transformer.resblocks.23.attn.in_proj_weight": (3072, 1024),
              "model.visual.transformer.resblocks.0.attn.out_proj.bias": (1024,), "model.visual.transformer.resblocks.5.mlp.c_fc.weight": (4096, 1024),
              "model.visual.transformer.resblocks.5.mlp.c_proj.weight": (1024, 4096), "model.visual.ln_post.weight": (1024,), "model.visual.ln_pre.bias": (1024,)}
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == s, k
    assert all(not p.requires_grad for p in v.parameters())


def test_encoders_pass_embeddings_through_and_refuse_cpu_work():
    t, v = FrozenCLIPEmbedder(), FrozenClipImageEmbedder2()
    e = torch.randn(2, 77, 768)
    assert t.encode(e) is e and v.encode(torch.randn(2, 9, 768)).shape == (2, 9, 768)
    with pytest.raises(RuntimeError, match="not loaded"):
        t(torch.zeros(2, 77, dtype=torch.long))
    with pytest.raises(RuntimeError, match="not loaded"):
        v(torch.zeros(1, 9, 3, 224, 224))
    t = FrozenCLIPEmbedder(arch=TINY_TEXT).materialize()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t(torch.zeros(2, 77, dtype=torch.long))
    v = FrozenClipImageEmbedder2(arch=TINY_VIS).materialize()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v(torch.zeros(1, 2, 3, 56, 56))


def test_engine_programs_record_without_a_device():
    """Host logic of the engines: weight packing ([hi | lo] planes, fused q|k|v, K 588 -> 592) and the launch sequence."""
    from upgpt_b200 import _C
    from upgpt_b200.clip_engine import ClipTextEngine, ClipVisionEngine
    t = FrozenCLIPEmbedder(arch=TINY_TEXT).materialize()
    e = ClipTextEngine(t, 2, 77, dry=True)
    assert e.launches == 1 + TINY_TEXT["layers"] * 8 + 1
    assert tuple(e.w["l0.qkv.weight"].shape) == (3 * 128, 2 * 128) and e.w["l0.qkv.weight"].dtype == torch.float16
    w = t.transformer.text_model.encoder.layers[0].self_attn.k_proj.weight
    hi, lo = e.w["l0.qkv.weight"][128:256, :128].float(), e.w["l0.qkv.weight"][128:256, 128:].float()
    assert float((hi + lo - w).abs().max()) < 1e-6 * float(w.abs().max()) + 1e-9
    with pytest.raises(_C.UpgptError):
        e.forward(torch.zeros(2, 77, dtype=torch.long))
    v = FrozenClipImageEmbedder2(arch=TINY_VIS).materialize()
    ev = ClipVisionEngine(v, 3, dry=True)
    assert ev.T == 17 and ev.Kp == 592 and ev.launches == 4 + TINY_VIS["layers"] * 8 + 2
    assert tuple(ev.w["patch.weight"].shape) == (128, 2 * 592)
    assert float(ev.w["patch.weight"][:, 588:592].abs().max()) == 0.0
    with pytest.raises(_C.UpgptError):
        ClipVisionEngine(v, 3)        # not on a CUDA device and not a dry recording
    # single-plane fp16 mode (the reference's own arithmetic class for the image tower): same program, K-wide operands
    e16, ev16 = ClipTextEngine(t, 2, 77, dry=True, precision="fp16"), ClipVisionEngine(v, 3, dry=True, precision="fp16")
    assert e16.launches == e.launches and ev16.launches == ev.launches
    assert tuple(ev16.w["patch.weight"].shape) == (128, 592) and tuple(e16.w["l0.qkv.weight"].shape) == (3 * 128, 128)
    L = _C.lib()
    for eng, x3 in ((e, True), (ev, True), (e16, False), (ev16, False)):
        gemms = [a[0]._obj for f, a in eng.prog.kernel_calls() if f is L.upgpt_gemm]
        assert gemms and all(bool(g.flags & _C.GEMM_F_X3) == x3 for g in gemms)
        assert all(g.a and g.w and (g.out32 or g.out16) for g in gemms), "every GEMM has its operands and an output"
        lns = [a for f, a in eng.prog.kernel_calls() if f in (L.upgpt_layernorm, L.upgpt_layernorm_split3)]
        assert lns and all(a[0] and a[4] and a[5] and a[7] for a in lns), "every LayerNorm launch carries x, gamma, beta, out"


def test_checkpoint_with_tower_weights_materialises_the_parameter_trees():
    """load_state_dict of a parent module whose checkpoint carries `cond_stage_model.transformer...` / `...model.visual...` keys
    creates the lazily built trees first (the reference builds them in __init__ from the network); without such keys nothing is built."""
    class Parent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.cond_stage_model = FrozenCLIPEmbedder(arch=TINY_TEXT)
            self.style = FrozenClipImageEmbedder2(arch=TINY_VIS)
    sd_t = synth.synth_state_dict(FrozenCLIPEmbedder(arch=TINY_TEXT).materialize().state_dict(), 7)
    sd_v = synth.synth_state_dict(FrozenClipImageEmbedder2(arch=TINY_VIS).materialize().state_dict(), 8)
    full = {**{"cond_stage_model." + k: v for k, v in sd_t.items()}, **{"style." + k: v for k, v in sd_v.items()}}
    p = Parent()
    assert len(p.state_dict()) == 0
    v0 = p.cond_stage_model._weights_version
    res = p.load_state_dict(full, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    got = p.state_dict()
    assert set(got) == set(full) and all(torch.equal(got[k], full[k]) for k in full)
    assert p.cond_stage_model._weights_version > v0
    q = Parent()
    q.load_state_dict({}, strict=False)
    assert q.cond_stage_model.transformer is None and q.style.model is None
