"""GPU parity of the CLIP ViT-L/14 conditioning towers through the C ABI: the CUDA-core kernels of
csrc/clip.cu per op, then both towers at a tiny shape against the committed outputs of the `transformers` implementations
(tests/golden/clip_golden.npz) and at the real ViT-L/14 shapes against the CPU oracle.  Tolerance: 1e-3 max-rel like the eps bound
(the reference computes the image tower in fp16 on CUDA; the engine's fp16x3 operands are ~100x closer to fp32 than that).
(File name sorts after the hot-path tests on purpose: the towers run once per request, outside the denoising loop.)"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import clip_oracle as CO
from oracle.make_clip_golden import TINY_TEXT, TINY_VIS
from upgpt_b200 import synth
from ldm.modules.encoders.modules import FrozenCLIPEmbedder, FrozenClipImageEmbedder2

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip_golden.npz"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from upgpt_b200 import _C
    _C.lib()
    return torch.device("cuda:0")


def relerr(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-9))


def _call(name, *args):
    from upgpt_b200 import _C, ops
    _C.check(getattr(_C.lib(), name)(*args, ops.stream()), name)
    torch.cuda.synchronize()


def _unplane(t, C_):
    t = t.float().cpu()
    return t[..., :C_] + t[..., C_:2 * C_]


def test_clip_small_kernels(dev):
    g = torch.Generator().manual_seed(0)
    # token + position embedding gather
    tok, pos = torch.randn(50, 64, generator=g), torch.randn(77, 64, generator=g)
    ids = torch.randint(0, 50, (3, 77), generator=g)
    out = torch.zeros(3 * 77, 64, device=dev)
    ids_d, tok_d, pos_d = ids.to(dev), tok.to(dev), pos.to(dev)          # keep the device copies alive across the launch
    _call("upgpt_embed_tokens", ids_d.data_ptr(), 3 * 77, 77, 50, tok_d.data_ptr(), pos_d.data_ptr(), 64, out.data_ptr())
    assert torch.equal(out.cpu().reshape(3, 77, 64), tok[ids] + pos[None])
    # im2col of the 14x14 stride-14 patch conv, K 588 -> 592, [hi | lo] planes
    img = torch.randn(2, 3, 56, 56, generator=g); w = torch.randn(32, 3, 14, 14, generator=g)
    pat = torch.full((2 * 16, 2 * 592), 7.0, device=dev, dtype=torch.half)
    img_d = img.to(dev)
    _call("upgpt_patchify", img_d.data_ptr(), 2, 3, 56, 14, 592, 1, pat.data_ptr(), 2 * 592)
    rows = _unplane(pat, 592)
    assert float(rows[:, 588:].abs().max()) == 0.0
    ref = F.conv2d(img, w, stride=14).reshape(2, 32, 16).permute(0, 2, 1).reshape(32, 32)
    assert relerr(rows[:, :588] @ w.reshape(32, -1).t(), ref) < 1e-5
    # class token + positional embedding
    patch, cls, pe = torch.randn(2, 16, 64, generator=g), torch.randn(64, generator=g), torch.randn(17, 64, generator=g)
    xo = torch.zeros(2, 17, 64, device=dev)
    patch_d, cls_d, pe_d = patch.to(dev), cls.to(dev), pe.to(dev)
    _call("upgpt_vit_assemble", patch_d.data_ptr(), cls_d.data_ptr(), pe_d.data_ptr(), 2, 17, 64, xo.data_ptr())
    assert torch.equal(xo.cpu(), torch.cat([cls.expand(2, 1, 64), patch], 1) + pe[None])
    # LayerNorm with fp32 output, strided rows (ln_post reads the class rows)
    for rows_, C_ in ((33, 768), (5, 1024), (7, 128)):
        x = torch.randn(rows_, 3, C_, generator=g) * 2 + 0.5; gm = 1 + 0.1 * torch.randn(C_, generator=g); bt = 0.1 * torch.randn(C_, generator=g)
        o = torch.zeros(rows_, C_, device=dev)
        x_d, gm_d, bt_d = x.to(dev), gm.to(dev), bt.to(dev)
        _call("upgpt_layernorm_f32", x_d.data_ptr(), 3 * C_, rows_, C_, gm_d.data_ptr(), bt_d.data_ptr(), 1e-5, o.data_ptr(), C_)
        assert relerr(o, F.layer_norm(x[:, 0], (C_,), gm, bt, 1e-5)) < 2e-6
    # QuickGELU -> operand planes
    x = torch.randn(37, 512, generator=g) * 3
    o16 = torch.zeros(37, 1024, device=dev, dtype=torch.half)
    x_d = x.to(dev)
    _call("upgpt_quick_gelu_cast", x_d.data_ptr(), 37, 512, 1, o16.data_ptr(), 1024)
    assert relerr(_unplane(o16, 512), CO.quick_gelu(x)) < 2e-6


@pytest.mark.parametrize("B,H,N,d,causal", [(2, 12, 77, 64, 1), (3, 2, 77, 64, 0), (1, 4, 20, 32, 1), (2, 3, 128, 48, 1)])
def test_attention_small_causal(dev, B, H, N, d, causal):
    """fp32 softmax attention of the text tower: q | k | v slices of one fp32 projection, causal mask, [hi | lo] output planes."""
    g = torch.Generator().manual_seed(N + d)
    Cc = H * d
    qkv = torch.randn(B, N, 3 * Cc, generator=g)
    q, k, v = [t.reshape(B, N, H, d).permute(0, 2, 1, 3) for t in qkv.split(Cc, -1)]
    s = torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5
    if causal:
        s = s + torch.full((N, N), float("-inf")).triu(1)
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, N, Cc)
    out = torch.zeros(B, N, 2 * Cc, device=dev, dtype=torch.half)
    qkv_d = qkv.to(dev)
    _call("upgpt_attention_small", qkv_d.data_ptr(), 3 * Cc, Cc, 2 * Cc, B, H, N, d, float(d) ** -0.5, causal, 1, out.data_ptr(), 2 * Cc)
    assert relerr(_unplane(out, Cc), ref) < 5e-6


def test_clip_text_tower_tiny_vs_transformers_golden(dev):
    host = FrozenCLIPEmbedder(arch=TINY_TEXT).materialize()
    sd = synth.synth_state_dict(host.transformer.state_dict(), 0)
    host.transformer.load_state_dict(sd); host.mark_weights_changed()
    host = host.to(dev)
    ids = torch.from_numpy(GOLD["text_tiny_ids"]).to(dev)
    y = host.encode(ids)
    assert tuple(y.shape) == (2, 77, 128)
    assert relerr(y, torch.from_numpy(GOLD["text_tiny_out"])) < 1e-3
    eng = host.engine(2, 77)
    eng.bufs["ids"].copy_(ids); eng.run(use_graph=False)
    assert torch.equal(eng.bufs["out"], y), "graph replay == eager program"


def test_clip_image_tower_tiny_vs_transformers_golden(dev):
    host = FrozenClipImageEmbedder2(arch=TINY_VIS).materialize()
    sd = synth.synth_state_dict(host.model.state_dict(), 1)
    host.model.load_state_dict(sd); host.mark_weights_changed()
    host = host.to(dev)
    img = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(1))
    y = host.encode(img[None].to(dev))                      # (b=1, n=3, c, h, w) like the style crops
    assert tuple(y.shape) == (1, 3, 96)
    assert relerr(y[0], torch.from_numpy(GOLD["vis_tiny_out"])) < 1e-3
    # opt-in single-plane fp16 operands (UPGPT_CLIP_PRECISION=fp16): the reference's own arithmetic class for this tower
    y16 = host.engine(3, precision="fp16").forward(img.to(dev))
    e16 = relerr(y16, torch.from_numpy(GOLD["vis_tiny_out"]))
    assert relerr(y[0], torch.from_numpy(GOLD["vis_tiny_out"])) < e16 < 5e-3


def test_clip_vit_l14_full_size_vs_oracle(dev):
    """The bbox.yaml towers at their real shapes: text (B=2, 77 tokens, 12 x 768) and image (2 x 3 style crops of 224x224, 24 x 1024)."""
    from upgpt_b200 import _C
    t = FrozenCLIPEmbedder().materialize()
    sd = synth.synth_state_dict(t.transformer.state_dict(), 2)
    t.transformer.load_state_dict(sd); t.mark_weights_changed()
    ids = torch.randint(0, 49408, (2, 77), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = CO.clip_text_forward(sd, 12, ids)
    l0 = _C.lib().upgpt_launch_count()
    y = t.to(dev).encode(ids.to(dev))
    assert _C.lib().upgpt_launch_count() - l0 >= 2 + 12 * 8
    assert tuple(y.shape) == (2, 77, 768) and relerr(y, ref) < 1e-3
    del t
    v = FrozenClipImageEmbedder2().materialize()
    sd = synth.synth_state_dict(v.model.state_dict(), 3)
    v.model.load_state_dict(sd); v.mark_weights_changed()
    crops = torch.randn(2, 3, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = CO.style_embed(sd, 16, crops)
    v = v.to(dev)
    y = v.encode(crops.to(dev))
    assert tuple(y.shape) == (2, 3, 768) and relerr(y, ref) < 1e-3
    y16 = v.engine(6, precision="fp16").forward(crops.reshape(6, 3, 224, 224).to(dev)).reshape(2, 3, 768)
    assert relerr(y16, ref) < 5e-3


def test_request_conditioning_through_latent_diffusion(dev):
    """The whole per-request conditioning step of the reference on the device (ddpm.py:553-565 get_learned_conditioning,
    ddpm.py:733-739 extra-cond loop): token ids -> text tower, style crops -> image tower, SMPL vector -> LinearProject,
    concatenated to the (B, 77 + 3 + 1, C) context the U-Net's cond-cache is built from; then one eps through that context."""
    import copy
    from conftest import tiny_ldm_config
    from ldm.util import instantiate_from_config
    cfg = copy.deepcopy(tiny_ldm_config())
    vis = dict(TINY_VIS, output_dim=128)
    cfg["params"]["cond_stage_config"]["params"] = {"arch": dict(TINY_TEXT)}
    cfg["params"]["extra_cond_stages"]["style_cond"] = {"target": "ldm.modules.encoders.modules.FrozenClipImageEmbedder2",
                                                        "cond_stage_key": "styles", "params": {"arch": vis}}
    model = instantiate_from_config(cfg)
    model.cond_stage_model.materialize(); model.extra_cond_models[0].materialize()
    sd = synth.synth_state_dict({k: v for k, v in model.state_dict().items()
                                 if k.startswith(("model.", "cond_stage_model.", "extra_cond_models."))}, 4)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(4)
    B = 2
    ids = torch.randint(0, TINY_TEXT["vocab"], (B, 77), generator=g)
    crops = torch.randn(B, 3, 3, 56, 56, generator=g)
    smpl = torch.randn(B, 1, 85, generator=g) * 0.5
    batch = {"txt": ids.to(dev), "styles": crops.to(dev), "smpl": smpl.to(dev)}
    c = model.get_learned_conditioning(batch["txt"])
    ctx = model.assemble_context(batch, c)
    assert tuple(ctx.shape) == (B, 77 + 3 + 1, 128)
    pre = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    with torch.no_grad():
        ref = torch.cat([CO.clip_text_forward(pre("cond_stage_model.transformer."), TINY_TEXT["heads"], ids),
                         CO.style_embed(pre("extra_cond_models.0.model."), vis["heads"], crops),
                         F.linear(smpl, sd["extra_cond_models.1.model.weight"], sd["extra_cond_models.1.model.bias"])], 1)
    assert relerr(ctx, ref) < 1e-3
    x, mask, _ = synth.synth_inputs(B, 16, 16, 81, 128, 4)
    eps = model.apply_model(x.to(dev), torch.full((B,), 500, dtype=torch.long, device=dev), {"c_crossattn": ctx, "c_concat": [mask.to(dev)]})
    assert tuple(eps.shape) == (B, 4, 16, 16) and torch.isfinite(eps).all()
