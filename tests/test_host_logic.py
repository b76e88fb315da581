"""CPU-only tests of the host side: config protocol, module tree / state_dict layout, schedules, weight packing,
conditioning routing, the C-ABI surface, and the no-fallback rule."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, tiny_ldm_config
from oracle import ldm_oracle as O
from oracle.ref_loader import BBOX_UNET_KW


def test_bbox_yaml_loads_unchanged_and_instantiates():
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    assert cfg.model.target == "ldm.models.diffusion.ddpm.LatentDiffusion"
    assert cfg.model.params.unet_config.params.model_channels == 224
    model = instantiate_from_config(cfg.model)
    assert type(model).__name__ == "LatentDiffusion"
    assert model.model.conditioning_key == "hybrid" and model.concat_key == "person_mask"
    assert model.extra_cond_keys == ["styles", "smpl"] and abs(model.scale_factor - 0.18215) < 1e-12
    unet = model.model.diffusion_model
    assert sum(p.numel() for p in unet.parameters()) == 425290884      # "DiffusionWrapper has 425.29 M params" (inference.ipynb)
    sd = model.state_dict()
    for k in ("model.diffusion_model.input_blocks.1.0.in_layers.2.weight", "model.diffusion_model.out.2.bias",
              "model.diffusion_model.middle_block.1.transformer_blocks.0.attn2.to_k.weight",
              "model.diffusion_model.output_blocks.11.1.proj_out.weight", "first_stage_model.decoder.up.3.block.0.conv1.weight",
              "first_stage_model.post_quant_conv.weight", "extra_cond_models.1.model.weight", "betas", "alphas_cumprod",
              "model_ema.decay"):
        assert k in sd, k
    assert sd["model.diffusion_model.middle_block.1.transformer_blocks.0.attn2.to_k.weight"].shape == (896, 768)
    # zero_module initialisation of the reference (openaimodel.py:229-231,685; attention.py:244-248)
    assert float(sd["model.diffusion_model.out.2.weight"].abs().max()) == 0.0
    assert float(sd["model.diffusion_model.input_blocks.1.1.proj_out.weight"].abs().max()) == 0.0


def test_unet_structure_matches_survey_table():
    from ldm.modules.diffusionmodules.openaimodel import UNetModel, ResBlock
    from ldm.modules.attention import SpatialTransformer
    m = UNetModel(**BBOX_UNET_KW)
    res = [x for x in m.modules() if isinstance(x, ResBlock)]
    st = [x for x in m.modules() if isinstance(x, SpatialTransformer)]
    assert len(res) == 22 and len(st) == 16
    assert sorted({s.d_head for s in st}) == [28, 56, 112]
    assert [b[0].channels for b in m.output_blocks] == [1792, 1792, 1792, 1792, 1792, 1344, 1344, 896, 672, 672, 448, 448]


def test_ddpm_schedule_buffers_match_oracle():
    from ldm.util import instantiate_from_config
    model = instantiate_from_config(tiny_ldm_config())
    sched = O.register_schedule(1000, 0.00085, 0.012)
    for k, v in sched.items():
        assert torch.equal(getattr(model, k), v), k
    assert model.num_timesteps == 1000


@pytest.mark.parametrize("S,eta", [(50, 0.0), (10, 1.0)])
def test_ddim_schedule_matches_reference_golden(golden, S, eta):
    from ldm.util import instantiate_from_config
    from ldm.models.diffusion.ddim import DDIMSampler
    model = instantiate_from_config(tiny_ldm_config())
    s = DDIMSampler(model)
    s.make_schedule(S, ddim_eta=eta, verbose=False)
    tag = f"ddim_S{S}_eta{int(eta)}"
    np.testing.assert_array_equal(s.ddim_timesteps, golden[tag + "_timesteps"])
    np.testing.assert_array_equal(s.ddim_alphas.astype(np.float64), golden[tag + "_alphas"])
    np.testing.assert_array_equal(s.ddim_alphas_prev, golden[tag + "_alphas_prev"])
    np.testing.assert_allclose(s.ddim_sigmas, golden[tag + "_sigmas"], rtol=1e-12)
    rows = s._coef_rows(False, 1.0)
    assert rows.shape == (S, 5) and rows.dtype == torch.float32
    assert float(rows[0, 1]) == float(np.float32(model.alphas_cumprod[0]))     # a_prev of the last step = alphas_cumprod[0]


def test_weight_packing_helpers():
    from upgpt_b200.unet_engine import pad_heads_rows, pad_heads_cols, pack_geglu, geglu_half, split3_w
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8 * 28, 224, generator=g)
    wp = pad_heads_rows(w, 8, 28, 64)
    assert wp.shape == (512, 224)
    assert torch.equal(wp.reshape(8, 64, 224)[:, :28], w.reshape(8, 28, 224)) and float(wp.reshape(8, 64, 224)[:, 28:].abs().max()) == 0
    wo = torch.randn(224, 8 * 28, generator=g)
    wop = pad_heads_cols(wo, 8, 28, 64)
    x = torch.randn(5, 8 * 28, generator=g)
    xp = torch.zeros(5, 8, 64); xp[:, :, :28] = x.reshape(5, 8, 28)
    torch.testing.assert_close(xp.reshape(5, 512) @ wop.t(), x @ wo.t())
    inner = 896
    half = geglu_half(inner)
    assert inner % half == 0 and half % 16 == 0
    w1, b1 = torch.randn(2 * inner, 224, generator=g), torch.randn(2 * inner, generator=g)
    wpk, bpk = pack_geglu(w1, b1, inner, half)
    y = x[:, :224] @ w1.t() + b1
    ypk = (x[:, :224] @ wpk.t() + bpk).reshape(5, inner // half, 2, half)
    torch.testing.assert_close(ypk[:, :, 0].reshape(5, inner), y[:, :inner])
    torch.testing.assert_close(ypk[:, :, 1].reshape(5, inner), y[:, inner:])
    w3 = split3_w(w1)
    k = 224
    rec = w3[:, :k].float() + w3[:, k:].float()            # planes [Wh | Wl]
    assert w3.shape[1] == 2 * k and float((rec - w1).abs().max()) < 1e-6 and torch.equal(w3[:, :k], w1.half())


def test_conditioning_routing():
    from upgpt_b200.sampler_engine import FusedSampler
    c = torch.zeros(2, 87, 8); m = torch.zeros(2, 1, 4, 4)
    cc, ct = FusedSampler.split_cond({"c_crossattn": c, "c_concat": [m]}, "hybrid")
    assert cc is c and ct is m
    cc, ct = FusedSampler.split_cond({"c_crossattn": [c, c], "c_concat": [m]}, "hybrid")
    assert cc.shape == (2, 174, 8)
    assert FusedSampler.split_cond({"c_crossattn": c}, "hybrid") is None
    cc, ct = FusedSampler.split_cond(c, "crossattn")
    assert cc is c and ct is None


def test_extra_cond_assembly_shapes():
    """c = cat(text(B,77,D), styles(B,9,D), Linear(smpl(B,1,85))) (ddpm.py:733-739) -- DummyModel/identity parts on CPU."""
    from ldm.modules.poses.poses import DummyModel
    from ldm.modules.encoders.modules import FrozenCLIPEmbedder, FrozenClipImageEmbedder2
    t = torch.randn(2, 77, 768)
    assert FrozenCLIPEmbedder().encode(t) is t
    s = torch.randn(2, 9, 768)
    assert FrozenClipImageEmbedder2()(s) is s and DummyModel()(s) is s
    with pytest.raises(RuntimeError):
        FrozenClipImageEmbedder2()(torch.randn(2, 9, 3, 224, 224))


def test_no_cpu_fallback():
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from oracle.make_golden import TINY_UNET_KW
    m = UNetModel(**TINY_UNET_KW)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 5, 16, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, 87, 128))
    from ldm.util import instantiate_from_config
    model = instantiate_from_config(tiny_ldm_config())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.decode_first_stage(torch.zeros(1, 4, 16, 16))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads (no GPU needed) and exports exactly the functions include/upgpt_b200.h declares."""
    from upgpt_b200 import _C
    hdr = open(os.path.join(ROOT, "include", "upgpt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(upgpt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _C.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_C.EXPORTS) == declared
    assert lib.upgpt_abi_version() == 1
    assert lib.upgpt_launch_count() == 0


def test_lr_scheduler_stub_matches_reference_formula():
    from ldm.lr_scheduler import LambdaLinearScheduler
    s = LambdaLinearScheduler(warm_up_steps=[100], f_min=[1.0], f_max=[1.0], f_start=[1e-6], cycle_lengths=[10 ** 13])
    assert abs(s(0) - 1e-6) < 1e-12 and abs(s(50) - (1e-6 + (1 - 1e-6) * 0.5)) < 1e-9 and s(1000) == 1.0


def test_upsample_fold_weights_reproduce_interpolate_conv():
    """up2_conv_w (UPGPT_GEMM_CONV3X3_UP2): four parity-wise 2x2 kernels over the low-resolution input == conv3x3(nearest x2 upsample),
    openaimodel.py:116-118 / model.py:49-52, including the zero padding at the borders of the UPSAMPLED image."""
    import torch.nn.functional as F
    from upgpt_b200.unet_engine import up2_conv_w
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 5, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 6, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wp = up2_conv_w(w)                                   # [4][Cout][4][Cin]
    assert wp.shape == (4, 4, 4, 6)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for a in (0, 1):
        for b in (0, 1):
            acc = 0
            for u in (0, 1):
                for v in (0, 1):
                    # source pixel (y + a - 1 + u, x + b - 1 + v); +1 for the zero pad of xp
                    src = xp[:, :, a + u:a + u + 5, b + v:b + v + 7]
                    acc = acc + torch.einsum("bchw,oc->bohw", src, wp[a * 2 + b, :, u * 2 + v, :])
            out[:, :, a::2, b::2] = acc
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("precision", ["fp16x3", "mixed", "fp16"])
def test_unet_program_operand_formats_are_consistent(precision, monkeypatch):
    """Host logic of the per-layer precision plan (recorded without a device): every tensor-core GEMM must find its A operand in the
    plane format ([hi | lo] vs single fp16) its producer wrote -- GroupNorm/prep, LayerNorm, attention output, or a GEMM's fp16 copy."""
    import torch
    from upgpt_b200 import _C
    from upgpt_b200 import unet_engine
    from upgpt_b200.unet_engine import UNetEngine
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from oracle.ref_loader import BBOX_UNET_KW
    bbox_arch = (BBOX_UNET_KW["model_channels"], tuple(BBOX_UNET_KW["channel_mult"]), BBOX_UNET_KW["num_res_blocks"],
                 tuple(BBOX_UNET_KW["attention_resolutions"]))
    assert unet_engine.MIXED_PROFILES[bbox_arch] == (64, 16)
    kw = dict(BBOX_UNET_KW, model_channels=64, num_heads=4, context_dim=64)     # same 4-level structure, narrow channels
    monkeypatch.setitem(unet_engine.MIXED_PROFILES, (64,) + bbox_arch[1:], (64, 16))
    unet = UNetModel(**kw).eval()
    eng = UNetEngine(unet, 1, 32, 32, 87, precision=precision, dry=True)
    L = _C.lib()
    fmt, n_x3, n_plain = {}, 0, 0
    for fn, args in eng.prog.kernel_calls():
        if fn in (L.upgpt_prep_operand, L.upgpt_groupnorm_prep):
            a = args[0]._obj
            fmt[a.out] = bool(a.split3)
            if a.raw:
                fmt[a.raw] = bool(a.split3) if a.raw_planes == 0 else a.raw_planes == 2
        elif fn in (L.upgpt_layernorm, L.upgpt_layernorm_split3):
            assert args[0] and args[4] and args[5] and args[7]
            fmt[args[7]] = fn is L.upgpt_layernorm_split3
        elif fn is L.upgpt_attention:
            a = args[0]._obj
            assert a.q and a.k and a.vt and a.out
            fmt[a.out] = bool(a.split3_out)
        elif fn is L.upgpt_gemm:
            a = args[0]._obj
            assert a.a and a.w and (a.out32 or a.out16), "every GEMM has its operands and an output"
            assert a.bias or (a.out16 and not a.out32), "only the attention q / k / v projections (fp16-only outputs) have no bias"
            x3 = bool(a.flags & _C.GEMM_F_X3)
            n_x3 += x3; n_plain += not x3
            assert a.a in fmt, "GEMM operand without a recorded producer"
            assert fmt[a.a] == x3, "operand planes do not match the GEMM's precision flag"
            if a.out16:
                fmt[a.out16] = bool(a.flags & _C.GEMM_F_SPLIT3OUT)
    assert n_x3 + n_plain == 193
    if precision == "fp16x3":
        assert n_plain == 0
    elif precision == "fp16":
        assert n_x3 == 0
    else:   # mixed: the 4x4 level entirely, the 8x8 level except skip / proj_in / proj_out and the convs sharing operands with a skip
        assert n_plain == 64 and eng.mixed
        assert eng.use_x3("conv", 1024) and not eng.use_x3("resid1x1", 16) and eng.use_x3("resid1x1", 64) and not eng.use_x3("tf", 64)
    if precision == "mixed":
        # calibrated plan deep+tf1 (upgpt_b200/precision.py): additionally every attention / feed-forward GEMM and the 8x8-level ResBlock
        # convs whose input also feeds a skip 1x1 GEMM on single planes; that GEMM keeps [hi | lo] planes of its own (raw_planes = 2)
        from upgpt_b200 import precision as P
        cands = dict(P.candidates(32, 32, 4))
        assert list(cands)[0] == "deep+tf1C" and list(cands)[-1] == "fp16x3" and set(cands) <= set(P.DESCRIPTIONS)
        eng2 = UNetEngine(unet, 1, 32, 32, 87, precision="mixed", dry=True, plan=dict(cands["deep+tf1C"], name="deep+tf1C"))
        fmt2, plain2, mixed_raw = {}, 0, 0
        for fn, args in eng2.prog.kernel_calls():
            if fn in (L.upgpt_prep_operand, L.upgpt_groupnorm_prep):
                a = args[0]._obj
                fmt2[a.out] = bool(a.split3)
                if a.raw:
                    fmt2[a.raw] = bool(a.split3) if a.raw_planes == 0 else a.raw_planes == 2
                    mixed_raw += fmt2[a.raw] != fmt2[a.out]
            elif fn is L.upgpt_attention:
                fmt2[args[0]._obj.out] = bool(args[0]._obj.split3_out)
            elif fn is L.upgpt_gemm:
                a = args[0]._obj
                x3 = bool(a.flags & _C.GEMM_F_X3)
                plain2 += not x3
                if a.a in fmt2:
                    assert fmt2[a.a] == x3, "operand planes do not match the GEMM's precision flag (calibrated plan)"
                if a.out16:
                    fmt2[a.out16] = bool(a.flags & _C.GEMM_F_SPLIT3OUT)
        # raw copy in [hi | lo] planes beside a single-plane operand: the skip-connected ResBlocks of the 8x8 level (4) + every decoder block above it (6)
        assert mixed_raw == 10 and plain2 > n_plain + 4, (mixed_raw, plain2, n_plain)
    with pytest.raises(_C.UpgptError):
        eng.run()
    # an architecture without a probed profile keeps fp16x3 everywhere in "mixed" (the probe shows its deep levels are not cheap in error)
    from oracle.make_golden import TINY_UNET_KW
    tiny = UNetEngine(UNetModel(**TINY_UNET_KW).eval(), 2, 16, 16, 87, precision="mixed", dry=True)
    assert not tiny.mixed and all(bool(a[0]._obj.flags & _C.GEMM_F_X3) for f, a in tiny.prog.kernel_calls() if f is L.upgpt_gemm)
    monkeypatch.setenv("UPGPT_MIXED_HW", "64,16")          # tuning override: thresholds for an architecture without a profile
    tiny = UNetEngine(UNetModel(**TINY_UNET_KW).eval(), 2, 16, 16, 87, precision="mixed", dry=True)
    assert tiny.mixed and not all(bool(a[0]._obj.flags & _C.GEMM_F_X3) for f, a in tiny.prog.kernel_calls() if f is L.upgpt_gemm)


def test_bbox_yaml_instantiates_the_way_inference_model_rewrites_it():
    """The reference's inference facade (ldm/data/generate_utils.py:131-146) rewrites the config before instantiating it: style_cond's
    target becomes ldm.modules.poses.poses.DummyModel while its params stay {'device': ...}, and cond_stage_config gets params
    {'device': ...}. Both constructors must accept that (DummyModel(*args, **kwargs) as in poses.py:11-13)."""
    import torch
    from ldm.util import load_config, instantiate_from_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    p = cfg["model"]["params"]
    p["extra_cond_stages"]["style_cond"]["params"] = {"device": "cpu"}
    style_enc = instantiate_from_config(p["extra_cond_stages"]["style_cond"])        # clip_image_encoder of the facade
    assert type(style_enc).__name__ == "FrozenClipImageEmbedder2"
    p["extra_cond_stages"]["style_cond"]["target"] = "ldm.modules.poses.poses.DummyModel"
    p["first_stage_config"]["params"]["ckpt_path"] = None
    p["cond_stage_config"]["params"] = {"device": "cpu"}
    p["use_ema"] = False
    model = instantiate_from_config(cfg["model"])
    assert type(model.extra_cond_models[0]).__name__ == "DummyModel"
    s = torch.randn(2, 9, 768)
    assert torch.equal(model.extra_cond_models[0](s), s)


def test_engine_cache_is_lru_bounded_and_shares_packed_weights(monkeypatch):
    """One packed weight copy per module (and weights tag) whatever the number of engines; the engine cache is LRU-bounded."""
    import torch
    from upgpt_b200.host import EngineHostMixin, WeightStore

    class Eng:
        def __init__(self, host):
            self.weights_version = -1
            self.packs = 0

        def pack_weights(self, host):
            self.packs += 1
            self.weights_version = host._weights_version

    class Host(EngineHostMixin):
        def __init__(self):
            self._host_init()

    monkeypatch.setenv("UPGPT_MAX_ENGINES", "3")
    h = Host()
    engs = [h._engine_get((b,), lambda: Eng(h)) for b in range(5)]
    assert len(h._engines) == 3 and list(h._engines)[0] == (2, "raw", 0)      # (key, weights tag, lane)
    assert h._engine_get((4,), lambda: None) is engs[4] and engs[4].packs == 1
    h.mark_weights_changed()
    assert h._engine_get((4,), lambda: None).packs == 2          # re-pack on a version change only
    h.use_weights_tag("ema")
    e_ema = h._engine_get((4,), lambda: Eng(h))
    assert e_ema is not engs[4] and list(h._engines)[-1] == (4, "ema", 0)
    h.use_weights_tag("raw")
    assert h._engine_get((4,), lambda: None) is engs[4] and engs[4].packs == 2, "leaving the EMA scope costs no re-pack"
    # a second batch in flight (upgpt_b200/lanes.py) gets its own engine -- buffers, programs, graphs -- for the same key
    from upgpt_b200 import lanes
    assert lanes.current() == 0 and lanes.branch_aux(0) == 0 and lanes.branch_aux(1) == 3
    lanes._current = 1
    try:
        e_l1 = h._engine_get((4,), lambda: Eng(h))
    finally:
        lanes._current = 0
    assert e_l1 is not engs[4] and list(h._engines)[-1] == (4, "raw", 1) and h._engine_get((4,), lambda: None) is engs[4]
    st = WeightStore()
    a = st.put("w", "raw", torch.ones(4), "cpu")
    b = st.put("w", "raw", torch.full((4,), 2.0), "cpu")
    assert a is b and float(a[0]) == 2.0, "re-pack updates in place (stable address for recorded programs / graphs)"
    assert st.put("w", "ema", torch.ones(4), "cpu") is not a


def test_inference_model_facade_builds_and_mixes_styles():
    """ldm/data/generate_utils.py mirror (reference generate_utils.py:131-190): config rewrite, create_batch, mix_style on pre-computed
    style embeddings (masked slots -> the empty style, per-slot overrides), bbox-mask interpolation helpers."""
    import torch
    from ldm.data.generate_utils import InferenceModel, style_names, interp_mask
    from ldm.util import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
    cfg["model"]["params"]["use_ema"] = False
    im = InferenceModel(cfg, None, "cpu")
    assert type(im.model.extra_cond_models[0]).__name__ == "DummyModel" and type(im.clip_image_encoder).__name__ == "FrozenClipImageEmbedder2"
    b = im.create_batch({"smpl": torch.zeros(1, 85), "txt": "a person"}, repeat=3)
    assert tuple(b["smpl"].shape) == (3, 1, 85) and b["txt"] == ["a person"] * 3
    s = torch.randn(9, 768)
    empty, hat = torch.zeros(768), torch.ones(768)
    out = im.mix_style(s, {"headwear": hat, "top": ""}, mask=["shoes"], empty_style=empty)
    assert tuple(out.shape) == (9, 768)
    assert torch.equal(out[style_names.index("headwear")], hat) and torch.equal(out[style_names.index("shoes")], empty)
    keep = [i for i, n in enumerate(style_names) if n not in ("headwear", "shoes")]
    assert torch.equal(out[keep], s[keep])
    with pytest.raises(NotImplementedError):
        im.mix_style(s, {"top": "a red shirt"})
    a = torch.full((1, 32, 24), -1.0); a[0, 4:20, 6:18] = -0.99215686
    c = torch.full((1, 32, 24), -1.0); c[0, 8:28, 2:10] = -0.99215686
    mid = interp_mask(a, c, 0.5)
    ys, xs = torch.nonzero(mid[0] > -1.0, as_tuple=True)
    assert (int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max())) == (6, 23, 4, 13)


def test_ema_scope_swaps_and_restores_parameters():
    """LitEma.store / copy_to / restore (reference ldm/modules/ema.py:55-76) as multi-tensor copies: inside the scope the module holds the
    EMA weights, afterwards exactly the training weights; the stash buffers are reused across scopes (no per-request allocation)."""
    from ldm.models.diffusion.ddpm import LitEma
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3))
    ema = LitEma(m)
    train = [p.detach().clone() for p in m.parameters()]
    shadow = dict(ema.named_buffers())
    for name, _ in m.named_parameters():
        shadow[ema.m_name2s_name[name]].mul_(0.5).add_(0.25)          # EMA weights that differ from the training weights
    want = [shadow[ema.m_name2s_name[n]].clone() for n, _ in m.named_parameters()]
    for rep in range(2):
        ema.store(m.parameters())
        stash = [c.data_ptr() for c in ema.collected_params]
        ema.copy_to(m)
        assert all(torch.equal(p, w) for p, w in zip(m.parameters(), want))
        ema.restore(m.parameters())
        assert all(torch.equal(p, w) for p, w in zip(m.parameters(), train))
        if rep:
            assert stash == stash0, "the second scope reuses the first scope's stash"
        stash0 = stash
