"""Dev probe (not a pytest module): which GEMM / conv groups of the bbox.yaml U-Net can take single-plane fp16 operands?
Emulates the engine's operand rounding inside the CPU oracle: in the selected groups both operands of conv / linear are rounded
to fp16 (products and sums stay fp32, like the tensor core's fp32 accumulation); everything else stays fp32 (the fp16x3 planes
carry ~22 bits). Prints eps max-rel error vs the fp32 oracle per group, to be added in quadrature to the fp16x3 floor.
usage: python tests/probe_mixed_precision.py [t ...]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ldm_oracle as O
from oracle.ref_loader import BBOX_UNET_KW
from upgpt_b200 import synth
from ldm.modules.diffusionmodules.openaimodel import UNetModel

torch.set_num_threads(8)
unet = UNetModel(**BBOX_UNET_KW); sd = synth.synth_state_dict(unet.state_dict(), 0); del unet
x, mask, ctx = synth.synth_inputs(1, 32, 32, 87, 768, 0)
xin = torch.cat([x, mask], 1)
_conv, _lin = O.conv, O.lin
SEL = {"fn": lambda kind, prefix, size: False}
SEEN = collections.OrderedDict()

def group_of(kind, prefix, size):
    if kind == "conv":
        name = ("conv3x3" if prefix.endswith(("in_layers.2", "out_layers.3", ".op", ".conv", "out.2")) else
                ("skip1x1" if "skip" in prefix else ("proj_in/out" if "proj_" in prefix else "conv_in")))
    else:
        tail = prefix.split(".transformer_blocks.")[-1] if ".transformer_blocks." in prefix else prefix
        if "attn1.to_q" in tail or "attn1.to_k" in tail or "attn1.to_v" in tail: name = "attn1_qkv"
        elif "attn2.to_q" in tail: name = "attn2_q"
        elif "attn2.to_k" in tail or "attn2.to_v" in tail: name = "attn2_kv(ctx)"
        elif "to_out" in tail: name = "attn_out"
        elif "ff.net.0" in tail: name = "ff1(geglu)"
        elif "ff.net.2" in tail: name = "ff2"
        else: name = "emb"
    return name, size


def conv(x, sd_, prefix, stride=1, padding=1):
    g = group_of("conv", prefix, x.shape[-1] * x.shape[-2] // (stride * stride))
    SEEN[g] = SEEN.get(g, 0) + 1
    if SEL["fn"](*g):
        w = sd_[prefix + ".weight"].half().float()
        return torch.nn.functional.conv2d(x.half().float(), w, sd_.get(prefix + ".bias"), stride=stride, padding=padding)
    return _conv(x, sd_, prefix, stride, padding)


def lin(x, sd_, prefix):
    size = x.shape[1] if x.dim() == 3 else 0
    if "attn2.to_k" in prefix or "attn2.to_v" in prefix: size = -1
    g = group_of("lin", prefix, size)
    SEEN[g] = SEEN.get(g, 0) + 1
    if SEL["fn"](*g):
        return torch.nn.functional.linear(x.half().float(), sd_[prefix + ".weight"].half().float(), sd_.get(prefix + ".bias"))
    return _lin(x, sd_, prefix)


O.conv, O.lin = conv, lin


def policy(deep_hw, full_hw):
    """The engine's `mixed` rule: levels with H*W <= full_hw run every GEMM / conv on single-plane fp16 operands; levels with
    H*W <= deep_hw all but the residual-path 1x1s (skip_connection, proj_in / proj_out) and the conv sharing its operand with a skip."""
    def sel(kind, prefix, x, stride):
        if kind == "conv":
            hw = x.shape[-1] * x.shape[-2] // (stride * stride)
            if hw <= full_hw: return True
            if hw > deep_hw: return False
            if "skip" in prefix or "proj_" in prefix: return False
            if prefix.endswith("in_layers.2") and (prefix[:-len("in_layers.2")] + "skip_connection.weight") in sd: return False
            return True
        if x.dim() != 3 or "attn2.to_k" in prefix or "attn2.to_v" in prefix or ".transformer_blocks." not in prefix: return False
        return x.shape[1] <= deep_hw
    return sel


def conv2(x, sd_, prefix, stride=1, padding=1):
    if POL["fn"]("conv", prefix, x, stride):
        return torch.nn.functional.conv2d(x.half().float(), sd_[prefix + ".weight"].half().float(), sd_.get(prefix + ".bias"), stride=stride, padding=padding)
    return _conv(x, sd_, prefix, stride, padding)


def lin2(x, sd_, prefix):
    if POL["fn"]("lin", prefix, x, 1):
        return torch.nn.functional.linear(x.half().float(), sd_[prefix + ".weight"].half().float(), sd_.get(prefix + ".bias"))
    return _lin(x, sd_, prefix)


POL = {"fn": None}
if sys.argv[1:2] == ["policy"]:
    O.conv, O.lin = conv2, lin2
    from oracle.make_golden import TINY_UNET_KW
    cases = [("bbox", BBOX_UNET_KW, sd, xin, ctx, [981, 481, 501, 1])]
    tu = UNetModel(**TINY_UNET_KW); sdt = synth.synth_state_dict(tu.state_dict(), 0); del tu
    xt, mt, ct = synth.synth_inputs(2, 16, 16, 87, 128, 0)
    cases.append(("tiny", TINY_UNET_KW, sdt, torch.cat([xt, mt], 1), ct, [981, 1]))
    xt, mt, ct = synth.synth_inputs(3, 16, 24, 20, 128, 1)
    tu = UNetModel(**TINY_UNET_KW); sdt1 = synth.synth_state_dict(tu.state_dict(), 1); del tu
    cases.append(("tinyrect", TINY_UNET_KW, sdt1, torch.cat([xt, mt], 1), ct, [500]))
    from oracle.ref_loader import UPSCALE_UNET_KW
    ukw = UPSCALE_UNET_KW
    tu = UNetModel(**ukw); sdu = synth.synth_state_dict(tu.state_dict(), 2); del tu
    xt, mt, ct = synth.synth_inputs(1, 32, 24, 86, ukw["context_dim"], 2, concat_channels=ukw["in_channels"] - ukw["out_channels"])
    cases.insert(0, ("upscale", ukw, sdu, torch.cat([xt[:, :ukw["out_channels"]], mt], 1), ct, [481]))
    if sys.argv[2:3] == ["upscale"]:
        cases = cases[:1]
    with torch.no_grad():
        for tag, kw, sd, xi, cx, tsl in cases:
            for t in tsl:
                tt = torch.full((xi.shape[0],), t, dtype=torch.long)
                POL["fn"] = lambda *a: False
                ref = O.unet_forward(sd, kw, xi, tt, cx)
                out = []
                for deep, full in ((64, 16), (64, 0), (16, 16), (0, 16), (256, 16), (64, 64)):
                    POL["fn"] = policy(deep, full)
                    y = O.unet_forward(sd, kw, xi, tt, cx)
                    out.append("deep<=%d full<=%d: %.2e" % (deep, full, float((y - ref).abs().max() / ref.abs().max())))
                print(tag, "t=%d" % t, " | ".join(out), flush=True)
    sys.exit(0)

ts = [int(a) for a in sys.argv[1:]] or [981, 481]
with torch.no_grad():
    for t in ts:
        tt = torch.full((1,), t, dtype=torch.long)
        SEEN.clear()
        SEL["fn"] = lambda *a: False
        ref = O.unet_forward(sd, BBOX_UNET_KW, xin, tt, ctx)
        err = lambda y: float((y - ref).abs().max() / ref.abs().max())
        SEL["fn"] = lambda name, size: name != "emb"
        print(f"t={t}: everything fp16: {err(O.unet_forward(sd, BBOX_UNET_KW, xin, tt, ctx)):.2e}")
        groups = sorted((g for g in SEEN if g[0] not in ("emb", "conv_in")), key=lambda g: (-g[1], g[0]))
        counts = dict(SEEN)
        for g in groups:
            SEL["fn"] = lambda name, size, g=g: (name, size) == g
            e = err(O.unet_forward(sd, BBOX_UNET_KW, xin, tt, ctx))
            print(f"   only {g[0]:14s} at {g[1]:5d} tokens/image ({counts[g] // 2:3d} GEMMs) in fp16: {e:.2e}", flush=True)
        for size in (1024, 256, 64, 16):
            SEL["fn"] = lambda name, s, size=size: s == size and name != "emb"
            print(f"   every GEMM at {size:5d} tokens/image in fp16: {err(O.unet_forward(sd, BBOX_UNET_KW, xin, tt, ctx)):.2e}", flush=True)
