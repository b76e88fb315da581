"""The shipped `mixed` precision plan of the bbox.yaml U-Net, checked on the CPU: the per-layer decisions are read from the engine itself
(recorded without a device) and emulated inside the oracle -- both operands of a GEMM / conv the plan runs on single-plane fp16 are
rounded to fp16, products and sums stay fp32 -- to bound the eps error the plan adds (B200 measurement: 3.2e-4 .. 3.4e-4 in total
against the reference's golden eps, profiles/r01_mixed_precision_check.json).  Guards MIXED_PROFILES against a careless edit."""
import torch
import torch.nn.functional as F

from oracle import ldm_oracle as O
from oracle.ref_loader import BBOX_UNET_KW
from upgpt_b200 import synth
from upgpt_b200.unet_engine import UNetEngine
from ldm.modules.diffusionmodules.openaimodel import UNetModel


def _plan(eng):
    """oracle weight prefix -> True when the engine runs that GEMM on single-plane fp16 operands."""
    def fp16(kind, hw):
        return not eng.use_x3(kind, hw)

    def decide(prefix):
        if ".transformer_blocks." in prefix:
            p = prefix.split(".transformer_blocks.")[0]
            if "attn2.to_k" in prefix or "attn2.to_v" in prefix:
                return False                                   # context K | V: engine default (fp16x3), once per request
            return fp16("tf", eng.layer_hw[p])
        for tail, kind in ((".in_layers.2", None), (".out_layers.3", "conv"), (".skip_connection", None), (".proj_in", "resid1x1"),
                           (".proj_out", "resid1x1"), (".op", "conv"), (".conv", "conv")):
            if prefix.endswith(tail):
                p = prefix[:-len(tail)]
                if p not in eng.layer_hw:
                    return False
                if kind is None:                               # conv1 and the skip GEMM share operand planes
                    kind = "conv_skipshared" if (p + ".skip.weight") in eng.w else "conv"
                    if tail == ".skip_connection":
                        kind = "conv_skipshared"
                return fp16(kind, eng.layer_hw[p])
        return False                                           # time_embed / emb_layers (fp32 GEMV), conv_in, out conv
    return decide


def test_shipped_mixed_plan_adds_less_than_3p5e4(monkeypatch):
    unet = UNetModel(**BBOX_UNET_KW).eval()
    sd = synth.synth_state_dict(unet.state_dict(), 0)
    eng = UNetEngine(unet, 1, 32, 32, 87, precision="mixed", dry=True)
    assert eng.mixed and eng.mixed_hw == (64, 16)
    decide = _plan(eng)
    x, mask, ctx = synth.synth_inputs(1, 32, 32, 87, 768, 0)
    xin, t = torch.cat([x, mask], 1), torch.full((1,), 981, dtype=torch.long)
    with torch.no_grad():
        ref = O.unet_forward(sd, BBOX_UNET_KW, xin, t, ctx)
    conv0, lin0, hits = O.conv, O.lin, {"n": 0}

    def conv(xx, sd_, prefix, stride=1, padding=1):
        if decide(prefix):
            hits["n"] += 1
            return F.conv2d(xx.half().float(), sd_[prefix + ".weight"].half().float(), sd_.get(prefix + ".bias"), stride=stride, padding=padding)
        return conv0(xx, sd_, prefix, stride, padding)

    def lin(xx, sd_, prefix):
        if decide(prefix):
            hits["n"] += 1
            return F.linear(xx.half().float(), sd_[prefix + ".weight"].half().float(), sd_.get(prefix + ".bias"))
        return lin0(xx, sd_, prefix)

    monkeypatch.setattr(O, "conv", conv); monkeypatch.setattr(O, "lin", lin)
    with torch.no_grad():
        y = O.unet_forward(sd, BBOX_UNET_KW, xin, t, ctx)
    added = float((y - ref).abs().max() / ref.abs().max())
    # the engine fuses q|k|v into one GEMM (3 oracle linears -> 1 launch): 64 single-plane launches = 64 + 2 * (self-attention blocks in the plan)
    assert hits["n"] >= 64
    assert 5e-5 < added < 3.5e-4, added
