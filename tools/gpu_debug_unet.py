"""Block-level debug: determinism across runs and per-block error vs the CPU oracle (tiny config)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import synth
from oracle import ldm_oracle as O
from oracle.make_golden import TINY_UNET_KW
from oracle.ref_loader import BBOX_UNET_KW
from ldm.modules.diffusionmodules.openaimodel import UNetModel

dev = torch.device("cuda:0")
cfgname = sys.argv[1] if len(sys.argv) > 1 else "tiny"
kw, B, H, W, L = (TINY_UNET_KW, 2, 16, 16, 87) if cfgname == "tiny" else (BBOX_UNET_KW, 1, 32, 32, 87)
m = UNetModel(**kw)
sd = synth.synth_state_dict(m.state_dict(), 0)
m.load_state_dict(sd); m = m.to(dev).eval()
x, mask, ctx = synth.synth_inputs(B, H, W, L, kw["context_dim"], 0)
xc = torch.cat([x, mask], 1)
tt = torch.full((B,), 981, dtype=torch.long)
taps = {}
with torch.no_grad():
    y_ref = O.unet_forward(sd, kw, xc, tt, ctx, taps)
for prec in ("fp16", "fp16x3"):
    eng = m.engine(B, H, W, L, precision=prec)
    eng.set_context(ctx.to(dev)); eng.stage_inputs(xc.to(dev), tt.to(dev))
    runs = []
    for r in range(3):
        eng.run(use_graph=False); torch.cuda.synchronize()
        runs.append({k: v.clone() for k, v in eng.bufs.items() if k.endswith(".out") or k in ("eps", "h_in0", "emb_all")})
    print(f"==== {prec}: block outputs: err vs oracle | run0-vs-run1 | run0-vs-run2")
    for k in runs[0]:
        a = runs[0][k].float()
        d1 = (a - runs[1][k].float()).abs().max().item() / max(a.abs().max().item(), 1e-9)
        d2 = (a - runs[2][k].float()).abs().max().item() / max(a.abs().max().item(), 1e-9)
        name = k[:-4] if k.endswith(".out") else k
        e = float("nan")
        if name in taps:
            r = taps[name]
            g = a.reshape(B, r.shape[2], r.shape[3], r.shape[1]).permute(0, 3, 1, 2).cpu()
            e = ((g - r).abs().max() / r.abs().max()).item()
        elif k == "eps":
            e = ((a.cpu() - y_ref).abs().max() / y_ref.abs().max()).item()
        elif k == "h_in0":
            r = taps["input_blocks.0.0"]
            g = a.reshape(B, H, W, -1).permute(0, 3, 1, 2).cpu(); e = ((g - r).abs().max() / r.abs().max()).item()
        print(f"{name:45s} err {e:9.3e} | {d1:9.3e} | {d2:9.3e}")
