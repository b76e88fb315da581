"""Runs a few eager U-Net steps at the BASELINE config-2 shape (B=8, 32x32, 87-token context) for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import synth
from oracle.ref_loader import BBOX_UNET_KW
from ldm.modules.diffusionmodules.openaimodel import UNetModel

B = int(os.environ.get("B", 8)); HW = int(os.environ.get("HW", 32)); steps = int(os.environ.get("STEPS", 2))
dev = torch.device("cuda:0")
m = UNetModel(**BBOX_UNET_KW)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
x, mask, ctx = synth.synth_inputs(B, HW, HW, 87, 768, 3)
eng = m.engine(B, HW, HW, 87, precision=os.environ.get("UPGPT_PRECISION", "fp16"))
eng.set_context(ctx.to(dev)); eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long).to(dev))
torch.cuda.synchronize()
for _ in range(steps):
    eng.run(use_graph=False)
torch.cuda.synchronize()
print("done", eng.launches_per_step)
