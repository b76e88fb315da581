"""Eager (no graph) pass of the hot path for ncu launch lists: STEPS U-Net steps at the BASELINE configs[1] shape, then one
VAE decode of the batch.  Prints the launch counts so the list can be cut into its parts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from upgpt_b200 import synth, _C

dev = torch.device("cuda:0")
B, HW, steps = int(os.environ.get("B", 8)), int(os.environ.get("HW", 32)), int(os.environ.get("STEPS", 2))
from upgpt_b200.unet_engine import default_precision
prec = default_precision()
model = bench.build_model(dev, prec)
unet = model.model.diffusion_model
x, mask, ctx = synth.synth_inputs(B, HW, HW, 87, 768, 3)
eng = unet.engine(B, HW, HW, 87, precision=prec)
eng.set_context(ctx.to(dev)); eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long).to(dev))
veng = model.first_stage_model.engine(B, HW, HW)
torch.cuda.synchronize()
L = _C.lib()
l0 = L.upgpt_launch_count()
for _ in range(steps):
    eng.run(use_graph=False)
torch.cuda.synchronize()
l1 = L.upgpt_launch_count()
veng.prog.run(eng._stream())
torch.cuda.synchronize()
l2 = L.upgpt_launch_count()
print("LAUNCHES unet_total=%d per_step=%d vae=%d" % (l1 - l0, eng.launches_per_step, l2 - l1))
