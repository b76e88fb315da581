"""Where the headline batch goes: U-Net step (graph replay) in both precision modes, VAE decode, sampler loop, at the
BASELINE configs[1] shape (B=8, 32x32x4 latent, 87x768 context).  CUDA events on the launching stream."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from upgpt_b200 import synth, ops
from ldm.models.diffusion.ddim import DDIMSampler

dev = torch.device("cuda:0")
B, HW = int(os.environ.get("B", 8)), int(os.environ.get("HW", 32))
model = bench.build_model(dev, "fp16x3")
unet = model.model.diffusion_model
x, mask, ctx = synth.synth_inputs(B, HW, HW, 87, 768, 3)
res = {}


def ev_time(fn, reps):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for prec in ("fp16x3", "fp16"):
    eng = unet.engine(B, HW, HW, 87, precision=prec)
    eng.set_context(ctx.to(dev)); eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long).to(dev))
    res["unet_step_graph_ms_" + prec] = ev_time(lambda: eng.run(use_graph=True), 20)
    res["unet_step_eager_ms_" + prec] = ev_time(lambda: eng.run(use_graph=False), 5)
    res["unet_launches_" + prec] = eng.launches_per_step
z = torch.randn(B, 4, HW, HW, device=dev)
res["vae_decode_ms"] = ev_time(lambda: model.decode_first_stage(z), 5)
img = model.decode_first_stage(z)
res["to_uint8_ms"] = ev_time(lambda: ops.to_uint8_nhwc(img), 5)
sampler = DDIMSampler(model)
cond = {"c_crossattn": ctx.to(dev), "c_concat": [mask.to(dev)]}
xd = x.to(dev)
for prec in ("fp16x3", "fp16"):
    os.environ["UPGPT_PRECISION"] = prec
    res["ddim50_ms_" + prec] = ev_time(lambda: sampler.sample(50, B, (4, HW, HW), conditioning=cond, eta=1.0, x_T=xd, verbose=False, log_every_t=1000), 2)
print(json.dumps(res, indent=1))
