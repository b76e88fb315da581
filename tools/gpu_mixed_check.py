"""eps error vs the reference golden and graph-replay step time of the bbox.yaml U-Net per precision mode (mixed / fp16x3 / fp16)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from upgpt_b200 import synth
from upgpt_b200.unet_engine import UNetEngine
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.util import load_config
dev = torch.device("cuda:0")
golden = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))
cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
kw = dict(cfg.model.params.unet_config.params)
m = UNetModel(**kw); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
modes = sys.argv[1:] or ["mixed", "fp16x3", "fp16"]
res = {}
with torch.no_grad():
    for prec in modes:
        x, mask, ctx = synth.synth_inputs(1, 32, 32, 87, 768, 0)
        eng = m.engine(1, 32, 32, 87, precision=prec)
        eng.set_context(ctx.to(dev))
        errs = {}
        for t in (981, 481):
            eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((1,), t, dtype=torch.long, device=dev))
            y = eng.run(use_graph=False).clone().cpu()
            ref = torch.from_numpy(golden[f"bbox_eps_t{t}"])
            errs[t] = float((y - ref).abs().max() / ref.abs().max())
        B = 8
        x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
        e8 = m.engine(B, 32, 32, 87, precision=prec)
        e8.set_context(ctx.to(dev)); e8.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
        for _ in range(3): e8.run(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): e8.run(True)
        e1.record(); torch.cuda.synchronize()
        res[prec] = {"eps_max_rel": errs, "ms_per_unet_step_b8": e0.elapsed_time(e1) / 20, "launches": e8.launches_per_step}
        print(prec, res[prec], flush=True)
        del eng, e8; m._engines.clear(); torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mixed_check.json"), "w"), indent=1)
