"""Does a high-priority main stream (aux branch = default priority) help the forked skip GEMMs hide in the chain's bubbles?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import synth
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.util import load_config
dev = torch.device("cuda:0")
cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
kw = dict(cfg.model.params.unet_config.params)
m = UNetModel(**kw); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
B = 8
x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
res = {}
for name, prio in (("default", None), ("high", -1), ("default2", None), ("high2", -1)):
    s = torch.cuda.Stream(priority=prio) if prio is not None else torch.cuda.current_stream()
    with torch.no_grad(), torch.cuda.stream(s):
        m._engines.clear()
        e8 = m.engine(B, 32, 32, 87, precision="mixed")
        e8.set_context(ctx.to(dev)); e8.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
        for _ in range(3): e8.run(True)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10): e8.run(True)
            e1.record(s); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        res[name] = sorted(ts)[2]
print(json.dumps({"par_skip": os.environ.get("UPGPT_PAR_SKIP", "1"), "ms_per_step": res}), flush=True)
