"""Every candidate of the precision calibration (upgpt_b200/precision.py) on the bbox.yaml U-Net: deviation of eps from the fp16x3 eps of
the same weights (B = 1, t in {981, 481}: what the calibration measures), eps vs the reference golden, and the B = 8 step time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from upgpt_b200 import synth, precision as P
from upgpt_b200.host import WeightStore
from upgpt_b200.unet_engine import UNetEngine
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.util import load_config
dev = torch.device("cuda:0")
golden = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))
cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
kw = dict(cfg.model.params.unet_config.params)
m = UNetModel(**kw); m.load_state_dict(synth.synth_state_dict(m.state_dict(), int(os.environ.get("WSEED", "0")))); m = m.to(dev).eval()
x1, mask1, ctx1 = synth.synth_inputs(1, 32, 32, 87, 768, 0)
x8, mask8, ctx8 = synth.synth_inputs(8, 32, 32, 87, 768, 3)
cands = P.candidates(32, 32, len(m.channel_mult))
if os.environ.get("EXTRA_PLANS"):
    # experiments: JSON {name: {"base": candidate name, ...plan keys to override / add}}
    base = dict(cands)
    extra = json.loads(os.environ["EXTRA_PLANS"])
    cands = [(n, dict(base[v.pop("base")], **v)) for n, v in extra.items()] + [cands[-1]]
ref = None
with torch.no_grad():
    for name, plan in reversed(cands):
        prec = "mixed" if plan["mixed_hw"] is not None else "fp16x3"
        e1 = UNetEngine(m, 1, 32, 32, 87, precision=prec, plan=dict(plan, name=name), store=WeightStore())
        e1.set_context(ctx1.to(dev))
        outs, errs = [], {}
        for t in (981, 481):
            e1.stage_inputs(torch.cat([x1, mask1], 1).to(dev), torch.full((1,), t, dtype=torch.long, device=dev))
            y = e1.run(use_graph=False).clone(); outs.append(y)
            if int(os.environ.get("WSEED", "0")) == 0:
                g = torch.from_numpy(golden[f"bbox_eps_t{t}"]).to(dev)
                errs[t] = float((y - g).abs().max() / g.abs().max())
        if ref is None:
            ref = outs
        devn = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, ref))
        del e1
        e8 = UNetEngine(m, 8, 32, 32, 87, precision=prec, plan=dict(plan, name=name), store=WeightStore())
        e8.set_context(ctx8.to(dev)); e8.stage_inputs(torch.cat([x8, mask8], 1).to(dev), torch.full((8,), 500, dtype=torch.long, device=dev))
        for _ in range(3): e8.run(True)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): e8.run(True)
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 10)
        del e8; torch.cuda.empty_cache()
        print(json.dumps({"plan": name, "deviation_vs_fp16x3": devn, "eps_vs_reference_b1": errs, "ms_per_unet_step_b8": sorted(ts)[2]}), flush=True)
