"""Round-2 probe: eps error vs the reference golden (B=1) and graph-replay U-Net step time (B=8) of the bbox.yaml U-Net under the
scheduling / precision knobs of the environment (UPGPT_PDL, UPGPT_PAR_SKIP, UPGPT_TF_PLANES, ...). One JSON line per run."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from upgpt_b200 import synth
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.util import load_config
dev = torch.device("cuda:0")
golden = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))
cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
kw = dict(cfg.model.params.unet_config.params)
m = UNetModel(**kw); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
prec = os.environ.get("UPGPT_PRECISION", "mixed")
knobs = {k: v for k, v in os.environ.items() if k.startswith("UPGPT_")}
res = {"knobs": knobs}
with torch.no_grad():
    x, mask, ctx = synth.synth_inputs(1, 32, 32, 87, 768, 0)
    eng = m.engine(1, 32, 32, 87, precision=prec)
    eng.set_context(ctx.to(dev))
    errs = {}
    for t in (981, 481):
        eng.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((1,), t, dtype=torch.long, device=dev))
        y = eng.run(use_graph=False).clone().cpu()
        ref = torch.from_numpy(golden[f"bbox_eps_t{t}"])
        errs[t] = float((y - ref).abs().max() / ref.abs().max())
    res["eps_max_rel_b1"] = errs
    B = 8
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
    e8 = m.engine(B, 32, 32, 87, precision=prec)
    e8.set_context(ctx.to(dev)); e8.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
    y_eager = e8.run(False).clone()
    for _ in range(3): e8.run(True)
    assert torch.equal(e8.run(True), y_eager), "graph replay != eager"
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): e8.run(True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    res["ms_per_unet_step_b8"] = sorted(ts)[len(ts) // 2]
    res["launches"] = e8.launches_per_step
print(json.dumps(res), flush=True)
