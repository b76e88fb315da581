"""In-graph launch trace of one bbox.yaml U-Net step at B = 8 (upgpt_trace_set): the EFFECTIVE cost of every launch inside the
replayed CUDA graph = difference of consecutive 'predecessor drained' stamps, joined with the recorded program (kernel + shape).
ncu's launch list is serialised and cold-cache; this is the dependent chain as it really runs.

    UPGPT_PAR_SKIP=0 python tools/gpu_trace_step.py [out.txt]      (a forked branch would make the launch order ambiguous)
"""
import os, sys, json, ctypes as C
os.environ.setdefault("UPGPT_PAR_SKIP", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from upgpt_b200 import synth, _C
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.util import load_config
from dump_program import describe

dev = torch.device("cuda:0")
cfg = load_config(os.path.join(ROOT, "configs", "deepfashion", "bbox.yaml"))
kw = dict(cfg.model.params.unet_config.params)
m = UNetModel(**kw); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
B = int(os.environ.get("TRACE_B", "8"))
L = _C.lib()
with torch.no_grad():
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
    e8 = m.engine(B, 32, 32, 87, precision=os.environ.get("UPGPT_PRECISION", "mixed"))
    e8.set_context(ctx.to(dev)); e8.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
    for _ in range(5): e8.run(True)
    torch.cuda.synchronize()
    cap = 4096
    rows = [describe(fn, args) for fn, args in e8.prog.kernel_calls()]
    runs = []
    for rep in range(5):
        buf = torch.zeros(cap + 2, dtype=torch.int64, device=dev); buf[1] = cap
        torch.cuda.synchronize()
        _C.check(L.upgpt_trace_set(buf.data_ptr()), "trace_set")
        e8.run(True); e8.run(True)      # two back-to-back replays: the second one's first launch has a predecessor
        torch.cuda.synchronize()
        _C.check(L.upgpt_trace_set(None), "trace_set")
        h = buf.cpu().numpy().astype(np.uint64)
        n = int(h[0]); st = h[2:2 + n]
        kind = (st & np.uint64(3)).astype(np.int64); t = (st >> np.uint64(2)).astype(np.int64)
        ent = t[kind == 0]; wai = t[kind == 1]
        runs.append((ent, wai))
    nk = len(rows)
    ok = all(len(w) == 2 * nk and len(e) == 2 * nk for e, w in runs)
    out = []
    out.append("in-graph launch trace, bbox.yaml U-Net step, B=%d, %d launches; stamps per replay pair: %s" % (B, nk, [len(w) for _, w in runs]))
    if not ok:
        out.append("stamp count != 2 x program launches: cannot attribute"); print("\n".join(out)); sys.exit(0)
    # effective cost of launch i of the SECOND replay = wait[i+1] - wait[i]; the last one closes with the median gap
    eff = np.zeros((len(runs), nk)); lead = np.zeros((len(runs), nk))
    for r, (ent, wai) in enumerate(runs):
        w = wai[nk:]; e = ent[nk:]
        d = np.diff(w).astype(np.float64) / 1e3
        eff[r, :nk - 1] = d; eff[r, nk - 1] = np.median(d)
        lead[r] = (w - e) / 1e3      # how long the kernel sat resident before its predecessor drained (PDL overlap)
    effm = np.median(eff, 0); leadm = np.median(lead, 0)
    step_us = float(np.median([(w[-1] - w[nk]) / 1e3 for _, w in runs]))
    out.append("sum of effective costs %.1f us (first->last drained stamp of the replay: %.1f us)" % (effm.sum(), step_us))
    agg = {}
    for i, (k, desc, fl) in enumerate(rows):
        key = (k, desc)
        a = agg.setdefault(key, [0, 0.0, 0.0, fl]); a[0] += 1; a[1] += effm[i]; a[2] += leadm[i]
    bykind = {}
    for (k, desc), (n, tt, ld, fl) in agg.items():
        b = bykind.setdefault(k, [0, 0.0]); b[0] += n; b[1] += tt
    out.append("== by kernel")
    for k, (n, tt) in sorted(bykind.items(), key=lambda x: -x[1][1]):
        out.append("  %-28s n=%4d %9.1f us %5.1f%%  avg %6.2f us" % (k, n, tt, 100 * tt / effm.sum(), tt / n))
    out.append("== by shape (effective us per launch in the chain; lead = resident-before-predecessor-drained)")
    out.append("%-60s %4s %9s %8s %8s %9s" % ("launch", "n", "total us", "avg us", "lead us", "MMA TF/s"))
    for (k, desc), (n, tt, ld, fl) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-60s %4d %9.1f %8.2f %8.2f %9.0f" % ((k + " " + desc)[:60], n, tt, tt / n, ld / n, (fl * n / (tt * 1e-6) / 1e12) if (tt and fl) else 0))
    out.append("== in launch order")
    for i, (k, desc, fl) in enumerate(rows):
        out.append("%3d %-58s %7.2f  (lead %5.2f)" % (i, (k + " " + desc)[:58], effm[i], leadm[i]))
txt = "\n".join(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")
print("\n".join(out[:80]))
