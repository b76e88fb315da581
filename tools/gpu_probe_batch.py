"""U-Net step time (graph replay) vs batch size, and two concurrent half-batch chains: how much of the B=8 step is a latency chain?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import synth, ops, _C
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from oracle.ref_loader import BBOX_UNET_KW
import ctypes as C
dev = torch.device("cuda:0")
m = UNetModel(**BBOX_UNET_KW); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
res = {}


def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[len(ts) // 2]


with torch.no_grad():
    engs = {}
    for B in (1, 2, 4, 8):
        x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
        e = m.engine(B, 32, 32, 87)
        e.set_context(ctx.to(dev)); e.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
        res["step_ms_B%d" % B] = timeit(lambda: e.run(True))
        engs[B] = e
    # two B=4 chains concurrently: a second engine object (own activations), one graph with a forked branch
    from upgpt_b200.unet_engine import UNetEngine
    for Bh in (4, 2):
        ea = engs[Bh]
        eb = UNetEngine(m, Bh, 32, 32, 87, precision=ea.precision)
        x, mask, ctx = synth.synth_inputs(Bh, 32, 32, 87, 768, 5)
        eb.set_context(ctx.to(dev)); eb.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((Bh,), 500, dtype=torch.long, device=dev))
        L = _C.lib()

        def body():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            aux = C.c_void_p(L.upgpt_aux_stream(1))
            _C.check(L.upgpt_stream_fork(s, 1), "fork")
            for fn, args in eb.prog.calls:
                _C.check(fn(*args, aux), "b")
            for fn, args in ea.prog.calls:
                _C.check(fn(*args, s), "a")
            _C.check(L.upgpt_stream_join(s, 1), "join")
        body(); torch.cuda.synchronize()
        g = ops.Graph().capture(body)
        res["two_chains_B%d_ms" % Bh] = timeit(lambda: g.launch())
        del eb
print(json.dumps(res), flush=True)
