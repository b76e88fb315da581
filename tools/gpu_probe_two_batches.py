"""Two independent B = 8 U-Net steps in flight (one graph, two streams) vs one after the other: do two latency-bound chains fill each
other's bubbles? (UPGPT_PAR_SKIP=0: a chain's own forked branch would share the auxiliary stream with the other chain.)"""
import os, sys, json
os.environ.setdefault("UPGPT_PAR_SKIP", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, ctypes as C
from upgpt_b200 import synth, ops, _C
from upgpt_b200.unet_engine import UNetEngine
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from oracle.ref_loader import BBOX_UNET_KW
dev = torch.device("cuda:0")
m = UNetModel(**BBOX_UNET_KW); m.load_state_dict(synth.synth_state_dict(m.state_dict(), 0)); m = m.to(dev).eval()
L = _C.lib()

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[2]

res = {}
with torch.no_grad():
    B = 8
    x, mask, ctx = synth.synth_inputs(B, 32, 32, 87, 768, 3)
    ea = m.engine(B, 32, 32, 87)
    ea.set_context(ctx.to(dev)); ea.stage_inputs(torch.cat([x, mask], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
    res["one_chain_ms"] = timeit(lambda: ea.run(True))
    ya = ea.run(True).clone()
    eb = UNetEngine(m, B, 32, 32, 87, precision=ea.precision, plan=dict(m._plans[("raw", 32, 32, 87)][1]) if hasattr(m, "_plans") and m._plans else None)
    x2, mask2, ctx2 = synth.synth_inputs(B, 32, 32, 87, 768, 5)
    eb.set_context(ctx2.to(dev)); eb.stage_inputs(torch.cat([x2, mask2], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
    yb_ref = eb.run(False).clone()

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
    res["two_chains_ms"] = timeit(lambda: g.launch())
    g.launch(); torch.cuda.synchronize()
    res["results_identical"] = bool(torch.equal(ea.bufs["eps"], ya) and torch.equal(eb.bufs["eps"], yb_ref))
    res["throughput_gain"] = 2 * res["one_chain_ms"] / res["two_chains_ms"]
    # three chains: main + both auxiliary streams
    ec = UNetEngine(m, B, 32, 32, 87, precision=ea.precision, plan=dict(m._plans[("raw", 32, 32, 87)][1]) if hasattr(m, "_plans") and m._plans else None)
    x3, mask3, ctx3 = synth.synth_inputs(B, 32, 32, 87, 768, 6)
    ec.set_context(ctx3.to(dev)); ec.stage_inputs(torch.cat([x3, mask3], 1).to(dev), torch.full((B,), 500, dtype=torch.long, device=dev))
    yc_ref = ec.run(False).clone()

    def body3():
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        a0, a1 = C.c_void_p(L.upgpt_aux_stream(0)), C.c_void_p(L.upgpt_aux_stream(1))
        _C.check(L.upgpt_stream_fork(s, 0), "fork"); _C.check(L.upgpt_stream_fork(s, 1), "fork")
        for fn, args in ec.prog.calls:
            _C.check(fn(*args, a0), "c")
        for fn, args in eb.prog.calls:
            _C.check(fn(*args, a1), "b")
        for fn, args in ea.prog.calls:
            _C.check(fn(*args, s), "a")
        _C.check(L.upgpt_stream_join(s, 0), "join"); _C.check(L.upgpt_stream_join(s, 1), "join")
    body3(); torch.cuda.synchronize()
    g3 = ops.Graph().capture(body3)
    res["three_chains_ms"] = timeit(lambda: g3.launch())
    g3.launch(); torch.cuda.synchronize()
    res["results_identical_3"] = bool(torch.equal(ea.bufs["eps"], ya) and torch.equal(eb.bufs["eps"], yb_ref) and torch.equal(ec.bufs["eps"], yc_ref))
    res["throughput_gain_3"] = 3 * res["one_chain_ms"] / res["three_chains_ms"]
print(json.dumps(res), flush=True)
