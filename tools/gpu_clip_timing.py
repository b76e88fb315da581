"""Time of the two CLIP ViT-L/14 conditioning towers per request of 8 samples (graph replay, CUDA events): text (8 x 77 tokens) and
image (8 x 9 style crops of 224 x 224), with their algorithmic GEMM + attention flops."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import synth
from ldm.modules.encoders.modules import FrozenCLIPEmbedder, FrozenClipImageEmbedder2
dev = torch.device("cuda:0")
res = {}


def timed(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t = FrozenCLIPEmbedder().materialize()
t.transformer.load_state_dict(synth.synth_state_dict(t.transformer.state_dict(), 2)); t = t.to(dev)
ids = torch.randint(0, 49408, (8, 77), device=dev)
eng = t.engine(8, 77); eng.bufs["ids"].copy_(ids)
ms = timed(lambda: eng.run(True))
M, C, L = 8 * 77, 768, 12
gf = L * (2 * M * C * (3 * C + C + 8 * C) + 4 * 8 * 77 * 77 * C) / 1e9
res["text_tower"] = {"shape": "8 x 77 tokens, 12 layers x 768", "ms": ms, "launches": eng.launches, "algorithmic_gflop": gf, "tflops": gf / ms}
del t, eng; torch.cuda.empty_cache()
v = FrozenClipImageEmbedder2().materialize()
v.model.load_state_dict(synth.synth_state_dict(v.model.state_dict(), 3)); v = v.to(dev)
n = 72
eng = v.engine(n); eng.bufs["img"].normal_()
ms = timed(lambda: eng.run(True), 5)
M, C, L, T = n * 257, 1024, 24, 257
gf = (L * (2 * M * C * (3 * C + C + 8 * C) + 4 * n * T * T * C) + 2 * n * 256 * 588 * C + 2 * n * C * 768) / 1e9
res["image_tower"] = {"shape": "72 crops (8 x 9) of 224x224 -> 257 tokens, 24 layers x 1024", "ms": ms, "launches": eng.launches,
                      "algorithmic_gflop": gf, "tflops": gf / ms}
eng16 = v.engine(n, precision="fp16"); eng16.bufs["img"].copy_(eng.bufs["img"])
ms16 = timed(lambda: eng16.run(True), 5)
res["image_tower_fp16"] = {"shape": res["image_tower"]["shape"], "ms": ms16, "tflops": gf / ms16,
                           "max_rel_vs_fp16x3": float((eng16.bufs["out"] - eng.bufs["out"]).abs().max() / eng.bufs["out"].abs().max())}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "clip_timing.json"), "w"), indent=1)
