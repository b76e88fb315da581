"""Where a request through LatentDiffusion.log_images spends its time (the reference's inference entry, generate_utils.py:159-163):
host-side phases with a device synchronize between them; cProfile of one request (top host functions)."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ldm.modules.poses.poses import DummyModel
dev = torch.device("cuda:0")
from upgpt_b200.unet_engine import default_precision
model = bench.build_model(dev, default_precision(), use_ema=True)
model.extra_cond_models[0] = DummyModel()
B, LAT = 8, 32
g = torch.Generator().manual_seed(0)
batch = {"txt": torch.randn(B, 77, 768, generator=g).to(dev), "styles": torch.randn(B, 9, 768, generator=g).to(dev),
         "smpl": (torch.randn(B, 1, 85, generator=g) * 0.5).to(dev), "person_mask": torch.full((B, 1, LAT, LAT), -1.0).to(dev)}
model.image_size = [LAT, LAT]
def request():
    out = model.log_images(batch, N=B, ddim_steps=50, ddim_eta=1.0)
    return torch.clamp(out["samples"], -1., 1.).cpu()
for _ in range(3): request()
torch.cuda.synchronize()
sync = torch.cuda.synchronize
def T(f):
    sync(); t = time.perf_counter(); r = f(); sync(); return r, (time.perf_counter() - t) * 1e3
for rep in range(2):
    (zc), t_in = T(lambda: model.get_input(batch, model.first_stage_key, bs=B))
    z, c = zc
    ctx = model.ema_scope()
    _, t_enter = T(lambda: ctx.__enter__())
    (si), t_s = T(lambda: model.sample_log(cond=c, batch_size=B, ddim=True, ddim_steps=50, eta=1.0, x_T=None))
    _, t_exit = T(lambda: ctx.__exit__(None, None, None))
    img, t_dec = T(lambda: model.decode_first_stage(si[0]))
    _, t_cpu = T(lambda: torch.clamp(img, -1., 1.).cpu())
    print("phases ms: get_input %.2f  ema enter %.2f  sample_log %.2f  ema exit %.2f  decode %.2f  clamp+cpu %.2f  | sum %.2f" %
          (t_in, t_enter, t_s, t_exit, t_dec, t_cpu, t_in + t_enter + t_s + t_exit + t_dec + t_cpu))
t0 = time.perf_counter()
for _ in range(3): request()
sync(); print("request ms (3 back to back): %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
pr = cProfile.Profile(); pr.enable(); request(); sync(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
