"""Like-for-like GPU baseline: the reference algorithm as plain PyTorch eager ops ON THE B200 -- the CPU oracle's
functional restatement moved to cuda:0 -- at BASELINE configs[1] (bbox.yaml U-Net, B=8, 32x32 latent; KL-f8 decode of the batch), in
strict fp32, with TF32 allowed, and under fp16 autocast.  Not a pytest module and not part of the product: it only puts a number
beside bench.py's (`python tests/gpu_eager_baseline.py`, needs a GPU)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ldm_oracle as O
from oracle.ref_loader import BBOX_UNET_KW, BBOX_VAE_KW
from upgpt_b200 import synth
from ldm.modules.diffusionmodules.openaimodel import UNetModel
from ldm.models.autoencoder import AutoencoderKL

dev = torch.device("cuda:0")
unet = UNetModel(**BBOX_UNET_KW); sd_u = {k: v.to(dev) for k, v in synth.synth_state_dict(unet.state_dict(), 0).items()}; del unet
ae = AutoencoderKL(BBOX_VAE_KW, embed_dim=4); sd_v = {k: v.to(dev) for k, v in synth.synth_state_dict(ae.state_dict(), 0).items()}; del ae
x, mask, ctx = [t.to(dev) for t in synth.synth_inputs(8, 32, 32, 87, 768, 0)]
xin, t = torch.cat([x, mask], 1), torch.full((8,), 501, dtype=torch.long, device=dev)
torch.set_default_device(dev)     # the oracle builds its small tables (timestep frequencies) on the default device


def timed(fn, n):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
with torch.no_grad():
    for name, tf32, amp in (("fp32", False, False), ("tf32", True, False), ("fp16_autocast", True, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32; torch.backends.cudnn.allow_tf32 = tf32
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            ms_u = timed(lambda: O.unet_forward(sd_u, BBOX_UNET_KW, xin, t, ctx), 5)
            ms_v = timed(lambda: O.decode_first_stage(sd_v, BBOX_VAE_KW, x, 0.18215), 3)
        res[name] = {"unet_step_ms_b8": ms_u, "vae_decode_ms_b8": ms_v, "images_per_s_50_step_ddim": 8.0 / ((50 * ms_u + ms_v) * 1e-3)}
        print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "eager_baseline.json"), "w"), indent=1)
