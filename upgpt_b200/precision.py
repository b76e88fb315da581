"""Load-time calibration of the "mixed" precision plan (which GEMMs / convs may run on single-plane fp16 operands).

The eps error of a plan depends on the WEIGHTS, not only on the architecture (profiles/r01_precision_sensitivity.txt: the same
thresholds cost 2.9e-4 on one U-Net and 5.3e-4 on another), so a plan tuned on synthetic weights proves nothing about a real
checkpoint.  BASELINE.json's tolerance is 1e-3 on eps; uniform fp16x3 (every operand as [hi | lo] planes) sits at 1e-4 .. 2e-4.
At the first engine request after the weights changed, the module therefore measures -- on the device, B = 1, two timesteps, seeded
synthetic latent / context -- the eps of each candidate plan against the fp16x3 eps of the same weights and keeps the FASTEST
candidate whose deviation stays under `LIMIT`:

    deep+tf1C : single plane in the two deepest levels (weight-bandwidth bound; incl. the ResBlock convs whose input also feeds a skip 1x1
                GEMM -- that GEMM keeps [hi | lo] planes of its own), in the decoder ResBlocks' first conv (input [h | skip]: the most
                expensive fp16x3 launches) and in every attention projection / feed-forward GEMM except the attention out-projections and
                ff2 of the full-resolution level (the worst error per microsecond saved, profiles/r01_precision_sensitivity.txt)
    deep+tf1sC: the same, but every attention / feed-forward GEMM of the full-resolution level keeps [hi | lo] planes
    deep+tf1s : as deep+tf1sC without the decoder convs
    deep     : single plane in the two deepest levels only (the round-1 plan)
    deepest  : single plane in the deepest level only
    fp16x3   : error-compensated operands everywhere

so a checkpoint whose eps is more sensitive falls back automatically instead of silently exceeding the tolerance.
UPGPT_CALIBRATE=0 disables the measurement (static profile of unet_engine.MIXED_PROFILES); UPGPT_CALIBRATE_LIMIT overrides LIMIT.
"""
import os

import torch

LIMIT = 6.5e-4        # max |eps_plan - eps_fp16x3| / max |eps_fp16x3|; + fp16x3's own <= 2e-4 stays inside the 1e-3 tolerance (the
                      # per-sample maximum at B = 8 runs ~1.3x the B = 1 probe: 7.5e-4 for a probe deviation of 5.8e-4)
TIMESTEPS = (981, 481)


# what each plan runs on single fp16 planes (everything else: [hi | lo] planes, 3 MMAs per product); bench.py quotes these
DESCRIPTIONS = {
    "deep+tf1C": "single f16 plane in the weight-bound 8x8 / 4x4 levels, in the decoder ResBlocks' first conv (input [h | skip]) of every level and in "
                 "every attention projection / feed-forward GEMM except the attention out-projections and ff2 of the 32x32 level",
    "deep+tf1sC": "single f16 plane in the weight-bound 8x8 / 4x4 levels, in the decoder ResBlocks' first conv (input [h | skip]) of every level and in "
                  "the attention projection / feed-forward GEMMs below the 32x32 level",
    "deep+tf1s": "single f16 plane in the weight-bound 8x8 / 4x4 levels and in the attention projection / feed-forward GEMMs below the 32x32 level",
    "deep": "single f16 plane in the weight-bound 8x8 / 4x4 levels",
    "deepest": "single f16 plane in the deepest level",
    "tf1": "single f16 plane in every attention projection / feed-forward GEMM",
    "fp16x3": "nowhere",
}


def candidates(H, W, n_levels):
    """Fastest first. Measured on the bbox.yaml U-Net, synthetic weights, B200 (tools/gpu_probe_candidates.py, profiles/r02_precision_
    candidates.jsonl): deviation from fp16x3 / U-Net step at B = 8 = 6.2e-4 / 3.98 ms (deep+tf1C), ~5e-4 / 4.07 (deep+tf1sC), 4.0e-4 /
    4.17 (deep+tf1s), 3.6e-4 / 4.42 (deep), 1.7e-4 / 4.67 (deepest), 0 / 4.78 (fp16x3). (The first r2 plan, every transformer GEMM +
    the two deep levels on single planes, sat at 6.3e-4 .. 6.8e-4 / 4.04 ms: dominated by deep+tf1C, which moves the attention
    out-projections and ff2 of the full-resolution level -- the worst error per microsecond saved -- back to [hi | lo] planes and
    spends that budget on the decoder's concatenated-input convs, the most expensive fp16x3 launches.)"""
    hw0 = H * W
    deep = (max(hw0 // 16, 1), max(hw0 // 64, 1)) if n_levels >= 3 else None
    out = []
    if deep is not None:
        keep = {"tf_out": hw0, "tf_ff2": hw0}
        out.append(("deep+tf1C", dict(mixed_hw=deep, tf_x1=True, skip_x1=True, tf_keep_x3=keep, concat_x1_hw=hw0)))
        out.append(("deep+tf1sC", dict(mixed_hw=deep, tf_x1=True, skip_x1=True, tf_hw=max(hw0 // 4, 1), concat_x1_hw=hw0)))
        out.append(("deep+tf1s", dict(mixed_hw=deep, tf_x1=True, skip_x1=True, tf_hw=max(hw0 // 4, 1))))
        out.append(("deep", dict(mixed_hw=deep, tf_x1=False)))
        out.append(("deepest", dict(mixed_hw=(deep[1], deep[1]), tf_x1=False)))
    else:
        out.append(("tf1", dict(mixed_hw=(0, 0), tf_x1=True)))
    out.append(("fp16x3", dict(mixed_hw=None, tf_x1=False)))
    return out


def enabled():
    return os.environ.get("UPGPT_CALIBRATE", "1") != "0"


@torch.no_grad()
def calibrate(unet, H, W, ctx_len, limit=None):
    """-> (plan dict for UNetEngine(plan=...), report dict). Engines built here pack into a throw-away weight store."""
    from .host import WeightStore
    from .unet_engine import UNetEngine
    limit = float(os.environ.get("UPGPT_CALIBRATE_LIMIT", LIMIT if limit is None else limit))
    dev = next(unet.parameters()).device
    g = torch.Generator().manual_seed(20261017)
    lat = min(unet.in_channels, unet.out_channels)
    x = torch.randn(1, lat, H, W, generator=g).to(dev)
    xc = torch.full((1, max(unet.in_channels - lat, 1), H, W), -1.0, device=dev)
    ctx = torch.randn(1, ctx_len, unet.context_dim, generator=g).to(dev)

    def eps_of(plan):
        store = WeightStore()
        eng = UNetEngine(unet, 1, H, W, ctx_len, precision="mixed" if plan["mixed_hw"] is not None else "fp16x3", plan=plan, store=store)
        eng.set_context(ctx)
        outs = []
        for t in TIMESTEPS:
            eng.stage_inputs(x, torch.full((1,), t, dtype=torch.long, device=dev), xc if unet.in_channels > lat else None)
            outs.append(eng.run(use_graph=False).clone())
        del eng, store
        return outs

    cands = candidates(H, W, len(unet.channel_mult))
    ref = eps_of(cands[-1][1])
    scale = [float(r.abs().max()) for r in ref]
    report = {"limit": limit, "timesteps": list(TIMESTEPS), "deviation_vs_fp16x3": {}}
    chosen = cands[-1]
    if min(scale) == 0.0:
        # zero_module-initialised U-Net (openaimodel.py:229-231,685): eps == 0 whatever the precision -- nothing to calibrate against
        report["degenerate"] = "eps is identically 0 with these weights"
        chosen = cands[0]
    else:
        for name, plan in cands[:-1]:
            got = eps_of(plan)
            dev_ = max(float((a - b).abs().max()) / s for a, b, s in zip(got, ref, scale))
            report["deviation_vs_fp16x3"][name] = dev_
            if dev_ <= limit:
                chosen = (name, plan)
                break
    report["chosen"] = chosen[0]
    torch.cuda.empty_cache()
    return dict(chosen[1], name=chosen[0]), report
