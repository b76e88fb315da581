"""Lists the U-Net engine's program (one line per launch: kernel + shape) in launch order -- CPU only, no kernels run.
Joined with an ncu launch list it attributes time to layers: python tools/dump_program.py [launches.csv]"""
import csv, os, re, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from upgpt_b200 import _C


def describe(fn, args):
    name = fn.__name__ if hasattr(fn, "__name__") else str(fn)
    if name == "upgpt_gemm":
        a = C.cast(args[0], C.POINTER(_C.GemmArgs)).contents
        mode = {0: "gemm", 1: "conv3x3", 2: "conv3x3s2", 3: "conv1x1", 4: "conv3x3s2a", 5: "conv3x3up2"}[a.mode]
        M = a.M * max(a.batch, 1) if a.mode == 0 else a.n_imgs * a.H * a.W * (4 if a.mode == 5 else 1)
        taps = 9 if a.mode in (1, 2, 4) else (4 if a.mode == 5 else 1)
        fl = 2.0 * M * a.N * a.K * taps
        x3 = bool(a.flags & _C.GEMM_F_X3)
        return "gemm", "%-10s M=%5d N=%4d K=%4d%s%s%s" % (mode, M, a.N, a.K * taps, " x3" if x3 else "", " geglu" if a.flags & 2 else "", " chw" if a.flags & 4 else "") + (" +res" if a.res32 else "") + (" +rv" if a.rowvec else "") + (" ln" if a.ln_stats else "") + (" rs" if a.rowstats_out else "") + (" h16" if (a.out16 and a.out32) else (" f16" if a.out16 else "")), fl * (3 if x3 else 1)
    if name == "upgpt_attention":
        a = C.cast(args[0], C.POINTER(_C.AttnArgs)).contents
        return "attention", "B=%d H=%2d Nq=%4d Nk=%4d dpad=%3d" % (a.B, a.H, a.Nq, a.Nk, a.dpad), 4.0 * a.B * a.H * a.Nq * a.Nk * a.dpad
    if name in ("upgpt_groupnorm_prep", "upgpt_prep_operand"):
        a = C.cast(args[0], C.POINTER(_C.PrepArgs)).contents
        return name.replace("upgpt_", ""), "C=%4d+%4d HW=%4d layout=%d silu=%d%s" % (a.C1, a.C2, a.H * a.W, a.layout, a.silu, " x3" if a.split3 else ""), 0.0
    return name.replace("upgpt_", ""), "", 0.0


def main():
    # build the engine on the meta level: we only need the recorded program, which needs device buffers -> requires CUDA
    assert torch.cuda.is_available(), "needs a GPU (buffers are device tensors)"
    import bench
    from upgpt_b200 import synth
    dev = torch.device("cuda:0")
    from upgpt_b200.unet_engine import default_precision
    model = bench.build_model(dev, default_precision())
    eng = model.model.diffusion_model.engine(8, 32, 32, 87)
    rows = [describe(fn, args) for fn, args in eng.prog.kernel_calls()]
    durs = None
    if len(sys.argv) > 1:
        lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
        d = [float(r["Metric Value"]) / 1e3 for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
        n_vae = int(sys.argv[2]) if len(sys.argv) > 2 else 95
        durs = d[-n_vae - len(rows):-n_vae]
    agg = {}
    for i, (k, desc, fl) in enumerate(rows):
        t = durs[i] if durs else 0.0
        if k == "gemm":
            key = desc
            a = agg.setdefault(key, [0, 0.0, fl])
            a[0] += 1; a[1] += t
    print("%-52s %4s %9s %8s %10s" % ("GEMM shape (B=8, 32x32 latent)", "n", "total us", "avg us", "MMA TF/s"))
    for key, (n, t, fl) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-52s %4d %9.1f %8.1f %10.0f" % (key, n, t, t / n if n else 0, (fl * n / (t * 1e-6) / 1e12) if t else 0))


if __name__ == "__main__":
    main()
