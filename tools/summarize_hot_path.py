"""Cuts an ncu launch list of tools/prof_hot_path.py (2 eager U-Net steps + 1 VAE decode) into its parts and prints per-kernel shares.
usage: python tools/summarize_hot_path.py <csv> [unet_launches_per_step=297] [vae_launches=95]"""
import collections, csv, re, sys
path = sys.argv[1]
n_unet = int(sys.argv[2]) if len(sys.argv) > 2 else 297
n_vae = int(sys.argv[3]) if len(sys.argv) > 3 else 95
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
d = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), float(r["Metric Value"]) / 1e3) for r in rows]
vae, unet = d[-n_vae:], d[-n_vae - n_unet:-n_vae]
print("ncu launch list %s: per-launch durations are cold-cache and serialised -- compare SHARES, not absolutes" % path)
for name, part in (("U-Net step (last of the eager steps)", unet), ("VAE decode of the batch", vae)):
    tot = sum(t for _, t in part)
    print("== %s: %d launches, sum %.1f us" % (name, len(part), tot))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, t in part:
        agg[k][0] += 1; agg[k][1] += t
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("  %-36s n=%4d %9.1f us %5.1f%%  avg %6.1f us" % (k[:36], n, t, 100 * t / tot, t / n))
