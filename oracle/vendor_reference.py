"""TEST / BASELINE INFRASTRUCTURE ONLY -- makes the reference's own hot-path modules available where /root/reference is not mounted.

`bench.py --impl reference` times the UNMODIFIED reference code (its UNetModel, DDIMSampler and VAE Decoder) on the host cores. The
GPU box has no /root/reference, so in the build container this script copies the seven reference files that path needs -- untouched --
into `baseline/_ref/` (git-ignored: the reference's sources never enter this repository's history; gpurun snapshots carry the
directory to the GPU box like the built .so files). `__graft_entry__.build()` runs it whenever /root/reference is present.

    python oracle/vendor_reference.py          # -> baseline/_ref/ldm/...
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("UPGPT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "ldm/util.py",
    "ldm/modules/attention.py",
    "ldm/modules/diffusionmodules/util.py",
    "ldm/modules/diffusionmodules/openaimodel.py",
    "ldm/modules/diffusionmodules/model.py",
    "ldm/models/diffusion/ddim.py",
    "ldm/models/diffusion/plms.py",
]


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "ldm")):
        return False
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
    with open(os.path.join(DST, "VENDORED_FROM"), "w") as fh:
        fh.write("soon-yau/upgpt, copied unmodified from %s by oracle/vendor_reference.py (git-ignored)\n" % SRC)
    if verbose:
        print("vendored %d reference files into %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
