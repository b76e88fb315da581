"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference modules (soon-yau/upgpt @ /root/reference).

Used by oracle/make_golden.py (build container) to (1) validate the CPU restatement in oracle/ldm_oracle.py and (2) generate the
committed golden vectors under tests/golden/, and by `bench.py --impl reference` to time the reference's own code on the host cores
(on the GPU box from the unmodified copy oracle/vendor_reference.py leaves under the git-ignored baseline/_ref/).  Never imported by
the product package.

Shims:
  * omegaconf is absent -> stub `omegaconf.listconfig.ListConfig` (openaimodel.py:476 imports it)
  * DDIMSampler.register_buffer hard-codes torch.device("cuda") (ddim.py:19-23) -> plain setattr
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_VENDORED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # oracle/vendor_reference.py (git-ignored, travels with gpurun)
REF_ROOT = os.environ.get("UPGPT_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/ldm") else _VENDORED)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ldm"))


class _RefImport:
    """Context manager that makes `import ldm...` resolve to the reference tree (and only inside)."""

    def __enter__(self):
        self._saved = {k: v for k, v in sys.modules.items() if k == "ldm" or k.startswith("ldm.")}
        for k in self._saved:
            del sys.modules[k]
        sys.path.insert(0, REF_ROOT)
        # the reference's `ldm` is a namespace package (no __init__.py) and would lose to this repo's regular `ldm`
        # package: pin a synthetic top-level package whose search path is the reference tree.
        import importlib.machinery
        import importlib.util
        spec = importlib.machinery.ModuleSpec("ldm", None, is_package=True)
        spec.submodule_search_locations = [os.path.join(REF_ROOT, "ldm")]
        sys.modules["ldm"] = importlib.util.module_from_spec(spec)
        if "omegaconf" not in sys.modules:
            oc, lc = types.ModuleType("omegaconf"), types.ModuleType("omegaconf.listconfig")
            lc.ListConfig = type("ListConfig", (list,), {})
            oc.listconfig = lc
            sys.modules.update({"omegaconf": oc, "omegaconf.listconfig": lc})
            self._stub = True
        else:
            self._stub = False
        return self

    def __exit__(self, *exc):
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "ldm" or k.startswith("ldm.")]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        return False  # the omegaconf stub stays: UNetModel.__init__ imports it lazily (openaimodel.py:476)


def load_reference():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present (expected at %s)" % REF_ROOT)
    with _RefImport():
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        from ldm.modules.diffusionmodules.model import Decoder, Encoder
        from ldm.models.diffusion.ddim import DDIMSampler
        from ldm.models.diffusion.plms import PLMSSampler
        from ldm.modules.diffusionmodules import util as dutil
        from ldm.modules import attention as attn
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)
    PLMSSampler.register_buffer = lambda self, n, a: setattr(self, n, a)     # same hard-coded "cuda" (plms.py:18-22)
    ns = types.SimpleNamespace(UNetModel=UNetModel, Decoder=Decoder, Encoder=Encoder, DDIMSampler=DDIMSampler, PLMSSampler=PLMSSampler,
                               util=dutil, attention=attn)
    return ns


BBOX_UNET_KW = dict(image_size=32, in_channels=5, out_channels=4, model_channels=224,
                    attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                    num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                    use_checkpoint=False, legacy=False)
BBOX_VAE_KW = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                   ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)

# models/upgpt/upscale/config.yaml:37-55 (the second U-Net family UPGPT ships: 4x super-resolution in a KL-f4 latent; 6 = latent 3 +
# low-resolution concat 3 input channels, 86-token context = 77 text + 9 style, no SMPL token)
UPSCALE_UNET_KW = dict(image_size=32, in_channels=6, out_channels=3, model_channels=256, attention_resolutions=[2, 4, 8],
                       num_res_blocks=2, channel_mult=[1, 2, 2, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                       context_dim=768, use_checkpoint=True, legacy=False)
UPSCALE_VAE_KW = dict(double_z=True, z_channels=3, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4],
                      num_res_blocks=2, attn_resolutions=[], dropout=0.0)
