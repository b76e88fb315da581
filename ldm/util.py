"""Config-instantiation protocol of the reference (ldm/util.py:78-93) -- the drop-in boundary of this package.

`instantiate_from_config({"target": "pkg.mod.Class", "params": {...}})` imports the dotted path and calls it with the
params.  OmegaConf is optional: `load_config` parses the reference's YAML files with PyYAML into attribute-access dicts
that behave like the subset of DictConfig the hot path touches (`cfg.model.params.unet_config`, `.get`, `in`, iteration).
"""
import importlib


class AttrDict(dict):
    """dict with attribute access, recursively applied (a minimal stand-in for omegaconf.DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return [AttrDict.wrap(v) for v in obj]
        return obj


def load_config(path):
    """Reads a reference YAML config (e.g. configs/deepfashion/bbox.yaml) unchanged."""
    try:
        from omegaconf import OmegaConf  # used when available, exactly like the reference's callers
        if hasattr(OmegaConf, "load"):
            return OmegaConf.load(path)
    except Exception:
        pass
    import yaml
    with open(path) as fh:
        return AttrDict.wrap(yaml.safe_load(fh))


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    mod = importlib.import_module(module)
    if reload:
        mod = importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", None)
    return get_obj_from_str(config["target"])(**(dict(params) if params is not None else {}))


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) else d


def count_params(model, verbose=False):
    total = sum(p.numel() for p in model.parameters())
    if verbose:
        print(f"{model.__class__.__name__} has {total * 1.e-6:.2f} M params.")
    return total


def ismap(x):
    import torch
    return isinstance(x, torch.Tensor) and x.dim() == 4 and x.shape[1] > 3


def isimage(x):
    import torch
    return isinstance(x, torch.Tensor) and x.dim() == 4 and x.shape[1] in (1, 3)
