"""Host-side bookkeeping shared by the nn.Modules that hand their arithmetic to an engine (UNetModel, AutoencoderKL, the CLIP towers).

* `WeightStore`: ONE packed copy of a module's weights per (operand format, weights tag), shared by every engine of the module.
  Engines differ in batch / spatial size / context length, i.e. in their activation buffers and recorded programs -- not in their
  weights -- so a second batch size (a partial last batch, classifier-free guidance's 2B engine, an app with variable N) costs
  activations only. Tensors are updated in place on re-pack: their addresses, and therefore recorded programs and captured CUDA
  graphs, stay valid across `load_state_dict`.
* weights tag: `ema_scope` (ddpm.py:179-192) swaps EMA weights into the module for the duration of a sampling call. The packed copy
  of the EMA weights lives under its own tag ("ema") beside the training weights ("raw"), so entering / leaving the scope switches
  between two resident packed sets instead of re-packing 425 M parameters (and re-capturing the step graph) twice per request.
* engine cache: LRU-bounded (UPGPT_MAX_ENGINES, default 8) -- a long-running service with many distinct shapes cannot grow without bound.
"""
import os
from collections import OrderedDict

import torch


class WeightStore:
    def __init__(self):
        self.tensors = {}     # (name, tag, shape, dtype) -> device tensor
        self.plans = {}       # (plan signature, tag) -> {"version": v, "w": {name: tensor}}

    def put(self, name, tag, t, dev):
        t = t.contiguous()
        key = (name, tag, tuple(t.shape), t.dtype)
        cur = self.tensors.get(key)
        if cur is not None and cur.device == torch.device(dev):
            cur.copy_(t)      # keep the address: recorded programs / captured graphs stay valid across re-packs
            return cur
        self.tensors[key] = t.to(dev).clone() if t.device == torch.device(dev) else t.to(dev)
        return self.tensors[key]

    def clear(self):
        self.tensors.clear()
        self.plans.clear()

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors.values())


class EngineHostMixin:
    """Mixed into an nn.Module: `_host_init()` in __init__, `_engine_get(key, make)` to fetch / build an engine."""

    def _host_init(self):
        self._engines = OrderedDict()
        self._weights_version = 0
        self._weights_tag = "raw"
        self._wstore = WeightStore()

    def mark_weights_changed(self):
        """The engines keep packed fp16 copies of the weights; they re-pack (in place) when this version moves."""
        self._weights_version += 1

    def use_weights_tag(self, tag):
        """Selects which resident packed weight set following calls run on ("raw" | "ema"); see the module docstring."""
        self._weights_tag = tag

    def _host_reset(self):
        self._engines = OrderedDict()
        self._wstore.clear()
        self.mark_weights_changed()

    def _engine_get(self, key, make):
        from . import lanes
        key = tuple(key) + (self._weights_tag, lanes.current())      # engines (buffers, programs, graphs) exist per lane; weights are shared
        eng = self._engines.get(key)
        if eng is None:
            cap = max(1, int(os.environ.get("UPGPT_MAX_ENGINES", "8")))
            while len(self._engines) >= cap:
                self._engines.popitem(last=False)          # least recently used: its activation buffers / graphs are freed
            eng = make()
        self._engines[key] = eng
        self._engines.move_to_end(key)
        if eng.weights_version != self._weights_version:
            eng.pack_weights(self)
            # upgpt_gemm(UPGPT_GEMM_F_W_STATIC) streams weights ahead of its stream dependency: the (rare) re-pack must have landed
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        return eng
