"""K / V cond-cache of the cross-attention layers as a first-class object.

The context (B, L, ctx_dim) = [77 CLIP-text tokens | 9 style tokens | 1 SMPL token] (ddpm.py:733-739) is timestep-invariant, so the
to_k / to_v projections of every SpatialTransformer (attention.py:162-163,175-176) are computed once per context, not once per
denoising step.  UPGPT's own callers change the context a row at a time:

  * the interpolation flow (app.py:296-301) lerps the SMPL vector between two poses: consecutive keyframes differ in row 86 only;
  * `InferenceModel.mix_style` (generate_utils.py:172-190) replaces single style slots (rows 77..85) by a text embedding or an
    empty style.

`CondCache` keeps the last context resident, finds the rows that changed (on the device), and refreshes only those rows of every
layer's K | V cache: one upgpt_gemm per (layer, changed row) with M = B rows addressed through row strides (A row b = token `row` of
sample b, output row b = the same token's K | V row) -- 87x less projection work than a rebuild for a keyframe step.  A full rebuild
remains one k | v GEMM per layer.  The cache lives with the engine (hence with the model), so it is handed from one
`DDIMSampler` / `PLMSSampler` instance to the next and across requests of the same batch shape.
"""
import ctypes as C

import torch

from . import _C


class CondCache:
    MAX_ROW_UPDATES = 4      # more changed rows than this: one full k | v GEMM per layer is cheaper than 16 launches per row

    def __init__(self, eng, layers):
        """layers: [(name, packed k|v weight, cache buffer [B*L][2*HD] fp16, HD)] in program order."""
        self.eng = eng
        self.layers = layers
        self.B, self.L, self.D = eng.B, eng.ctx_len, eng.ctx_dim
        self.valid = False
        self.row_version = [0] * self.L
        self.stats = {"full_rebuilds": 0, "row_updates": 0, "unchanged": 0, "gemm_launches": 0}
        self._ref = None

    def invalidate(self):
        """Weights changed: the cached projections are stale whatever the context."""
        self.valid = False

    # ---- change detection ----
    def _changed_rows(self, context):
        cur = self.eng.bufs["ctx32"]
        diff = (context != cur).any(dim=2).any(dim=0)          # (L,) on the device; one small D2H read per set_context
        return torch.nonzero(diff).flatten().tolist()

    # ---- refresh ----
    def set_context(self, context, force=False):
        """Makes the cache hold the projections of `context` (B, L, ctx_dim) fp32. Returns the list of refreshed rows (None = all)."""
        eng = self.eng
        assert tuple(context.shape) == (self.B, self.L, self.D), (tuple(context.shape), (self.B, self.L, self.D))
        if not self.valid or force:
            rows = None
        else:
            rows = self._changed_rows(context)
            if not rows:
                self.stats["unchanged"] += 1
                return []
            if len(rows) > self.MAX_ROW_UPDATES:
                rows = None
        eng.bufs["ctx32"].copy_(context)
        stream = eng._stream()
        eng.ctx_prep.run(stream)                                # fp32 context -> fp16 operand planes (one launch)
        if rows is None:
            eng.ctx_prog.run(stream)
            self.stats["full_rebuilds"] += 1
            self.stats["gemm_launches"] += len(self.layers)
            self.row_version = [v + 1 for v in self.row_version]
        else:
            self.update_rows(rows, _staged=True)
        self.valid = True
        self._ref = context       # keep the tensor alive: callers key on it (see UNetEngine.set_context)
        return rows

    def update_rows(self, rows, context=None, _staged=False):
        """Recomputes K | V of the context rows `rows` (token indices) in every layer. `context`: new full context to take the rows from
        (default: what is staged in the engine)."""
        eng = self.eng
        stream = eng._stream()
        if context is not None:
            eng.bufs["ctx32"][:, rows] = context[:, rows]
        if not _staged:
            eng.ctx_prep.run(stream)
        kx = eng.kx
        ctx16 = eng.bufs["ctx16"]
        K = self.D
        for r in rows:
            assert 0 <= r < self.L
            for name, w, kvc, HD in self.layers:
                a = _C.GemmArgs()
                a.a = ctx16.data_ptr() + r * K * kx * 2
                a.lda = self.L * K * kx                       # row b of the operand = token r of sample b
                a.w = w.data_ptr()
                a.mode, a.M, a.N, a.K = _C.GEMM_PLAIN, self.B, 2 * HD, K
                a.out16 = kvc.data_ptr() + r * 2 * HD * 2
                a.ld16 = self.L * 2 * HD                      # output row b = the cache row of token r of sample b
                a.flags = eng.x3
                rc = eng.L.upgpt_gemm(C.byref(a), stream)
                if rc != 0:
                    raise _C.UpgptError("cond-cache row update failed (%d): %s" % (rc, eng.L.upgpt_last_error().decode()))
            self.row_version[r] += 1
        self.stats["row_updates"] += len(rows)
        self.stats["gemm_launches"] += len(rows) * len(self.layers)

    def set_style_slot(self, slot, emb, n_text=77):
        """`mix_style` (generate_utils.py:172-190): replaces style slot `slot` (context row n_text + slot) of every sample by `emb`
        ((ctx_dim,) or (B, ctx_dim)) and refreshes that row of the cache."""
        row = n_text + int(slot)
        assert self.valid and n_text <= row < self.L
        self.eng.bufs["ctx32"][:, row] = emb.to(self.eng.bufs["ctx32"].dtype)
        self.update_rows([row])
