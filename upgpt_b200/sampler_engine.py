"""Fused sampling loops: one captured CUDA graph per denoising step, replayed S times.

Step graph = [gather this step's timestep-embedding row from a per-schedule table] -> [U-Net program] -> [DDIM / DDPM update in
place on the staged latent] -> [advance the device-side step counter].  Schedule coefficients are a device table indexed by the
counter, so replays need no host data (reference per-step host work: ddim.py:142,189-192; ddpm.py:1157-1185).
"""
import ctypes as C

import numpy as np
import torch

from . import _C, ops


class FusedSampler:
    MAX_STEPS = 1024

    def __init__(self, ldm_model):
        self.model = ldm_model
        self.unet = ldm_model.model.diffusion_model
        self._graphs = {}
        self._emb_tables = {}

    # ---- conditioning normalisation (DiffusionWrapper.forward routing, ddpm.py:1557-1577) ----
    @staticmethod
    def split_cond(cond, conditioning_key):
        """-> (context (B,L,D) or None, c_concat (B,Cc,H,W) or None); None if the form is not fusable."""
        if cond is None:
            return None
        if not isinstance(cond, dict):
            cond = {("c_concat" if conditioning_key == "concat" else "c_crossattn"): cond if isinstance(cond, list) else [cond]}
        cc = cond.get("c_crossattn")
        ct = cond.get("c_concat")
        if isinstance(cc, (list, tuple)):
            cc = cc[0] if len(cc) == 1 else torch.cat(list(cc), 1)
        if isinstance(ct, (list, tuple)):
            ct = [c for c in ct if c is not None]
            ct = None if not ct else (ct[0] if len(ct) == 1 else torch.cat(ct, 1))
        if conditioning_key in ("crossattn", "hybrid") and cc is None:
            return None
        if conditioning_key in ("concat", "hybrid") and ct is None:
            return None
        if conditioning_key == "crossattn":
            ct = None
        if conditioning_key == "concat":
            return None   # concat-only U-Nets have no context: not a UPGPT configuration
        return cc, ct

    def _prepare(self, x_T, cond):
        cc, ct = self.split_cond(cond, self.model.model.conditioning_key)
        B, _, H, W = x_T.shape
        eng = self.unet.engine(B, H, W, cc.shape[1])
        eng.set_context(cc.contiguous().float())
        eng.stage_inputs(x_T.contiguous().float(), None, None if ct is None else ct.contiguous().float())
        dev = x_T.device
        eng.buf("s_step", (1,), torch.int32)
        eng.buf("s_ttable", (self.MAX_STEPS,), torch.int64)
        eng.buf("s_coef", (self.MAX_STEPS, 6), torch.float32)
        eng.buf("s_noise1", tuple(x_T.shape), torch.float32)
        eng.buf("s_pred_x0", tuple(x_T.shape), torch.float32)
        return eng

    def _emb_table(self, eng, t_loop):
        """[S][emb_total] table of the ResBlocks' timestep-embedding projections, one row per loop step (all samples of a batch
        share t: ddim.py:142, ddpm.py:1271). Computed with the U-Net program's own embedding launches, once per (schedule, weights)."""
        t_key = np.ascontiguousarray(t_loop, dtype=np.int64).tobytes()
        key = (id(eng), eng.weights_version, t_key)
        tab = self._emb_tables.get(key)
        if tab is None:
            if len(self._emb_tables) > 8:
                self._emb_tables.clear()
            S = len(t_loop)
            tab = torch.empty(S, eng.emb_total, device=eng.dev, dtype=torch.float32)
            for i, t in enumerate(np.asarray(t_loop).tolist()):
                eng.bufs["t_in"].fill_(int(t))
                eng.run_calls(0, eng.n_emb_calls)
                tab[i].copy_(eng.bufs["emb_all"][0])
            self._emb_tables[key] = tab
        return tab

    def _step_graph(self, eng, kind, noise_mode, noise_buf, emb_tab):
        """kind: 'ddim' | 'ddpm'; noise_mode: 0 none, 1 per-step buffer refreshed by the host loop, 2 strided table."""
        key = (id(eng), kind, noise_mode, 0 if noise_buf is None else noise_buf.data_ptr(), eng.weights_version, emb_tab.data_ptr())
        g = self._graphs.get(key)
        if g is not None:
            return g
        L = _C.lib()
        b = eng.bufs
        x = b["x_lat"]
        n = x.numel()
        step_fn = L.upgpt_ddim_step if kind == "ddim" else L.upgpt_ddpm_step
        ncols = 5 if kind == "ddim" else 6
        coef = b["s_coef"]
        assert coef.is_contiguous()
        noise_ptr = 0 if noise_mode == 0 else noise_buf.data_ptr()
        stride = n if noise_mode == 2 else 0

        def body():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            # this step's timestep-embedding rows from the per-schedule table (replaces 4 launches + 43 MB of fp32 weights per step)
            _C.check(L.upgpt_gather_step_row(emb_tab.data_ptr(), emb_tab.shape[1], b["s_step"].data_ptr(), b["emb_all"].data_ptr(),
                                             x.shape[0], emb_tab.shape[1], s), "gather_step_row")
            eng.run_calls(eng.n_emb_calls, None, s)
            _C.check(step_fn(x.data_ptr(), b["eps"].data_ptr(), noise_ptr, stride, coef.data_ptr(), b["s_step"].data_ptr(), 0,
                             x.data_ptr(), b["s_pred_x0"].data_ptr(), n, s), "sampler_step")
            _C.check(L.upgpt_step_state(b["s_step"].data_ptr(), 1, 1, 0, 0, 0, s), "step_state")

        self._coef_cols = ncols
        body()   # warm-up outside capture (also validates arguments)
        g = ops.Graph().capture(body)
        self._graphs[key] = g
        return g

    def _loop(self, eng, kind, S, t_loop, coef_loop, x_noise, log_idx, callback, img_callback, intermediates):
        b = eng.bufs
        assert S <= self.MAX_STEPS
        dev = b["x_lat"].device
        b["s_ttable"][:S].copy_(torch.as_tensor(np.ascontiguousarray(t_loop), dtype=torch.int64))
        ncols = coef_loop.shape[1]
        # the coefficient table is addressed with a row stride of `ncols`: keep it dense at the front of the buffer
        flat = b["s_coef"].reshape(-1)
        flat[:S * ncols].copy_(coef_loop.reshape(-1).to(dev))
        noise_mode, noise_buf = 0, None
        if x_noise is not None:
            if isinstance(x_noise, torch.Tensor):
                noise_mode, noise_buf = 2, x_noise.contiguous().float()
            else:   # True -> draw on the fly into a single-step buffer
                noise_mode, noise_buf = 1, b["s_noise1"]
        x_saved = b["x_lat"].clone()
        emb_tab = self._emb_table(eng, t_loop)
        ops.step_state(b["s_step"], 0, 0)
        g = self._step_graph(eng, kind, noise_mode, noise_buf, emb_tab)   # warm-up run inside mutates x_lat / step: restore
        b["x_lat"].copy_(x_saved)
        ops.step_state(b["s_step"], 0, 0)
        for i in range(S):
            if noise_mode == 1:
                b["s_noise1"].normal_()
            g.launch()
            if callback:
                callback(i)
            if img_callback:
                img_callback(b["s_pred_x0"].clone(), i)
            if i in log_idx:
                intermediates["x_inter"].append(b["x_lat"].clone())
                intermediates["pred_x0"].append(b["s_pred_x0"].clone())
        return b["x_lat"].clone(), intermediates

    # ---- DDIM (ddim.py:114-163) ----
    def run_ddim(self, x_T, cond, time_range, coef_by_index, x_noise, log_every_t, callback, img_callback, intermediates):
        eng = self._prepare(x_T, cond)
        S = len(time_range)
        coef_loop = coef_by_index[:S].flip(0).contiguous()     # loop i uses index S-1-i
        log_idx = {i for i in range(S) if (S - i - 1) % log_every_t == 0 or (S - i - 1) == S - 1}
        return self._loop(eng, "ddim", S, np.asarray(time_range), coef_loop, x_noise, log_idx, callback, img_callback, intermediates)

    # ---- DDPM ancestral (ddpm.py:1244-1292) ----
    def run_ddpm(self, x_T, cond, timesteps, coef_by_t, x_noise, log_every_t, callback, img_callback, intermediates):
        eng = self._prepare(x_T, cond)
        t_loop = np.arange(timesteps)[::-1]
        coef_loop = coef_by_t[torch.as_tensor(t_loop.copy())].contiguous()
        log_idx = {i for i, t in enumerate(t_loop) if t % log_every_t == 0 or t == timesteps - 1}
        return self._loop(eng, "ddpm", int(timesteps), t_loop, coef_loop, True if x_noise is None else x_noise, log_idx,
                          callback, img_callback, intermediates)

    @property
    def launches_per_step(self):
        eng = next(iter(self.unet._engines.values()))
        return eng.launches_per_step - eng.n_emb_calls + 3     # gather + U-Net body + update + step++
