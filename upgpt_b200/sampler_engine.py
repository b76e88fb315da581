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
    MAX_GRAPHS = 32

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

    def cfg_fusable(self, cond, ucond):
        """Classifier-free guidance runs as ONE 2B-batch step graph when both conditionings have the same fusable form."""
        key = self.model.model.conditioning_key
        a, b = self.split_cond(cond, key), self.split_cond(ucond, key)
        if a is None or b is None or a[0] is None or b[0] is None or a[0].shape != b[0].shape:
            return False
        return (a[1] is None) == (b[1] is None) and (a[1] is None or a[1].shape == b[1].shape)

    def _prepare(self, x_T, cond, ucond=None):
        """Stages latent / context / concat channels into the engine for this batch. With `ucond` (classifier-free guidance,
        ddim.py:171-178) the engine runs 2B samples per pass: rows [0, B) = unconditional, rows [B, 2B) = conditional."""
        cc, ct = self.split_cond(cond, self.model.model.conditioning_key)
        B, _, H, W = x_T.shape
        x_T = x_T.contiguous().float()
        if ucond is not None:
            uc, ut = self.split_cond(ucond, self.model.model.conditioning_key)
            cc = torch.cat([uc.float(), cc.float()], 0)
            ct = None if ct is None else torch.cat([ut.float(), ct.float()], 0)
            x_T = torch.cat([x_T, x_T], 0)
        eng = self.unet.engine(x_T.shape[0], H, W, cc.shape[1])
        eng.set_context(cc.contiguous().float())
        eng.stage_inputs(x_T, None, None if ct is None else ct.contiguous().float())
        sfx = "" if ucond is None else "_cfg"       # per-sample buffers of the guided half-batch have their own names
        eng.buf("s_step", (1,), torch.int32)
        eng.buf("s_ttable", (self.MAX_STEPS,), torch.int64)
        eng.buf("s_coef", (self.MAX_STEPS, 6), torch.float32)
        eng.buf("s_noise1" + sfx, (B,) + tuple(x_T.shape[1:]), torch.float32)
        eng.buf("s_pred_x0" + sfx, (B,) + tuple(x_T.shape[1:]), torch.float32)
        if ucond is not None:
            eng.buf("s_eps_cfg", (B,) + tuple(eng.bufs["eps"].shape[1:]), torch.float32)
        return eng

    def _emb_table(self, eng, t_loop):
        """[S][emb_total] table of the ResBlocks' timestep-embedding projections, one row per loop step (all samples of a batch
        share t: ddim.py:142, ddpm.py:1271). Computed with the U-Net program's own embedding launches, once per (schedule, weights)."""
        t_key = np.ascontiguousarray(t_loop, dtype=np.int64).tobytes()
        key = (id(eng), eng.weights_version, t_key)
        tab = self._emb_tables.get(key)
        if tab is None:
            S = len(t_loop)
            # same (engine, schedule) at a new weights version: refill the old table in place, so its address -- baked into the
            # captured step graph -- stays valid
            stale = [k for k in self._emb_tables if k[0] == key[0] and k[2] == key[2]]
            tab = self._emb_tables.pop(stale[0]) if stale else None
            if tab is None:
                if len(self._emb_tables) > 8:
                    self._emb_tables.clear()
                tab = torch.empty(S, eng.emb_total, device=eng.dev, dtype=torch.float32)
            for i, t in enumerate(np.asarray(t_loop).tolist()):
                eng.bufs["t_in"].fill_(int(t))
                eng.run_calls(0, eng.n_emb_calls)
                tab[i].copy_(eng.bufs["emb_all"][0])
            self._emb_tables[key] = tab
        return tab

    def _step_graph(self, eng, kind, noise_mode, noise_buf, emb_tab, cfg_scale=None, blend=None):
        """kind: 'ddim' | 'ddpm'; noise_mode: 0 none, 1 per-step buffer refreshed by the host loop, 2 strided table.
        cfg_scale: classifier-free guidance -- the engine batch is [unconditional | conditional]; after the U-Net pass
        eps[:B] <- s*eps_c + (1-s)*eps_u, the update runs on the first half of the latent and is mirrored into the second."""
        # weights are re-packed IN PLACE (upgpt_b200/host.py), so a captured graph stays valid across weight versions: the key holds
        # addresses only (the timestep-embedding table, whose VALUES depend on the weights, is keyed on the version in _emb_table)
        key = (id(eng), kind, noise_mode, 0 if noise_buf is None else noise_buf.data_ptr(), emb_tab.data_ptr(), cfg_scale,
               None if blend is None else tuple(t.data_ptr() for t in blend))
        g = self._graphs.get(key)
        if g is not None:
            return g
        L = _C.lib()
        b = eng.bufs
        x = b["x_lat"]
        cfg = cfg_scale is not None
        n = x.numel() // 2 if cfg else x.numel()     # elements of the B samples the update kernel advances
        pred_x0 = b["s_pred_x0_cfg" if cfg else "s_pred_x0"]
        eps = b["eps"]
        assert x.is_contiguous() and eps.is_contiguous() and eps.numel() == x.numel()
        step_fn = L.upgpt_ddim_step if kind == "ddim" else L.upgpt_ddpm_step
        ncols = 5 if kind == "ddim" else 6
        coef = b["s_coef"]
        assert coef.is_contiguous()
        noise_ptr = 0 if noise_mode == 0 else noise_buf.data_ptr()
        stride = n if noise_mode == 2 else 0

        def body():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if blend is not None:
                # known-region blend in front of the step (ddim.py:144-147): x <- q_sample(x0, t) * mask + (1 - mask) * x, with t and this
                # step's q_sample noise found through the device-side step counter
                bx0, bnoise, bmask, sa, s1m = blend
                Bn, Cn = x.shape[0], x.shape[1]
                _C.check(L.upgpt_qsample_blend(bx0.data_ptr(), bnoise.data_ptr(), n, bmask.data_ptr(), bmask.shape[1], x.data_ptr(), x.data_ptr(),
                                               sa.data_ptr(), s1m.data_ptr(), 0, b["s_ttable"].data_ptr(), b["s_step"].data_ptr(), 0, Bn, Cn,
                                               x.numel() // (Bn * Cn), s), "qsample_blend")
            # this step's timestep-embedding rows from the per-schedule table (replaces 4 launches + 43 MB of fp32 weights per step)
            _C.check(L.upgpt_gather_step_row(emb_tab.data_ptr(), emb_tab.shape[1], b["s_step"].data_ptr(), b["emb_all"].data_ptr(),
                                             x.shape[0], emb_tab.shape[1], s), "gather_step_row")
            eng.run_calls(eng.n_emb_calls, None, s)
            e_ptr = eps.data_ptr()
            if cfg:   # e = s e_c + (1 - s) e_u  (= e_u + s (e_c - e_u), the same expression as the general loop)
                e_ptr = b["s_eps_cfg"].data_ptr()
                _C.check(L.upgpt_axpby(eps.data_ptr() + 4 * n, float(cfg_scale), eps.data_ptr(), 1.0 - float(cfg_scale), e_ptr, n, s),
                         "cfg_combine")
            _C.check(step_fn(x.data_ptr(), e_ptr, noise_ptr, stride, coef.data_ptr(), b["s_step"].data_ptr(), 0,
                             x.data_ptr(), pred_x0.data_ptr(), n, s), "sampler_step")
            if cfg:   # both halves of the batch carry the same latent
                _C.check(L.upgpt_axpby(x.data_ptr(), 1.0, 0, 0.0, x.data_ptr() + 4 * n, n, s), "cfg_mirror")
            _C.check(L.upgpt_step_state(b["s_step"].data_ptr(), 1, 1, 0, 0, 0, s), "step_state")

        self._coef_cols = ncols
        body()   # warm-up outside capture (also validates arguments)
        g = ops.Graph().capture(body)
        if len(self._graphs) >= self.MAX_GRAPHS:      # a graph bakes the noise / table addresses: bound the cache for long-running callers
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = g
        return g

    def _loop(self, eng, kind, S, t_loop, coef_loop, x_noise, log_idx, callback, img_callback, intermediates, cfg_scale=None, blend=None):
        b = eng.bufs
        cfg = cfg_scale is not None
        nB = b["x_lat"].shape[0] // 2 if cfg else b["x_lat"].shape[0]
        noise1, pred_x0 = b["s_noise1_cfg" if cfg else "s_noise1"], b["s_pred_x0_cfg" if cfg else "s_pred_x0"]
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
                n_lat = b["x_lat"].numel() // (2 if cfg else 1)
                if x_noise.numel() < S * n_lat or tuple(x_noise.shape[1:]) != (nB,) + tuple(b["x_lat"].shape[1:]):
                    raise ValueError("x_noise must be (S, B, C, H, W) = %s with S >= %d, got %s (the step kernel indexes noise + step * B*C*H*W)"
                                     % ((S, nB) + tuple(b["x_lat"].shape[1:]), S, tuple(x_noise.shape)))
                noise_mode, noise_buf = 2, x_noise.contiguous().float()
            else:   # True -> draw on the fly into a single-step buffer
                noise_mode, noise_buf = 1, noise1
        x_saved = b["x_lat"].clone()
        emb_tab = self._emb_table(eng, t_loop)
        ops.step_state(b["s_step"], 0, 0)
        if blend is not None:
            bx0, bnoise, bmask = blend
            shp = tuple(b["x_lat"].shape)
            if tuple(bnoise.shape) != (S,) + shp or tuple(bx0.shape) != shp or bmask.shape[1] not in (1, shp[1]):
                raise ValueError("mask / x0 blend: x0 %s, x0_noise (S,)+%s and mask (B, 1|C, H, W) expected, got %s / %s / %s"
                                 % (shp, shp, tuple(bx0.shape), tuple(bnoise.shape), tuple(bmask.shape)))
            m = self.model
            blend = (bx0.contiguous().float(), bnoise.contiguous().float(), bmask.expand(shp[0], *bmask.shape[1:]).contiguous().float(),
                     m.sqrt_alphas_cumprod.contiguous(), m.sqrt_one_minus_alphas_cumprod.contiguous())
            self._blend_ref = blend       # the graph bakes these addresses: keep the tensors alive with the sampler
        g = self._step_graph(eng, kind, noise_mode, noise_buf, emb_tab, cfg_scale, blend)   # warm-up run inside mutates x_lat / step: restore
        b["x_lat"].copy_(x_saved)
        ops.step_state(b["s_step"], 0, 0)
        for i in range(S):
            if noise_mode == 1:
                noise1.normal_()
            g.launch()
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0.clone(), i)
            if i in log_idx:
                intermediates["x_inter"].append(b["x_lat"][:nB].clone())
                intermediates["pred_x0"].append(pred_x0.clone())
        return b["x_lat"][:nB].clone(), intermediates

    # ---- DDIM (ddim.py:114-163) ----
    def run_ddim(self, x_T, cond, time_range, coef_by_index, x_noise, log_every_t, callback, img_callback, intermediates,
                 ucond=None, cfg_scale=None, mask=None, x0=None, x0_noise=None):
        eng = self._prepare(x_T, cond, ucond)
        S = len(time_range)
        coef_loop = coef_by_index[:S].flip(0).contiguous()     # loop i uses index S-1-i
        log_idx = {i for i in range(S) if (S - i - 1) % log_every_t == 0 or (S - i - 1) == S - 1}
        blend = None
        if mask is not None:
            assert ucond is None, "the known-region blend with guidance runs on the general loop"
            blend = (x0, x0_noise, mask)
        return self._loop(eng, "ddim", S, np.asarray(time_range), coef_loop, x_noise, log_idx, callback, img_callback, intermediates,
                          cfg_scale if ucond is not None else None, blend)

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
