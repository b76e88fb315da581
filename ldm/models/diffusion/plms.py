"""PLMSSampler (reference ldm/models/diffusion/plms.py) on the B200 path -- SURVEY.md 8(f) rank 3.

Same constructor / `sample` surface as the reference. eps comes from `LatentDiffusion.apply_model` (the U-Net engine's graph
replay); the pseudo linear multistep combinations of the eps history run in one kernel (upgpt_lincomb4) and the x_prev / pred_x0
formulas in the DDIM update kernel (sigma = 0: PLMS requires eta = 0, plms.py:25-26). No torch arithmetic on the latents."""
import numpy as np
import torch

from ldm.models.diffusion.ddim import DDIMSampler


class PLMSSampler(DDIMSampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        return super().make_schedule(ddim_num_steps, ddim_discretize=ddim_discretize, ddim_eta=0., verbose=verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.plms_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None):
        """plms.py:114-170."""
        from upgpt_b200 import ops
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        img = img.contiguous().clone()
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = list(reversed(range(0, timesteps))) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        coef = self._coef_rows(ddim_use_original_steps, temperature)
        old_eps = []
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            ts_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                blended = torch.empty_like(img)
                ops.mask_blend(img_orig, img, mask, blended) if hasattr(ops, "mask_blend") else blended.copy_(img_orig * mask + (1. - mask) * img)
                img = blended
            img, pred_x0, e_t = self.p_sample_plms(img, cond, ts, index=index, use_original_steps=ddim_use_original_steps,
                                                   quantize_denoised=quantize_denoised, temperature=temperature,
                                                   noise_dropout=noise_dropout, score_corrector=score_corrector,
                                                   corrector_kwargs=corrector_kwargs,
                                                   unconditional_guidance_scale=unconditional_guidance_scale,
                                                   unconditional_conditioning=unconditional_conditioning, old_eps=old_eps,
                                                   t_next=ts_next, _coef=coef)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_plms(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, old_eps=None, t_next=None, _coef=None):
        """plms.py:172-236."""
        from upgpt_b200 import _C, ops
        if quantize_denoised:
            raise NotImplementedError("quantize_denoised needs a VQ first stage (not used by UPGPT's KL-f8 configs)")
        coef = _coef if _coef is not None else self._coef_rows(use_original_steps, temperature)
        assert float(coef[index, 2]) == 0., "PLMS runs with eta = 0"
        x = x.contiguous().float()

        def get_model_output(xx, tt):
            if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
                e = self.model.apply_model(xx, tt, c)
            else:   # classifier-free guidance, dict-cond aware (two U-Net passes; plms.py:180-185)
                e_u = self.model.apply_model(xx, tt, unconditional_conditioning)
                e_c = self.model.apply_model(xx, tt, c)
                e = torch.empty_like(e_c)
                ops.axpby(e_c, unconditional_guidance_scale, e_u, 1. - unconditional_guidance_scale, e)
            if score_corrector is not None:
                assert self.model.parameterization == "eps"
                e = score_corrector.modify_score(self.model, e, xx, tt, c, **corrector_kwargs)
            return e.contiguous().float()

        def get_x_prev_and_pred_x0(e, idx):
            x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
            ops.ddim_step(x, e, coef, x_prev, pred_x0, noise=None, step_imm=int(idx))
            return x_prev, pred_x0

        def lincomb(terms, den):
            out = torch.empty_like(x)
            ptrs = [(tt.data_ptr(), float(w)) for tt, w in terms] + [(0, 0.)] * (4 - len(terms))
            args = []
            for pp, w in ptrs:
                args += [pp, w]
            _C.check(_C.lib().upgpt_lincomb4(*args, float(den), out.data_ptr(), out.numel(), ops.stream()), "upgpt_lincomb4")
            return out

        e_t = get_model_output(x, t)
        if len(old_eps) == 0:       # pseudo improved Euler (2nd order)
            x_prev, pred_x0 = get_x_prev_and_pred_x0(e_t, index)
            e_t_next = get_model_output(x_prev, t_next)
            e_t_prime = lincomb([(e_t, 1.), (e_t_next, 1.)], 2.)
        elif len(old_eps) == 1:     # 2nd order Adams-Bashforth
            e_t_prime = lincomb([(e_t, 3.), (old_eps[-1], -1.)], 2.)
        elif len(old_eps) == 2:     # 3rd order
            e_t_prime = lincomb([(e_t, 23.), (old_eps[-1], -16.), (old_eps[-2], 5.)], 12.)
        else:                       # 4th order
            e_t_prime = lincomb([(e_t, 55.), (old_eps[-1], -59.), (old_eps[-2], 37.), (old_eps[-3], -9.)], 24.)
        x_prev, pred_x0 = get_x_prev_and_pred_x0(e_t_prime, index)
        return x_prev, pred_x0, e_t
