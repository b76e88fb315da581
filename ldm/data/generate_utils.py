"""Inference facade of the reference (ldm/data/generate_utils.py:131-190) on the B200 engines: `InferenceModel(config, ckpt, device)`
with `create_batch`, `generate`, `mix_style` -- the caller of the hot path used by app.py and the notebooks -- plus the bbox-mask
interpolation helpers of the SMPL-interpolation flow (generate_utils.py:101-128; app.py:296-301).

Mirror of the class surface only: the data pipeline around it (datasets, segmentation, plotting) is out of scope.
`mix_style` works on the style tokens of the K | V cond-cache: masked slots and text overrides replace single context rows, which
`upgpt_b200.cond_cache.CondCache` refreshes row-wise instead of re-projecting all 87 rows.
"""
import numpy as np
import torch

from ldm.util import instantiate_from_config

style_names = ['face', 'hair', 'headwear', 'background', 'top', 'outer', 'bottom', 'shoes', 'accesories']


def load_model_from_config(config, ckpt, verbose=False):
    """generate_utils.py:32-48; ckpt None keeps the constructor's weights (tests, benchmarks)."""
    cfg_model = config["model"] if isinstance(config, dict) else config.model
    model = instantiate_from_config(cfg_model)
    if ckpt is not None:
        print(f"Loading model from {ckpt}")
        pl_sd = torch.load(ckpt, map_location="cpu")
        if "global_step" in pl_sd:
            print(f"Global Step: {pl_sd['global_step']}")
        m, u = model.load_state_dict(pl_sd["state_dict"], strict=False)
        if verbose:
            print("missing keys:", m)
            print("unexpected keys:", u)
    model.eval()
    return model


def get_coord(batch_mask):
    """Bounding box (xmin, xmax, ymin, ymax) of the person mask (generate_utils.py:101-110)."""
    mask = batch_mask[0].detach().cpu().numpy().copy()
    mask[mask == -1] = 0
    x = np.nonzero(np.mean(mask, 1))[0]
    y = np.nonzero(np.mean(mask, 0))[0]
    return np.array([x[0], x[-1], y[0], y[-1]])


def get_mask(mask, coord):
    """-1 outside / -0.99215686 inside the box (generate_utils.py:112-118)."""
    xmin, xmax, ymin, ymax = coord
    new_mask = np.ones_like(mask.detach().cpu().numpy()) * (-1)
    new_mask[0, xmin:xmax + 1, ymin:ymax + 1] = -0.99215686
    return torch.tensor(new_mask).to(mask.device)


def interp_mask(src_mask, dst_mask, alpha):
    """Box-coordinate interpolation between two person masks (generate_utils.py:120-128)."""
    coord = (alpha * get_coord(src_mask) + (1 - alpha) * get_coord(dst_mask)).astype(np.int32)
    return get_mask(src_mask, coord)


class InferenceModel:
    """generate_utils.py:131-169. The config rewrite is the reference's: style_cond becomes DummyModel (style embeddings are computed once
    by `mix_style` and passed through), the first stage loads no separate checkpoint, cond_stage gets the device."""

    def __init__(self, config, ckpt, device, clip_text_encoder=None):
        self.device = device
        params = config['model']['params']
        style_cond_config = params['extra_cond_stages']['style_cond']
        style_cond_config['params'] = {'device': device}
        self.clip_image_encoder = instantiate_from_config(style_cond_config)       # FrozenClipImageEmbedder2 on the CLIP engine
        # pooled, projected CLIP text features for per-slot text overrides (encoders/modules.py:165-198): supplied by the caller
        # (pre-computed (n, 768) embeddings work without it)
        self.clip_text_encoder = clip_text_encoder
        params['extra_cond_stages']['style_cond']['target'] = 'ldm.modules.poses.poses.DummyModel'
        params['first_stage_config']['params']['ckpt_path'] = None
        params['cond_stage_config']['params'] = {'device': device}
        self.model = load_model_from_config(config, ckpt).to(device)

    def create_batch(self, batch, repeat=1):
        for k, v in batch.items():
            if type(v) == torch.Tensor:
                temp = batch[k].unsqueeze(0)
                repeat_list = [1] * len(temp.shape)
                repeat_list[0] = repeat
                batch[k] = temp.repeat(repeat_list).to(self.device)
            else:
                batch[k] = [batch[k]] * repeat
        return batch

    def generate(self, batch, steps=200, repeat=1, use_ema=True):
        with torch.no_grad():
            images = self.model.log_images(batch, ddim_steps=steps, use_ema=use_ema, unconditional_guidance_scale=3.,
                                           unconditional_guidance_label=[""])
        for k in images:
            images[k] = torch.clamp(images[k].detach(), -1., 1.).cpu().numpy().transpose(0, 2, 3, 1) * 0.5 + 0.5
        return images

    def mix_style(self, s, w, mask=[], empty_style=None):
        """s: the 9 style crops (9, 3, 224, 224) -- or their embeddings (9, 768); w: {style name: text prompt or (768,) embedding};
        mask: style names replaced by the empty style. Returns the (9, 768) style tokens (generate_utils.py:172-190)."""
        style2id = dict(zip(style_names, range(len(style_names))))
        s = s.clone()
        pre_embedded = s.dim() == 2
        for m in mask:
            if empty_style is None:
                raise ValueError("mix_style(mask=...) needs the empty style (the CLIP-preprocessed black crop, or its embedding)")
            s[style2id[m]] = empty_style
        with torch.no_grad():
            image_emb = s.to(self.device)[None] if pre_embedded else self.clip_image_encoder(s.unsqueeze(0).to(self.device))
            image_emb = image_emb.clone()
            for k, v in w.items():
                if isinstance(v, str):
                    if v == '':
                        continue
                    if self.clip_text_encoder is None:
                        raise NotImplementedError("text overrides need a pooled CLIP text encoder (clip_text_encoder=...) or pre-computed embeddings")
                    v = self.clip_text_encoder([[v]])[0, 0]
                image_emb[0, style2id[k]] = v.to(self.device)
        return image_emb.squeeze(0)

    def update_style_slots(self, slots, emb, batch_size, H, W, ctx_len=87):
        """Applies changed style tokens to the resident cond-cache of the (batch_size, H, W) engine: only the rows of `slots` are
        re-projected in the 16 cross-attention layers (CondCache.set_style_slot)."""
        eng = self.model.model.diffusion_model.engine(batch_size, H, W, ctx_len)
        for sl in slots:
            eng.cond.set_style_slot(sl, emb[sl])
        return eng
