"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain fp32 PyTorch functional ops) of the two CLIP ViT-L/14 conditioning towers the
reference runs once per request.  Only tests/ and the golden generator import this module.

The arithmetic of both towers lives in third-party dependencies of the reference, not in /root/reference itself:
  * text : `transformers.CLIPTextModel` ("openai/clip-vit-large-patch14"), called by FrozenCLIPEmbedder.forward
           (ldm/modules/encoders/modules.py:137-162); pinned `transformers==4.19.2` (environment.yaml), this image has 5.5 with
           the same CLIP text architecture.  State-dict keys: `text_model.embeddings.token_embedding.weight`, ...
  * image: `clip.load("ViT-L/14")` -> `clip.model.VisionTransformer` (git+https://github.com/openai/CLIP.git, un-pinned in
           environment.yaml; absent from this image), called by FrozenClipImageEmbedder2.forward (modules.py:234-256).  Its
           published algorithm is restated below with OpenAI's state-dict keys (`visual.conv1.weight`, `visual.class_embedding`,
           `visual.transformer.resblocks.N.attn.in_proj_weight`, ...).

Parity is PINNED against the `transformers` implementations executed in the build container (oracle/make_clip_golden.py:
CLIPTextModel for the text tower; CLIPVisionModelWithProjection -- the HF port of the same VisionTransformer -- for the image
tower through the key map `openai_to_hf_vision`), golden vectors in tests/golden/clip_golden.npz.
"""
import torch
import torch.nn.functional as F


def quick_gelu(x):
    """clip/model.py QuickGELU = transformers `quick_gelu`: x * sigmoid(1.702 x)."""
    return x * torch.sigmoid(1.702 * x)


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, heads, causal):
    """Multi-head self-attention, scale d^-1/2, optional causal mask (key j <= query i)."""
    B, N, C = x.shape
    d = C // heads
    split = lambda t: t.reshape(B, N, heads, d).permute(0, 2, 1, 3)
    q, k, v = split(F.linear(x, wq, bq)), split(F.linear(x, wk, bk)), split(F.linear(x, wv, bv))
    s = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    if causal:
        s = s + torch.full((N, N), float("-inf")).triu(1)
    o = torch.einsum("bhij,bhjd->bhid", s.softmax(-1), v).permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(o, wo, bo)


def clip_text_forward(sd, heads, ids):
    """transformers CLIPTextTransformer.forward -> last_hidden_state (what FrozenCLIPEmbedder returns, modules.py:154-159):
    token + position embeddings; per layer x += attn(ln1(x)) [causal], x += fc2(quick_gelu(fc1(ln2(x)))); final_layer_norm.
    sd keys as in CLIPTextModel.state_dict(); ids (B, L) int64."""
    p = "text_model."
    L = ids.shape[1]
    x = sd[p + "embeddings.token_embedding.weight"][ids] + sd[p + "embeddings.position_embedding.weight"][:L][None]
    i = 0
    while f"{p}encoder.layers.{i}.layer_norm1.weight" in sd:
        q = f"{p}encoder.layers.{i}"
        a = q + ".self_attn."
        x = x + _mha(_ln(x, sd, q + ".layer_norm1"), sd[a + "q_proj.weight"], sd[a + "q_proj.bias"], sd[a + "k_proj.weight"], sd[a + "k_proj.bias"],
                     sd[a + "v_proj.weight"], sd[a + "v_proj.bias"], sd[a + "out_proj.weight"], sd[a + "out_proj.bias"], heads, True)
        h = F.linear(_ln(x, sd, q + ".layer_norm2"), sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc1.bias"])
        x = x + F.linear(quick_gelu(h), sd[q + ".mlp.fc2.weight"], sd[q + ".mlp.fc2.bias"])
        i += 1
    return _ln(x, sd, p + "final_layer_norm")


def clip_vision_forward(sd, heads, images, prefix="visual."):
    """clip.model.VisionTransformer.forward (OpenAI CLIP): conv1 (patch, stride = patch, no bias) -> [class_embedding ; patches] +
    positional_embedding -> ln_pre -> resblocks (x += attn(ln_1(x)); x += c_proj(QuickGELU(c_fc(ln_2(x))))) -> ln_post(x[:, 0]) @ proj.
    images (n, 3, S, S) fp32 already normalised; returns (n, output_dim).  (clip.load on CUDA holds fp16 weights; the oracle is fp32.)"""
    p = prefix
    w = sd[p + "conv1.weight"]
    x = F.conv2d(images, w, None, stride=w.shape[-1])
    n, C = x.shape[0], x.shape[1]
    x = x.reshape(n, C, -1).permute(0, 2, 1)
    x = torch.cat([sd[p + "class_embedding"].expand(n, 1, C), x], 1) + sd[p + "positional_embedding"][None]
    x = _ln(x, sd, p + "ln_pre")
    i = 0
    while f"{p}transformer.resblocks.{i}.ln_1.weight" in sd:
        q = f"{p}transformer.resblocks.{i}"
        wi, bi = sd[q + ".attn.in_proj_weight"], sd[q + ".attn.in_proj_bias"]
        x = x + _mha(_ln(x, sd, q + ".ln_1"), wi[:C], bi[:C], wi[C:2 * C], bi[C:2 * C], wi[2 * C:], bi[2 * C:],
                     sd[q + ".attn.out_proj.weight"], sd[q + ".attn.out_proj.bias"], heads, False)
        h = F.linear(_ln(x, sd, q + ".ln_2"), sd[q + ".mlp.c_fc.weight"], sd[q + ".mlp.c_fc.bias"])
        x = x + F.linear(quick_gelu(h), sd[q + ".mlp.c_proj.weight"], sd[q + ".mlp.c_proj.bias"])
        i += 1
    return _ln(x[:, 0], sd, p + "ln_post") @ sd[p + "proj"]


def style_embed(sd, heads, crops, prefix="visual."):
    """FrozenClipImageEmbedder2.forward (modules.py:250-255): (b, n, c, h, w) -> encode_image((b n) c h w) -> (b, n, width)."""
    b, n = crops.shape[:2]
    return clip_vision_forward(sd, heads, crops.reshape((b * n,) + tuple(crops.shape[2:])), prefix).reshape(b, n, -1)


def openai_to_hf_vision(sd, prefix="visual."):
    """OpenAI CLIP visual state dict -> transformers CLIPVisionModelWithProjection state dict (same architecture, other names)."""
    p, out = prefix, {}
    v = "vision_model."
    out[v + "embeddings.class_embedding"] = sd[p + "class_embedding"]
    out[v + "embeddings.patch_embedding.weight"] = sd[p + "conv1.weight"]
    out[v + "embeddings.position_embedding.weight"] = sd[p + "positional_embedding"]
    for a, b in (("ln_pre", "pre_layrnorm"), ("ln_post", "post_layernorm")):
        out[v + b + ".weight"], out[v + b + ".bias"] = sd[p + a + ".weight"], sd[p + a + ".bias"]
    out["visual_projection.weight"] = sd[p + "proj"].t().contiguous()
    i = 0
    while f"{p}transformer.resblocks.{i}.ln_1.weight" in sd:
        q, h = f"{p}transformer.resblocks.{i}", f"{v}encoder.layers.{i}"
        C = sd[q + ".ln_1.weight"].shape[0]
        wi, bi = sd[q + ".attn.in_proj_weight"], sd[q + ".attn.in_proj_bias"]
        for j, nme in enumerate(("q_proj", "k_proj", "v_proj")):
            out[f"{h}.self_attn.{nme}.weight"], out[f"{h}.self_attn.{nme}.bias"] = wi[j * C:(j + 1) * C], bi[j * C:(j + 1) * C]
        out[h + ".self_attn.out_proj.weight"], out[h + ".self_attn.out_proj.bias"] = sd[q + ".attn.out_proj.weight"], sd[q + ".attn.out_proj.bias"]
        for a, b in ((".ln_1", ".layer_norm1"), (".ln_2", ".layer_norm2"), (".mlp.c_fc", ".mlp.fc1"), (".mlp.c_proj", ".mlp.fc2")):
            out[h + b + ".weight"], out[h + b + ".bias"] = sd[q + a + ".weight"], sd[q + a + ".bias"]
        i += 1
    return out
