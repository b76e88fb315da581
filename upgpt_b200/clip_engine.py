"""B200 execution engines for the two CLIP ViT-L/14 conditioning towers.

Same construction as the U-Net engine: the parameter tree is packed once into fp16 [hi | lo] operand planes, the forward pass is a
static program of C-ABI launches over pre-allocated buffers, replayed as one CUDA graph.

  text  (FrozenCLIPEmbedder -> transformers.CLIPTextModel, reference ldm/modules/encoders/modules.py:137-162):
      embed_tokens -> 12 x [ LN -> q|k|v GEMM(+bias, fp32) -> causal attention (fp32, 77 tokens: upgpt_attention_small)
                             -> out GEMM(+bias +residual) -> LN -> fc1 GEMM(+bias) -> QuickGELU cast -> fc2 GEMM(+bias +residual) ]
      -> final LayerNorm (fp32)                                                           => (B, 77, 768)
  image (FrozenClipImageEmbedder2 -> clip VisionTransformer, modules.py:234-256):
      patchify (im2col, K 588 -> 592) -> patch GEMM -> [class ; patches] + positional -> ln_pre (fp32 residual stream)
      -> 24 x [ LN -> q|k|v GEMM(+bias, fp16) -> flash attention on tcgen05 (257 tokens, 16 heads of 64; V read row-major)
                -> out GEMM(+bias +residual) -> LN -> fc1 GEMM(+bias) -> QuickGELU cast -> fc2 GEMM(+bias +residual) ]
      -> ln_post on the class rows -> projection GEMM 1024 -> 768                          => (n, 768)

Operands are error-compensated fp16x3 planes by default (within 1e-3 of the fp32 oracle); UPGPT_CLIP_PRECISION=fp16 switches both
towers to single-plane operands -- the arithmetic class of the reference's own image tower, whose clip.load model holds fp16 weights
on CUDA (44.4 -> 21.7 ms for the 72 style crops of a request of 8, 1.1e-3 from the fp16x3 result).
"""
import ctypes as C
import os

import torch

from . import _C
from .unet_engine import EngineBase, _Program, split3_w, _round_up


def clip_precision():
    """"fp16x3" (default: error-compensated operand planes, within 1e-3 of the fp32 oracle) or "fp16" (single operand plane = the
    arithmetic class of the reference itself, whose clip.load model holds fp16 weights on CUDA; ~2.5x faster image tower)."""
    return os.environ.get("UPGPT_CLIP_PRECISION", "fp16x3")


class _ClipEngineBase(EngineBase):
    def __init__(self, device, dry=False, precision=None, host=None):
        """dry: record the program over host buffers without a device (CPU unit tests of the host logic); it can never run."""
        if device.type != "cuda" and not dry:
            raise _C.UpgptError("CLIP engines need the module on a CUDA device (no CPU fallback)")
        precision = precision or clip_precision()
        assert precision in ("fp16x3", "fp16"), precision
        super().__init__(device, precision, None if dry else getattr(host, "_wstore", None), getattr(host, "_weights_tag", "raw"))
        self.s3 = int(self.split3)             # producers emit [hi | lo] planes
        self.dry = dry
        self.weights_version = -1
        self.graph = None

    def _w16(self, w):
        return split3_w(w) if self.split3 else w.half()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(self, use_graph=True):
        if self.dry or self.dev.type != "cuda":
            raise _C.UpgptError("CLIP engine recorded without a CUDA device: nothing to run (no CPU fallback)")
        if use_graph:
            if self.graph is None:
                from .ops import Graph
                self.prog.run(self._stream())
                self.graph = Graph().capture(lambda: self.prog.run(self._stream()))
            self.graph.launch()
        else:
            self.prog.run(self._stream())

    @property
    def launches(self):
        return len(self.prog.calls)

    # ---- shared transformer layer: x (fp32 residual stream [M][Cc]) updated in place over two buffers ----
    def _mlp(self, q, cur, nxt, M, Cc, inner, tok16, h32, h16, ln2):
        """x += fc2(QuickGELU(fc1(LN(x)))): returns (cur, nxt) swapped."""
        g = lambda n: self.w.get(q + n)
        self.e_layernorm(cur, M, Cc, g(ln2 + ".weight"), g(ln2 + ".bias"), tok16)
        self.e_gemm(a=tok16, w=g(".fc1.weight"), mode=_C.GEMM_PLAIN, M=M, N=inner, K=Cc, out32=h32, bias=g(".fc1.bias"), flags=self.x3)
        if not self._sizing:
            self.prog.add(self.L.upgpt_quick_gelu_cast, h32.data_ptr(), M, inner, self.s3, h16.data_ptr(), self.kx * inner)
        self.e_gemm(a=h16, w=g(".fc2.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=inner, out32=nxt, bias=g(".fc2.bias"), res32=cur, flags=self.x3)
        return nxt, cur


class ClipTextEngine(_ClipEngineBase):
    def __init__(self, host, B, L, dry=False, precision=None):
        tm = host.transformer.text_model
        super().__init__(tm.final_layer_norm.weight.device, dry, precision, host)
        a = host.arch
        self.B, self.seq = B, L
        self.Cc, self.heads, self.inner, self.vocab = a["width"], a["heads"], a["mlp"], a["vocab"]
        self.n_layers = len(tm.encoder.layers)
        assert self.Cc % self.heads == 0 and L <= tm.embeddings.position_embedding.weight.shape[0] and L <= 128
        self.pack_weights(host)
        self._emit()
        self.finish_sizing()
        self._emit()

    def pack_weights(self, host):
        if self.shared_pack(host._weights_version):
            self.weights_version = host._weights_version
            return
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in host.transformer.state_dict().items()}
        put, p = self.put, "text_model."
        put("tok", sd[p + "embeddings.token_embedding.weight"]); put("pos", sd[p + "embeddings.position_embedding.weight"])
        for i in range(self.n_layers):
            s, q = f"{p}encoder.layers.{i}", f"l{i}"
            at = s + ".self_attn."
            put(q + ".qkv.weight", self._w16(torch.cat([sd[at + "q_proj.weight"], sd[at + "k_proj.weight"], sd[at + "v_proj.weight"]], 0)))
            put(q + ".qkv.bias", torch.cat([sd[at + "q_proj.bias"], sd[at + "k_proj.bias"], sd[at + "v_proj.bias"]], 0))
            put(q + ".out.weight", self._w16(sd[at + "out_proj.weight"])); put(q + ".out.bias", sd[at + "out_proj.bias"])
            for n in ("fc1", "fc2"):
                put(f"{q}.{n}.weight", self._w16(sd[f"{s}.mlp.{n}.weight"])); put(f"{q}.{n}.bias", sd[f"{s}.mlp.{n}.bias"])
            for n in ("layer_norm1", "layer_norm2"):
                put(f"{q}.{n}.weight", sd[f"{s}.{n}.weight"]); put(f"{q}.{n}.bias", sd[f"{s}.{n}.bias"])
        put("lnf.weight", sd[p + "final_layer_norm.weight"]); put("lnf.bias", sd[p + "final_layer_norm.bias"])
        self.weights_version = host._weights_version
        self.publish_pack(self.weights_version)

    def _emit(self):
        B, L, Cc, Hh, inner = self.B, self.seq, self.Cc, self.heads, self.inner
        M, d = B * L, self.Cc // self.heads
        if not self._sizing:
            self.prog = _Program()
        ids = self.buf("ids", (B, L), torch.int64)
        xa, xb = self.buf("xa", (M, Cc)), self.buf("xb", (M, Cc))
        out = self.buf("out", (B, L, Cc))
        tok16 = self.scratch("tok16", M * Cc * 2, torch.float16)
        qkv32 = self.scratch("qkv32", M * 3 * Cc, torch.float32)
        att16 = self.scratch("att16", M * Cc * 2, torch.float16)
        h32 = self.scratch("h32", M * inner, torch.float32)
        h16 = self.scratch("h16", M * inner * 2, torch.float16)
        if not self._sizing:
            self.prog.add(self.L.upgpt_embed_tokens, ids.data_ptr(), M, L, self.vocab, self.w["tok"].data_ptr(), self.w["pos"].data_ptr(), Cc,
                          xa.data_ptr())
        cur, nxt = xa, xb
        for i in range(self.n_layers):
            q = f"l{i}"
            g = lambda n: self.w.get(q + n)
            self.e_layernorm(cur, M, Cc, g(".layer_norm1.weight"), g(".layer_norm1.bias"), tok16)
            self.e_gemm(a=tok16, w=g(".qkv.weight"), mode=_C.GEMM_PLAIN, M=M, N=3 * Cc, K=Cc, out32=qkv32, bias=g(".qkv.bias"), flags=self.x3)
            if not self._sizing:   # causal mask: CLIPTextTransformer builds it for every input (attention_mask=None in the reference call)
                self.prog.add(self.L.upgpt_attention_small, qkv32.data_ptr(), 3 * Cc, Cc, 2 * Cc, B, Hh, L, d, float(d) ** -0.5, 1, self.s3,
                              att16.data_ptr(), self.kx * Cc)
            self.e_gemm(a=att16, w=g(".out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out32=nxt, bias=g(".out.bias"), res32=cur, flags=self.x3)
            cur, nxt = nxt, cur
            cur, nxt = self._mlp(q, cur, nxt, M, Cc, inner, tok16, h32, h16, ".layer_norm2")
        if not self._sizing:
            self.prog.add(self.L.upgpt_layernorm_f32, cur.data_ptr(), Cc, M, Cc, self.w["lnf.weight"].data_ptr(), self.w["lnf.bias"].data_ptr(), 1e-5,
                          out.data_ptr(), Cc)

    def forward(self, ids, use_graph=True):
        """ids (B, L) int64 on the device -> last_hidden_state (B, L, width) fp32."""
        assert tuple(ids.shape) == (self.B, self.seq)
        self.bufs["ids"].copy_(ids)
        self.run(use_graph)
        return self.bufs["out"].clone()


class ClipVisionEngine(_ClipEngineBase):
    def __init__(self, host, n, dry=False, precision=None):
        v = host.model.visual
        super().__init__(v.proj.device, dry, precision, host)
        a = host.arch
        self.n, self.Cc, self.heads, self.patch, self.S, self.odim = n, a["width"], a["heads"], a["patch"], a["resolution"], a["output_dim"]
        self.n_layers = len(v.transformer.resblocks)
        self.G = self.S // self.patch
        self.T = self.G * self.G + 1
        self.Kp = _round_up(3 * self.patch * self.patch, 8)          # 588 -> 592: 16-byte TMA rows
        assert self.Cc // self.heads == 64, "the attention kernel takes 64-wide (or 128-wide) heads; ViT-L/14 has 16 heads of 64"
        self.pack_weights(host)
        self._emit()
        self.finish_sizing()
        self._emit()

    def pack_weights(self, host):
        if self.shared_pack(host._weights_version):
            self.weights_version = host._weights_version
            return
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in host.model.visual.state_dict().items()}
        put = self.put
        w = sd["conv1.weight"].reshape(self.Cc, -1)
        wp = w.new_zeros(self.Cc, self.Kp); wp[:, :w.shape[1]] = w
        put("patch.weight", self._w16(wp))
        put("cls", sd["class_embedding"]); put("pos", sd["positional_embedding"])
        for n in ("ln_pre", "ln_post"):
            put(n + ".weight", sd[n + ".weight"]); put(n + ".bias", sd[n + ".bias"])
        put("proj.weight", self._w16(sd["proj"].t().contiguous()))                # [output_dim][width]
        for i in range(self.n_layers):
            s, q = f"transformer.resblocks.{i}", f"l{i}"
            put(q + ".qkv.weight", self._w16(sd[s + ".attn.in_proj_weight"])); put(q + ".qkv.bias", sd[s + ".attn.in_proj_bias"])
            put(q + ".out.weight", self._w16(sd[s + ".attn.out_proj.weight"])); put(q + ".out.bias", sd[s + ".attn.out_proj.bias"])
            put(q + ".fc1.weight", self._w16(sd[s + ".mlp.c_fc.weight"])); put(q + ".fc1.bias", sd[s + ".mlp.c_fc.bias"])
            put(q + ".fc2.weight", self._w16(sd[s + ".mlp.c_proj.weight"])); put(q + ".fc2.bias", sd[s + ".mlp.c_proj.bias"])
            for nm in ("ln_1", "ln_2"):
                put(f"{q}.{nm}.weight", sd[f"{s}.{nm}.weight"]); put(f"{q}.{nm}.bias", sd[f"{s}.{nm}.bias"])
        self.weights_version = host._weights_version
        self.publish_pack(self.weights_version)

    def _emit(self):
        n, Cc, Hh, T, G, Kp = self.n, self.Cc, self.heads, self.T, self.G, self.Kp
        M, Mp, inner, d = n * T, n * G * G, 4 * self.Cc, 64
        HD = Hh * d
        if not self._sizing:
            self.prog = _Program()
        img = self.buf("img", (n, 3, self.S, self.S))
        xa, xb = self.buf("xa", (M, Cc)), self.buf("xb", (M, Cc))
        cls16 = self.buf("cls16", (n, self.kx * Cc), torch.float16)
        out = self.buf("out", (n, self.odim))
        pat16 = self.scratch("pat16", Mp * Kp * 2, torch.float16)
        pat32 = self.scratch("h32", max(Mp * Cc, M * inner), torch.float32)       # patch embeddings share the MLP's fp32 scratch
        tok16 = self.scratch("tok16", M * Cc * 2, torch.float16)
        qkv16 = self.scratch("qkv16", M * 3 * HD, torch.float16)
        att16 = self.scratch("att16", M * HD * 2, torch.float16)
        h32 = pat32
        h16 = self.scratch("h16", M * inner * 2, torch.float16)
        if not self._sizing:
            P = self.prog
            P.add(self.L.upgpt_patchify, img.data_ptr(), n, 3, self.S, self.patch, Kp, self.s3, pat16.data_ptr(), self.kx * Kp)
        self.e_gemm(a=pat16, w=self.w.get("patch.weight"), mode=_C.GEMM_PLAIN, M=Mp, N=Cc, K=Kp, out32=pat32, flags=self.x3)
        if not self._sizing:
            P.add(self.L.upgpt_vit_assemble, pat32.data_ptr(), self.w["cls"].data_ptr(), self.w["pos"].data_ptr(), n, T, Cc, xb.data_ptr())
            P.add(self.L.upgpt_layernorm_f32, xb.data_ptr(), Cc, M, Cc, self.w["ln_pre.weight"].data_ptr(), self.w["ln_pre.bias"].data_ptr(), 1e-5,
                  xa.data_ptr(), Cc)
        cur, nxt = xa, xb
        for i in range(self.n_layers):
            q = f"l{i}"
            g = lambda nme: self.w.get(q + nme)
            self.e_layernorm(cur, M, Cc, g(".ln_1.weight"), g(".ln_1.bias"), tok16)
            # one q | k | v projection (+bias) straight to fp16; the flash kernel reads V row-major as an MN-major operand
            self.e_gemm(a=tok16, w=g(".qkv.weight"), mode=_C.GEMM_PLAIN, M=M, N=3 * HD, K=Cc, out16=qkv16, bias=g(".qkv.bias"), flags=self.x3)
            kptr = None if self._sizing else qkv16[HD:]
            vptr = None if self._sizing else qkv16[2 * HD:]
            self.e_attention(q=qkv16, ldq=3 * HD, k=kptr, ldk=3 * HD, k_batch_stride=T * 3 * HD, vt=vptr, ldvt=3 * HD, v_rowmajor=1,
                             v_batch_stride=T * 3 * HD, out=att16, ldo=self.kx * HD, B=n, H=Hh, Nq=T, Nk=T, dpad=d, scale=float(d) ** -0.5,
                             split3_out=self.s3)
            self.e_gemm(a=att16, w=g(".out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=HD, out32=nxt, bias=g(".out.bias"), res32=cur, flags=self.x3)
            cur, nxt = nxt, cur
            cur, nxt = self._mlp(q, cur, nxt, M, Cc, inner, tok16, h32, h16, ".ln_2")
        # ln_post on the class token of every image (row stride T*Cc), then the 1024 -> 768 projection (no bias)
        self.e_layernorm(cur, n, Cc, self.w.get("ln_post.weight"), self.w.get("ln_post.bias"), cls16, ldx=T * Cc)
        self.e_gemm(a=cls16, w=self.w.get("proj.weight"), mode=_C.GEMM_PLAIN, M=n, N=self.odim, K=Cc, out32=out, flags=self.x3)

    def forward(self, images, use_graph=True):
        """images (n, 3, S, S) fp32 on the device, already normalised as the caller of encode_image does -> (n, output_dim) fp32."""
        assert tuple(images.shape) == (self.n, 3, self.S, self.S), (tuple(images.shape), (self.n, 3, self.S, self.S))
        self.bufs["img"].copy_(images)
        self.run(use_graph)
        return self.bufs["out"].clone()
