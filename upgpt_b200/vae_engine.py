"""B200 execution engine for the KL-f8 VAE decode (AutoencoderKL.decode -> Decoder.forward, model.py:535-568).

Same building blocks as the U-Net engine: GroupNorm(eps 1e-6)+swish -> fp16 operand, tcgen05 implicit-GEMM 3x3 convs,
nearest-x2 folded into the operand-prep kernel, and the middle single-head attention (d = 512) as
QK^T GEMM -> row softmax -> PV GEMM with the batched tcgen05 GEMM.
"""
import ctypes as C

import torch

from . import _C
from .unet_engine import EngineBase, _Program, default_precision, up2_conv_w, UP2_FOLD


class VAEDecoderEngine(EngineBase):
    def __init__(self, ae, B, H, W, precision=None):
        dev = next(ae.parameters()).device
        if dev.type != "cuda":
            raise _C.UpgptError("VAEDecoderEngine needs the module on a CUDA device (no CPU fallback)")
        super().__init__(dev, precision or default_precision(), getattr(ae, '_wstore', None), getattr(ae, '_weights_tag', 'raw'))
        self.B, self.H, self.W = B, H, W
        self.dec = ae.decoder
        self.zc = ae.post_quant_conv.in_channels
        self.weights_version = -1
        self.graph = None
        self.pack_weights(ae)
        self._emit()
        self.finish_sizing()
        self._emit()

    def _conv_w(self, w):
        return w.permute(0, 2, 3, 1).contiguous().reshape(w.shape[0], 9, w.shape[1]).half()

    def pack_weights(self, ae):
        if self.shared_pack(ae._weights_version):
            self.weights_version = ae._weights_version
            return
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in ae.state_dict().items()}
        put = self.put
        w = sd["post_quant_conv.weight"]
        put("pq.weight", w.permute(1, 2, 3, 0).reshape(-1, w.shape[0])); put("pq.bias", sd["post_quant_conv.bias"])
        w = sd["decoder.conv_in.weight"]
        put("conv_in.weight", w.permute(1, 2, 3, 0).reshape(-1, w.shape[0])); put("conv_in.bias", sd["decoder.conv_in.bias"])
        for k, v in sd.items():
            if not k.startswith("decoder."):
                continue
            n = k[len("decoder."):]
            if n.startswith("conv_in."):
                continue
            if v.dim() == 4 and v.shape[-1] == 3 and ".upsample." in n and UP2_FOLD:
                put(n, up2_conv_w(v).half())
            elif v.dim() == 4 and v.shape[-1] == 3:
                put(n, self._conv_w(v))
            elif v.dim() == 4:
                put(n, v.reshape(v.shape[0], v.shape[1]).half())
            else:
                put(n, v)
        self.weights_version = ae._weights_version
        self.publish_pack(self.weights_version)

    def _resnet(self, p, cin, cout, x, B, H, W, out):
        HW = H * W
        has_skip = (p + ".nin_shortcut.weight") in self.w
        op, raw = self.norm_operand(x, cin, None, 0, B, H, W, p + ".norm1", 1e-6, True, want_raw=has_skip)
        h32 = self.scratch("res_h", B * HW * cout, torch.float32)
        self.e_gemm(a=op, w=self.w.get(p + ".conv1.weight"), mode=_C.GEMM_CONV3X3, N=cout, K=cin, n_imgs=B, H=H, W=W, out32=h32,
                    bias=self.w.get(p + ".conv1.bias"))
        op2, _ = self.norm_operand(h32, cout, None, 0, B, H, W, p + ".norm2", 1e-6, True)
        if has_skip:
            skip32 = self.scratch("res_skip", B * HW * cout, torch.float32)
            self.e_gemm(a=raw, w=self.w.get(p + ".nin_shortcut.weight"), mode=_C.GEMM_PLAIN, M=B * HW, N=cout, K=cin, out32=skip32,
                        bias=self.w.get(p + ".nin_shortcut.bias"))
            res = skip32
        else:
            res = x
        self.e_gemm(a=op2, w=self.w.get(p + ".conv2.weight"), mode=_C.GEMM_CONV3X3, N=cout, K=cout, n_imgs=B, H=H, W=W, out32=out,
                    bias=self.w.get(p + ".conv2.bias"), res32=res)

    def _attn(self, p, Cc, x, B, H, W, out):
        N = H * W
        M = B * N
        op, _ = self.norm_operand(x, Cc, None, 0, B, H, W, p + ".norm", 1e-6, False)
        q16 = self.scratch("vq16", M * Cc, torch.float16)
        k16 = self.scratch("vk16", M * Cc, torch.float16)
        vt16 = self.scratch("vvt16", M * Cc, torch.float16)
        s32 = self.scratch("vs32", B * N * N, torch.float32)
        p16 = self.scratch("vp16", B * N * N, torch.float16)
        o16 = self.scratch("vo16", M * Cc, torch.float16)
        g = lambda n: self.w.get(p + n)
        self.e_gemm(a=op, w=g(".q.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out16=q16, bias=g(".q.bias"))
        self.e_gemm(a=op, w=g(".k.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out16=k16, bias=g(".k.bias"))
        self.e_gemm(a=op, w=g(".v.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out16=vt16, bias=g(".v.bias"), rows_per_group=N,
                    ldT=N, flags=_C.GEMM_F_CHW)
        # scores[b] = Q[b] K[b]^T  (model.py:186-189), softmax over keys with scale c^-0.5, O[b] = P[b] V[b]
        self.e_gemm(a=q16, w=k16, mode=_C.GEMM_PLAIN, M=N, N=N, K=Cc, batch=B, out32=s32)
        if not self._sizing:
            self.prog.add(self.L.upgpt_softmax_rows, s32.data_ptr(), N, B * N, N, float(Cc) ** -0.5, p16.data_ptr(), N)
        self.e_gemm(a=p16, w=vt16, mode=_C.GEMM_PLAIN, M=N, N=Cc, K=N, batch=B, out16=o16)
        self.e_gemm(a=o16, w=g(".proj_out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out32=out, bias=g(".proj_out.bias"), res32=x)

    def _emit(self):
        B, H, W = self.B, self.H, self.W
        dec = self.dec
        if not self._sizing:
            self.prog = _Program()
        z_in = self.buf("z_in", (B, self.zc, H, W))
        z_pq = self.buf("z_pq", (B, dec.z_channels, H, W))
        scale = self.buf("in_scale_host", (1,))   # placeholder to keep naming uniform (scale passed by value)
        ch = dec.ch * dec.ch_mult[-1]
        h = self.buf("h_conv_in", (B, H * W, ch))
        self._pq_call_index = None
        if not self._sizing:
            P = self.prog
            self._pq_call_index = len(P.calls)
            P.add(self.L.upgpt_conv_small_cin, z_in.data_ptr(), self.zc, 0, 0, 1.0, B, H, W, 1, self.w["pq.weight"].data_ptr(),
                  self.w["pq.bias"].data_ptr(), dec.z_channels, z_pq.data_ptr(), 1)
            P.add(self.L.upgpt_conv_small_cin, z_pq.data_ptr(), dec.z_channels, 0, 0, 1.0, B, H, W, 3, self.w["conv_in.weight"].data_ptr(),
                  self.w["conv_in.bias"].data_ptr(), ch, h.data_ptr(), 0)
        hh, ww = H, W

        def newbuf(name, c):
            return self.buf(name, (B, hh * ww, c))

        out = newbuf("mid.block_1.out", ch); self._resnet("mid.block_1", ch, ch, h, B, hh, ww, out); h = out
        out = newbuf("mid.attn_1.out", ch); self._attn("mid.attn_1", ch, h, B, hh, ww, out); h = out
        out = newbuf("mid.block_2.out", ch); self._resnet("mid.block_2", ch, ch, h, B, hh, ww, out); h = out
        for i_level in reversed(range(dec.num_resolutions)):
            cout = dec.ch * dec.ch_mult[i_level]
            for i_block in range(dec.num_res_blocks + 1):
                p = f"up.{i_level}.block.{i_block}"
                # ping-pong two buffers per level to bound memory at 256x256
                out = self.buf(f"lvl{i_level}.pp{i_block & 1}.c{cout}", (B, hh * ww, cout))
                self._resnet(p, ch, cout, h, B, hh, ww, out)
                h, ch = out, cout
            if i_level != 0:
                p = f"up.{i_level}.upsample.conv"
                out = self.buf(f"lvl{i_level}.up", (B, hh * ww * 4, ch))
                if UP2_FOLD:
                    # Upsample (model.py:49-52) without the x2 replicate: four parity-wise 2x2 convolutions over the low-resolution operand
                    op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=0)
                    self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3_UP2, N=ch, K=ch, n_imgs=B, H=hh, W=ww, out32=out,
                                bias=self.w.get(p + ".bias"))
                    hh, ww = hh * 2, ww * 2
                else:
                    op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=1)
                    hh, ww = hh * 2, ww * 2
                    self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3, N=ch, K=ch, n_imgs=B, H=hh, W=ww, out32=out,
                                bias=self.w.get(p + ".bias"))
                h = out
        img = self.buf("img", (B, dec.out_ch, hh, ww))
        op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, "norm_out", 1e-6, True)
        self.e_gemm(a=op, w=self.w.get("conv_out.weight"), mode=_C.GEMM_CONV3X3, N=dec.out_ch, K=ch, n_imgs=B, H=hh, W=ww, block_n=16,
                    splits=1, out32=img, bias=self.w.get("conv_out.bias"), flags=_C.GEMM_F_CHW)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def set_in_scale(self, in_scale):
        """1/scale_factor is a by-value argument of the post-quant conv launch; patch it (graphs are re-captured)."""
        fn, args = self.prog.calls[self._pq_call_index]
        if args[4] != float(in_scale):
            args = args[:4] + (float(in_scale),) + args[5:]
            self.prog.calls[self._pq_call_index] = (fn, args)
            self.graph = None

    def run(self, use_graph=True):
        if use_graph:
            if self.graph is None:
                from .ops import Graph
                self.prog.run(self._stream())
                self.graph = Graph().capture(lambda: self.prog.run(self._stream()))
            self.graph.launch()
        else:
            self.prog.run(self._stream())
        return self.bufs["img"]

    def decode(self, z, in_scale=1.0, use_graph=True):
        self.set_in_scale(in_scale)
        self.bufs["z_in"].copy_(z)
        return self.run(use_graph).clone()

    @property
    def launches_per_decode(self):
        return len(self.prog.calls)


class VAEEncoderEngine(VAEDecoderEngine):
    """AutoencoderKL.encode -> Encoder.forward -> quant_conv (autoencoder.py:324-328; model.py:368-459) on the same kernels as the
    decoder: 3->ch input conv as a direct small-Cin convolution from the NCHW image, ResnetBlocks (fused GroupNorm+swish -> tcgen05
    implicit-GEMM convs), Downsample = stride-2 conv over the (0,1,0,1)-padded input as 4 stride-2 phase planes + TMA zero fill
    (UPGPT_GEMM_CONV3X3_S2PHASE_ASYM), middle single-head attention, conv_out stored channel-major (NCHW moments), 1x1 quant_conv.
    SURVEY.md 8(f) rank 2 ('next' row): precision is the decoder's (single fp16 operand plane, fp32 everything else)."""

    def __init__(self, ae, B, H, W, precision=None):
        dev = next(ae.parameters()).device
        if dev.type != "cuda":
            raise _C.UpgptError("VAEEncoderEngine needs the module on a CUDA device (no CPU fallback)")
        EngineBase.__init__(self, dev, "fp16", getattr(ae, '_wstore', None), getattr(ae, '_weights_tag', 'raw'))
        self.B, self.H, self.W = B, H, W
        self.enc = ae.encoder
        nres = self.enc.num_resolutions
        assert H % (2 ** (nres - 1)) == 0 and W % (2 ** (nres - 1)) == 0, "image size must divide by the encoder stride"
        self.weights_version = -1
        self.graph = None
        self.pack_weights(ae)
        self._emit()
        self.finish_sizing()
        self._emit()

    def pack_weights(self, ae):
        if self.shared_pack(ae._weights_version):
            self.weights_version = ae._weights_version
            return
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in ae.state_dict().items()}
        put = self.put
        w = sd["encoder.conv_in.weight"]
        put("conv_in.weight", w.permute(1, 2, 3, 0).reshape(-1, w.shape[0])); put("conv_in.bias", sd["encoder.conv_in.bias"])
        w = sd["quant_conv.weight"]
        put("quant.weight", w.permute(1, 2, 3, 0).reshape(-1, w.shape[0])); put("quant.bias", sd["quant_conv.bias"])
        for k, v in sd.items():
            if not k.startswith("encoder.") or k.startswith("encoder.conv_in."):
                continue
            n = k[len("encoder."):]
            if v.dim() == 4 and v.shape[-1] == 3:
                put(n, self._conv_w(v))
            elif v.dim() == 4:
                put(n, v.reshape(v.shape[0], v.shape[1]).half())
            else:
                put(n, v)
        self.weights_version = ae._weights_version
        self.publish_pack(self.weights_version)

    def _emit(self):
        B, H, W = self.B, self.H, self.W
        enc = self.enc
        if not self._sizing:
            self.prog = _Program()
        x_in = self.buf("x_in", (B, enc.in_channels, H, W))
        ch = enc.ch
        h = self.buf("h_conv_in", (B, H * W, ch))
        if not self._sizing:
            self.prog.add(self.L.upgpt_conv_small_cin, x_in.data_ptr(), enc.in_channels, 0, 0, 1.0, B, H, W, 3, self.w["conv_in.weight"].data_ptr(),
                          self.w["conv_in.bias"].data_ptr(), ch, h.data_ptr(), 0)
        hh, ww = H, W
        for i_level in range(enc.num_resolutions):
            cout = enc.ch * enc.ch_mult[i_level]
            for i_block in range(enc.num_res_blocks):
                p = f"down.{i_level}.block.{i_block}"
                out = self.buf(f"lvl{i_level}.pp{i_block & 1}.c{cout}", (B, hh * ww, cout))
                self._resnet(p, ch, cout, h, B, hh, ww, out)
                h, ch = out, cout
            if i_level != enc.num_resolutions - 1:
                p = f"down.{i_level}.downsample.conv"
                op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=2)
                hh, ww = hh // 2, ww // 2
                out = self.buf(f"lvl{i_level}.down", (B, hh * ww, ch))
                self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3_S2PHASE_ASYM, N=ch, K=ch, n_imgs=B, H=hh, W=ww, out32=out,
                            bias=self.w.get(p + ".bias"))
                h = out

        def newbuf(name, c):
            return self.buf(name, (B, hh * ww, c))

        out = newbuf("mid.block_1.out", ch); self._resnet("mid.block_1", ch, ch, h, B, hh, ww, out); h = out
        out = newbuf("mid.attn_1.out", ch); self._attn("mid.attn_1", ch, h, B, hh, ww, out); h = out
        out = newbuf("mid.block_2.out", ch); self._resnet("mid.block_2", ch, ch, h, B, hh, ww, out); h = out
        zc2 = enc.conv_out.out_channels
        h_out = self.buf("h_out", (B, zc2, hh, ww))           # NCHW
        op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, "norm_out", 1e-6, True)
        self.e_gemm(a=op, w=self.w.get("conv_out.weight"), mode=_C.GEMM_CONV3X3, N=zc2, K=ch, n_imgs=B, H=hh, W=ww, block_n=16,
                    splits=1, out32=h_out, bias=self.w.get("conv_out.bias"), flags=_C.GEMM_F_CHW)
        nq = self.w["quant.bias"].shape[0] if "quant.bias" in self.w else zc2
        moments = self.buf("moments", (B, nq, hh, ww))         # NCHW {mean | logvar}
        if not self._sizing:
            self.prog.add(self.L.upgpt_conv_small_cin, h_out.data_ptr(), zc2, 0, 0, 1.0, B, hh, ww, 1, self.w["quant.weight"].data_ptr(),
                          self.w["quant.bias"].data_ptr(), nq, moments.data_ptr(), 1)

    def run(self, use_graph=True):
        if use_graph:
            if self.graph is None:
                from .ops import Graph
                self.prog.run(self._stream())
                self.graph = Graph().capture(lambda: self.prog.run(self._stream()))
            self.graph.launch()
        else:
            self.prog.run(self._stream())
        return self.bufs["moments"]

    def encode(self, x, use_graph=True):
        """x (B, 3, H, W) fp32 NCHW -> moments (B, 2*embed_dim, H/8, W/8) fp32 NCHW."""
        self.bufs["x_in"].copy_(x)
        return self.run(use_graph).clone()
