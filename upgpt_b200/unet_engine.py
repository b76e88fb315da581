"""B200 execution engine for ldm.modules.diffusionmodules.openaimodel.UNetModel.

The engine turns the module tree into a static *program*: a list of C-ABI kernel launches over pre-allocated device
buffers (NHWC fp32 residual stream, fp16 tensor-core operands), recorded once per (batch, H, W, context length).
Replaying the list is the eager path; capturing one replay in a CUDA graph is the fast path used by the samplers.

Program per ResBlock (openaimodel.py:255-275):
    gn_stats -> prep(GN+SiLU -> fp16) -> conv3x3[tcgen05](+bias +timestep-emb row vector) -> gn_stats -> prep
    -> [skip 1x1 GEMM] -> conv3x3(+bias +residual)
per SpatialTransformer (attention.py:250-261,211-215):
    fused GN -> proj_in GEMM -> LN -> q|k|v GEMM -> flash attention (V row-major, MN-major operand) -> to_out GEMM(+res)
    -> LN -> Q GEMM -> flash attention over the cached context K/V -> to_out GEMM(+res)
    -> LN -> GEGLU GEMM -> ff2 GEMM(+res, +fp16 copy) -> proj_out GEMM(+block input)
The skip concat th.cat([h, hs.pop()]) (openaimodel.py:736) never exists in fp32: gn_stats / prep read both tensors.
"""
import ctypes as C
import os

import torch

from . import _C
from ldm.modules.attention import SpatialTransformer
from ldm.modules.diffusionmodules import openaimodel as om


def default_precision():
    """Operand precision of the tensor-core GEMMs / convs.
      "fp16x3"          : every operand is split into hi + lo fp16 planes and each product is formed as
                          Ah*Wh + Al*Wh + Ah*Wl with fp32 accumulation -> eps within ~1.5e-4 of the fp32 reference
                          (BASELINE.json's tolerance is 1e-3);
      "mixed" (default) : fp16x3 everywhere except the deep low-resolution levels of a U-Net with a probed profile (MIXED_PROFILES:
                          the bbox.yaml architecture; any other U-Net runs fp16x3 throughout), whose GEMMs / convs are
                          weight-bandwidth bound and contribute least to the eps error (CPU probe tests/probe_mixed_precision.py
                          predicted +2.9e-4 in quadrature at bbox.yaml; measured on B200: eps 3.2e-4 .. 3.4e-4, U-Net step 5.03 ->
                          4.64 ms): levels with H*W <= 16 run single-plane fp16 throughout, levels with H*W <= 64 all but the
                          residual-path 1x1s (skip_connection, proj_in, proj_out) and a conv that shares its operand with a skip GEMM;
      "fp16"            : single fp16 plane, ~1.5x faster end to end, eps within ~1.3e-3 .. 1.7e-3 (opt-in fast mode)."""
    return os.environ.get("UPGPT_PRECISION", "mixed")


# "mixed" precision profiles: U-Net architecture (model_channels, channel_mult, num_res_blocks, attention_resolutions) ->
# (deep_hw, full_hw) thresholds on the tokens per image of a level. How much of eps flows through the deep levels depends on the
# architecture (the same thresholds cost +2.9e-4 at bbox.yaml but +4.6e-4 on the upscale U-Net fed a 32x24 latent), so a profile is
# only applied to an architecture it was probed on (tests/probe_mixed_precision.py) and verified against the reference's golden
# eps on the GPU; every other U-Net keeps fp16x3 throughout.
MIXED_PROFILES = {
    (224, (1, 2, 4, 4), 2, (4, 2, 1)): (64, 16),      # configs/deepfashion/bbox.yaml: eps 3.2e-4 .. 3.4e-4, step 5.03 -> 4.64 ms at B=8
}


def _round_up(x, m):
    return (x + m - 1) // m * m


def pad_heads_rows(w, heads, d, dpad):
    """[heads*d, K] -> [heads*dpad, K] with zero rows appended per head."""
    K = w.shape[1]
    out = w.new_zeros(heads, dpad, K)
    out[:, :d] = w.reshape(heads, d, K)
    return out.reshape(heads * dpad, K)


def pad_heads_cols(w, heads, d, dpad):
    """[N, heads*d] -> [N, heads*dpad] with zero columns appended per head."""
    N = w.shape[0]
    out = w.new_zeros(N, heads, dpad)
    out[:, :, :d] = w.reshape(N, heads, d)
    return out.reshape(N, heads * dpad)


def head_pad(d, heads):
    """Padded head width of the attention operands: 32 (two heads share one 64-wide row of q / k / v; needs an even head count),
    64 or 128 columns. d = 28 / 56 / 112 at bbox.yaml -> 32 / 64 / 128."""
    if d <= 32 and heads % 2 == 0 and os.environ.get("UPGPT_ATTN_D32", "1") != "0":
        return 32
    assert d <= 128, "head dim > 128 not supported by the attention kernel"
    return 64 if d <= 64 else 128


def geglu_half(inner, x3=False):
    """Half-width of a GEGLU accumulator tile ([x | gate] = 2*half columns). In the fp16x3 mode a 256-column tile leaves shared
    memory for a single {hi, lo} slot pair (no load / MMA overlap: measured 47 us for M=8192, N=1792, K=224), so 224 is preferred;
    a half of 128 gets the TMA-store epilogue (two 64-column fp16 chunks, one per epilogue group)."""
    order = (112, 96, 64, 128, 80, 48, 32, 16) if x3 else (128, 112, 96, 80, 64, 48, 32, 16)
    for h in order:
        if inner % h == 0:
            return h
    raise ValueError("GEGLU inner dim %d not a multiple of 16" % inner)


def pack_geglu(w, b, inner, half):
    """Rows [x(inner) ; gate(inner)] -> per tile [x(half) | gate(half)] so one accumulator tile holds matching columns."""
    idx = []
    for t in range(inner // half):
        idx += list(range(t * half, (t + 1) * half)) + list(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx, device=w.device)
    return w[idx].contiguous(), b[idx].contiguous()


def split3_w(w):
    """Error-compensated fp16 weights: planes [Wh | Wl] along K (w ~= Wh + Wl to ~22 bits).  The GEMM (UPGPT_GEMM_F_X3)
    forms Ah*Wh + Al*Wh + Ah*Wl from these and the activation planes [Ah | Al]."""
    wh = w.half()
    wl = (w - wh.float()).half()
    return torch.cat([wh, wl], dim=-1)


def up2_conv_w(w):
    """[Cout, Cin, 3, 3] fp32 weights of the conv behind a nearest-x2 upsample (openaimodel.py:116-118, model.py:49-52) -> the four
    parity-wise 2x2 kernels [4 (a*2+b), Cout, 4 (u*2+v), Cin] of UPGPT_GEMM_CONV3X3_UP2: output pixel (2y+a, 2x+b) reads the upsampled
    rows 2y+a-1 .. 2y+a+1, i.e. source rows y+a-1 and y+a; the 3x3 taps that land on the same source pixel are summed (in fp32)."""
    rows = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}      # parity -> 3x3 tap indices of source offset u = 0 / 1
    out = w.new_zeros(4, w.shape[0], 4, w.shape[1])
    for a in (0, 1):
        for b in (0, 1):
            for u in (0, 1):
                for v in (0, 1):
                    acc = 0
                    for i in rows[a][u]:
                        for j in rows[b][v]:
                            acc = acc + w[:, :, i, j]
                    out[a * 2 + b, :, u * 2 + v, :] = acc
    return out


UP2_FOLD = os.environ.get("UPGPT_UP2_FOLD", "1") != "0"


class _OnAux:
    """A launch that goes to one of the library's auxiliary streams (a parallel branch of the program, between fork and join)."""

    def __init__(self, fn, idx):
        self.fn, self.idx, self.__name__ = fn, idx, fn.__name__

    def __call__(self, *a):
        return self.fn(*a[:-1], C.c_void_p(_C.lib().upgpt_aux_stream(self.idx)))


class _Edge:
    """fork: the auxiliary stream waits for the main stream; join: the main stream waits for the auxiliary stream."""

    def __init__(self, kind, idx):
        self.kind, self.idx, self.__name__ = kind, idx, "upgpt_stream_" + kind

    def __call__(self, stream):
        L = _C.lib()
        return (L.upgpt_stream_fork if self.kind == "fork" else L.upgpt_stream_join)(stream, self.idx)


class _Program:
    """Recorded list of (function, argument tuple) launches; fork / join edges and auxiliary-stream launches mark parallel branches."""

    def __init__(self):
        self.calls = []
        self.keep = []   # ctypes structs must outlive the program
        self.aux = None  # index of the auxiliary stream the following launches go to (between fork and join_point)

    def add(self, fn, *args):
        self.calls.append((fn if self.aux is None else _OnAux(fn, self.aux), args))

    def add_struct(self, fn, struct):
        self.keep.append(struct)
        self.add(fn, C.byref(struct))

    def fork(self, idx=0):
        self.calls.append((_Edge("fork", idx), ()))

    def join(self, idx=0):
        self.calls.append((_Edge("join", idx), ()))

    def kernel_calls(self):
        """(C-ABI function, args) of every kernel launch, whatever stream it goes to (stream edges and memset nodes are not kernels)."""
        for fn, args in self.calls:
            if isinstance(fn, _Edge) or getattr(fn, "__name__", "") == "upgpt_zero":
                continue
            yield (fn.fn if isinstance(fn, _OnAux) else fn), args

    @property
    def n_kernels(self):
        return sum(1 for _ in self.kernel_calls())

    def run(self, stream):
        for fn, args in self.calls:
            rc = fn(*args, stream)
            if rc != 0:
                raise _C.UpgptError("%s failed (%d): %s" % (fn.__name__, rc, _C.lib().upgpt_last_error().decode()))


class EngineBase:
    """Buffer bookkeeping + emitters shared by the U-Net and VAE engines."""

    def __init__(self, device, precision, store=None, tag="raw"):
        self.dev = device
        self.precision = precision
        self.store, self.tag = store, tag      # shared packed-weight store of the host module (upgpt_b200/host.py)
        assert precision in ("fp16x3", "mixed", "fp16"), precision
        self.split3 = precision in ("fp16x3", "mixed")   # engine-wide default operand format ("mixed" overrides it per layer)
        self.x3 = _C.GEMM_F_X3 if self.split3 else 0     # GEMM flag: operands carry [hi | lo] planes
        self.kx = 2 if self.split3 else 1                # storage planes per fp16 operand
        self.L = _C.lib()
        self.w = {}
        self.bufs = {}
        self._scratch_need = {}
        self._scratch = {}
        self._sizing = True
        self.prog = _Program()
        from . import lanes
        self.lane = lanes.current()                       # the lane (upgpt_b200/lanes.py) this engine's buffers and programs belong to
        self.branch_aux = lanes.branch_aux(self.lane)     # auxiliary stream of its forked branches
        # GroupNorm statistics from the producing GEMM's epilogue (include/upgpt_b200.h: gn_acc): out32 pointer -> recorded args struct of
        # the launch that produces the tensor; accumulator slots of [B][32][2] int64 from one pool that the program zeroes first
        self.gn_epi = os.environ.get("UPGPT_GN_EPILOGUE", "0") != "0"
        self._producers = {}
        self._gn_slots = 0

    # ---- memory ----
    def put(self, name, t):
        t = t.contiguous()
        if self.store is not None:
            # one packed copy per (module, operand-format plan, weights tag), shared by every engine with that plan
            self.w[name] = self.store.put((self.plan_signature(), name), self.tag, t, self.dev)
        elif name in self.w and self.w[name].shape == t.shape and self.w[name].dtype == t.dtype:
            self.w[name].copy_(t)      # keep the address: captured graphs stay valid across re-packs
        else:
            self.w[name] = t.to(self.dev)
        return self.w[name]

    def plan_signature(self):
        """Engines with equal signatures pack identical weight sets (same names, formats, values)."""
        return (type(self).__name__, self.precision)

    def shared_pack(self, version):
        """True if another engine of the module already packed this weight version in this engine's formats: bind and skip."""
        if self.store is None:
            return False
        rec = self.store.plans.get((self.plan_signature(), self.tag))
        if rec is None or rec["version"] != version:
            return False
        self.w = dict(rec["w"])
        return True

    def publish_pack(self, version):
        if self.store is not None:
            self.store.plans[(self.plan_signature(), self.tag)] = {"version": version, "w": dict(self.w)}

    def buf(self, name, shape, dtype=torch.float32):
        if name not in self.bufs:
            self.bufs[name] = torch.zeros(shape, device=self.dev, dtype=dtype)
        return self.bufs[name]

    def scratch(self, role, numel, dtype):
        """Role-shared scratch: sized to the max request during the sizing pass; pointer handed out afterwards."""
        key = (role, dtype)
        if self._sizing:
            self._scratch_need[key] = max(self._scratch_need.get(key, 0), int(numel))
            return None
        return self._scratch[key]

    def finish_sizing(self):
        for key, n in self._scratch_need.items():
            self._scratch[key] = torch.zeros(n, device=self.dev, dtype=key[1])
        self._sizing = False

    @staticmethod
    def p(t):
        return 0 if t is None else t.data_ptr()

    # ---- emitters ----
    def e_gn_affine(self, x1, C1, x2, C2, B, HW, stats, gamma, beta, eps, ss):
        if self._sizing:
            return
        self.prog.add(self.L.upgpt_groupnorm_affine, self.p(x1), C1, self.p(x2), C2, B, HW, 32, self.p(stats), self.p(gamma), self.p(beta),
                      eps, self.p(ss))

    def e_prep(self, x1, C1, x2, C2, B, H, W, stats, gamma, beta, eps, silu, layout, out, raw=None, split3=False, ss=None, gn_acc=0, raw_split3=None):
        if self._sizing:
            return
        a = _C.PrepArgs()
        a.scale_shift = self.p(ss)
        a.gn_acc = gn_acc
        a.raw_planes = 0 if raw_split3 is None else (2 if raw_split3 else 1)
        a.x1, a.C1, a.x2, a.C2 = self.p(x1), C1, self.p(x2), C2
        a.B, a.H, a.W, a.groups = B, H, W, 32
        a.stats, a.gamma, a.beta, a.eps = self.p(stats), self.p(gamma), self.p(beta), eps
        a.silu, a.layout, a.split3 = int(silu), layout, int(split3)
        a.out, a.ldo, a.raw, a.ldraw = self.p(out), 0, self.p(raw), 0
        self.prog.add_struct(self.L.upgpt_prep_operand, a)

    def e_gn_prep(self, x1, C1, x2, C2, B, H, W, stats, gamma, beta, eps, silu, layout, out, raw=None, split3=False, raw_split3=None):
        if self._sizing:
            return
        a = _C.PrepArgs()
        a.raw_planes = 0 if raw_split3 is None else (2 if raw_split3 else 1)
        a.x1, a.C1, a.x2, a.C2 = self.p(x1), C1, self.p(x2), C2
        a.B, a.H, a.W, a.groups = B, H, W, 32
        a.stats, a.scale_shift = 0, 0
        a.gamma, a.beta, a.eps = self.p(gamma), self.p(beta), eps
        a.silu, a.layout, a.split3 = int(silu), layout, int(split3)
        a.out, a.ldo, a.raw, a.ldraw = self.p(out), 0, self.p(raw), 0
        self.prog.keep.append(a)
        self.prog.add(self.L.upgpt_groupnorm_prep, C.byref(a), C.c_void_p(self.p(stats)))

    def e_gemm(self, **kw):
        """Records one upgpt_gemm launch; returns the number of N tiles the library picks for it (= rowstats slots per row)."""
        if self._sizing:
            return 1
        a = _C.GemmArgs()
        for k, v in kw.items():
            setattr(a, k, self.p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
        if a.batch <= 1:
            # every unbatched GEMM of the engines multiplies by model weights (only the VAE attention's q k^T / p v products have an
            # activation as W): the kernel may stream W ahead of its dependency wait
            a.flags |= _C.GEMM_F_W_STATIC
        n_tiles = 1
        if a.rowstats_out and not getattr(self, "dry", False):
            plan = (C.c_int * 8)()
            if self.L.upgpt_gemm_plan(C.byref(a), C.byref(plan)) != 0:
                a.splits = 1        # the row statistics need the TMA-store or the cluster split-K epilogue: drop split-K if the pick has neither
                _C.check(self.L.upgpt_gemm_plan(C.byref(a), C.byref(plan)), "upgpt_gemm_plan")
            n_tiles = int(plan[1])
            a.rowstats_slots = n_tiles      # a later launch that would tile differently fails loudly instead of mis-feeding the consumers
        self.prog.add_struct(self.L.upgpt_gemm, a)
        if a.out32:
            self._producers[a.out32] = a
        return n_tiles

    GN_MAX_SLOTS = 96

    def gn_begin(self):
        """Start of a recorded pass: one memset node re-arms the GroupNorm moment accumulators (its byte count is patched by gn_end)."""
        if self._sizing or not self.gn_epi or getattr(self, "dry", False):
            return
        self._producers, self._gn_slots = {}, 0
        pool = self.buf("gn_acc_pool", (self.GN_MAX_SLOTS, self.B, 32, 2), torch.int64)
        self._gn_zero_idx = len(self.prog.calls)
        self.prog.add(self.L.upgpt_zero, pool.data_ptr(), pool.numel() * 8)

    def gn_end(self):
        if self._sizing or not self.gn_epi or getattr(self, "dry", False):
            return
        fn, args = self.prog.calls[self._gn_zero_idx]
        if self._gn_slots == 0:
            del self.prog.calls[self._gn_zero_idx]
        else:
            self.prog.calls[self._gn_zero_idx] = (fn, (args[0], self._gn_slots * self.B * 32 * 2 * 8))

    def _gn_from_epilogues(self, sources, B, HW):
        """sources: [(tensor, channels, channel offset in the GroupNorm's input)]. If every source tensor was produced by a recorded
        upgpt_gemm launch that can emit GroupNorm moments from its epilogue, wires them to a fresh accumulator slot and returns its
        address; else None (the caller uses the one-launch fused GroupNorm kernel)."""
        if not self.gn_epi or getattr(self, "dry", False) or self._gn_slots >= self.GN_MAX_SLOTS:
            return None
        Cc = sum(c for _, c, _ in sources)
        if Cc % 32:
            return None
        prods = []
        for t, c, off in sources:
            a = self._producers.get(self.p(t))
            if a is None or (a.gn_acc and a.gn_acc2) or HW % 4:
                return None
            prods.append((a, off))
        if "gn_acc_pool" not in self.bufs:
            return None       # no gn_begin() in this program
        acc = self.bufs["gn_acc_pool"].data_ptr() + self._gn_slots * B * 32 * 2 * 8
        saved = []
        ok = True
        for a, off in prods:
            saved.append((a, a.gn_acc, a.gn_groups, a.gn_cpg, a.gn_choff, a.gn_acc2, a.gn_cpg2, a.gn_choff2, a.rows_per_group, a.splits, a.block_n))
            if a.mode == _C.GEMM_PLAIN:
                a.rows_per_group = HW
            if not a.gn_acc:
                a.gn_acc, a.gn_groups, a.gn_cpg, a.gn_choff = acc, 32, Cc // 32, off
            else:
                a.gn_acc2, a.gn_cpg2, a.gn_choff2 = acc, Cc // 32, off
            plan = (C.c_int * 8)()
            rc = self.L.upgpt_gemm_plan(C.byref(a), C.byref(plan))
            if rc != 0 and a.splits <= 0:
                # e.g. a split factor of 3 / 5 / 6 / 7: the moments need equal power-of-two row slices -> pin the next lower power of two
                probe = _C.GemmArgs.from_buffer_copy(a)
                probe.gn_acc, probe.gn_acc2 = 0, 0
                if self.L.upgpt_gemm_plan(C.byref(probe), C.byref(plan)) == 0 and plan[2] > 1:
                    sp = 1
                    while sp * 2 <= plan[2]:
                        sp *= 2
                    a.splits, a.block_n = sp, int(plan[0])     # (a pinned split alone would re-pick the tile width for an unsplit launch)
                    rc = self.L.upgpt_gemm_plan(C.byref(a), C.byref(plan))
            if rc != 0:
                ok = False
                break
        if not ok:
            for a, g0, g1, g2, g3, g4, g5, g6, rpg, sp, bn in saved:
                a.gn_acc, a.gn_groups, a.gn_cpg, a.gn_choff, a.gn_acc2, a.gn_cpg2, a.gn_choff2, a.rows_per_group, a.splits, a.block_n = g0, g1, g2, g3, g4, g5, g6, rpg, sp, bn
            return None
        self._gn_slots += 1
        return acc

    def e_layernorm(self, x, rows, Cc, gamma, beta, out16, split3=None, ldx=None):
        if self._sizing:
            return
        fn = self.L.upgpt_layernorm_split3 if (self.split3 if split3 is None else split3) else self.L.upgpt_layernorm
        self.prog.add(fn, self.p(x), ldx or Cc, rows, Cc, self.p(gamma), self.p(beta), 1e-5, self.p(out16), 0)

    def e_attention(self, **kw):
        if self._sizing:
            return
        a = _C.AttnArgs()
        for k, v in kw.items():
            setattr(a, k, self.p(v) if (isinstance(v, torch.Tensor) or v is None) else v)
        self.prog.add_struct(self.L.upgpt_attention, a)

    # GroupNorm(+SiLU) -> fp16 operand of the (possibly concatenated) input; returns (operand, raw16)
    def norm_operand(self, x1, C1, x2, C2, B, H, W, gname, eps, silu, want_raw=False, layout=0, split3=False, raw_split3=None):
        """raw_split3: plane format of the un-normalised copy (None = as the operand's): the skip GEMM that reads it may run fp16x3
        while the convolution on the normalised operand takes a single plane."""
        Cc = C1 + C2
        stats = self.buf("gn_stats", (B, 32, 2), torch.float64)
        mult = 4 if layout == 1 else 1
        rs3 = split3 if raw_split3 is None else raw_split3
        op = self.scratch("op16", B * H * W * mult * Cc * (2 if split3 else 1), torch.float16)
        raw = self.scratch("raw16", B * H * W * Cc * (2 if rs3 else 1), torch.float16) if want_raw else None
        if gname is not None:
            acc = None
            if not self._sizing:
                acc = self._gn_from_epilogues([(x1, C1, 0)] + ([(x2, C2, C1)] if C2 else []), B, H * W)
            if acc is not None:
                # the moments come from the epilogues of the launches that produced x1 / x2: GroupNorm(+SiLU) + cast is apply-only
                self.e_prep(x1, C1, x2, C2, B, H, W, None, self.w.get(gname + ".weight"), self.w.get(gname + ".bias"), eps, silu, layout, op,
                            raw, split3, gn_acc=acc, raw_split3=raw_split3)
            else:
                # GroupNorm(+SiLU) + cast: one fused cluster launch per GroupNorm (two launches internally for VAE-sized images)
                self.e_gn_prep(x1, C1, x2, C2, B, H, W, stats, self.w.get(gname + ".weight"), self.w.get(gname + ".bias"), eps, silu, layout, op,
                               raw, split3, raw_split3=raw_split3)
        else:
            self.e_prep(x1, C1, x2, C2, B, H, W, None, None, None, 0.0, False, layout, op, raw, split3, raw_split3=raw_split3)
        return op, raw


class UNetEngine(EngineBase):
    def __init__(self, unet, B, H, W, ctx_len, precision=None, dry=False, plan=None, store=None):
        """dry: record the program over host buffers without a device (CPU unit tests of the host logic); it can never run.
        plan: {"mixed_hw": (deep_hw, full_hw) | None, "tf_x1": bool[, "tf_hw": int]} of the "mixed" precision mode (upgpt_b200/precision.py calibrates
        it per checkpoint); default: the static profile of MIXED_PROFILES / the UPGPT_MIXED_HW, UPGPT_TF_PLANES overrides.
        store: packed-weight store to use instead of the module's (throw-away engines of the calibration)."""
        dev = next(unet.parameters()).device
        if dev.type != "cuda" and not dry:
            raise _C.UpgptError("UNetEngine needs the module on a CUDA device (no CPU fallback)")
        super().__init__(dev, precision or default_precision(), store if store is not None else getattr(unet, "_wstore", None),
                         getattr(unet, "_weights_tag", "raw"))
        self.dry = dry
        self.B, self.H, self.W, self.ctx_len = B, H, W, ctx_len
        self.mc = unet.model_channels
        self.in_ch, self.out_ch = unet.in_channels, unet.out_channels
        self.ctx_dim = unet.context_dim
        nlevels = len(unet.channel_mult)
        assert H % (2 ** (nlevels - 1)) == 0 and W % (2 ** (nlevels - 1)) == 0, "latent size must divide by the U-Net stride"
        self.unet_ref = unet
        self.weights_version = -1
        self.graph = None
        self._ctx_key = None
        arch = (unet.model_channels, tuple(unet.channel_mult), unet.num_res_blocks, tuple(unet.attention_resolutions))
        self.mixed_hw = MIXED_PROFILES.get(arch) if self.precision == "mixed" else None
        if self.precision == "mixed" and os.environ.get("UPGPT_MIXED_HW"):    # tuning override "deep_hw,full_hw" for any architecture
            self.mixed_hw = tuple(int(v) for v in os.environ["UPGPT_MIXED_HW"].split(","))
            assert len(self.mixed_hw) == 2, "UPGPT_MIXED_HW=deep_hw,full_hw"
        tf_x1 = os.environ.get("UPGPT_TF_PLANES", "x3") == "x1"
        if plan is not None and self.precision == "mixed":
            self.mixed_hw, tf_x1 = plan["mixed_hw"], bool(plan["tf_x1"])
        # tf_hw: the attention / feed-forward GEMMs run on single planes up to this many tokens per image (None = every level)
        self.tf_hw = (plan or {}).get("tf_hw") if self.precision == "mixed" else None
        self.skip_x1 = bool((plan or {}).get("skip_x1", os.environ.get("UPGPT_SKIP_X1", "0") == "1")) and self.precision == "mixed"
        # finer allocation of the eps budget (precision.py, plans deep+tf1c*): concat_x1_hw = the decoder ResBlocks' first conv (input =
        # [h | skip], K = 9 (C1 + C2): the most expensive fp16x3 launches) takes single planes up to this many tokens per image;
        # tf_keep_x3 = {sub-kind: hw}: attention out-projections ("tf_out") / ff2 ("tf_ff2") keep [hi | lo] planes from that resolution up
        # (their error per microsecond saved is the worst of the transformer GEMMs, profiles/r01_precision_sensitivity.txt)
        self.concat_x1_hw = int((plan or {}).get("concat_x1_hw", 0)) if self.precision == "mixed" else 0
        self.tf_keep_x3 = dict((plan or {}).get("tf_keep_x3", {})) if self.precision == "mixed" else {}
        self.force_x1 = frozenset((k, int(h)) for k, h in (plan or {}).get("x1", ())) if self.precision == "mixed" else frozenset()
        self.force_x3 = frozenset((k, int(h)) for k, h in (plan or {}).get("x3", ())) if self.precision == "mixed" else frozenset()
        self.mixed = self.mixed_hw is not None
        # "mixed" only: attention projections + feed-forward GEMMs of every level on single fp16 planes (see default_precision)
        self.tf_x1 = self.mixed and tf_x1
        self.plan_name = (plan or {}).get("name", "static")
        # a ResBlock's skip 1x1 GEMM runs on an auxiliary stream beside conv1 + GroupNorm (fork / join edges in the step graph)
        self.par_skip = os.environ.get("UPGPT_PAR_SKIP", "1") != "0"
        # LayerNorm folded into the GEMMs around it (include/upgpt_b200.h: rowstats_out / ln_stats): no LayerNorm launches
        self.ln_fold = os.environ.get("UPGPT_LN_FOLD", "1") != "0"
        self.layer_hw = self._layer_resolutions(unet)
        self.emb_off, off = {}, 0      # column of each ResBlock's timestep-embedding projection in the concatenated GEMV
        for name, mod in unet.named_modules():
            if isinstance(mod, om.ResBlock):
                self.emb_off[name] = off
                off += mod.out_channels
        self.emb_total = off
        self.pack_weights(unet)
        # two passes over the same emitter: sizing, then recording
        self._emit(unet)
        self.finish_sizing()
        self._emit(unet)
        self._emit_context(unet)

    # ------------------------------------------------------------------------------------------------ precision plan
    def _layer_resolutions(self, unet):
        """module path -> tokens per image (H*W) the layer's GEMMs produce (Downsample / Upsample: their output size)."""
        hw, res = self.H * self.W, {}

        def walk(prefix, layers, hw):
            for j, mod in enumerate(layers):
                if isinstance(mod, om.Downsample):
                    hw //= 4
                elif isinstance(mod, om.Upsample):
                    hw *= 4
                res[f"{prefix}.{j}"] = hw
            return hw
        for i in range(1, len(unet.input_blocks)):
            hw = walk(f"input_blocks.{i}", unet.input_blocks[i], hw)
        hw = walk("middle_block", unet.middle_block, hw)
        for i, blk in enumerate(unet.output_blocks):
            hw = walk(f"output_blocks.{i}", blk, hw)
        return res

    def use_x3(self, kind, hw):
        """Operand format of one GEMM: True = error-compensated [hi | lo] planes (fp16x3), False = single fp16 plane.
        kind: "conv" (3x3), "conv_skipshared" (3x3 whose operand planes also feed a skip_connection GEMM), "resid1x1" (skip_connection,
        proj_in, proj_out: they write the residual stream directly), "tf" (attention projections and the feed-forward)."""
        if not self.mixed:
            return self.split3
        deep_hw, full_hw = self.mixed_hw
        # explicit per-(kind, tokens per image) overrides of a plan: {"x1": [[kind, hw], ...], "x3": [...]} (the sub-kind first, then its parent)
        if (kind, hw) in self.force_x1:
            return False
        if (kind, hw) in self.force_x3:
            return True
        if kind in ("tf_out", "tf_ff2"):       # sub-kinds of "tf" that a plan may keep error-compensated at the high resolutions
            if hw > full_hw and kind in self.tf_keep_x3 and hw >= self.tf_keep_x3[kind]:
                return True
            kind = "tf"
        if kind == "conv_concat":              # a decoder ResBlock's first conv (always with a skip connection)
            if hw <= self.concat_x1_hw:
                return False
            kind = "conv_skipshared"
        if kind == "conv_updown":              # Downsample / Upsample convs
            kind = "conv"
        if (kind, hw) in self.force_x1:
            return False
        if (kind, hw) in self.force_x3:
            return True
        if hw <= full_hw or (kind == "tf" and self.tf_x1 and (self.tf_hw is None or hw <= self.tf_hw)):
            return False
        if hw <= deep_hw:
            # conv_skipshared: the un-normalised copy that feeds the skip GEMM keeps its own [hi | lo] planes (raw_planes), so the 3x3 conv
            # on the normalised operand need not follow the skip GEMM's format (skip_x1, plans deep+tf1 / deep+tf1s)
            return kind == "resid1x1" or (kind == "conv_skipshared" and not self.skip_x1)
        return True

    @staticmethod
    def _conv1_kind(path, has_skip):
        """Precision-plan kind of a ResBlock's first conv: output_blocks.N.0 reads the concatenation [h | skip] (openaimodel.py:736)."""
        if path.startswith("output_blocks.") and has_skip:
            return "conv_concat"
        return "conv_skipshared" if has_skip else "conv"

    # ------------------------------------------------------------------------------------------------ weights
    def _w16(self, w, x3=None):
        """fp32 weight [..., K] -> fp16 operand; as an fp16x3 operand the K axis carries the planes [Wh | Wl]."""
        return split3_w(w) if (self.split3 if x3 is None else x3) else w.half()

    def _conv_w(self, w, x3=None):
        """[Cout, Cin, 3, 3] fp32 -> [Cout, 9, Cin] fp16 (x2 planes as an fp16x3 operand)."""
        return self._w16(w.permute(0, 2, 3, 1).contiguous().reshape(w.shape[0], 9, w.shape[1]), x3)

    def plan_signature(self):
        if getattr(self, "_plan_sig", None) is None:
            self._plan_sig = (type(self).__name__, self.precision, self.mixed_hw, self.tf_x1, self.tf_hw, self.skip_x1, self.concat_x1_hw,
                              tuple(sorted(self.tf_keep_x3.items())), tuple(sorted(self.force_x1)), tuple(sorted(self.force_x3)), self.ln_fold,
                              tuple(sorted(self.layer_hw.items())) if self.mixed else None)
        return self._plan_sig

    def pack_weights(self, unet):
        self._ctx_key = None   # cached context K/V depends on the weights
        if getattr(self, "cond", None) is not None:
            self.cond.invalidate()
        if self.shared_pack(unet._weights_version):
            self.weights_version = unet._weights_version
            return
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in unet.state_dict().items()}
        put = self.put
        for k in ("time_embed.0", "time_embed.2"):
            put(k + ".weight", sd[k + ".weight"]); put(k + ".bias", sd[k + ".bias"])
        # boundary convs
        w0 = sd["input_blocks.0.0.weight"]
        put("conv_in.weight", w0.permute(1, 2, 3, 0).reshape(-1, w0.shape[0]))
        put("conv_in.bias", sd["input_blocks.0.0.bias"])
        emb_w, emb_b = [], []
        for name, mod in unet.named_modules():
            if isinstance(mod, om.ResBlock):
                p = name
                hw = self.layer_hw[p]
                has_skip = (p + ".skip_connection.weight") in sd
                x3a = self.use_x3(self._conv1_kind(p, has_skip), hw)                 # conv1
                x3s = x3a or self.use_x3("resid1x1", hw)                             # the skip GEMM (on the un-normalised copy's planes)
                for g in (".in_layers.0", ".out_layers.0"):
                    put(p + g + ".weight", sd[p + g + ".weight"]); put(p + g + ".bias", sd[p + g + ".bias"])
                put(p + ".conv1.weight", self._conv_w(sd[p + ".in_layers.2.weight"], x3a)); put(p + ".conv1.bias", sd[p + ".in_layers.2.bias"])
                put(p + ".conv2.weight", self._conv_w(sd[p + ".out_layers.3.weight"], self.use_x3("conv", hw)))
                put(p + ".conv2.bias", sd[p + ".out_layers.3.bias"])
                if has_skip:
                    ws = sd[p + ".skip_connection.weight"]
                    put(p + ".skip.weight", self._w16(ws.reshape(ws.shape[0], ws.shape[1]), x3s)); put(p + ".skip.bias", sd[p + ".skip_connection.bias"])
                emb_w.append(sd[p + ".emb_layers.1.weight"]); emb_b.append(sd[p + ".emb_layers.1.bias"])
            elif isinstance(mod, (om.Downsample, om.Upsample)):
                sub = ".op" if isinstance(mod, om.Downsample) else ".conv"
                w = sd[name + sub + ".weight"]
                if isinstance(mod, om.Upsample) and UP2_FOLD:
                    put(name + ".weight", self._w16(up2_conv_w(w), self.use_x3("conv_updown", self.layer_hw[name])))
                else:
                    put(name + ".weight", self._conv_w(w, self.use_x3("conv_updown", self.layer_hw[name])))
                put(name + ".bias", sd[name + sub + ".bias"])
            elif isinstance(mod, SpatialTransformer):
                p = name
                Cc, Hh, d = mod.in_channels, mod.n_heads, mod.d_head
                dpad = head_pad(d, Hh)
                x3p, x3t = self.use_x3("resid1x1", self.layer_hw[p]), self.use_x3("tf", self.layer_hw[p])
                x3o, x3f = self.use_x3("tf_out", self.layer_hw[p]), self.use_x3("tf_ff2", self.layer_hw[p])    # attention out-projections ; ff2
                put(p + ".norm.weight", sd[p + ".norm.weight"]); put(p + ".norm.bias", sd[p + ".norm.bias"])
                wpi = sd[p + ".proj_in.weight"]
                put(p + ".proj_in.weight", self._w16(wpi.reshape(wpi.shape[0], wpi.shape[1]), x3p)); put(p + ".proj_in.bias", sd[p + ".proj_in.bias"])
                wpo = sd[p + ".proj_out.weight"]
                put(p + ".proj_out.weight", self._w16(wpo.reshape(wpo.shape[0], wpo.shape[1]), x3p)); put(p + ".proj_out.bias", sd[p + ".proj_out.bias"])
                for bi in range(len(mod.transformer_blocks)):
                    q = f"{p}.transformer_blocks.{bi}"
                    for n in ("norm1", "norm2", "norm3"):
                        put(f"{q}.{n}.weight", sd[f"{q}.{n}.weight"]); put(f"{q}.{n}.bias", sd[f"{q}.{n}.bias"])
                    wq = pad_heads_rows(sd[q + ".attn1.to_q.weight"], Hh, d, dpad)
                    wk = pad_heads_rows(sd[q + ".attn1.to_k.weight"], Hh, d, dpad)
                    wv = pad_heads_rows(sd[q + ".attn1.to_v.weight"], Hh, d, dpad)
                    wqkv = torch.cat([wq, wk, wv], 0)                                            # one fused q | k | v projection
                    wq2 = pad_heads_rows(sd[q + ".attn2.to_q.weight"], Hh, d, dpad)
                    if self.ln_fold:
                        # LN(x) W^T = rstd (x (gamma o W)^T - mean colsum) + W beta: gamma goes into the weights, beta into a bias
                        wqkv = self._put_ln_folded(q + ".attn1.qkv", wqkv, None, sd[q + ".norm1.weight"], sd[q + ".norm1.bias"], x3t)
                        wq2 = self._put_ln_folded(q + ".attn2.q", wq2, None, sd[q + ".norm2.weight"], sd[q + ".norm2.bias"], x3t)
                    put(q + ".attn1.qkv.weight", self._w16(wqkv, x3t))
                    put(q + ".attn1.out.weight", self._w16(pad_heads_cols(sd[q + ".attn1.to_out.0.weight"], Hh, d, dpad), x3o))
                    put(q + ".attn1.out.bias", sd[q + ".attn1.to_out.0.bias"])
                    put(q + ".attn2.q.weight", self._w16(wq2, x3t))
                    put(q + ".attn2.kv.weight", self._w16(torch.cat([pad_heads_rows(sd[q + ".attn2.to_k.weight"], Hh, d, dpad),
                                                                     pad_heads_rows(sd[q + ".attn2.to_v.weight"], Hh, d, dpad)], 0)))
                    put(q + ".attn2.out.weight", self._w16(pad_heads_cols(sd[q + ".attn2.to_out.0.weight"], Hh, d, dpad), x3o))
                    put(q + ".attn2.out.bias", sd[q + ".attn2.to_out.0.bias"])
                    inner = sd[q + ".ff.net.2.weight"].shape[1]
                    half = geglu_half(inner, x3t)
                    w1, b1 = sd[q + ".ff.net.0.proj.weight"], sd[q + ".ff.net.0.proj.bias"]
                    if self.ln_fold:
                        g3, be3 = sd[q + ".norm3.weight"], sd[q + ".norm3.bias"]
                        b1 = (b1.double() + w1.double() @ be3.double()).float()
                        w1 = w1 * g3[None, :]
                    w1, b1 = pack_geglu(w1, b1, inner, half)
                    put(q + ".ff1.weight", self._w16(w1, x3t)); put(q + ".ff1.bias", b1)
                    if self.ln_fold:
                        put(q + ".ff1.colsum", self.w[q + ".ff1.weight"].double().sum(-1).float())
                    put(q + ".ff2.weight", self._w16(sd[q + ".ff.net.2.weight"], x3f)); put(q + ".ff2.bias", sd[q + ".ff.net.2.bias"])
        from .ops import timestep_freqs
        put("temb.freqs", timestep_freqs(self.mc))
        put("emb_all.weight", torch.cat(emb_w, 0)); put("emb_all.bias", torch.cat(emb_b, 0))
        put("out.0.weight", sd["out.0.weight"]); put("out.0.bias", sd["out.0.bias"])
        put("out.conv.weight", self._conv_w(sd["out.2.weight"])); put("out.conv.bias", sd["out.2.bias"])
        self.weights_version = unet._weights_version
        self.publish_pack(self.weights_version)

    def _put_ln_folded(self, name, w, b, gamma, beta, x3):
        """LayerNorm(gamma, beta) folded into the Linear (w, b) that consumes it: stores bias' = b + W beta and the column sums of the
        gamma-scaled weight, returns that weight for the caller to pack."""
        bias = w.double() @ beta.double()
        if b is not None:
            bias = bias + b.double()
        self.put(name + ".bias", bias.float())
        w = w * gamma[None, :]
        # column sums of the weights AS THE TENSOR CORE SEES THEM (fp16-rounded planes), so the mean term cancels exactly
        self.put(name + ".colsum", self._w16(w, x3).double().sum(-1).float())
        return w

    # ------------------------------------------------------------------------------------------------ program
    def _res_block(self, p, mod, x1, C1, x2, C2, B, H, W, out):
        Cin, Cout = C1 + C2, mod.out_channels
        HW = H * W
        has_skip = (p + ".skip.weight") in self.w
        x3a = self.use_x3(self._conv1_kind(p, has_skip), HW)                 # conv1
        x3s = x3a or self.use_x3("resid1x1", HW)                             # the skip GEMM, on the un-normalised copy's own planes
        x3b = self.use_x3("conv", HW)                                        # conv2
        fa, fb, fs = (_C.GEMM_F_X3 if x3a else 0), (_C.GEMM_F_X3 if x3b else 0), (_C.GEMM_F_X3 if x3s else 0)
        op, raw = self.norm_operand(x1, C1, x2, C2, B, H, W, p + ".in_layers.0", 1e-5, True, want_raw=has_skip, split3=x3a, raw_split3=x3s)
        h32 = self.scratch("res_h", B * HW * Cout, torch.float32)
        emb = None if self._sizing else self.bufs["emb_all"][:, self.emb_off[p]:]
        par_skip = has_skip and self.par_skip and not self._sizing
        if has_skip:
            skip32 = self.scratch("res_skip", B * HW * Cout, torch.float32)
        if par_skip:
            # the skip 1x1 (openaimodel.py:241) only needs the un-normalised operand planes: it runs on an auxiliary stream beside
            # conv1 and the second GroupNorm and is joined in front of conv2, which consumes it as its residual
            self.prog.fork(self.branch_aux)
            self.prog.aux = self.branch_aux
            self.e_gemm(a=raw, w=self.w.get(p + ".skip.weight"), mode=_C.GEMM_PLAIN, M=B * HW, N=Cout, K=Cin, out32=skip32,
                        bias=self.w.get(p + ".skip.bias"), flags=fs)
            self.prog.aux = None
        self.e_gemm(a=op, w=self.w.get(p + ".conv1.weight"), mode=_C.GEMM_CONV3X3, N=Cout, K=Cin, n_imgs=B, H=H, W=W,
                    out32=h32, bias=self.w.get(p + ".conv1.bias"), rowvec=emb, ld_rowvec=self.emb_total, flags=fa)
        op2, _ = self.norm_operand(h32, Cout, None, 0, B, H, W, p + ".out_layers.0", 1e-5, True, split3=x3b)
        if par_skip:
            self.prog.join(self.branch_aux)
            res = skip32
        elif has_skip:
            self.e_gemm(a=raw, w=self.w.get(p + ".skip.weight"), mode=_C.GEMM_PLAIN, M=B * HW, N=Cout, K=Cin, out32=skip32,
                        bias=self.w.get(p + ".skip.bias"), flags=fs)
            res = skip32
        else:
            assert x2 is None or self._sizing or C2 == 0
            res = x1
        self.e_gemm(a=op2, w=self.w.get(p + ".conv2.weight"), mode=_C.GEMM_CONV3X3, N=Cout, K=Cout, n_imgs=B, H=H, W=W,
                    out32=out, bias=self.w.get(p + ".conv2.bias"), res32=res, flags=fb)

    def _transformer(self, p, mod, x, Cc, B, H, W, out):
        HW, M = H * W, B * H * W
        Hh, d = mod.n_heads, mod.d_head
        dpad = head_pad(d, Hh)
        HD = Hh * dpad
        L, Lp = self.ctx_len, _round_up(self.ctx_len, 8)
        x3p, x3t = self.use_x3("resid1x1", HW), self.use_x3("tf", HW)    # proj_in / proj_out ; attention projections + feed-forward
        x3o, x3f = self.use_x3("tf_out", HW), self.use_x3("tf_ff2", HW)  # the attention out-projections / ff2 may keep [hi | lo] planes
        kx = 2 if (x3p or x3t or x3o) else 1   # operand planes [hi | lo] of an error-compensated operand (scratch sizing)
        kxt = 2 if x3t else 1
        kxo, kxf = (2 if x3o else 1), (2 if x3f else 1)
        x3 = _C.GEMM_F_X3 if x3t else 0
        fo, ff_ = (_C.GEMM_F_X3 if x3o else 0), (_C.GEMM_F_X3 if x3f else 0)
        s3f = _C.GEMM_F_SPLIT3OUT if x3f else 0
        fp = _C.GEMM_F_X3 if x3p else 0
        s3 = _C.GEMM_F_SPLIT3OUT if x3t else 0
        op, _ = self.norm_operand(x, Cc, None, 0, B, H, W, p + ".norm", 1e-6, False, split3=x3p)
        tokA = self.scratch("tokA", M * Cc, torch.float32)
        tokB = self.scratch("tokB", M * Cc, torch.float32)
        tok16 = self.scratch("tok16", M * Cc * kx, torch.float16)
        qkv16 = self.scratch("qkv16", M * 3 * HD, torch.float16)
        att16 = self.scratch("att16", M * HD * kx, torch.float16)
        fold = self.ln_fold
        # folded LayerNorm: the GEMM that writes the residual stream also emits its raw fp16 planes and per-row {sum, sumsq}; the GEMM
        # that consumes LN(x) reads those planes with gamma-scaled weights and applies mean / rstd in its epilogue
        lnst = self.scratch("lnstats", M * 16 * 2, torch.float32) if fold else None
        rawkw = (lambda: dict(out16=tok16, rowstats_out=lnst)) if fold else (lambda: {})
        slots = self.e_gemm(a=op, w=self.w.get(p + ".proj_in.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out32=tokA,
                            bias=self.w.get(p + ".proj_in.bias"), flags=fp | (s3 if fold else 0), **rawkw())
        lnkw = lambda n, sl: dict(ln_stats=lnst, ln_slots=sl, ln_eps=1e-5, ln_colsum=self.w.get(n + ".colsum"), bias=self.w.get(n + ".bias"))
        cur, nxt = tokA, tokB
        for bi in range(len(mod.transformer_blocks)):
            q = f"{p}.transformer_blocks.{bi}"
            g = lambda n: self.w.get(q + n)
            inner = mod.transformer_blocks[bi].ff.net[2].in_features
            ff16 = self.scratch("ff16", M * inner * kxf, torch.float16)
            last = bi == len(mod.transformer_blocks) - 1
            assert slots <= 16
            # --- self attention ---
            if not fold:
                self.e_layernorm(cur, M, Cc, g(".norm1.weight"), g(".norm1.bias"), tok16, x3t)
            # one q | k | v projection (row-major fp16); the attention kernel reads V row-major as an MN-major operand (no V^T)
            self.e_gemm(a=tok16, w=g(".attn1.qkv.weight"), mode=_C.GEMM_PLAIN, M=M, N=3 * HD, K=Cc, out16=qkv16, flags=x3,
                        **(lnkw(q + ".attn1.qkv", slots) if fold else {}))
            kptr = None if self._sizing else qkv16[HD:]
            vptr = None if self._sizing else qkv16[2 * HD:]
            self.e_attention(q=qkv16, ldq=3 * HD, k=kptr, ldk=3 * HD, k_batch_stride=HW * 3 * HD, vt=vptr, ldvt=3 * HD, v_rowmajor=1,
                             v_batch_stride=HW * 3 * HD, out=att16, ldo=HD * kxo, B=B, H=Hh, Nq=HW, Nk=HW, dpad=dpad,
                             scale=float(d) ** -0.5, split3_out=int(x3o))
            slots = self.e_gemm(a=att16, w=g(".attn1.out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=HD, out32=nxt,
                                bias=g(".attn1.out.bias"), res32=cur, flags=fo | (s3 if fold else 0), **rawkw())
            cur, nxt = nxt, cur
            # --- cross attention over the cached context K / V^T ---
            if not fold:
                self.e_layernorm(cur, M, Cc, g(".norm2.weight"), g(".norm2.bias"), tok16, x3t)
            self.e_gemm(a=tok16, w=g(".attn2.q.weight"), mode=_C.GEMM_PLAIN, M=M, N=HD, K=Cc, out16=qkv16, flags=x3,
                        **(lnkw(q + ".attn2.q", slots) if fold else {}))
            kvc = self.buf(q + ".ctx_kv", (B * L, 2 * HD), torch.float16)       # cond-cache: K | V of the context, row-major
            vcp = None if self._sizing else kvc.reshape(-1)[HD:]
            self.e_attention(q=qkv16, ldq=HD, k=kvc, ldk=2 * HD, k_batch_stride=L * 2 * HD, vt=vcp, ldvt=2 * HD, v_rowmajor=1,
                             v_batch_stride=L * 2 * HD, out=att16, ldo=HD * kxo, B=B, H=Hh, Nq=HW, Nk=L, dpad=dpad,
                             scale=float(d) ** -0.5, split3_out=int(x3o))
            slots = self.e_gemm(a=att16, w=g(".attn2.out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=HD, out32=nxt,
                                bias=g(".attn2.out.bias"), res32=cur, flags=fo | (s3 if fold else 0), **rawkw())
            cur, nxt = nxt, cur
            # --- GEGLU feed-forward ---
            if not fold:
                self.e_layernorm(cur, M, Cc, g(".norm3.weight"), g(".norm3.bias"), tok16, x3t)
            self.e_gemm(a=tok16, w=g(".ff1.weight"), mode=_C.GEMM_PLAIN, M=M, N=2 * inner, K=Cc, block_n=2 * geglu_half(inner, x3t),
                        out16=ff16, flags=_C.GEMM_F_GEGLU | s3f | x3,
                        **(lnkw(q + ".ff1", slots) if fold else dict(bias=g(".ff1.bias"))))
            # the last block's feed-forward also emits the fp16 operand of proj_out, in proj_out's operand format; an inner block's emits
            # the raw planes + row statistics of the next block's first LayerNorm
            if last:
                self.e_gemm(a=ff16, w=g(".ff2.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=inner, out32=nxt, bias=g(".ff2.bias"),
                            res32=cur, out16=tok16, flags=(_C.GEMM_F_SPLIT3OUT if x3p else 0) | ff_)
            else:
                slots = self.e_gemm(a=ff16, w=g(".ff2.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=inner, out32=nxt, bias=g(".ff2.bias"),
                                    res32=cur, flags=ff_ | (s3 if fold else 0), **rawkw())
            cur, nxt = nxt, cur
        self.e_gemm(a=tok16, w=self.w.get(p + ".proj_out.weight"), mode=_C.GEMM_PLAIN, M=M, N=Cc, K=Cc, out32=out,
                    bias=self.w.get(p + ".proj_out.bias"), res32=x, flags=fp)

    def _emit(self, unet):
        B, H, W = self.B, self.H, self.W
        L_ = self.L
        if not self._sizing:
            self.prog = _Program()
        # staged NCHW inputs: the latent and the c_concat channels live in separate buffers so the samplers can update
        # the latent in place; the input conv reads both (DiffusionWrapper's th.cat never materialises)
        self.lat_ch = min(self.in_ch, self.out_ch)
        x_lat = self.buf("x_lat", (B, self.lat_ch, H, W))
        x_cat = self.buf("x_cat", (B, max(self.in_ch - self.lat_ch, 1), H, W))
        t_in = self.buf("t_in", (B,), torch.int64)
        temb = self.buf("temb", (B, self.mc))
        emb1 = self.buf("emb1", (B, 4 * self.mc))
        emb = self.buf("emb", (B, 4 * self.mc))
        emb_all = self.buf("emb_all", (B, self.emb_total))
        eps = self.buf("eps", (B, self.out_ch, H, W))
        if not self._sizing:
            P = self.prog
            P.add(L_.upgpt_timestep_embedding, t_in.data_ptr(), B, self.mc, 10000.0, self.w["temb.freqs"].data_ptr(), temb.data_ptr())
            P.add(L_.upgpt_linear_small_m, temb.data_ptr(), self.mc, B, self.w["time_embed.0.weight"].data_ptr(),
                  self.w["time_embed.0.bias"].data_ptr(), 4 * self.mc, self.mc, 0, 1, emb1.data_ptr(), 4 * self.mc)
            P.add(L_.upgpt_linear_small_m, emb1.data_ptr(), 4 * self.mc, B, self.w["time_embed.2.weight"].data_ptr(),
                  self.w["time_embed.2.bias"].data_ptr(), 4 * self.mc, 4 * self.mc, 0, 1, emb.data_ptr(), 4 * self.mc)
            # `emb` only ever feeds the ResBlocks' emb_layers = Linear(SiLU(emb)) (openaimodel.py:218-224): the SiLU is applied once
            # at the output of time_embed (silu_out above) instead of once per output feature of the 22 projections
            P.add(L_.upgpt_linear_small_m, emb.data_ptr(), 4 * self.mc, B, self.w["emb_all.weight"].data_ptr(),
                  self.w["emb_all.bias"].data_ptr(), self.emb_total, 4 * self.mc, 0, 0, emb_all.data_ptr(), self.emb_total)
        # the launches above depend on the timestep only: samplers may run them once per schedule (sampler_engine.py)
        self.n_emb_calls = 0 if self._sizing else len(self.prog.calls)     # (no fork / join edges among them)
        self.gn_begin()
        # ---- input blocks ----
        hs = []
        h = self.buf("h_in0", (B, H * W, self.mc))
        if not self._sizing:
            ncat = self.in_ch - self.lat_ch
            self.prog.add(L_.upgpt_conv_small_cin, x_lat.data_ptr(), self.lat_ch, x_cat.data_ptr() if ncat else 0, ncat, 1.0, B, H, W, 3,
                          self.w["conv_in.weight"].data_ptr(), self.w["conv_in.bias"].data_ptr(), self.mc, h.data_ptr(), 0)
        ch, ch_h, ch_w = self.mc, H, W
        hs.append((h, ch, ch_h, ch_w))

        def run_layers(prefix, layers, h, ch, hh, ww, skip=None):
            """Runs one TimestepEmbedSequential; `skip` = (tensor, channels) concatenated onto the first ResBlock input."""
            for j, mod in enumerate(layers):
                p = f"{prefix}.{j}"
                if isinstance(mod, om.ResBlock):
                    out = self.buf(p + ".out", (B, hh * ww, mod.out_channels))
                    if skip is not None:
                        self._res_block(p, mod, h, ch, skip[0], skip[1], B, hh, ww, out)
                        skip = None
                    else:
                        self._res_block(p, mod, h, ch, None, 0, B, hh, ww, out)
                    h, ch = out, mod.out_channels
                elif isinstance(mod, SpatialTransformer):
                    out = self.buf(p + ".out", (B, hh * ww, ch))
                    self._transformer(p, mod, h, ch, B, hh, ww, out)
                    h = out
                elif isinstance(mod, om.Downsample):
                    x3c = self.use_x3("conv_updown", (hh // 2) * (ww // 2))
                    op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=2, split3=x3c)
                    hh, ww = hh // 2, ww // 2
                    out = self.buf(p + ".out", (B, hh * ww, mod.out_channels))
                    self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3_S2PHASE, N=mod.out_channels, K=ch, n_imgs=B,
                                H=hh, W=ww, out32=out, bias=self.w.get(p + ".bias"), flags=_C.GEMM_F_X3 if x3c else 0)
                    h, ch = out, mod.out_channels
                elif isinstance(mod, om.Upsample):
                    x3c = self.use_x3("conv_updown", hh * ww * 4)
                    out = self.buf(p + ".out", (B, hh * ww * 4, mod.out_channels))
                    if UP2_FOLD:
                        # the x2 replicate is never materialised: four parity-wise 2x2 convolutions over the low-resolution operand
                        op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=0, split3=x3c)
                        self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3_UP2, N=mod.out_channels, K=ch, n_imgs=B, H=hh,
                                    W=ww, out32=out, bias=self.w.get(p + ".bias"), flags=_C.GEMM_F_X3 if x3c else 0)
                        hh, ww = hh * 2, ww * 2
                    else:
                        op, _ = self.norm_operand(h, ch, None, 0, B, hh, ww, None, 0.0, False, layout=1, split3=x3c)
                        hh, ww = hh * 2, ww * 2
                        self.e_gemm(a=op, w=self.w.get(p + ".weight"), mode=_C.GEMM_CONV3X3, N=mod.out_channels, K=ch, n_imgs=B, H=hh,
                                    W=ww, out32=out, bias=self.w.get(p + ".bias"), flags=_C.GEMM_F_X3 if x3c else 0)
                    h, ch = out, mod.out_channels
                else:
                    raise NotImplementedError(type(mod))
            return h, ch, hh, ww

        for i in range(1, len(unet.input_blocks)):
            h, ch, ch_h, ch_w = run_layers(f"input_blocks.{i}", unet.input_blocks[i], h, ch, ch_h, ch_w)
            hs.append((h, ch, ch_h, ch_w))
        h, ch, ch_h, ch_w = run_layers("middle_block", unet.middle_block, h, ch, ch_h, ch_w)
        for i, blk in enumerate(unet.output_blocks):
            s, sc, sh, sw = hs.pop()
            assert (sh, sw) == (ch_h, ch_w)
            h, ch, ch_h, ch_w = run_layers(f"output_blocks.{i}", blk, h, ch, ch_h, ch_w, skip=(s, sc))
        # ---- out: GN + SiLU + conv3x3 -> eps (NCHW fp32) ----
        op, _ = self.norm_operand(h, ch, None, 0, B, H, W, "out.0", 1e-5, True, split3=self.split3)
        self.e_gemm(a=op, w=self.w.get("out.conv.weight"), mode=_C.GEMM_CONV3X3, N=self.out_ch, K=ch,
                    n_imgs=B, H=H, W=W, block_n=16, splits=1, out32=eps, bias=self.w.get("out.conv.bias"), flags=_C.GEMM_F_CHW | self.x3)
        self.gn_end()

    def _emit_context(self, unet):
        """Program that fills the per-layer context K / V^T caches (timestep-invariant: attention.py:162-163,175-176)."""
        B, L, Lp = self.B, self.ctx_len, _round_up(self.ctx_len, 8)
        main = self.prog
        self.prog = _Program()
        kx, x3 = self.kx, self.x3
        ctx32 = self.buf("ctx32", (B, L, self.ctx_dim))
        ctx16 = self.buf("ctx16", (B * L, self.ctx_dim * kx), torch.float16)
        self.e_prep(ctx32, self.ctx_dim, None, 0, B, 1, L, None, None, None, 0.0, False, 0, ctx16, split3=self.split3)
        self.ctx_prep = self.prog          # fp32 context -> fp16 operand planes (shared by full rebuilds and row refreshes)
        self.prog = _Program()
        layers = []
        for name, mod in unet.named_modules():
            if isinstance(mod, SpatialTransformer):
                dpad = head_pad(mod.d_head, mod.n_heads)
                HD = mod.n_heads * dpad
                for bi in range(len(mod.transformer_blocks)):
                    q = f"{name}.transformer_blocks.{bi}"
                    self.e_gemm(a=ctx16, w=self.w[q + ".attn2.kv.weight"], mode=_C.GEMM_PLAIN, M=B * L, N=2 * HD, K=self.ctx_dim,
                                out16=self.bufs[q + ".ctx_kv"], flags=x3)
                    layers.append((q, self.w[q + ".attn2.kv.weight"], self.bufs[q + ".ctx_kv"], HD))
        self.ctx_prog = self.prog
        self.prog = main
        from .cond_cache import CondCache
        self.cond = CondCache(self, layers)

    # ------------------------------------------------------------------------------------------------ execution
    def _stream(self):
        if self.dry or self.dev.type != "cuda":
            raise _C.UpgptError("UNetEngine recorded without a CUDA device: nothing to run (no CPU fallback)")
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def set_context(self, context, force=False):
        """Makes the cross-attention K | V cond-cache (upgpt_b200/cond_cache.py) hold the projections of `context` (B, L, ctx_dim):
        nothing when the same tensor is passed again, a refresh of the changed rows only when it differs from the resident context in a
        few rows (the SMPL token of an interpolation keyframe, a mixed style slot), else one k | v GEMM per layer."""
        key = (context.data_ptr(), context._version, tuple(context.shape))
        if not force and key == self._ctx_key and self.cond.valid:
            return
        if os.environ.get("UPGPT_COND_CACHE_ROWS", "1") == "0":
            self.cond.valid = False        # (benchmark knob: rebuild every layer's K | V on any change)
        self.cond.set_context(context, force)
        self._ctx_key = key
        # the key is only sound while the keyed tensor is alive: a freed context's address is handed to the next one by the caching
        # allocator (same ptr, version 0, same shape -> a stale cond-cache for the next keyframe of an interpolation sequence)
        self._ctx_ref = context

    def stage_inputs(self, x, timesteps, c_concat=None):
        """x: (B, in_ch, H, W) already concatenated, or the (B, lat_ch, H, W) latent with c_concat given separately."""
        if c_concat is None and x.shape[1] == self.in_ch and self.in_ch > self.lat_ch:
            self.bufs["x_lat"].copy_(x[:, :self.lat_ch])
            self.bufs["x_cat"].copy_(x[:, self.lat_ch:])
        else:
            self.bufs["x_lat"].copy_(x)
            if c_concat is not None:
                self.bufs["x_cat"].copy_(c_concat)
        if timesteps is not None:
            self.bufs["t_in"].copy_(timesteps)

    def run(self, use_graph=True):
        """Executes one U-Net pass over the staged inputs; result in self.bufs['eps']."""
        if use_graph:
            if self.graph is None:
                from .ops import Graph
                self.prog.run(self._stream())   # warm-up (lazy module load, attribute setup) outside capture
                self.graph = Graph().capture(lambda: self.prog.run(self._stream()))
            self.graph.launch()
        else:
            self.prog.run(self._stream())
        return self.bufs["eps"]

    def run_calls(self, lo, hi, stream=None):
        """Runs launches [lo, hi) of the recorded program (hi None = to the end)."""
        stream = stream or self._stream()
        for fn, args in self.prog.calls[lo:hi]:
            rc = fn(*args, stream)
            if rc != 0:
                raise _C.UpgptError("%s failed (%d): %s" % (fn.__name__, rc, _C.lib().upgpt_last_error().decode()))

    def forward(self, x, timesteps, context, use_graph=None):
        """x (B, in_channels, H, W) fp32 NCHW (latent already concatenated with c_concat), timesteps (B,), context (B,L,D)."""
        if use_graph is None:
            use_graph = os.environ.get("UPGPT_NO_GRAPH", "0") != "1"
        self.set_context(context)
        self.stage_inputs(x, timesteps)
        return self.run(use_graph).clone()

    @property
    def launches_per_step(self):
        return self.prog.n_kernels
