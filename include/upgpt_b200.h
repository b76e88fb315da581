/* upgpt_b200 C ABI — B200 (sm_100a) kernels for UPGPT's denoising hot path.
 *
 * The reference (soon-yau/upgpt) is pure Python/PyTorch and has no FFI of its own; the seam it offers is the
 * config-instantiation protocol (ldm/util.py:78-93) plus the nn.Module call surface.  This header is the C ABI the
 * Python mirror of that surface (package `ldm/` + `upgpt_b200/`) binds with ctypes: plain pointers and sizes, no torch
 * types.  Every entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated; tensors are dense row-major; "NHWC" = [n][h][w][c]
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it (CUDA-graph capturable)
 *   - return 0 on success, negative on error; upgpt_last_error() returns the message of the last failure
 *   - fp16 = IEEE binary16 (`__half`), the tensor-core operand type; accumulation and all statistics are fp32
 */
#ifndef UPGPT_B200_H_
#define UPGPT_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* upgpt_last_error(void);
/* ABI version of this header; bumped on any struct change. */
int upgpt_abi_version(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches` evidence) */
long long upgpt_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * tcgen05 GEMM / implicit-GEMM convolution
 *   replaces: F.conv2d at openaimodel.py:116-118,151-153,204,230,241,519,685; nn.Linear / 1x1 conv at
 *   attention.py:41,54-62,161-168,233-248; VAE decoder convs at model.py:42-57,82-141,150-202,535-568.
 * ------------------------------------------------------------------------------------------------------------------ */
enum {
  UPGPT_GEMM_PLAIN = 0,           /* A [batch][M][K] tokens, W [batch?][N][K] */
  UPGPT_GEMM_CONV3X3 = 1,         /* A NHWC [n_imgs][H][W][K], W [N][9][K], stride 1, zero pad 1 */
  UPGPT_GEMM_CONV3X3_S2PHASE = 2, /* stride-2 conv; A holds the 4 stride-2 phases [4][n_imgs][H][W][K] (H,W = OUTPUT size) */
  UPGPT_GEMM_CONV1X1 = 3,         /* A NHWC, W [N][K] (image addressing, used when the epilogue needs per-image rows) */
  UPGPT_GEMM_CONV3X3_S2PHASE_ASYM = 4, /* stride-2 conv with zero pad (0,1,0,1) = right/bottom only (VAE Encoder Downsample, model.py:59-79):
                                      out(y,x) tap (r,s) reads in(2y+r, 2x+s); A = the 4 stride-2 phases as in S2PHASE */
  UPGPT_GEMM_CONV3X3_UP2 = 5      /* nearest x2 upsample + 3x3 conv (Upsample, openaimodel.py:116-118; model.py:49-52) WITHOUT the upsampled
                                      tensor: A = the LOW-resolution NHWC operand [n_imgs][H][W][K] (H, W = INPUT size), the result is
                                      [n_imgs][2H * 2W][N]. Output pixel (2y+a, 2x+b) only sees the 2x2 source pixels (y+a-1+u, x+b-1+v):
                                      four parity-wise 2x2 convolutions, 16 instead of 36 tap products per source pixel (2.25x fewer MACs).
                                      W = [4 parities a*2+b][N][4 taps u*2+v][K]: the sums of the 3x3 taps that land on one source pixel
                                      (rows {0 | 1+2} for a = 0, {0+1 | 2} for a = 1; same for columns), made by the caller */
};
enum {
  UPGPT_GEMM_F_GEGLU = 1u << 1, /* W rows packed per tile as [x | gate]; out16 = x * gelu(gate)   (attention.py:37-44) */
  UPGPT_GEMM_F_CHW = 1u << 2,   /* store channel-major: out[(group*N + n)*ldT + row_in_group] (NCHW images, V^T for attention) */
  UPGPT_GEMM_F_SPLIT3OUT = 1u << 4, /* out16 rows = [hi | lo] fp16 planes (N columns each, ld16 default 2N; x ~= hi + lo to ~22 bits):
                                      the A operand of a following UPGPT_GEMM_F_X3 GEMM */
  UPGPT_GEMM_F_W_STATIC = 1u << 6,  /* w holds model weights, i.e. nothing a preceding launch of the stream writes: the kernel may fetch it
                                      before its programmatic-dependency wait (weight stream overlaps the predecessor's tail) */
  UPGPT_GEMM_F_X3 = 1u << 5         /* error-compensated fp16x3 product (the precision mode that meets the 1e-3 eps tolerance):
                                      A rows = [Ah | Al] planes of K columns (lda default 2K), W rows per tap = [Wh | Wl] (ldw
                                      default 2K); D = Ah*Wh + Al*Wh + Ah*Wl in fp32: 3 MMAs per k-step on 2 loaded plane pairs */
};
typedef struct upgpt_gemm_args {
  const void* a;            /* fp16 activations */
  const void* w;            /* fp16 weights */
  int mode;                 /* UPGPT_GEMM_* */
  int M, N, K;              /* PLAIN: rows per batch entry / output cols / reduction. CONV: M ignored, K = Cin */
  int batch;                /* PLAIN only (>=1) */
  long long a_batch_stride; /* elements; 0 = M*lda */
  long long w_batch_stride; /* elements between per-batch W matrices; 0 = N*taps*ldw (dense) */
  int lda;                  /* elements between A rows / pixels (0 = K) */
  int ldw;                  /* elements between W taps (0 = K) */
  int n_imgs, H, W;         /* CONV: output images and spatial size */
  int block_n;              /* UMMA N tile (0 = auto) */
  int splits;               /* split-K factor (0 = auto; >1 requires fp32-only output, zero-filled by the caller) */
  float* out32;             /* fp32 result [rows][ld32] or NULL */
  int ld32;
  void* out16;              /* fp16 copy of the result or NULL */
  int ld16;
  const float* bias;        /* [N] or NULL */
  const float* rowvec;      /* [groups][ld_rowvec] added to every row of a group (timestep embedding, openaimodel.py:264-273) */
  int ld_rowvec;
  int rows_per_group;       /* 0 = H*W (conv) or M (plain) */
  const float* res32;       /* residual added to the result [rows][ldres] or NULL */
  int ldres;
  int ldT;                  /* CHW stores: elements between channels (0 = rows_per_group) */
  unsigned flags;           /* UPGPT_GEMM_F_* */
  float out_scale;          /* accumulator scale before bias (0 = 1.0) */
  /* ---- LayerNorm folded into the GEMMs around it (attention.py:203-205,211-215: x + attn(LN(x)), x + ff(LN(x))) ----
   * producer (the GEMM that writes the residual stream x): rowstats_out[row][t][2] = {sum, sum of squares} of the fp32 result row over
   * the columns of N tile t (t < n_tiles of upgpt_gemm_plan), written with plain stores in a fixed order (bit-reproducible).
   * consumer (the GEMM that reads LN(x)): a = fp16 planes of the RAW x, w = gamma-scaled weights W'[n][k] = gamma[k] W[n][k],
   * ln_colsum[n] = sum_k W'[n][k] (of the fp16-rounded planes), bias[n] = b[n] + sum_k beta[k] W[n][k]; per row the epilogue forms
   *   mean = sum / K, rstd = rsqrt(sumsq / K - mean^2 + ln_eps),   D = rstd * (acc - mean * ln_colsum[n]) + bias[n]
   * which equals LN(x) W^T + b with the statistics of the fp32 x: no LayerNorm kernel and no normalised copy of x exist. */
  float* rowstats_out;      /* [rows][n_tiles][2] or NULL */
  const float* ln_stats;    /* the producer's rowstats_out, or NULL (no folded LayerNorm) */
  int ln_slots;             /* partial slots per row of ln_stats (= the producer's n_tiles) */
  float ln_eps;
  const float* ln_colsum;   /* [N] */
  /* ---- GroupNorm statistics from the epilogue (openaimodel.py:201-203,225-227; attention.py:254): the kernel that PRODUCES a tensor a
   * GroupNorm will read adds the per-(image, group) moments of its fp32 result to 64-bit fixed-point accumulators
   *   gn_acc[(img * gn_groups + g) * 2 + {0, 1}] += {sum * 2^24, sumsq * 2^20},   g = (gn_choff + n) / gn_cpg,  img = row / rows_per_group
   * (per-tile partial sums in fp32 in a fixed order, then integer atomics: associative, hence bit-reproducible). The GroupNorm itself
   * is then upgpt_prep_operand(gn_acc = ...): apply only, no statistics pass. gn_acc2: a second consumer with its own grouping (an
   * encoder output also feeds a decoder ResBlock's concatenated input). Needs an fp32 row-major result and rows_per_group = H*W. */
  long long* gn_acc;
  int gn_groups, gn_cpg, gn_choff;
  long long* gn_acc2;
  int gn_cpg2, gn_choff2;
  int rowstats_slots;       /* 0, or the number of N tiles (= partial slots per row of rowstats_out) the caller sized its consumers for: the
                               launch fails if the tiling picked now differs (e.g. the process-wide tiling objective was switched after the
                               program was recorded) instead of feeding a consumer the wrong number of partials */
} upgpt_gemm_args;
int upgpt_gemm(const upgpt_gemm_args* args, void* stream);
/* the tiling upgpt_gemm picks for these arguments on the current device, without launching: plan[0] = block_n, plan[1] = n_tiles
 * (N tiles = rowstats slots per row), plan[2] = split-K factor, plan[3] = pipeline stages, plan[4] = epilogue path (1 TMA-store,
 * 2 cluster split-K reduction, 0 other: only 1 and 2 can emit rowstats_out / gn_acc), plan[5..7] reserved */
int upgpt_gemm_plan(const upgpt_gemm_args* args, int plan[8]);
/* Tiling objective of every following upgpt_gemm / upgpt_gemm_plan of the process: 0 (default) = the latency of one launch that has the
 * GPU to itself; w > 0 = latency x (1 + w x CTAs / SMs): SM time counts, for several independent batches in flight on one GPU
 * (upgpt_b200/lanes.py sets 8). Set it before programs are recorded / graphs captured: a captured graph keeps the tiling it was built with. */
int upgpt_gemm_set_sm_weight(double w);
/* Programmatic dependent launch of every following launch of the library (default on; UPGPT_PDL=0): a successor's CTAs become resident
 * while their predecessor drains -- good for one chain's latency, but with several batches in flight those waiting CTAs hold SMs another
 * batch's kernel could use (throughput mode switches it off: +1.8 %). Captured graphs keep what they were built with. */
int upgpt_set_pdl(int on);
/* bring-up instrumentation: CTA c of every following upgpt_gemm stamps %globaltimer (ns) into buf[c*16 + slot]; NULL = off */
int upgpt_debug_set_gemm_timestamps(long long* buf);
/* bring-up instrumentation: launch trace inside dependent chains / graph replays. buf (device memory, 8-byte words): buf[0] = counter
 * (zero it), buf[1] = capacity, buf[2 + i] = (%globaltimer ns << 2) | kind; kind 0 = block 0 of a kernel entered, kind 1 = its
 * griddepcontrol.wait returned (predecessor drained). Differences of consecutive kind-1 stamps = effective cost per launch. NULL = off */
int upgpt_trace_set(unsigned long long* buf);

/* ------------------------------------------------------------------------------------------------------------------
 * Normalisation / operand preparation (HBM-bound elementwise + reductions)
 *   replaces: GroupNorm32 (util.py:199-216), Normalize eps 1e-6 (attention.py:76-77, model.py:38-39), SiLU / swish
 *   (openaimodel.py:201-203,225-227; model.py:33-35), LayerNorm (attention.py:203-205), nearest x2 upsample
 *   (openaimodel.py:116; model.py:53), the skip concat th.cat([h, hs.pop()],1) (openaimodel.py:736), softmax (model.py:184)
 * ------------------------------------------------------------------------------------------------------------------ */
/* stats[b][g] = {sum, sum of squares} (double) over group g of image b of the channel-concat [x1 | x2] (x2 may be NULL). */
int upgpt_groupnorm_stats(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                          void* stream);
/* same, and also scale_shift[b] = {scale[C], shift[C]} with y = x*scale + shift == GroupNorm(x; gamma, beta, eps) */
int upgpt_groupnorm_affine(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                           const float* gamma, const float* beta, float eps, float* scale_shift, void* stream);
typedef struct upgpt_prep_args {
  const float* x1; int C1;   /* fp32 NHWC [B][H][W][C1] */
  const float* x2; int C2;   /* optional second tensor, concatenated on channels */
  int B, H, W;
  int groups;                /* GroupNorm groups (32) */
  const double* stats;       /* from upgpt_groupnorm_stats, or NULL = no normalisation (plain cast) */
  const float* gamma; const float* beta; float eps;
  int silu;                  /* apply x*sigmoid(x) after the affine */
  int layout;                /* 0 same; 1 nearest-x2 upsampled [B][2H][2W][C]; 2 stride-2 phases [4][B][H/2][W/2][C] */
  int split3;                /* emit error-compensated operand planes [hi | lo] (2C channels) */
  void* out; int ldo;        /* fp16 output, ldo elements per pixel (0 = C or 2C) */
  void* raw; int ldraw;      /* optional un-normalised fp16 copy (layout 0) */
  const float* scale_shift;  /* optional [B][2][C] affine from upgpt_groupnorm_affine (then stats/gamma/beta/eps are ignored) */
  const long long* gn_acc;   /* optional [B][groups][2] fixed-point group moments {sum * 2^24, sumsq * 2^20} accumulated by the kernels that
                                PRODUCED x1 / x2 (upgpt_gemm's gn_acc, upgpt_gn_accumulate): GroupNorm(gamma, beta, eps) is applied from
                                them -- no statistics pass over the tensor, no reduction in this kernel */
  int raw_planes;            /* format of `raw`: 0 = as `out` (split3), 1 = one fp16 plane, 2 = [hi | lo] planes. A ResBlock's skip 1x1 GEMM
                                (openaimodel.py:241) runs error-compensated on the raw copy while conv1 may take a single-plane operand */
} upgpt_prep_args;
int upgpt_prep_operand(const upgpt_prep_args* args, void* stream);
/* Adds the per-(image, group) moments of x [B][HW][C] (channels choff .. choff + C of a GroupNorm over `groups` groups of `cpg` channels)
 * to acc[b][g] = {sum * 2^24, sumsq * 2^20} with 64-bit integer atomics: integer addition is associative, so the accumulated moments
 * are bit-reproducible whatever the arrival order. For tensors no GEMM epilogue produced (the input convolution). */
int upgpt_gn_accumulate(const float* x, int C, int B, int HW, int groups, int cpg, int choff, long long* acc, void* stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (graph-capturable): re-arms the moment accumulators at the start of a pass */
int upgpt_zero(void* ptr, long long bytes, void* stream);
/* GroupNorm(gamma, beta, eps) [+ SiLU] + cast of `args` in ONE call (args->stats / scale_shift are ignored; `stats` receives the
 * {sum, sumsq} moments). One launch when an image's [H*W][C] fp32 tile fits the shared memory of a thread-block cluster (one
 * cluster per image: TMA bulk load, moments exchanged over distributed shared memory, normalised out of shared memory), else
 * upgpt_groupnorm_affine + upgpt_prep_operand internally.   replaces: GroupNorm32 + SiLU (openaimodel.py:201-203,225-227) */
int upgpt_groupnorm_prep(const upgpt_prep_args* args, double* stats, void* stream);
int upgpt_layernorm(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                    void* out16, int ldo, void* stream);
/* same with out16 rows = [hi | lo] planes of C columns (ldo default 2C) */
int upgpt_layernorm_split3(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                           void* out16, int ldo, void* stream);
/* out16[r][i] = softmax_i(scale * x[r][i]) */
int upgpt_softmax_rows(const float* x, int ldx, long long rows, int n, float scale, void* out16, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Flash-style attention on tcgen05/TMEM:  out = softmax(scale * Q K^T) V per (batch, head)
 *   replaces: CrossAttention.forward einsum/softmax/einsum (attention.py:178-192), self (attn1) and cross (attn2)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct upgpt_attn_args {
  const void* q;  int ldq;   /* fp16 [B][Nq][ldq], head h at columns [h*dpad, (h+1)*dpad) */
  const void* k;  int ldk;   /* fp16 [B][Nk][ldk], same head layout */
  long long k_batch_stride;  /* elements between batches of K (0 = Nk*ldk) */
  const void* vt; int ldvt;  /* fp16 V transposed [B][H*dpad][ldvt], keys contiguous (ldvt >= Nk, multiple of 8) */
  void* out;      int ldo;   /* fp16 [B][Nq][ldo] */
  int B, H, Nq, Nk;
  int dpad;                  /* head dim padded with zero columns to 64 or 128 */
  float scale;               /* dim_head ** -0.5 (attention.py:157) */
  int split3_out;            /* out rows = [hi | lo] planes of H*dpad columns each (ldo >= 2*H*dpad) */
  int v_rowmajor;            /* 1: `vt` holds V row-major [B][Nk][ldvt] with the head layout of K (e.g. a slice of one fused q|k|v
                                projection); the P V product then reads it as an MN-major tensor-core operand, no transpose anywhere */
  long long v_batch_stride;  /* v_rowmajor: elements between batches of V (0 = Nk*ldvt) */
} upgpt_attn_args;
int upgpt_attention(const upgpt_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Boundary convolutions, timestep-embedding MLP, sampler updates, device step state, CUDA-graph helpers
 * ------------------------------------------------------------------------------------------------------------------ */
/* Direct k x k (k = 1 or 3, zero pad k/2) convolution for Cin = C1 + C2 <= 8. Inputs NCHW fp32 (x1 scaled by in_scale,
 * x2 = concatenated extra channels, may be NULL); weights k-major [Cin*k*k][Cout] fp32; out NHWC (or NCHW) fp32.
 *   replaces: DiffusionWrapper hybrid concat + input conv (ddpm.py:1567-1570, openaimodel.py:519); z/scale_factor +
 *   post_quant_conv (ddpm.py:779, autoencoder.py:330-333); VAE Decoder.conv_in (model.py:491-495) */
int upgpt_conv_small_cin(const float* x1, int C1, const float* x2, int C2, float in_scale, int B, int H, int W, int ksize,
                         const float* wt_kmajor, const float* bias, int Cout, float* out, int out_nchw, void* stream);
/* out[b] = [cos(t_b f_k) | sin(t_b f_k)]  (util.py:151-171); t is int64 on the device. freqs: optional device table of the
 * dim/2 fp32 frequencies (pass the host-computed torch table for bit-identical arguments), NULL = computed on the device */
int upgpt_timestep_embedding(const long long* t, int B, int dim, float max_period, const float* freqs, float* out, void* stream);
/* out[r][n] = act_out(sum_k act_in(x[r][k]) W[n][k] + bias[n]); fp32, rows <= 64 (time_embed, emb_layers, LinearProject) */
int upgpt_linear_small_m(const float* x, int ldx, int rows, const float* w, const float* bias, int N, int K, int silu_in,
                         int silu_out, float* out, int ldo, void* stream);
/* DDIM update (ddim.py:189-203). coef[step] = {a_t, a_prev, sigma_t, sqrt(1-a_t), temperature}; the row index is
 * *step_ptr when step_ptr != NULL (device int, for graph replay) else step_imm. noise may be NULL (eta = 0);
 * otherwise noise + step*noise_step_stride holds this step's N(0,1) draws. pred_x0 may be NULL. */
int upgpt_ddim_step(const float* x, const float* eps, const float* noise, long long noise_step_stride, const float* coef,
                    const int* step_ptr, int step_imm, float* x_prev, float* pred_x0, long long n, void* stream);
/* DDPM ancestral update (ddpm.py:224-237,1125-1185, clip_denoised=False). coef[step] = {sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped, t != 0}. */
int upgpt_ddpm_step(const float* x, const float* eps, const float* noise, long long noise_step_stride, const float* coef,
                    const int* step_ptr, int step_imm, float* x_prev, float* pred_x0, long long n, void* stream);
/* op 0: *step = value; op 1: *step += value. Then, if t_buf: t_buf[0..B) = t_table[*step] (ddim.py:142). */
int upgpt_step_state(int* step_ptr, int op, int value, long long* t_buf, int B, const long long* t_table, void* stream);
/* dst[b][0..n) = table[*step_ptr * row_stride + 0..n) for b < B (n, row_stride multiples of 4): the per-step row of a table computed
 * once per schedule, e.g. emb_layers(SiLU(time_embed(timestep_embedding(t_step)))) of all 22 ResBlocks (openaimodel.py:218-224,723-724) --
 * t is the same for every sample of a batch in ddim.py:142 / ddpm.py:1271, so the 4 embedding launches leave the step graph */
int upgpt_gather_step_row(const float* table, long long row_stride, const int* step_ptr, float* dst, int B, int n, void* stream);
/* q_sample with an optional known-region blend, one launch:
 *   q = sqrt_acp[t] * x0 + sqrt_1m_acp[t] * noise                  (DDPM.q_sample ddpm.py:281-284; stochastic_encode ddim.py:207-221)
 *   out = mask ? q * mask + (1 - mask) * img : q                    (ddim.py:144-147, ddpm.py:1281-1284)
 * x0 / noise / img / out: [B][C][HW] fp32; mask: [B][mask_c][HW] with mask_c = 1 (broadcast over channels) or C.
 * t: per sample from t_per_sample[B] (int64), else t_table[*step_ptr] (device-side step counter of the sampler graphs; noise is then
 * read at noise + *step_ptr * noise_step_stride), else t_imm. sqrt_acp / sqrt_1m_acp: the schedule buffers (device, fp32). */
int upgpt_qsample_blend(const float* x0, const float* noise, long long noise_step_stride, const float* mask, int mask_c,
                        const float* img, float* out, const float* sqrt_acp, const float* sqrt_1m_acp,
                        const long long* t_per_sample, const long long* t_table, const int* step_ptr, int t_imm, int B, int C, int HW,
                        void* stream);
/* out = a*sa + b*sb (b may be NULL): classifier-free guidance combine, latent mirror */
int upgpt_axpby(const float* a, float sa, const float* b, float sb, float* out, long long n, void* stream);
/* out = (wa*a + wb*b + wc*c + wd*d) / den, terms with a NULL pointer skipped: the pseudo linear multistep eps combinations of
 * PLMSSampler.p_sample_plms (plms.py:211-229), e.g. (55 e_t - 59 e_1 + 37 e_2 - 9 e_3) / 24 */
int upgpt_lincomb4(const float* a, float wa, const float* b, float wb, const float* c, float wc, const float* d, float wd, float den,
                   float* out, long long n, void* stream);
/* VAE posterior: moments NCHW [B][2C][HW] = {mean | logvar} -> out [B][C][HW] = (mean + exp(0.5*clamp(logvar,-30,20))*noise)*out_scale;
 * noise == NULL gives the mode (mean*out_scale).  replaces DiagonalGaussianDistribution.sample/.mode (distributions.py:24-37) and
 * the scale_factor multiply of get_first_stage_encoding (ddpm.py:569-576) */
int upgpt_gaussian_sample(const float* moments, const float* noise, float out_scale, float* out, int B, int C, int HW, void* stream);
/* clamp(-1,1)*0.5+0.5 -> uint8 NHWC (generate_utils.py:165-168) */
int upgpt_to_uint8_nhwc(const float* x, int B, int C, int HW, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * CLIP ViT-L/14 conditioning towers: the CUDA-core pieces around upgpt_gemm / upgpt_attention.
 *   replaces: FrozenCLIPEmbedder.forward -> transformers.CLIPTextModel (ldm/modules/encoders/modules.py:137-162) and
 *   FrozenClipImageEmbedder2.forward -> clip.model.VisionTransformer (modules.py:234-256; `clip` = openai/CLIP, not vendored)
 * ------------------------------------------------------------------------------------------------------------------ */
/* out[r] = tok_emb[ids[r]] + pos_emb[r % seq], r < rows = B*seq (CLIPTextEmbeddings); ids int64 on the device, clamped to [0, vocab) */
int upgpt_embed_tokens(const long long* ids, int rows, int seq, int vocab, const float* tok_emb, const float* pos_emb, int C,
                       float* out, void* stream);
/* im2col of the P x P stride-P patch conv: img NCHW fp32 [n][Cin][S][S] -> fp16 rows [n*(S/P)^2][Kpad], one patch per row in
 * (c, py, px) order (= conv1.weight.reshape(width, -1)), zero columns from Cin*P*P to Kpad (multiple of 8); split3: [hi | lo] planes */
int upgpt_patchify(const float* img, int n, int Cin, int S, int P, int Kpad, int split3, void* out16, int ldo, void* stream);
/* out[b][0] = cls + pos[0]; out[b][1+i] = patch[b][i] + pos[1+i] for T tokens of C channels (VisionTransformer.forward before ln_pre) */
int upgpt_vit_assemble(const float* patch, const float* cls, const float* pos, int n, int T, int C, float* out, void* stream);
/* LayerNorm with fp32 output (ln_pre / final_layer_norm) */
int upgpt_layernorm_f32(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps, float* out, int ldo,
                        void* stream);
/* out16 = x * sigmoid(1.702 x) (QuickGELU) as a GEMM operand; split3: [hi | lo] planes of C columns */
int upgpt_quick_gelu_cast(const float* x, long long rows, int C, int split3, void* out16, int ldo, void* stream);
/* fp32 softmax attention for short sequences (N <= 128, d <= 64): qkv fp32 [B][N][ld], q at column h*d, k at koff + h*d, v at
 * voff + h*d; causal: key j <= query i (CLIP text tower). out16 fp16 [B][N][ldo], head h at columns h*d; split3_out: [hi | lo] planes */
int upgpt_attention_small(const float* qkv, int ld, int koff, int voff, int B, int H, int N, int d, float scale, int causal,
                          int split3_out, void* out16, int ldo, void* stream);

/* Stream capture of a sequence of upgpt_* calls into an executable CUDA graph. */
int upgpt_capture_begin(void* stream);
int upgpt_capture_end(void* stream, void** graph_exec_out);
int upgpt_graph_launch(void* graph_exec, void* stream);
int upgpt_graph_destroy(void* graph_exec);
/* kernel nodes inside a captured graph (each launch of the graph adds this to upgpt_launch_count) */
long long upgpt_graph_kernel_count(void* graph_exec);

/* Parallel branches of a recorded program (e.g. a ResBlock's skip 1x1 GEMM beside conv1 + GroupNorm, openaimodel.py:241,275):
 * the library owns two auxiliary non-blocking streams per device. fork: the auxiliary stream waits for everything enqueued on
 * `main_stream` so far; join: `main_stream` waits for the auxiliary stream. Both are event edges, so they are captured into CUDA
 * graphs as dependencies. upgpt_gemm launches on an auxiliary stream use their own split-K workspace. */
void* upgpt_aux_stream(int aux_idx);
int upgpt_stream_fork(void* main_stream, int aux_idx);
int upgpt_stream_join(void* main_stream, int aux_idx);

#ifdef __cplusplus
}
#endif
#endif /* UPGPT_B200_H_ */
