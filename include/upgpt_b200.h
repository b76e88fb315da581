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
  UPGPT_GEMM_CONV1X1 = 3          /* A NHWC, W [N][K] (image addressing, used when the epilogue needs per-image rows) */
};
enum {
  UPGPT_GEMM_F_GEGLU = 1u << 1, /* W rows packed per tile as [x | gate]; out16 = x * gelu(gate)   (attention.py:37-44) */
  UPGPT_GEMM_F_CHW = 1u << 2    /* store channel-major: out[(group*N + n)*ldT + row_in_group] (NCHW images, V^T for attention) */
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
} upgpt_gemm_args;
int upgpt_gemm(const upgpt_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Normalisation / operand preparation (HBM-bound elementwise + reductions)
 *   replaces: GroupNorm32 (util.py:199-216), Normalize eps 1e-6 (attention.py:76-77, model.py:38-39), SiLU / swish
 *   (openaimodel.py:201-203,225-227; model.py:33-35), LayerNorm (attention.py:203-205), nearest x2 upsample
 *   (openaimodel.py:116; model.py:53), the skip concat th.cat([h, hs.pop()],1) (openaimodel.py:736), softmax (model.py:184)
 * ------------------------------------------------------------------------------------------------------------------ */
/* stats[b][g] = {sum, sum of squares} (double) over group g of image b of the channel-concat [x1 | x2] (x2 may be NULL). */
int upgpt_groupnorm_stats(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                          void* stream);
typedef struct upgpt_prep_args {
  const float* x1; int C1;   /* fp32 NHWC [B][H][W][C1] */
  const float* x2; int C2;   /* optional second tensor, concatenated on channels */
  int B, H, W;
  int groups;                /* GroupNorm groups (32) */
  const double* stats;       /* from upgpt_groupnorm_stats, or NULL = no normalisation (plain cast) */
  const float* gamma; const float* beta; float eps;
  int silu;                  /* apply x*sigmoid(x) after the affine */
  int layout;                /* 0 same; 1 nearest-x2 upsampled [B][2H][2W][C]; 2 stride-2 phases [4][B][H/2][W/2][C] */
  int split3;                /* emit error-compensated operand planes [hi | lo | hi] (3C channels) */
  void* out; int ldo;        /* fp16 output, ldo elements per pixel (0 = C or 3C) */
  void* raw; int ldraw;      /* optional un-normalised fp16 copy (layout 0) */
} upgpt_prep_args;
int upgpt_prep_operand(const upgpt_prep_args* args, void* stream);
int upgpt_layernorm(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                    void* out16, int ldo, void* stream);
/* out16[r][i] = softmax_i(scale * x[r][i]) */
int upgpt_softmax_rows(const float* x, int ldx, long long rows, int n, float scale, void* out16, int ldo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPGPT_B200_H_ */
