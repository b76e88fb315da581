// Normalisation / operand-preparation kernels (HBM-bound, coalesced, vectorised).
//
//   gn_stats   : per-(image, group) sum / sum-of-squares of an NHWC fp32 tensor (optionally the channel concat of two)
//   prep       : [GroupNorm-apply] [SiLU] fp32 NHWC -> fp16 tensor-core operand, optionally nearest-x2 upsampled,
//                split into the 4 stride-2 phases, or emitted as error-compensated hi/lo planes
//   layernorm  : per-token LayerNorm fp32 -> fp16 operand
//   softmax    : row softmax fp32 -> fp16 (VAE single-head attention)
//
// Replaces: GroupNorm32 / normalization (util.py:199-216), Normalize eps=1e-6 (attention.py:76-77, model.py:38-39),
// nn.SiLU / nonlinearity (openaimodel.py:201-203,225-227; model.py:33-35), nn.LayerNorm (attention.py:203-205),
// F.interpolate nearest x2 (openaimodel.py:116; model.py:53), th.cat([h, hs.pop()]) (openaimodel.py:736),
// softmax (model.py:184).
#include "common.cuh"
#include "../../include/upgpt_b200.h"

namespace upgpt {

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm statistics. stats[b][g] = {sum, sumsq} in double (zeroed by the launcher).
// grid = (pixel chunks, B); thread t owns channels t, t+blockDim, ... and walks the chunk's pixels (coalesced rows).
// ------------------------------------------------------------------------------------------------------------------
template <int MAXC_PER_THREAD>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int HW, int chunk,
                int groups, double* __restrict__ stats, double* __restrict__ partials, int* __restrict__ counters,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ scale_shift) {
  extern __shared__ float shc[];  // per-channel {sum, sumsq}: [2][C]
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_red[4][64];
  const int C = C1 + C2;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int nchunks = gridDim.x;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, HW);
  float s[MAXC_PER_THREAD], q[MAXC_PER_THREAD];
#pragma unroll
  for (int j = 0; j < MAXC_PER_THREAD; ++j) { s[j] = 0.f; q[j] = 0.f; }
  // 4 pixels per trip: 4 x MAXC independent loads in flight per thread
  int p = p0;
  for (; p + 4 <= p1; p += 4) {
    float v[4][MAXC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t row = (size_t)b * HW + p + u;
#pragma unroll
      for (int j = 0; j < MAXC_PER_THREAD; ++j) {
        const int c = threadIdx.x + j * 256;
        v[u][j] = 0.f;
        if (c < C) v[u][j] = c < C1 ? __ldg(x1 + row * C1 + c) : __ldg(x2 + row * C2 + (c - C1));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < MAXC_PER_THREAD; ++j) { s[j] += v[u][j]; q[j] = fmaf(v[u][j], v[u][j], q[j]); }
  }
  for (; p < p1; ++p) {
    const size_t row = (size_t)b * HW + p;
#pragma unroll
    for (int j = 0; j < MAXC_PER_THREAD; ++j) {
      const int c = threadIdx.x + j * 256;
      if (c < C) {
        const float v = c < C1 ? __ldg(x1 + row * C1 + c) : __ldg(x2 + row * C2 + (c - C1));
        s[j] += v;
        q[j] = fmaf(v, v, q[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXC_PER_THREAD; ++j) {
    const int c = threadIdx.x + j * 256;
    if (c < C) { shc[c] = s[j]; shc[C + c] = q[j]; }
  }
  __syncthreads();
  // one thread per (group, moment): fixed-order double sum over the group's channels -> this chunk's partial
  const int nout = 2 * groups;
  double* my = partials + ((size_t)b * nchunks + blockIdx.x) * nout;
  for (int i = threadIdx.x; i < nout; i += blockDim.x) {
    const int g = i >> 1, which = i & 1;
    const float* src = shc + which * C + g * cpg;
    double acc = 0.0;
    for (int k = 0; k < cpg; ++k) acc += (double)src[k];
    if (nchunks == 1) stats[(size_t)b * nout + i] = acc; else __stcg(my + i, acc);
  }
  if (nchunks > 1) {
    // deterministic cross-chunk reduction: the last chunk to arrive sums all partials of image b in a fixed order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int prev = atomicAdd(&counters[b], 1);
      s_last = prev == nchunks - 1;
      if (s_last) counters[b] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double* base = partials + (size_t)b * nchunks * nout;
    // 4 threads per output, each over an interleaved quarter of the chunks (4 loads in flight), combined in fixed order
    if (nout <= 64) {
      const int i = threadIdx.x & 63, part = threadIdx.x >> 6;
      double a0 = 0.0, a1 = 0.0;
      if (i < nout) {
        int k = part;
        for (; k + 4 < nchunks; k += 8) {
          a0 += __ldcg(base + (size_t)k * nout + i);
          a1 += __ldcg(base + (size_t)(k + 4) * nout + i);
        }
        if (k < nchunks) a0 += __ldcg(base + (size_t)k * nout + i);
      }
      s_red[part][i] = a0 + a1;
      __syncthreads();
      if (threadIdx.x < nout) stats[(size_t)b * nout + threadIdx.x] = (s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + (s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
    } else {
      for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        double a0 = 0.0;
        for (int k = 0; k < nchunks; ++k) a0 += __ldcg(base + (size_t)k * nout + i);
        stats[(size_t)b * nout + i] = a0;
      }
    }
  }
  if (!scale_shift) return;
  // per-(image, channel) affine of the normalisation, so the apply kernel is a pure fma: y = x * scale + shift
  __syncthreads();
  const double inv_n = 1.0 / ((double)cpg * HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double su = stats[(size_t)b * nout + 2 * g];
    const double sq = stats[(size_t)b * nout + 2 * g + 1];
    const double mean = su * inv_n;
    double var = sq * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    scale_shift[(size_t)b * 2 * C + c] = rstd * ga;
    scale_shift[(size_t)b * 2 * C + C + c] = be - (float)mean * rstd * ga;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// prep: normalise (optional) + SiLU (optional) + cast to fp16 with a layout transform.
// grid = (pixel chunks, B); 4 channels per thread (float4 in, 8-byte out).
// ------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* x1; int C1;
  const float* x2; int C2;
  int H, W;
  int chunk;
  int groups;
  const double* stats;     // null -> no normalisation
  const float* scale_shift;   // optional [B][2][C] precomputed affine (from gn_stats); replaces stats/gamma/beta
  const float* gamma; const float* beta;
  float eps;
  int silu;
  int layout;              // 0 same, 1 nearest-up x2, 2 stride-2 phases
  int split3;              // output channels = 2C: [hi | lo]
  __half* out; int ldo;    // elements per output pixel (>= C or 3C)
  __half* raw; int ldraw;  // optional un-normalised fp16 copy (layout 0)
  int B;
};

__global__ void __launch_bounds__(256)
prep_kernel(const PrepParams p) {
  extern __shared__ float shf[];  // scale[C], shift[C]
  pdl_launch_dependents();
  pdl_wait();
  const int C = p.C1 + p.C2;
  const int HW = p.H * p.W;
  float* scale = shf;
  float* shift = shf + C;
  const int b = blockIdx.y;
  if (p.scale_shift) {
    const float* ss = p.scale_shift + (size_t)b * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) shf[c] = ss[c];
    __syncthreads();
  } else if (p.stats) {
    const int cpg = C / p.groups;
    const double inv_n = 1.0 / ((double)cpg * HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = c / cpg;
      const double su = p.stats[((size_t)b * p.groups + g) * 2];
      const double sq = p.stats[((size_t)b * p.groups + g) * 2 + 1];
      const double mean = su * inv_n;
      double var = sq * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      const float ga = p.gamma ? p.gamma[c] : 1.f;
      const float be = p.beta ? p.beta[c] : 0.f;
      scale[c] = rstd * ga;
      shift[c] = be - (float)mean * rstd * ga;
    }
    __syncthreads();
  }
  const int C4 = C >> 2;
  const int p0 = blockIdx.x * p.chunk;
  const int p1 = min(p0 + p.chunk, HW);
  const int total = (p1 - p0) * C4;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int pp = p0 + idx / C4;
    const int c = (idx % C4) << 2;
    const size_t row = (size_t)b * HW + pp;
    float4 v = c < p.C1 ? *(const float4*)(p.x1 + row * p.C1 + c) : *(const float4*)(p.x2 + row * p.C2 + (c - p.C1));
    if (p.raw) {
      __half2 r0 = __floats2half2_rn(v.x, v.y), r1 = __floats2half2_rn(v.z, v.w);
      uint2 pk = make_uint2(*(uint32_t*)&r0, *(uint32_t*)&r1);
      *(uint2*)(p.raw + row * p.ldraw + c) = pk;
      if (p.split3) {   // the un-normalised copy feeds the skip 1x1 GEMM: same [hi | lo] planes
        const float2 f0 = __half22float2(r0), f1 = __half22float2(r1);
        __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
        *(uint2*)(p.raw + row * p.ldraw + C + c) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
      }
    }
    if (p.stats || p.scale_shift) {
      v.x = fmaf(v.x, scale[c], shift[c]);
      v.y = fmaf(v.y, scale[c + 1], shift[c + 1]);
      v.z = fmaf(v.z, scale[c + 2], shift[c + 2]);
      v.w = fmaf(v.w, scale[c + 3], shift[c + 3]);
    }
    if (p.silu) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const uint2 hi = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
    uint2 lo = make_uint2(0, 0);
    if (p.split3) {
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
      lo = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
    }
    if (p.layout == 0) {
      __half* o = p.out + row * p.ldo + c;
      *(uint2*)o = hi;
      if (p.split3) *(uint2*)(o + C) = lo;
    } else if (p.layout == 1) {
      const int y = pp / p.W, x = pp % p.W;
      const int W2 = 2 * p.W;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          __half* o = p.out + ((size_t)b * 4 * HW + (size_t)(2 * y + i) * W2 + (2 * x + j)) * p.ldo + c;
          *(uint2*)o = hi;
          if (p.split3) *(uint2*)(o + C) = lo;
        }
    } else {
      const int y = pp / p.W, x = pp % p.W;
      const int ph = (y & 1) * 2 + (x & 1);
      const int Ho = p.H >> 1, Wo = p.W >> 1;
      __half* o = p.out + ((((size_t)ph * p.B + b) * Ho + (y >> 1)) * Wo + (x >> 1)) * p.ldo + c;
      *(uint2*)o = hi;
      if (p.split3) *(uint2*)(o + C) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 4*32*NV), one warp per token row, exact two-pass statistics in registers.
// ------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, int rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out, int ldo, int split3) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int C4 = C >> 2;
  const float4* xr = (const float4*)(x + (size_t)warp * ldx);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + j * 32;
    if (i < C4) { v[j] = xr[i]; s += v[j].x + v[j].y + v[j].z + v[j].w; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + j * 32;
    if (i < C4) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  __half* orow = out + (size_t)warp * ldo;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + j * 32;
    if (i < C4) {
      const float4 g = ((const float4*)gamma)[i], be = ((const float4*)beta)[i];
      const float y0 = (v[j].x - mean) * rstd * g.x + be.x, y1 = (v[j].y - mean) * rstd * g.y + be.y;
      const float y2 = (v[j].z - mean) * rstd * g.z + be.z, y3 = (v[j].w - mean) * rstd * g.w + be.w;
      __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
      const uint2 hi = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
      *(uint2*)(orow + 4 * i) = hi;
      if (split3) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        __half2 l0 = __floats2half2_rn(y0 - f0.x, y1 - f0.y), l1 = __floats2half2_rn(y2 - f1.x, y3 - f1.y);
        *(uint2*)(orow + C + 4 * i) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Row softmax fp32 -> fp16, one CTA (256 threads) per row; scale applied to the logits first.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, int ldx, int n, float scale, __half* __restrict__ out, int ldo) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const float* xr = x + (size_t)blockIdx.x * ldx;
  __half* orow = out + (size_t)blockIdx.x * ldo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, xr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += expf((xr[i] - m) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < n; i += 256) orow[i] = __float2half_rn(expf((xr[i] - m) * scale) * inv);
}

static int pick_chunk(int HW, int B) {
  // aim for >= ~4 CTAs per SM without making chunks tiny
  int chunk = (int)(((long long)HW * B + 591) / 592);
  if (chunk < 4) chunk = 4;
  if (chunk > 64) chunk = 64;
  if (chunk > HW) chunk = HW;
  return chunk;
}

}  // namespace upgpt

using namespace upgpt;

static double* g_gn_partials = nullptr;   // [B][chunks][2*groups] per-chunk partial moments (graph-stable address)
static int* g_gn_counters = nullptr;
static constexpr size_t kGnPartialDoubles = (size_t)1 << 21;   // 16 MB
static constexpr int kGnMaxBatch = 4096;

static int groupnorm_stats_impl(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                                const float* gamma, const float* beta, float eps, float* scale_shift, cudaStream_t stream) {
  const int C = C1 + C2;
  UPGPT_REQUIRE(x1 && stats && C > 0 && C % groups == 0, "groupnorm_stats: bad args (C=%d groups=%d)", C, groups);
  UPGPT_REQUIRE(C <= 2048 && B <= kGnMaxBatch, "groupnorm_stats: C=%d > 2048 or B=%d too large", C, B);
  if (!g_gn_partials) {
    UPGPT_CHECK_CUDA(cudaMalloc(&g_gn_partials, kGnPartialDoubles * sizeof(double)));
    UPGPT_CHECK_CUDA(cudaMalloc(&g_gn_counters, kGnMaxBatch * sizeof(int)));
    UPGPT_CHECK_CUDA(cudaMemset(g_gn_counters, 0, kGnMaxBatch * sizeof(int)));
  }
  // ~2 CTAs per SM in total, at least 8 pixels per CTA, at most 256 chunks per image (cross-chunk reduction cost)
  int per_img = (2 * 148 + B - 1) / B;
  if (per_img > 256) per_img = 256;
  int chunk = (HW + per_img - 1) / per_img;
  if (chunk < 8) chunk = 8;
  if (chunk > HW) chunk = HW;
  while ((size_t)B * ((HW + chunk - 1) / chunk) * 2 * groups > kGnPartialDoubles) chunk *= 2;
  dim3 grid((HW + chunk - 1) / chunk, B);
  const size_t sm = sizeof(float) * 2 * C;
#define UPGPT_GN_LAUNCH(MC) UPGPT_CHECK_CUDA(launch_k(gn_stats_kernel<MC>, grid, dim3(256), sm, stream, x1, C1, x2, C2, HW, chunk, groups, stats, \
                                                    g_gn_partials, g_gn_counters, gamma, beta, eps, scale_shift))
  if (C <= 256) UPGPT_GN_LAUNCH(1);
  else if (C <= 512) UPGPT_GN_LAUNCH(2);
  else if (C <= 1024) UPGPT_GN_LAUNCH(4);
  else UPGPT_GN_LAUNCH(8);
#undef UPGPT_GN_LAUNCH
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_groupnorm_stats(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups,
                                     double* stats, void* stream_) {
  return groupnorm_stats_impl(x1, C1, x2, C2, B, HW, groups, stats, nullptr, nullptr, 0.f, nullptr, (cudaStream_t)stream_);
}

extern "C" int upgpt_groupnorm_affine(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                                      const float* gamma, const float* beta, float eps, float* scale_shift, void* stream_) {
  UPGPT_REQUIRE(scale_shift, "groupnorm_affine: scale_shift is null");
  return groupnorm_stats_impl(x1, C1, x2, C2, B, HW, groups, stats, gamma, beta, eps, scale_shift, (cudaStream_t)stream_);
}

extern "C" int upgpt_prep_operand(const upgpt_prep_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(a && a->x1 && a->out, "prep_operand: null");
  const int C = a->C1 + a->C2;
  UPGPT_REQUIRE(a->C1 % 4 == 0 && a->C2 % 4 == 0 && C > 0, "prep_operand: channels must be multiples of 4");
  UPGPT_REQUIRE(!a->stats || a->scale_shift || (a->groups > 0 && C % a->groups == 0), "prep_operand: bad groups");
  UPGPT_REQUIRE(a->layout != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "prep_operand: stride-2 phases need even H, W");
  UPGPT_REQUIRE(!a->raw || a->layout == 0, "prep_operand: raw copy only with layout 0");
  PrepParams p{};
  p.x1 = a->x1; p.C1 = a->C1; p.x2 = a->x2; p.C2 = a->C2; p.H = a->H; p.W = a->W; p.B = a->B;
  p.scale_shift = a->scale_shift;
  p.groups = a->groups; p.stats = a->stats; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
  p.layout = a->layout; p.split3 = a->split3;
  p.out = (__half*)a->out; p.ldo = a->ldo > 0 ? a->ldo : (a->split3 ? 2 * C : C);
  p.raw = (__half*)a->raw; p.ldraw = a->ldraw > 0 ? a->ldraw : (a->split3 ? 2 * C : C);
  UPGPT_REQUIRE(p.ldo % 4 == 0 && p.ldraw % 4 == 0, "prep_operand: ld must be a multiple of 4");
  const int HW = a->H * a->W;
  p.chunk = pick_chunk(HW, a->B);
  dim3 grid((HW + p.chunk - 1) / p.chunk, a->B);
  UPGPT_CHECK_CUDA(launch_k(prep_kernel, grid, dim3(256), sizeof(float) * 2 * C, stream, p));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int layernorm_impl(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps, void* out16,
                          int ldo, int split3, cudaStream_t stream) {
  UPGPT_REQUIRE(x && out16 && gamma && beta && C % 4 == 0 && C <= 2048, "layernorm: bad args (C=%d)", C);
  if (ldx <= 0) ldx = C;
  if (ldo <= 0) ldo = split3 ? 2 * C : C;
  UPGPT_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "layernorm: ld must be multiple of 4");
  const int warps_per_block = 8;
  dim3 grid((rows + warps_per_block - 1) / warps_per_block);
  __half* o = (__half*)out16;
  if (C <= 256) UPGPT_CHECK_CUDA(launch_k(layernorm_kernel<2>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, o, ldo, split3));
  else if (C <= 512) UPGPT_CHECK_CUDA(launch_k(layernorm_kernel<4>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, o, ldo, split3));
  else if (C <= 1024) UPGPT_CHECK_CUDA(launch_k(layernorm_kernel<8>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, o, ldo, split3));
  else UPGPT_CHECK_CUDA(launch_k(layernorm_kernel<16>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, o, ldo, split3));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_layernorm(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                               void* out16, int ldo, void* stream_) {
  return layernorm_impl(x, ldx, rows, C, gamma, beta, eps, out16, ldo, 0, (cudaStream_t)stream_);
}

extern "C" int upgpt_layernorm_split3(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                                      void* out16, int ldo, void* stream_) {
  return layernorm_impl(x, ldx, rows, C, gamma, beta, eps, out16, ldo, 1, (cudaStream_t)stream_);
}

extern "C" int upgpt_softmax_rows(const float* x, int ldx, long long rows, int n, float scale, void* out16, int ldo,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && out16 && n > 0 && rows > 0, "softmax_rows: bad args");
  UPGPT_CHECK_CUDA(launch_k(softmax_rows_kernel, dim3((unsigned)rows), dim3(256), 0, stream, x, ldx > 0 ? ldx : n, n, scale, (__half*)out16, ldo > 0 ? ldo : n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}
