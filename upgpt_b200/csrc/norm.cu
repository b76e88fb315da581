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
#include "tc_gemm.cuh"   // FastDiv
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "../../include/upgpt_b200.h"

namespace upgpt {

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm statistics. stats[b][g] = {sum, sumsq} in double (zeroed by the launcher).
// grid = (pixel chunks, B); thread = (channel quad, row lane): float4 loads, 4 rows in flight, fixed-order reductions.
// ------------------------------------------------------------------------------------------------------------------
template <int NQ>   // channel quads (float4) per thread: C/4 <= 256 * NQ
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int HW, int chunk,
                int groups, double* __restrict__ stats, double* __restrict__ partials, int* __restrict__ counters,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ scale_shift) {
  extern __shared__ float shc[];  // per-channel {sum, sumsq}: [2][C], then per-row-lane partials [lanes][2][C]
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_red[4][64];
  const int C = C1 + C2;
  const int C4 = C >> 2;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int nchunks = gridDim.x;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, HW);
  // Thread layout: (channel quad, row lane). With C/4 < 256 the CTA's threads are spread over `lanes` pixel rows as well, every
  // thread streams float4 loads (16 B) and keeps 4 rows x NQ quads in flight: a CTA has up to 16 KB outstanding instead of the
  // 256 x 4 x 4 B of a scalar channel-per-thread walk (which ran the VAE's 256x256 levels at 1/5 of HBM speed).
  const int lanes = NQ > 1 ? 1 : (256 / C4 < 1 ? 1 : 256 / C4);
  const int qi = NQ > 1 ? threadIdx.x : threadIdx.x % C4;
  const int lane = NQ > 1 ? 0 : threadIdx.x / C4;
  float4 s[NQ], q[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) { s[j] = make_float4(0.f, 0.f, 0.f, 0.f); q[j] = s[j]; }
  auto ld4 = [&](int row_in_img, int c4) -> float4 {
    const size_t row = (size_t)b * HW + row_in_img;
    const int c = c4 << 2;
    return c < C1 ? __ldg((const float4*)(x1 + row * C1 + c)) : __ldg((const float4*)(x2 + row * C2 + (c - C1)));
  };
  if (lane < lanes) {
    int p = p0 + lane;
    constexpr int R = NQ > 1 ? 4 : 8;   // rows in flight per thread
    for (; p + (R - 1) * lanes < p1; p += R * lanes) {
      float4 v[R][NQ];
#pragma unroll
      for (int u = 0; u < R; ++u)
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          const int c4 = qi + j * 256;
          v[u][j] = c4 < C4 ? ld4(p + u * lanes, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < R; ++u)
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          s[j].x += v[u][j].x; s[j].y += v[u][j].y; s[j].z += v[u][j].z; s[j].w += v[u][j].w;
          q[j].x = fmaf(v[u][j].x, v[u][j].x, q[j].x); q[j].y = fmaf(v[u][j].y, v[u][j].y, q[j].y);
          q[j].z = fmaf(v[u][j].z, v[u][j].z, q[j].z); q[j].w = fmaf(v[u][j].w, v[u][j].w, q[j].w);
        }
    }
    for (; p < p1; p += lanes) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const int c4 = qi + j * 256;
        if (c4 < C4) {
          const float4 v = ld4(p, c4);
          s[j].x += v.x; s[j].y += v.y; s[j].z += v.z; s[j].w += v.w;
          q[j].x = fmaf(v.x, v.x, q[j].x); q[j].y = fmaf(v.y, v.y, q[j].y); q[j].z = fmaf(v.z, v.z, q[j].z); q[j].w = fmaf(v.w, v.w, q[j].w);
        }
      }
    }
    // per-lane per-channel partials -> smem [lane][2][C]
    float* mine = shc + 2 * C + (size_t)lane * 2 * C;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const int c4 = qi + j * 256;
      if (c4 < C4) { *(float4*)(mine + 4 * c4) = s[j]; *(float4*)(mine + C + 4 * c4) = q[j]; }
    }
  }
  __syncthreads();
  // fixed-order sum over the row lanes -> per-channel {sum, sumsq} of this chunk
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += shc[2 * C + (size_t)l * 2 * C + i];
    shc[i] = acc;
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
  const long long* gn_acc;    // optional [B][groups][2] fixed-point moments accumulated by the producers of x1 / x2
  const float* gamma; const float* beta;
  float eps;
  int silu;
  int layout;              // 0 same, 1 nearest-up x2, 2 stride-2 phases
  int split3;              // output channels = 2C: [hi | lo]
  __half* out; int ldo;    // elements per output pixel (>= C or 3C)
  __half* raw; int ldraw;  // optional un-normalised fp16 copy (layout 0)
  int raw_split3;          // the raw copy carries [hi | lo] planes (may differ from split3: the skip GEMM runs fp16x3, conv1 need not)
  int B;
};

// One item = 4 consecutive channels of one pixel: optional raw fp16 copy, affine (GroupNorm apply), SiLU, fp16 [hi | lo] planes,
// stored in the requested layout. Shared by prep_kernel and the fused GroupNorm kernel.
__device__ __forceinline__ void prep_emit(const PrepParams& p, int b, int pp, int c, float4 v, const float* scale, const float* shift,
                                          bool affine) {
  const int C = p.C1 + p.C2;
  const int HW = p.H * p.W;
  const size_t row = (size_t)b * HW + pp;
    if (p.raw) {
      __half2 r0 = __floats2half2_rn(v.x, v.y), r1 = __floats2half2_rn(v.z, v.w);
      uint2 pk = make_uint2(*(uint32_t*)&r0, *(uint32_t*)&r1);
      *(uint2*)(p.raw + row * p.ldraw + c) = pk;
      if (p.raw_split3) {   // the un-normalised copy feeds the skip 1x1 GEMM: [hi | lo] planes when that GEMM runs fp16x3
        const float2 f0 = __half22float2(r0), f1 = __half22float2(r1);
        __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
        *(uint2*)(p.raw + row * p.ldraw + C + c) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
      }
    }
    if (affine) {
      const float4 sc = *(const float4*)(scale + c), sh = *(const float4*)(shift + c);   // 16-byte aligned: C % 4 == 0
      v.x = fmaf(v.x, sc.x, sh.x);
      v.y = fmaf(v.y, sc.y, sh.y);
      v.z = fmaf(v.z, sc.z, sh.z);
      v.w = fmaf(v.w, sc.w, sh.w);
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
  } else if (p.stats || p.gn_acc) {
    const int cpg = C / p.groups;
    const double inv_n = 1.0 / ((double)cpg * HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = c / cpg;
      double su, sq;
      if (p.gn_acc) {     // fixed-point moments from the producers' epilogues (kGnSumScale / kGnSqScale)
        su = (double)p.gn_acc[((size_t)b * p.groups + g) * 2] * (1.0 / 16777216.0);
        sq = (double)p.gn_acc[((size_t)b * p.groups + g) * 2 + 1] * (1.0 / 1048576.0);
      } else {
        su = p.stats[((size_t)b * p.groups + g) * 2];
        sq = p.stats[((size_t)b * p.groups + g) * 2 + 1];
      }
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
  // four items per thread per trip, all loads issued before the first use (the stores of one item would otherwise fence the
  // loads of the next: the compiler cannot prove that out / raw do not alias x1 / x2). (pixel, channel-quad) of an item advance
  // incrementally: an integer division per item made this kernel instruction-bound (3.5 TB/s at the VAE's 256x256 level).
  const int dpx = (int)blockDim.x / C4, dc = (int)blockDim.x % C4;
  int px = (int)threadIdx.x / C4, c4 = (int)threadIdx.x % C4;     // item idx = px * C4 + c4, idx = threadIdx.x + k * blockDim.x
  const int npx = p1 - p0;
  while (px < npx) {
    float4 vin[4];
    int ipx[4], ic[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ipx[u] = px; ic[u] = c4 << 2;
      vin[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < npx) {
        const size_t row = (size_t)b * HW + p0 + px;
        vin[u] = ic[u] < p.C1 ? __ldg((const float4*)(p.x1 + row * p.C1 + ic[u])) : __ldg((const float4*)(p.x2 + row * p.C2 + (ic[u] - p.C1)));
      }
      px += dpx; c4 += dc;
      if (c4 >= C4) { c4 -= C4; ++px; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ipx[u] < npx) prep_emit(p, b, p0 + ipx[u], ic[u], vin[u], scale, shift, p.stats || p.scale_shift || p.gn_acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused GroupNorm (+SiLU) -> fp16 operand in ONE launch: one thread-block cluster per image.
//   Every CTA of the cluster bulk-copies its pixel chunk (contiguous in NHWC) into shared memory with the TMA engine, computes
//   the chunk's per-group moments from there, the cluster exchanges the 2*groups partial moments over distributed shared memory
//   (summed in rank order: bit-reproducible), and every CTA then normalises its chunk straight out of shared memory.
//   The tensor is read from HBM/L2 exactly once and the statistics never leave the chip; the two-kernel form (gn_stats with a
//   last-CTA cross-chunk reduction, then prep) cost ~9 us + ~9 us per GroupNorm at the U-Net's sizes, almost all of it latency.
// grid = (CL, B), cluster = (CL, 1, 1), 512 threads; dynamic smem = chunk + reduction scratch.
// ------------------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU)+cast with ONE CTA PER (image, 4 groups). The groups of a GroupNorm are independent: a CTA that owns whole groups
// needs no exchange with any other CTA -- no cluster, no distributed shared memory, no second pass over memory. It loads its
// HW x 4 cpg values into registers (float4 items, <= MAXI <= 7 per thread at 1024 threads), reduces the moments inside the block in a fixed order, and
// normalises out of registers through prep_emit. grid = (groups / 4, B). A pixel contributes a run of 4 cpg channels (112 B .. 896 B):
// full sectors at the 8x8 / 4x4 levels; at 32x32 (cpg = 7: 112-byte loads, 56-byte plane stores shared between CTAs) the partial
// sectors make this form slower than the pixel-parallel cluster kernel below, which keeps those levels (measured).
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kGroupThreads = 1024, kGroupG = 4;   // 32 warps: the emit is latency-bound per thread (SiLU, stores)
template <int MAXI>
__global__ void __launch_bounds__(kGroupThreads)
gn_group_kernel(const PrepParams p, double* __restrict__ stats_out, const FastDiv fd_ip, const FastDiv fd_cpg) {
  constexpr int G = kGroupG;
  __shared__ double red[kGroupThreads / 32][2 * G];
  __shared__ __align__(16) float sc[256], sh[256];
  __shared__ float ms[G][2];
  pdl_launch_dependents();
  const int C = p.C1 + p.C2, HW = p.H * p.W;
  const int cpg = C / p.groups;
  const int gq = blockIdx.x, b = blockIdx.y;
  const int cbase = gq * G * cpg;
  const int IP = (G * cpg) >> 2;            // float4 items per pixel
  const int n_items = HW * IP;
  pdl_wait();
  float4 v[MAXI];
  float s[G], q[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { s[g] = 0.f; q[g] = 0.f; }
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = (int)threadIdx.x + kGroupThreads * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n_items) {
      const int px = fd_ip.div(j);
      const int c4 = j - px * IP;
      const int c = cbase + 4 * c4;
      const size_t row = (size_t)b * HW + px;
      v[i] = c < p.C1 ? *(const float4*)(p.x1 + row * p.C1 + c) : *(const float4*)(p.x2 + row * p.C2 + (c - p.C1));
      const int g0 = fd_cpg.div(4 * c4), g3 = fd_cpg.div(4 * c4 + 3);
      if (g0 == g3) {       // the item lies inside one group (always when cpg % 4 == 0)
        const float vs = (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float vq = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, v[i].w * v[i].w)));
#pragma unroll
        for (int g = 0; g < G; ++g) { s[g] += g == g0 ? vs : 0.f; q[g] += g == g0 ? vq : 0.f; }
      } else {
        const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gk = fd_cpg.div(4 * c4 + k);
#pragma unroll
          for (int g = 0; g < G; ++g) { s[g] += g == gk ? e[k] : 0.f; q[g] += g == gk ? e[k] * e[k] : 0.f; }
        }
      }
    }
  }
  // block moments in a fixed order: butterfly inside the warp, then the 32 warp partials in double
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s[g] += __shfl_xor_sync(0xffffffffu, s[g], o); q[g] += __shfl_xor_sync(0xffffffffu, q[g], o); }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) { red[threadIdx.x >> 5][2 * g] = (double)s[g]; red[threadIdx.x >> 5][2 * g + 1] = (double)q[g]; }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double S = 0.0, Q = 0.0;
#pragma unroll
    for (int w = 0; w < kGroupThreads / 32; ++w) { S += red[w][2 * g]; Q += red[w][2 * g + 1]; }
    const double inv_n = 1.0 / ((double)cpg * HW);
    const double mean = S * inv_n;
    double var = Q * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    ms[g][0] = (float)mean;
    ms[g][1] = (float)(1.0 / sqrt(var + (double)p.eps));
    if (stats_out) {
      stats_out[((size_t)b * p.groups + gq * G + g) * 2] = S;
      stats_out[((size_t)b * p.groups + gq * G + g) * 2 + 1] = Q;
    }
  }
  __syncthreads();
  for (int cl = threadIdx.x; cl < G * cpg; cl += kGroupThreads) {
    const int g = fd_cpg.div(cl);
    const float ga = p.gamma ? p.gamma[cbase + cl] : 1.f, be = p.beta ? p.beta[cbase + cl] : 0.f;
    const float scl = ms[g][1] * ga;
    sc[cl] = scl;
    sh[cl] = be - ms[g][0] * scl;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = (int)threadIdx.x + kGroupThreads * i;
    if (j >= n_items) continue;
    const int px = fd_ip.div(j);
    const int c = cbase + 4 * (j - px * IP);
    prep_emit(p, b, px, c, v[i], sc - cbase, sh - cbase, true);      // (scale / shift tables are indexed by the global channel)
  }
}

static constexpr int kFusedThreads = 1024;   // 32 warps per SM: the emit phase is instruction-latency bound

__global__ void __launch_bounds__(kFusedThreads, 1)
gn_prep_fused_kernel(const PrepParams p, int cl, int px_per_cta, double* __restrict__ stats_out) {
  extern __shared__ __align__(128) uint8_t fsm[];
  const int C = p.C1 + p.C2;
  const int C4 = C >> 2;
  const int HW = p.H * p.W;
  const int b = blockIdx.y;
  const int rank = blockIdx.x;                   // == %cluster_ctarank for cluster dims (cl, 1, 1)
  const int p0 = rank * px_per_cta;
  const int npx = max(0, min(px_per_cta, HW - p0));
  const int groups = p.groups;
  const int cpg = C / groups;
  // smem carve-up
  const size_t bytes1 = (size_t)px_per_cta * p.C1 * 4, bytes2 = (size_t)px_per_cta * p.C2 * 4;
  float* sx1 = (float*)fsm;                       // [px][C1]
  float* sx2 = (float*)(fsm + bytes1);            // [px][C2]
  const int lanes = kFusedThreads / C4 < 1 ? 1 : kFusedThreads / C4;
  float* lane_part = (float*)(fsm + bytes1 + bytes2);          // [lanes][2][C]
  float* chan = lane_part + (size_t)lanes * 2 * C;              // [2][C]
  float* scale = chan + 2 * C;                                  // [C]
  float* shift = scale + C;                                     // [C]
  double* gs = (double*)(((uintptr_t)(shift + C) + 15) & ~(uintptr_t)15);   // [2*groups] this CTA's partial moments
  double* gtot = gs + 2 * groups;                               // [2*groups] cluster totals
  uint64_t* bar = (uint64_t*)(gtot + 2 * groups);

  pdl_launch_dependents();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  pdl_wait();
  // ---- phase 0: bulk copy of the chunk (TMA, no registers involved) ----
  if (threadIdx.x == 0) {
    const size_t n1 = (size_t)npx * p.C1 * 4, n2 = (size_t)npx * p.C2 * 4;
    mbar_arrive_expect_tx(bar, (uint32_t)(n1 + n2));
    const uint8_t* g1 = (const uint8_t*)(p.x1 + ((size_t)b * HW + p0) * p.C1);
    for (size_t o = 0; o < n1; o += 65536) bulk_g2s((uint8_t*)sx1 + o, g1 + o, (uint32_t)min((size_t)65536, n1 - o), bar);
    if (p.C2 > 0) {
      const uint8_t* g2 = (const uint8_t*)(p.x2 + ((size_t)b * HW + p0) * p.C2);
      for (size_t o = 0; o < n2; o += 65536) bulk_g2s((uint8_t*)sx2 + o, g2 + o, (uint32_t)min((size_t)65536, n2 - o), bar);
    }
  }
  mbar_wait(bar, 0);
  auto lds4 = [&](int px, int c) -> float4 {
    return c < p.C1 ? *(const float4*)(sx1 + (size_t)px * p.C1 + c) : *(const float4*)(sx2 + (size_t)px * p.C2 + (c - p.C1));
  };
  // ---- phase 1: per-channel moments of the chunk; thread = (channel quad, row lane) ----
  {
    const int qi = threadIdx.x % C4, lane = threadIdx.x / C4;
    if (lane < lanes && C4 <= kFusedThreads) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
      for (int px = lane; px < npx; px += lanes) {
        const float4 v = lds4(px, qi << 2);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
      float* mine = lane_part + (size_t)lane * 2 * C;
      *(float4*)(mine + 4 * qi) = s;
      *(float4*)(mine + C + 4 * qi) = q;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kFusedThreads) {   // fixed-order sum over the row lanes
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += lane_part[(size_t)l * 2 * C + i];
    chan[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += kFusedThreads) {   // (group, moment): double sum over the group's channels
    const int g = i >> 1, which = i & 1;
    const float* src = chan + which * C + g * cpg;
    double acc = 0.0;
    for (int k = 0; k < cpg; ++k) acc += (double)src[k];
    gs[i] = acc;
  }
  // ---- phase 2: cluster-wide moments over DSMEM, then the per-channel affine ----
  cluster_sync_all();
  for (int i = threadIdx.x; i < 2 * groups; i += kFusedThreads) {
    // all (<= 16) sibling loads are issued before the first add: one DSMEM round trip instead of `cl` serialised ones
    const uint32_t my = smem_u32(gs + i);
    double acc = 0.0;
#pragma unroll
    for (int r0 = 0; r0 < 16; r0 += 8) {     // 8 sibling loads in flight at a time; summed in rank order: bit-reproducible
      if (r0 >= cl) break;
      double t[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) t[r] = ld_cluster_f64(mapa_shared(my, (uint32_t)min(r0 + r, cl - 1)));
#pragma unroll
      for (int r = 0; r < 8; ++r) acc += (r0 + r) < cl ? t[r] : 0.0;
    }
    gtot[i] = acc;
    if (stats_out && rank == 0) stats_out[(size_t)b * 2 * groups + i] = acc;
  }
  cluster_arrive();   // my DSMEM reads are done (the matching wait is at the end: nobody exits while a sibling may still read)
  __syncthreads();
  {
    const double inv_n = 1.0 / ((double)cpg * HW);
    for (int c = threadIdx.x; c < C; c += kFusedThreads) {
      const int g = c / cpg;
      const double mean = gtot[2 * g] * inv_n;
      double var = gtot[2 * g + 1] * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      const float ga = p.gamma ? p.gamma[c] : 1.f;
      const float be = p.beta ? p.beta[c] : 0.f;
      scale[c] = rstd * ga;
      shift[c] = be - (float)mean * rstd * ga;
    }
  }
  __syncthreads();
  // ---- phase 3: normalise (+SiLU) out of shared memory, emit the fp16 operand ----
  {
    const int dpx = kFusedThreads / C4, dc = kFusedThreads % C4;
    int px = (int)threadIdx.x / C4, c4 = (int)threadIdx.x % C4;
    while (px < npx) {
      prep_emit(p, b, p0 + px, c4 << 2, lds4(px, c4 << 2), scale, shift, true);
      px += dpx; c4 += dc;
      if (c4 >= C4) { c4 -= C4; ++px; }
    }
  }
  cluster_wait();
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

// ------------------------------------------------------------------------------------------------------------------
// Group moments of a tensor as 64-bit fixed-point accumulators (see include/upgpt_b200.h: upgpt_gn_accumulate).
// grid = (pixel chunks, B), 256 threads: thread = (channel quad, row lane) like gn_stats; per-channel chunk sums -> integer atomics.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_accumulate_kernel(const float* __restrict__ x, int C, int HW, int chunk, int groups, int cpg, int choff, long long* __restrict__ acc) {
  extern __shared__ float shc[];   // [lanes][2][C]
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C >> 2;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, HW);
  const int lanes = max(1, 256 / C4);
  const int qi = threadIdx.x % C4, lane = threadIdx.x / C4;
  if (lane < lanes && C4 <= 256) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int px = p0 + lane; px < p1; px += lanes) {
      const float4 v = __ldg((const float4*)(x + ((size_t)b * HW + px) * C + 4 * qi));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    float* mine = shc + (size_t)lane * 2 * C;
    *(float4*)(mine + 4 * qi) = s;
    *(float4*)(mine + C + 4 * qi) = q;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < lanes; ++l) { s += shc[(size_t)l * 2 * C + c]; q += shc[(size_t)l * 2 * C + C + c]; }
    const int g = (choff + c) / cpg;
    unsigned long long* a = (unsigned long long*)(acc + ((size_t)b * groups + g) * 2);
    atomicAdd(a, (unsigned long long)__float2ll_rn(s * 16777216.f));
    atomicAdd(a + 1, (unsigned long long)__float2ll_rn(q * 1048576.f));
  }
}

static int pick_chunk(int HW, int B) {
  // aim for >= ~4 CTAs per SM without making chunks tiny
  int chunk = (int)(((long long)HW * B + 591) / 592);
  static int cap = -1;
  if (cap < 0) { const char* e = getenv("UPGPT_PREP_CHUNK_MAX"); cap = e ? atoi(e) : 64; }
  if (chunk < 4) chunk = 4;
  if (chunk > cap) chunk = cap;
  if (chunk > HW) chunk = HW;
  return chunk;
}

}  // namespace upgpt

using namespace upgpt;

// per-device state (workspaces live on the device that runs the kernels; function attributes are per device)
static constexpr size_t kGnPartialDoubles = (size_t)1 << 21;   // 16 MB
static constexpr int kGnMaxBatch = 4096;
static constexpr size_t kFusedSsFloats = (size_t)1 << 20;
static constexpr int kNormMaxDevices = 64;
struct NormDev {
  // per stream slot (common.cuh: stream_slot), allocated on the slot's first use, then fixed (graph-stable addresses):
  double* gn_partials[kStreamSlots] = {};   // [B][chunks][2*groups] per-chunk partial moments
  int* gn_counters[kStreamSlots] = {};
  float* fused_ss[kStreamSlots] = {};       // scale/shift scratch of the two-launch fallback
  int fused_max_cl = -1;
  int fused_n16 = 0;               // co-resident 16-CTA clusters (occupancy API, full shared-memory carve-out)
  int fused_smem_optin = 0;
};
static NormDev g_ndev[kNormMaxDevices];
static std::mutex g_norm_mu;
static NormDev* norm_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kNormMaxDevices) return nullptr;
  return &g_ndev[dev];
}

static int groupnorm_stats_impl(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, double* stats,
                                const float* gamma, const float* beta, float eps, float* scale_shift, cudaStream_t stream) {
  const int C = C1 + C2;
  UPGPT_REQUIRE(x1 && stats && C > 0 && C % groups == 0, "groupnorm_stats: bad args (C=%d groups=%d)", C, groups);
  UPGPT_REQUIRE(C <= 2048 && B <= kGnMaxBatch, "groupnorm_stats: C=%d > 2048 or B=%d too large", C, B);
  NormDev* nd = norm_dev();
  UPGPT_REQUIRE(nd, "groupnorm_stats: no current CUDA device");
  const int slot = stream_slot(stream);
  {
    std::lock_guard<std::mutex> lk(g_norm_mu);
    if (!nd->gn_partials[slot]) {
      UPGPT_CHECK_CUDA(cudaMalloc(&nd->gn_partials[slot], kGnPartialDoubles * sizeof(double)));
      UPGPT_CHECK_CUDA(cudaMalloc(&nd->gn_counters[slot], kGnMaxBatch * sizeof(int)));
      UPGPT_CHECK_CUDA(cudaMemset(nd->gn_counters[slot], 0, kGnMaxBatch * sizeof(int)));
    }
  }
  double* g_gn_partials = nd->gn_partials[slot];
  int* g_gn_counters = nd->gn_counters[slot];
  // ~2 CTAs per SM in total, at least 8 pixels per CTA, at most 256 chunks per image (cross-chunk reduction cost)
  int per_img = (4 * 148 + B - 1) / B;
  if (per_img > 256) per_img = 256;
  int chunk = (HW + per_img - 1) / per_img;
  if (chunk < 8) chunk = 8;
  if (chunk > HW) chunk = HW;
  while ((size_t)B * ((HW + chunk - 1) / chunk) * 2 * groups > kGnPartialDoubles) chunk *= 2;
  dim3 grid((HW + chunk - 1) / chunk, B);
  const int gn_lanes = C / 4 > 256 ? 1 : (256 / (C / 4) < 1 ? 1 : 256 / (C / 4));
  const size_t sm = sizeof(float) * 2 * C * (1 + gn_lanes);
#define UPGPT_GN_LAUNCH(MC) UPGPT_CHECK_CUDA(launch_k(gn_stats_kernel<MC>, grid, dim3(256), sm, stream, x1, C1, x2, C2, HW, chunk, groups, stats, \
                                                    g_gn_partials, g_gn_counters, gamma, beta, eps, scale_shift))
  UPGPT_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0, "groupnorm_stats: channel counts must be multiples of 4 (C1=%d C2=%d)", C1, C2);
  if (C <= 1024) UPGPT_GN_LAUNCH(1);
  else UPGPT_GN_LAUNCH(2);
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
  UPGPT_REQUIRE(!(a->stats || a->gn_acc) || a->scale_shift || (a->groups > 0 && C % a->groups == 0), "prep_operand: bad groups");
  UPGPT_REQUIRE(a->layout != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "prep_operand: stride-2 phases need even H, W");
  UPGPT_REQUIRE(!a->raw || a->layout == 0, "prep_operand: raw copy only with layout 0");
  PrepParams p{};
  p.x1 = a->x1; p.C1 = a->C1; p.x2 = a->x2; p.C2 = a->C2; p.H = a->H; p.W = a->W; p.B = a->B;
  p.scale_shift = a->scale_shift;
  p.gn_acc = a->gn_acc;
  p.groups = a->groups; p.stats = a->stats; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
  p.layout = a->layout; p.split3 = a->split3;
  p.out = (__half*)a->out; p.ldo = a->ldo > 0 ? a->ldo : (a->split3 ? 2 * C : C);
  p.raw_split3 = a->raw_planes == 0 ? a->split3 : (a->raw_planes == 2);
  p.raw = (__half*)a->raw; p.ldraw = a->ldraw > 0 ? a->ldraw : (p.raw_split3 ? 2 * C : C);
  UPGPT_REQUIRE(p.ldo % 4 == 0 && p.ldraw % 4 == 0, "prep_operand: ld must be a multiple of 4");
  const int HW = a->H * a->W;
  p.chunk = pick_chunk(HW, a->B);
  dim3 grid((HW + p.chunk - 1) / p.chunk, a->B);
  UPGPT_CHECK_CUDA(launch_k(prep_kernel, grid, dim3(256), sizeof(float) * 2 * C, stream, p));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// GroupNorm(+SiLU) + cast in one call. Fused single-launch cluster kernel when one image's [HW][C] fp32 tile fits the shared
// memory of a cluster (<= 16 CTAs); otherwise statistics (gn_stats) and apply (prep) as two launches.

extern "C" int upgpt_groupnorm_prep(const upgpt_prep_args* a, double* stats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(a && a->x1 && a->out && stats, "groupnorm_prep: null");
  const int C = a->C1 + a->C2;
  const int HW = a->H * a->W;
  UPGPT_REQUIRE(a->groups > 0 && C > 0 && C % a->groups == 0 && a->C1 % 4 == 0 && a->C2 % 4 == 0, "groupnorm_prep: bad channels (C1=%d C2=%d groups=%d)", a->C1, a->C2, a->groups);
  {
    // one CTA per (image, 4 groups) when that fits the registers of a block and the level is small enough for the short channel runs
    // to be whole sectors (UPGPT_GN_GROUP_HW: largest H*W that takes this form, 0 = never; default from the measurements in DESIGN.md)
    static const int group_hw = getenv("UPGPT_GN_GROUP_HW") ? atoi(getenv("UPGPT_GN_GROUP_HW")) : 64;
    const int cpg = C / a->groups;
    const long long n_items = (long long)HW * cpg;      // = HW * (4 cpg) / 4
    if (HW <= group_hw && a->groups % kGroupG == 0 && kGroupG * cpg <= 256 && n_items <= (long long)kGroupThreads * 7 &&
        ((uintptr_t)a->x1 & 15) == 0 && (!a->x2 || ((uintptr_t)a->x2 & 15) == 0) && a->B <= 65535) {
      UPGPT_REQUIRE(a->layout != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "groupnorm_prep: stride-2 phases need even H, W");
      UPGPT_REQUIRE(!a->raw || a->layout == 0, "groupnorm_prep: raw copy only with layout 0");
      PrepParams p{};
      p.x1 = a->x1; p.C1 = a->C1; p.x2 = a->x2; p.C2 = a->C2; p.H = a->H; p.W = a->W; p.B = a->B;
      p.groups = a->groups; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
      p.layout = a->layout; p.split3 = a->split3;
      p.out = (__half*)a->out; p.ldo = a->ldo > 0 ? a->ldo : (a->split3 ? 2 * C : C);
      p.raw_split3 = a->raw_planes == 0 ? a->split3 : (a->raw_planes == 2);
  p.raw = (__half*)a->raw; p.ldraw = a->ldraw > 0 ? a->ldraw : (p.raw_split3 ? 2 * C : C);
      UPGPT_REQUIRE(p.ldo % 4 == 0 && p.ldraw % 4 == 0, "groupnorm_prep: ld must be a multiple of 4");
      const FastDiv fd_ip = make_fastdiv(cpg), fd_cpg = make_fastdiv(cpg);      // items per pixel = 4 cpg / 4 = cpg
      const dim3 grid(a->groups / kGroupG, a->B), block(kGroupThreads);
      const int per = (int)((n_items + kGroupThreads - 1) / kGroupThreads);
      cudaError_t e;
      if (per <= 1) e = launch_k(gn_group_kernel<1>, grid, block, 0, stream, p, stats, fd_ip, fd_cpg);
      else if (per <= 4) e = launch_k(gn_group_kernel<4>, grid, block, 0, stream, p, stats, fd_ip, fd_cpg);
      else e = launch_k(gn_group_kernel<7>, grid, block, 0, stream, p, stats, fd_ip, fd_cpg);
      UPGPT_CHECK_CUDA(e);
      count_launch();
      UPGPT_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  NormDev* nd = norm_dev();
  UPGPT_REQUIRE(nd, "groupnorm_prep: no current CUDA device");
  std::unique_lock<std::mutex> lk(g_norm_mu);
  int& g_fused_max_cl = nd->fused_max_cl;
  int& g_fused_n16 = nd->fused_n16;
  int& g_fused_smem_optin = nd->fused_smem_optin;
  float*& g_fused_ss = nd->fused_ss[stream_slot(stream)];
  if (!g_fused_ss) UPGPT_CHECK_CUDA(cudaMalloc(&g_fused_ss, kFusedSsFloats * sizeof(float)));
  if (g_fused_max_cl < 0) {
    int dev = 0;
    UPGPT_CHECK_CUDA(cudaGetDevice(&dev));
    UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_fused_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    UPGPT_CHECK_CUDA(cudaFuncSetAttribute(gn_prep_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_fused_smem_optin));
    UPGPT_CHECK_CUDA(cudaFuncSetAttribute(gn_prep_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // can a 16-CTA cluster with the full shared-memory carve-out be scheduled on this part at all?
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(16, 8); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = g_fused_smem_optin;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    const bool ok16 = cudaOccupancyMaxActiveClusters(&n, gn_prep_fused_kernel, &cfg) == cudaSuccess && n >= 1;
    (void)cudaGetLastError();
    g_fused_n16 = ok16 ? n : 0;
    g_fused_max_cl = (ok16 && getenv("UPGPT_NO_CLUSTER16") == nullptr) ? 16 : 8;
    if (getenv("UPGPT_NO_FUSED_GN")) g_fused_max_cl = 0;
    if (getenv("UPGPT_GN_VERBOSE")) fprintf(stderr, "[upgpt] fused GroupNorm: max cluster %d, co-resident 16-CTA clusters %d\n", g_fused_max_cl, g_fused_n16);
  }
  lk.unlock();
  // cluster size: the smallest power of two whose chunk fits, but at least 8 px per CTA and preferably >= 8 CTAs per image
  int cl = 0, px_per_cta = 0;
  size_t smem = 0;
  const int lanes = kFusedThreads / (C / 4) < 1 ? 1 : kFusedThreads / (C / 4);
  const size_t scratch = (size_t)lanes * 2 * C * 4 + (size_t)4 * C * 4 + 16 + (size_t)4 * a->groups * 8 + 64;
  if (C / 4 <= kFusedThreads) {
    // the largest cluster that keeps >= 8 pixels per CTA (more SMs pulling and normalising); 16-CTA clusters (non-portable size)
    // only while all images' clusters are co-resident
    for (int c = 1; c <= g_fused_max_cl; c *= 2) {
      const int px = (HW + c - 1) / c;
      const size_t need = (size_t)px * C * 4 + scratch;
      if (c > 1 && px < 8) break;
      if (c == 16 && a->B > g_fused_n16 && cl != 0 && getenv("UPGPT_GN_FORCE16") == nullptr) break;
      if (need + 256 > (size_t)g_fused_smem_optin) continue;
      cl = c; px_per_cta = px; smem = need + 128;
    }
  }
  if (cl == 0) {
    // two-launch fallback (VAE-sized images)
    UPGPT_REQUIRE((size_t)a->B * 2 * C <= kFusedSsFloats, "groupnorm_prep: scale/shift scratch too small");
    int rc = groupnorm_stats_impl(a->x1, a->C1, a->x2, a->C2, a->B, HW, a->groups, stats, a->gamma, a->beta, a->eps, g_fused_ss, stream);
    if (rc) return rc;
    upgpt_prep_args b = *a;
    b.stats = nullptr; b.scale_shift = g_fused_ss;
    return upgpt_prep_operand(&b, stream_);
  }
  UPGPT_REQUIRE(a->layout != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "groupnorm_prep: stride-2 phases need even H, W");
  UPGPT_REQUIRE(!a->raw || a->layout == 0, "groupnorm_prep: raw copy only with layout 0");
  PrepParams p{};
  p.x1 = a->x1; p.C1 = a->C1; p.x2 = a->x2; p.C2 = a->C2; p.H = a->H; p.W = a->W; p.B = a->B;
  p.groups = a->groups; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps; p.silu = a->silu;
  p.layout = a->layout; p.split3 = a->split3;
  p.out = (__half*)a->out; p.ldo = a->ldo > 0 ? a->ldo : (a->split3 ? 2 * C : C);
  p.raw_split3 = a->raw_planes == 0 ? a->split3 : (a->raw_planes == 2);
  p.raw = (__half*)a->raw; p.ldraw = a->ldraw > 0 ? a->ldraw : (p.raw_split3 ? 2 * C : C);
  UPGPT_REQUIRE(p.ldo % 4 == 0 && p.ldraw % 4 == 0, "groupnorm_prep: ld must be a multiple of 4");
  UPGPT_REQUIRE((((uintptr_t)a->x1) & 15) == 0 && (!a->x2 || (((uintptr_t)a->x2) & 15) == 0), "groupnorm_prep: inputs must be 16-byte aligned");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cl, a->B); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  UPGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gn_prep_fused_kernel, p, cl, px_per_cta, stats));
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

extern "C" int upgpt_gn_accumulate(const float* x, int C, int B, int HW, int groups, int cpg, int choff, long long* acc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && acc && C > 0 && C % 4 == 0 && C <= 1024 && B > 0 && HW > 0 && groups > 0 && cpg > 0, "gn_accumulate: bad args (C=%d)", C);
  UPGPT_REQUIRE((choff + C + cpg - 1) / cpg <= groups, "gn_accumulate: channels %d..%d exceed %d groups of %d", choff, choff + C, groups, cpg);
  int per_img = (2 * 148 + B - 1) / B;
  int chunk = (HW + per_img - 1) / per_img;
  if (chunk < 16) chunk = 16;
  if (chunk > HW) chunk = HW;
  dim3 grid((HW + chunk - 1) / chunk, B);
  const int lanes = 256 / (C / 4) < 1 ? 1 : 256 / (C / 4);
  UPGPT_CHECK_CUDA(launch_k(gn_accumulate_kernel, grid, dim3(256), sizeof(float) * 2 * C * lanes, stream, x, C, HW, chunk, groups, cpg, choff, acc));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_zero(void* ptr, long long bytes, void* stream_) {
  UPGPT_REQUIRE(ptr && bytes > 0, "upgpt_zero: bad args");
  UPGPT_CHECK_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream_));
  return 0;
}

UPGPT_TRACE_TU(norm)
