// Flash-style multi-head attention on tcgen05 / TMEM for sm_100a.
//
//   O = softmax(scale * Q K^T) V        per (batch, head), Q:[Nq, d]  K:[Nk, d]  V:[Nk, d]
//
// One CTA owns a 128-query tile of one (batch, head) and streams 128-key tiles:
//   warp0  : TMA producer (Q once; K tile + V^T tile per step, 2-stage ring)
//   warp1  : TMEM allocator + tcgen05.mma issuer:  S = Q K^T  (TMEM cols [0,128)),  O += P V (TMEM cols [128,128+d))
//   warps2-5 : online softmax, one query row per thread (TMEM lane == row, so row max / row sum need no shuffles):
//              tcgen05.ld S -> running max -> exp2 -> P (fp16) written to smem in the 128B-swizzled K-major layout the
//              PV MMA consumes; rescales the O accumulator in TMEM when the running max moves; final 1/l and fp16 store.
// Operands are fp16, accumulation / softmax statistics fp32. Head dim is padded to 64 or 128 (zero columns).
// V is consumed transposed (V^T [d][Nk], written that way by the projection GEMM's channel-major epilogue) so that every
// tensor-core operand in the library uses the one K-major SWIZZLE_128B layout.
//
// Replaces CrossAttention.forward's einsum / softmax / einsum (attention.py:178-192) for both attn1 (self) and attn2
// (keys = [text | style | SMPL] context, attention.py:213).
#include "common.cuh"
#include <cstdlib>
#include "../../include/upgpt_b200.h"

namespace upgpt {

struct AttnParams {
  int Nq, Nk, H, B;
  int dpad;            // 64 or 128
  int num_q_tiles;
  float scale_log2e;   // softmax scale * log2(e)
  __half* out;         // [B][Nq][ldo], head h at columns h*dpad
  int ldo;
  int plane;           // > 0: also write the lo plane at column offset plane (fp16x3 operand layout [hi | lo])
  int v_mn;            // V is row-major [Nk][d] (same layout as K): the P V MMA reads it as an MN-major B operand
  int k_stages;        // K ring depth: 2, or 1 when all keys fit one tile (cross-attention over the 87-token context, 8x8 / 4x4 self)
  int v_stages;        // V ring depth (1 in the compact two-CTAs-per-SM forms)
  int s_bufs;          // S accumulators in TMEM: 2 (ping-pong) or 1 (compact forms: the co-resident CTA fills the gap instead)
  int d32;             // head dim padded to 32: q / k / v hold head PAIRS in 64-wide rows (head h at columns 32 h); the S MMA of head h
                       // runs over the two 16-wide k-steps of its half of the row, P V over the pair's 64 V columns (its own 32 are kept)
  uint32_t tmem_cols;  // TMEM columns to allocate: 512 (two S buffers + O), or 256 for the compact forms (S + O) so that two CTAs
  uint32_t o_col;      // share an SM; first column of the O accumulator
};

static constexpr int kAttnThreads = 320;     // warp0 TMA, warp1 MMA, warps2-9 softmax (two threads per query row)
static constexpr int kTileBytes = 128 * 64 * 2;  // one [128][64] fp16 chunk

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MINB = 2: the single-key-tile form (80 KB of shared memory, 256 TMEM columns, <= 96 registers): two CTAs per SM overlap each other's
// load -> S -> softmax -> PV -> store latency chain, which is all such a CTA consists of.
template <int MINB>
__global__ void __launch_bounds__(kAttnThreads, MINB)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const int dch = p.dpad >> 6;                       // 64-wide d chunks (1 or 2)
  const uint32_t qk_bytes = (uint32_t)dch * kTileBytes;      // Q tile or K tile
  const uint32_t vt_chunk = (uint32_t)p.dpad * 128u;          // V^T chunk: [dpad rows][64 keys]
  const uint32_t vt_bytes = 2u * vt_chunk;
  const uint32_t nks = (uint32_t)p.k_stages, nvs = (uint32_t)p.v_stages;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + qk_bytes;                        // nks stages
  uint8_t* sV = sK + nks * qk_bytes;                  // nvs stages
  uint8_t* sP = sV + nvs * vt_bytes;                  // [2 key chunks][128][64] fp16
  const uint32_t off_bar = (1 + nks) * qk_bytes + nvs * vt_bytes + 2 * kTileBytes;
  uint64_t* bars = (uint64_t*)(smem + off_bar);
  uint64_t* bar_q = bars;
  uint64_t* bar_k_full = bars + 1;     // [2]
  uint64_t* bar_k_empty = bars + 3;    // [2] K slot read by the S MMA
  uint64_t* bar_v_full = bars + 5;     // [2]
  uint64_t* bar_v_empty = bars + 7;    // [2] V slot read by the P V MMA
  uint64_t* bar_s = bars + 9;          // [2] S buffer b written by the tensor core
  uint64_t* bar_sfree = bars + 11;     // [2] S buffer b consumed by all softmax warps
  uint64_t* bar_p = bars + 13;         // P tile staged (and O rescaled)
  uint64_t* bar_o = bars + 14;         // O += P V finished
  uint32_t* tmem_base_smem = (uint32_t*)(bars + 15);
  float* xch = (float*)(smem + off_bar + 128);   // [2 tiles parity][2 halves][128 rows] row-max exchange, then row-sum exchange
  // ring slot / use count of tile j in a ring of `n` (1 or 2) slots
  auto slot = [](int j, int n) { return n == 2 ? (j & 1) : 0; };
  auto use = [](int j, int n) { return n == 2 ? (j >> 1) : j; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int qt = blockIdx.x % p.num_q_tiles;
  const int bh = blockIdx.x / p.num_q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int hc = p.d32 ? (h >> 1) : h;      // head coordinate of the TMA boxes (a 64-wide box holds a head pair when d32)
  const int q0 = qt * 128;
  const int n_kv = (p.Nk + 127) >> 7;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_k_full[s], 1); mbar_init(&bar_k_empty[s], 1);
      mbar_init(&bar_v_full[s], 1); mbar_init(&bar_v_empty[s], 1);
      mbar_init(&bar_s[s], 1); mbar_init(&bar_sfree[s], 8);
    }
    mbar_init(bar_p, 8);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_wait();
  const uint32_t tO = tmem_base + p.o_col;    // dpad columns; S buffers at columns [0,128) and (two-stage form) [128,256)

  if (warp == 0) {
    // two independent producer lanes: lane 0 feeds Q and the K ring, lane 1 the V ring (a V slot only frees when its P V MMA has
    // completed, long after the next K tile is wanted)
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, qk_bytes);
      for (int c = 0; c < dch; ++c) tma_load_4d(sQ + c * kTileBytes, &tmQ, bar_q, c * 64, hc, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = slot(j, p.k_stages);
        mbar_wait(&bar_k_empty[s], (use(j, p.k_stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_k_full[s], qk_bytes);
        for (int c = 0; c < dch; ++c)
          tma_load_4d(sK + s * qk_bytes + c * kTileBytes, &tmK, &bar_k_full[s], c * 64, hc, j * 128, b);
      }
    } else if (lane == 1) {
      for (int j = 0; j < n_kv; ++j) {
        const int s = slot(j, p.v_stages);
        mbar_wait(&bar_v_empty[s], (use(j, p.v_stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_v_full[s], vt_bytes);
        if (p.v_mn) {
          // row-major V: the same [128 keys][64 d] swizzled boxes as K
          for (int c = 0; c < dch; ++c)
            tma_load_4d(sV + s * vt_bytes + c * kTileBytes, &tmVt, &bar_v_full[s], c * 64, hc, j * 128, b);
        } else {
          for (int c = 0; c < 2; ++c)
            tma_load_3d(sV + s * vt_bytes + c * vt_chunk, &tmVt, &bar_v_full[s], j * 128 + c * 64, h * p.dpad, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128);
      const uint32_t idesc_o = make_idesc_f16(128, (uint32_t)p.dpad, false, p.v_mn != 0);
      auto issue_pv = [&](int jj) {   // O (+)= P(jj) V(jj)
        const int sp = slot(jj, p.v_stages);
        mbar_wait(&bar_v_full[sp], use(jj, p.v_stages) & 1);
        tc_fence_after();
        for (int c = 0; c < 2; ++c) {
          const uint64_t ad = make_desc_kmajor_sw128(smem_u32(sP + c * kTileBytes));
          if (p.v_mn) {
            // B = V tile [128 keys][dpad] as stored by TMA: MN-major (d contiguous in 128-byte swizzled rows, one row per key).
            // A 16-key k-step is two 8-row atoms (SBO = 1024 B); dpad = 128 adds a second 64-wide MN chunk kTileBytes further (LBO).
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t row_off = (uint32_t)(c * 64 + k * 16) * 128u;
              const uint64_t bd = make_desc_mnmajor_sw128(smem_u32(sV + sp * vt_bytes) + row_off, (uint32_t)kTileBytes, 1024u);
              tc_mma_f16_ss(tO, ad + 2 * k, bd, idesc_o, (jj > 0 || c > 0 || k > 0) ? 1u : 0u);
            }
          } else {
            const uint64_t bd = make_desc_kmajor_sw128(smem_u32(sV + sp * vt_bytes + c * vt_chunk));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_f16_ss(tO, ad + 2 * k, bd + 2 * k, idesc_o, (jj > 0 || c > 0 || k > 0) ? 1u : 0u);
          }
        }
        tc_commit(&bar_v_empty[sp]);
        tc_commit(bar_o);
      };
      mbar_wait(bar_q, 0);
      // d32: head h owns the 16-wide k-steps {2 (h & 1), 2 (h & 1) + 1} of the pair's 64-wide rows
      const int k_lo = p.d32 ? 2 * (h & 1) : 0, k_hi = p.d32 ? k_lo + 2 : 4;
      for (int j = 0; j < n_kv; ++j) {
        const int s = slot(j, p.k_stages), sb = slot(j, p.s_bufs);
        mbar_wait(&bar_k_full[s], use(j, p.k_stages) & 1);
        mbar_wait(&bar_sfree[sb], (use(j, p.s_bufs) & 1) ^ 1);     // softmax has loaded what S buffer sb held before
        tc_fence_after();
        // S(j) = Q K(j)^T into S buffer sb -- issued before waiting for P(j-1), so it overlaps softmax(j-1)
        const uint32_t tS = tmem_base + (uint32_t)(sb * 128);
        for (int c = 0; c < dch; ++c) {
          const uint64_t ad = make_desc_kmajor_sw128(smem_u32(sQ + c * kTileBytes));
          const uint64_t bd = make_desc_kmajor_sw128(smem_u32(sK + s * qk_bytes + c * kTileBytes));
          for (int k = k_lo; k < k_hi; ++k) tc_mma_f16_ss(tS, ad + 2 * k, bd + 2 * k, idesc_s, (c > 0 || k > k_lo) ? 1u : 0u);
        }
        tc_commit(&bar_k_empty[s]);
        tc_commit(&bar_s[sb]);
        if (j > 0) {
          mbar_wait(bar_p, (j - 1) & 1);
          tc_fence_after();
          issue_pv(j - 1);
        }
      }
      mbar_wait(bar_p, (n_kv - 1) & 1);
      tc_fence_after();
      issue_pv(n_kv - 1);
    }
  } else {
    // ================================ softmax / epilogue: 2 threads per query row ================================
    const int we = warp - 2;
    const int hf = we >> 2;                 // which 64-key half of each tile / which half of the O columns
    const int quad = warp & 3;              // TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_run = -INFINITY, l_part = 0.f;
    const float c2 = p.scale_log2e;
    const int ocols = p.dpad >> 1;          // O columns owned by this thread
    for (int j = 0; j < n_kv; ++j) {
      const int sb = slot(j, p.s_bufs);
      mbar_wait(&bar_s[sb], use(j, p.s_bufs) & 1);
      tc_fence_after();
      uint32_t v[64];
      {
        const uint32_t ts = tmem_base + (uint32_t)(sb * 128 + hf * 64) + lane_off;
        uint32_t a[32], bb[32];
        tmem_ld32(ts, a);
        tmem_ld32(ts + 32, bb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = a[i]; v[32 + i] = bb[i]; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_sfree[sb]);
      const int kbase = j * 128 + hf * 64;
      float m_loc = -INFINITY;
      if (kbase + 64 > p.Nk) {       // only the ragged last half-tile carries keys past Nk (warp-uniform branch)
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float sv = (kbase + i < p.Nk) ? __uint_as_float(v[i]) : -INFINITY;
          v[i] = __float_as_uint(sv);
          m_loc = fmaxf(m_loc, sv);
        }
      } else {
        // two independent max chains (a single 64-deep chain of dependent FMNMX exposes its latency at two warps per scheduler)
        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
        for (int i = 2; i < 64; i += 2) { m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1])); }
        m_loc = fmaxf(m0, m1);
      }
      float* mx = xch + (j & 1) * 256;
      mx[hf * 128 + r] = m_loc;
      named_bar_sync(2, 256);
      // The running reference max only moves when the row max grew by more than 2^8: until then P <= 256 (well inside fp16, same
      // relative precision) and l / O keep their scale, so the TMEM round trip that rescales O is skipped for most tiles.
      const float m_cand = fmaxf(m_run, fmaxf(mx[r], mx[128 + r]));
      const float m_new = ((m_cand - m_run) * c2 > 8.f) ? m_cand : m_run;
      const float alpha = ex2_approx((m_run - m_new) * c2);
      const float mc = m_new * c2;
      float l_tile = 0.f, l_tile1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c2, -mc));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c2, -mc));
        l_tile += p0; l_tile1 += p1;
        __half2 hh = __floats2half2_rn(p0, p1);
        pk[i >> 1] = *(uint32_t*)&hh;
      }
      l_tile += l_tile1;
      l_part = l_part * alpha + l_tile;
      m_run = m_new;
      // P smem and the O accumulator are busy until PV(j-1) has completed
      if (j > 0) {
        mbar_wait(bar_o, (j - 1) & 1);
        tc_fence_after();
      }
      {
        uint8_t* rowp = sP + hf * kTileBytes + r * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *(uint4*)(rowp + ((u ^ (r & 7)) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
      // rescale this thread's half of the O row only when some row of the warp moved its running max
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
        for (int cc = 0; cc < ocols; cc += 32) {
          uint32_t o[32];
          tmem_ld32(tO + lane_off + (uint32_t)(hf * ocols + cc), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tO + lane_off + (uint32_t)(hf * ocols + cc), o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> fp16 ----
    float* lx = xch + (n_kv & 1) * 256;
    lx[hf * 128 + r] = l_part;
    named_bar_sync(2, 256);
    const float inv_l = 1.f / (lx[r] + lx[128 + r]);
    mbar_wait(bar_o, (n_kv - 1) & 1);
    tc_fence_after();
    const int q = q0 + r;
    // d32: the O accumulator holds P_h times the V columns of BOTH heads of the pair; head h's are columns [32 (h & 1), +32), i.e.
    // exactly the half owned by the threads with hf == (h & 1) -- they store, the other half of the threads has nothing to write
    const bool writer = !p.d32 || hf == (h & 1);
    __half* orow = p.out + ((size_t)b * p.Nq + q) * p.ldo + (p.d32 ? h * 32 : h * p.dpad + hf * ocols);
#pragma unroll 1
    for (int cc = 0; cc < ocols; cc += 32) {
      uint32_t o[32];
      tmem_ld32(tO + lane_off + (uint32_t)(hf * ocols + cc), o);
      tmem_ld_wait();
      if (q < p.Nq && writer) {
        uint32_t pk[16], pl[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = __uint_as_float(o[i]) * inv_l, a1 = __uint_as_float(o[i + 1]) * inv_l;
          __half2 hh = __floats2half2_rn(a0, a1);
          pk[i >> 1] = *(uint32_t*)&hh;
          const float2 fh = __half22float2(hh);
          __half2 ll = __floats2half2_rn(a0 - fh.x, a1 - fh.y);
          pl[i >> 1] = *(uint32_t*)&ll;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 hi = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          *(uint4*)(orow + cc + u * 8) = hi;
          if (p.plane > 0) {
            *(uint4*)(orow + p.plane + cc + u * 8) = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static bool g_attn_attr[64] = {};   // per device: function attributes are per device

}  // namespace upgpt

using namespace upgpt;

extern "C" int upgpt_attention(const upgpt_attn_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(a && a->q && a->k && a->vt && a->out, "attention: null pointer");
  UPGPT_REQUIRE(a->dpad == 32 || a->dpad == 64 || a->dpad == 128, "attention: dpad must be 32, 64 or 128 (got %d)", a->dpad);
  const bool d32 = a->dpad == 32;
  UPGPT_REQUIRE(!d32 || (a->H % 2 == 0 && a->v_rowmajor), "attention: dpad 32 runs on head pairs (even H) with row-major V");
  const int dpad = d32 ? 64 : a->dpad;      // width of the loaded rows (a head pair when d32)
  const int Hc = d32 ? a->H / 2 : a->H;     // head coordinate range of the TMA boxes
  UPGPT_REQUIRE(a->Nq > 0 && a->Nk > 0 && a->H > 0 && a->B > 0, "attention: bad sizes");
  UPGPT_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldvt % 8 == 0 && a->ldo % 8 == 0, "attention: ld must be multiples of 8");
  int dev_ = 0;
  UPGPT_CHECK_CUDA(cudaGetDevice(&dev_));
  UPGPT_REQUIRE(dev_ >= 0 && dev_ < 64, "attention: device index %d out of range", dev_);
  if (!g_attn_attr[dev_]) {
    UPGPT_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    UPGPT_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    g_attn_attr[dev_] = true;
  }
  CUtensorMap tmQ, tmK, tmVt;
  {
    uint64_t dims[4] = {(uint64_t)dpad, (uint64_t)Hc, (uint64_t)a->Nq, (uint64_t)a->B};
    uint64_t str[3] = {(uint64_t)dpad * 2, (uint64_t)a->ldq * 2, (uint64_t)a->ldq * 2 * a->Nq};
    uint32_t box[4] = {64, 1, 128, 1};
    if (make_tmap_f16(&tmQ, a->q, 4, dims, str, box, true)) return -3;
  }
  {
    uint64_t dims[4] = {(uint64_t)dpad, (uint64_t)Hc, (uint64_t)a->Nk, (uint64_t)a->B};
    uint64_t str[3] = {(uint64_t)dpad * 2, (uint64_t)a->ldk * 2,
                       (uint64_t)(a->k_batch_stride > 0 ? a->k_batch_stride : (long long)a->ldk * a->Nk) * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    if (make_tmap_f16(&tmK, a->k, 4, dims, str, box, true)) return -3;
  }
  if (a->v_rowmajor) {
    // V row-major [B][Nk][ldvt], head h at columns [h*dpad, (h+1)*dpad): same addressing as K
    uint64_t dims[4] = {(uint64_t)dpad, (uint64_t)Hc, (uint64_t)a->Nk, (uint64_t)a->B};
    uint64_t str[3] = {(uint64_t)dpad * 2, (uint64_t)a->ldvt * 2,
                       (uint64_t)(a->v_batch_stride > 0 ? a->v_batch_stride : (long long)a->ldvt * a->Nk) * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    if (make_tmap_f16(&tmVt, a->vt, 4, dims, str, box, true)) return -3;
  } else {
    // V^T: [B][H*dpad][ldvt], valid keys = Nk
    uint64_t dims[3] = {(uint64_t)a->Nk, (uint64_t)a->H * dpad, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->ldvt * 2, (uint64_t)a->ldvt * 2 * a->H * dpad};
    uint32_t box[3] = {64, (uint32_t)dpad, 1};
    if (make_tmap_f16(&tmVt, a->vt, 3, dims, str, box, true)) return -3;
  }
  AttnParams p{};
  p.Nq = a->Nq; p.Nk = a->Nk; p.H = a->H; p.B = a->B; p.dpad = dpad; p.d32 = d32 ? 1 : 0;
  p.num_q_tiles = (a->Nq + 127) / 128;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.out = (__half*)a->out; p.ldo = a->ldo;
  p.plane = a->split3_out ? a->H * a->dpad : 0;      // (dpad 32: the planes are H * 32 columns apart)
  p.v_mn = a->v_rowmajor ? 1 : 0;
  UPGPT_REQUIRE(!a->split3_out || a->ldo >= 2 * a->H * a->dpad, "attention: split3_out needs ldo >= 2*H*dpad");
  const int dch = dpad / 64;
  const bool one_tile = a->Nk <= 128;      // all keys in one tile: no K/V ring, one S buffer
  // compact multi-tile form (64-wide rows): K ring of 2, ONE V slot, ONE S buffer -> 96 KB of shared memory and 256 TMEM columns, so two
  // CTAs share an SM and fill each other's S -> softmax -> P -> PV bubbles (the per-tile chain, not a pipe, bounded the one-CTA form)
  const bool compact = one_tile || (dpad == 64 && a->v_rowmajor && getenv("UPGPT_ATTN_NO_COMPACT") == nullptr);
  p.k_stages = one_tile ? 1 : 2;
  p.v_stages = one_tile ? 1 : (compact ? 1 : 2);
  p.s_bufs = compact ? 1 : 2;
  p.tmem_cols = compact ? 256u : 512u;
  p.o_col = compact ? 128u : 256u;
  const size_t smem = 1024 + (size_t)dch * kTileBytes * (1 + p.k_stages) + (size_t)p.v_stages * 2 * dpad * 128 + 2 * kTileBytes + 128 +
                      2 * 256 * 4 + 64;
  const unsigned grid = (unsigned)(p.num_q_tiles * a->H * a->B);
  if (compact) UPGPT_CHECK_CUDA(launch_k(attention_kernel<2>, dim3(grid), dim3(kAttnThreads), smem, stream, tmQ, tmK, tmVt, p));
  else UPGPT_CHECK_CUDA(launch_k(attention_kernel<1>, dim3(grid), dim3(kAttnThreads), smem, stream, tmQ, tmK, tmVt, p));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

UPGPT_TRACE_TU(attention)
