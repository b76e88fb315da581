// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N] = A[M, K] * W[N, K]^T   (fp16 operands, fp32 accumulation in TMEM)
//
// * A is fetched by TMA through a 4-D tensor map. For token matrices it is (K, M, 1, batch); for NHWC images it is
//   (C, W, H, n_imgs) and one M tile (128 accumulator lanes) is a box of whole image rows (or whole images when
//   H*W < 128). A 3x3 tap is then just the same box shifted by (dy, dx): TMA zero-fills out-of-bounds coordinates,
//   which *is* the convolution's zero padding, so im2col never exists in memory.
// * W is (K, taps, N, batch) - one box per (tap, k-block, n-tile); channel tails are zero-filled by TMA.
// * Warp roles: warp0 = TMA producer, warp1 = TMEM allocator + single-thread tcgen05.mma issuer,
//   warps2-5 = epilogue (tcgen05.ld -> bias / timestep-embedding row vector / residual / GEGLU -> global).
// * Persistent over tiles with a 2-deep TMEM accumulator ring so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Replaces (reference call sites): F.conv2d in ResBlock/Downsample/Upsample (openaimodel.py:116-118,151-153,204,230,241),
// nn.Linear / 1x1 conv in SpatialTransformer/CrossAttention/FeedForward (attention.py:37-64,161-168,233-248),
// and the VAE decoder convolutions (model.py:82-141,535-568).
#include "common.cuh"
#include "tc_gemm.cuh"
#include "../../include/upgpt_b200.h"

namespace upgpt {

static constexpr int kGemmThreads = 192;
static constexpr int kABytes = 128 * 64 * 2;  // smem slot for one A stage

__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * kABytes;
  uint64_t* bar_full = (uint64_t*)(sB + (size_t)p.stages * b_bytes);
  uint64_t* bar_empty = bar_full + p.stages;
  uint64_t* bar_tfull = bar_empty + p.stages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_base_smem = (uint32_t*)(bar_tempty + 2);
  uint32_t* split_flag = tmem_base_smem + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // accumulator ring: 2 stages of block_n columns
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.block_n) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int tiles_per_batch = tiles_mn * p.num_splits;
  const int num_tiles = tiles_per_batch * p.batch;
  const int k_iters_total = p.taps * p.kblocks_per_tap;
  const int k_per_split = (k_iters_total + p.num_splits - 1) / p.num_splits;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int bidx = tile / tiles_per_batch;
        int rem = tile - bidx * tiles_per_batch;
        const int split = rem / tiles_mn;
        rem -= split * tiles_mn;
        const int nt = rem / p.num_m_tiles;
        const int mt = rem - nt * p.num_m_tiles;
        int c1, c2, c3;  // A box origin (before tap shift)
        if (p.flags & GEMM_CONV) {
          if (p.tile_imgs > 1) {
            c1 = 0; c2 = 0; c3 = mt * p.tile_imgs;
          } else if (p.tiles_per_row > 1) {
            const int img = mt / p.tiles_per_img;
            const int t = mt - img * p.tiles_per_img;
            c1 = (t % p.tiles_per_row) * 128; c2 = t / p.tiles_per_row; c3 = img;
          } else {
            const int img = mt / p.tiles_per_img;
            c1 = 0; c2 = (mt - img * p.tiles_per_img) * p.tile_rows; c3 = img;
          }
        } else {
          c1 = mt * 128; c2 = 0; c3 = bidx;
        }
        const int k_begin = split * k_per_split;
        const int k_end = min(k_begin + k_per_split, k_iters_total);
        for (int kit = k_begin; kit < k_end; ++kit) {
          const int tap = kit / p.kblocks_per_tap;
          const int kb = kit - tap * p.kblocks_per_tap;
          mbar_wait(&bar_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_full[stage], p.a_bytes + b_bytes);
          tma_load_4d(sA + (size_t)stage * kABytes, &tmA, &bar_full[stage], kb * 64, c1 + p.tap_dx[tap],
                      c2 + p.tap_dy[tap], c3 + p.tap_dn[tap]);
          tma_load_4d(sB + (size_t)stage * b_bytes, &tmB, &bar_full[stage], kb * 64, tap, nt * p.block_n, bidx);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, (uint32_t)p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int rem = tile % tiles_per_batch;
        const int split = rem / tiles_mn;
        const int k_begin = split * k_per_split;
        const int k_end = min(k_begin + k_per_split, k_iters_total);
        mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        for (int kit = k_begin; kit < k_end; ++kit) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_desc_kmajor_sw128(smem_u32(sA + (size_t)stage * kABytes));
          const uint64_t bdesc = make_desc_kmajor_sw128(smem_u32(sB + (size_t)stage * b_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes (16 fp16) along K inside the 128-byte swizzle row: +2 in the encoded start address
            tc_mma_f16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (kit > k_begin || k > 0) ? 1u : 0u);
          }
          tc_commit(&bar_empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&bar_tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;       // accumulator row handled by this thread
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool conv = (p.flags & GEMM_CONV) != 0;
    const bool chw = (p.flags & GEMM_CHW) != 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int bidx = tile / tiles_per_batch;
      int rem = tile - bidx * tiles_per_batch;
      const int split = rem / tiles_mn;
      rem -= split * tiles_mn;
      const int nt = rem / p.num_m_tiles;
      const int mt = rem - nt * p.num_m_tiles;
      // ---- row bookkeeping: tile row -> (valid, global output row) ----
      auto row_of = [&](int rr, bool& ok, long long& gr) {
        if (conv) {
          if (p.tile_imgs > 1) {
            const int img = mt * p.tile_imgs + rr / p.HW;
            ok = (rr < p.tile_imgs * p.HW) && (img < p.n_imgs);
            gr = (long long)mt * p.tile_imgs * p.HW + rr;
          } else if (p.tiles_per_row > 1) {
            ok = true;                          // W % 128 == 0: every lane is a pixel
            gr = (long long)mt * 128 + rr;      // tiles enumerate the image in raster order
          } else {
            const int img = mt / p.tiles_per_img;
            const int y0 = (mt - img * p.tiles_per_img) * p.tile_rows;
            const int y = y0 + rr / p.W;
            ok = (rr < p.tile_rows * p.W) && (y < p.H);
            gr = (long long)img * p.HW + (long long)y0 * p.W + rr;
          }
        } else {
          const int m = mt * 128 + rr;
          ok = m < p.M_total;
          gr = (long long)bidx * p.M_total + m;
        }
      };
      bool valid;
      long long grow;  // global output row of this thread's accumulator lane
      row_of(r, valid, grow);
      const int group = (int)(grow / p.rows_per_group);
      const int rig = (int)(grow - (long long)group * p.rows_per_group);
      const float* rv = p.rowvec ? p.rowvec + (size_t)group * p.ld_rowvec : nullptr;
      const float* rs = p.res32 ? p.res32 + (size_t)grow * p.ldres : nullptr;
      const float* bs = p.bias;

      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * p.block_n) + ((uint32_t)(quad * 32) << 16);

      if (p.flags & GEMM_GEGLU) {
        const int half_n = p.block_n >> 1;
        for (int j0 = 0; j0 < half_n; j0 += 16) {
          uint32_t xr[16], gr[16];
          tmem_ld16(t_acc + (uint32_t)j0, xr);
          tmem_ld16(t_acc + (uint32_t)(half_n + j0), gr);
          tmem_ld_wait();
          if (valid) {
            const int ncol_x = nt * p.block_n + j0;            // packed column of x
            const int ncol_g = nt * p.block_n + half_n + j0;   // packed column of gate
            const int ocol = nt * half_n + j0;
            __align__(16) __half o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float xv = __uint_as_float(xr[i]) + (bs ? bs[ncol_x + i] : 0.f);
              float gv = __uint_as_float(gr[i]) + (bs ? bs[ncol_g + i] : 0.f);
              o[i] = __float2half_rn(xv * gelu_erf_f(gv));
            }
            uint4* dst = (uint4*)(p.out16 + (size_t)grow * p.ld16 + ocol);
            dst[0] = ((uint4*)o)[0];
            dst[1] = ((uint4*)o)[1];
          }
        }
      } else {
        // applies bias / row vector / residual to 16 consecutive columns of this thread's row and stores them
        auto store_cols = [&](int n0, float (&f)[16]) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + i;
            if (n < p.N_total) {
              if (p.bias) f[i] += p.bias[n];
              if (rv) f[i] += rv[n];
              if (rs) f[i] += rs[n];
            }
          }
          if (chw) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + i;
              if (n < p.N_total) {
                const size_t o = ((size_t)group * p.N_total + n) * (size_t)p.ldT + rig;
                if (p.out32) p.out32[o] = f[i];
                if (p.out16) p.out16[o] = __float2half_rn(f[i]);
              }
            }
          } else if (n0 + 16 <= p.N_total) {
            if (p.out32) {
              float* dst = p.out32 + (size_t)grow * p.ld32 + n0;
#pragma unroll
              for (int i = 0; i < 4; ++i) ((float4*)dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
            if (p.out16) {
              __align__(16) __half o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float2half_rn(f[i]);
              uint4* dst = (uint4*)(p.out16 + (size_t)grow * p.ld16 + n0);
              dst[0] = ((uint4*)o)[0];
              dst[1] = ((uint4*)o)[1];
            }
          } else {
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + i;
              if (n < p.N_total) {
                if (p.out32) p.out32[(size_t)grow * p.ld32 + n] = f[i];
                if (p.out16) p.out16[(size_t)grow * p.ld16 + n] = __float2half_rn(f[i]);
              }
            }
          }
        };
        if (p.num_splits == 1) {
          for (int j0 = 0; j0 < p.block_n; j0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_acc + (uint32_t)j0, v);
            tmem_ld_wait();
            if (valid) {
              float f[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
              store_cols(nt * p.block_n + j0, f);
            }
          }
        } else {
          // ---- deterministic split-K: partial tile -> workspace; the last split to arrive reduces in fixed order ----
          const size_t ws_split_stride = (size_t)p.ws_rows * p.ws_ld;
          float* wrow = p.ws + (size_t)split * ws_split_stride + (size_t)grow * p.ws_ld + nt * p.block_n;
          for (int j0 = 0; j0 < p.block_n; j0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_acc + (uint32_t)j0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                __stcg((float4*)(wrow + j0) + i, make_float4(__uint_as_float(v[4 * i]) * p.out_scale, __uint_as_float(v[4 * i + 1]) * p.out_scale,
                                                          __uint_as_float(v[4 * i + 2]) * p.out_scale, __uint_as_float(v[4 * i + 3]) * p.out_scale));
            }
          }
          // the accumulator stage can be recycled now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tempty[acc]);
          __threadfence();
          named_bar_sync(1, 128);
          if (r == 0) {
            int* ctr = p.counters + (bidx * tiles_mn + nt * p.num_m_tiles + mt);
            const int prev = atomicAdd(ctr, 1);
            const int last = prev == p.num_splits - 1;
            if (last) *ctr = 0;   // self-reset: the counters are zero again when the kernel ends
            *split_flag = (uint32_t)last;
          }
          named_bar_sync(1, 128);
          const bool is_last = *split_flag != 0;
          named_bar_sync(1, 128);   // everyone has read the flag before a later tile may overwrite it
          if (is_last) {
            // coalesced reduction: the 128 epilogue threads sweep the tile as a flat array of float4
            __threadfence();
            const int n4 = p.block_n >> 2;
            for (int idx = r; idx < 128 * n4; idx += 128) {
              const int rr = idx / n4;
              const int c = (idx - rr * n4) << 2;
              bool ok; long long gr;
              row_of(rr, ok, gr);
              const int n = nt * p.block_n + c;
              if (!ok || n >= p.N_total) continue;
              const float4* src = (const float4*)(p.ws + (size_t)gr * p.ws_ld + n);
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              int sp = 0;
              for (; sp + 4 <= p.num_splits; sp += 4) {   // 4 independent loads in flight, summed in split order
                const float4 t0 = __ldcg(src + (size_t)(sp + 0) * (ws_split_stride >> 2));
                const float4 t1 = __ldcg(src + (size_t)(sp + 1) * (ws_split_stride >> 2));
                const float4 t2 = __ldcg(src + (size_t)(sp + 2) * (ws_split_stride >> 2));
                const float4 t3 = __ldcg(src + (size_t)(sp + 3) * (ws_split_stride >> 2));
                a.x += t0.x; a.y += t0.y; a.z += t0.z; a.w += t0.w;
                a.x += t1.x; a.y += t1.y; a.z += t1.z; a.w += t1.w;
                a.x += t2.x; a.y += t2.y; a.z += t2.z; a.w += t2.w;
                a.x += t3.x; a.y += t3.y; a.z += t3.z; a.w += t3.w;
              }
              for (; sp < p.num_splits; ++sp) {
                const float4 t0 = __ldcg(src + (size_t)sp * (ws_split_stride >> 2));
                a.x += t0.x; a.y += t0.y; a.z += t0.z; a.w += t0.w;
              }
              if (p.bias) { const float4 b4 = *(const float4*)(p.bias + n); a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w; }
              if (p.rowvec) {
                const float4 b4 = *(const float4*)(p.rowvec + (size_t)(gr / p.rows_per_group) * p.ld_rowvec + n);
                a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
              }
              if (p.res32) { const float4 b4 = *(const float4*)(p.res32 + (size_t)gr * p.ldres + n); a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w; }
              if (p.out32) *(float4*)(p.out32 + (size_t)gr * p.ld32 + n) = a;
              if (p.out16) {
                __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
                *(uint2*)(p.out16 + (size_t)gr * p.ld16 + n) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
              }
            }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int g_smem_optin = 0;
static bool g_attr_set = false;
static float* g_ws = nullptr;                 // split-K partial-tile workspace (allocated once: graph-stable address)
static size_t g_ws_bytes = 0;
static int* g_counters = nullptr;
static constexpr int kMaxCounters = 1 << 16;

static int gemm_device_setup() {
  if (g_attr_set) return 0;
  int dev = 0;
  UPGPT_CHECK_CUDA(cudaGetDevice(&dev));
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  UPGPT_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
  g_ws_bytes = (size_t)96 << 20;
  UPGPT_CHECK_CUDA(cudaMalloc(&g_ws, g_ws_bytes));
  UPGPT_CHECK_CUDA(cudaMalloc(&g_counters, kMaxCounters * sizeof(int)));
  UPGPT_CHECK_CUDA(cudaMemset(g_counters, 0, kMaxCounters * sizeof(int)));
  g_attr_set = true;
  return 0;
}

// Tile-width heuristic: the widest legal UMMA N (multiple of 16, <= 256) that divides N when the M tiles alone fill the
// machine; otherwise the narrowest divisor (>= 64) whose tile count still fits in one wave, so that small-M layers
// (weight-bandwidth bound) spread their weight stream over all SMs before split-K has to.
static int pick_block_n(int N, int base_tiles, int num_sms) {
  static const int cands[] = {256, 224, 192, 160, 128, 112, 96, 80, 64};
  int widest = 0;
  for (int c : cands) if (N % c == 0) { widest = c; break; }
  if (!widest) {
    int n = ((N + 15) / 16) * 16;
    return n <= 256 ? n : 128;   // ragged N: TMA zero-fills the weight tail
  }
  if (base_tiles * (N / widest) >= num_sms) return widest;
  int best = widest;
  for (int c : cands) {
    if (N % c) continue;
    if (base_tiles * (N / c) <= num_sms) best = c; else break;
  }
  return best;
}

}  // namespace upgpt

using namespace upgpt;

extern "C" int upgpt_gemm(const upgpt_gemm_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (gemm_device_setup()) return -2;
  UPGPT_REQUIRE(a && a->a && a->w, "upgpt_gemm: null operand");
  UPGPT_REQUIRE(a->out32 || a->out16, "upgpt_gemm: no output");
  UPGPT_REQUIRE(a->K > 0 && a->N > 0, "upgpt_gemm: bad K/N");
  UPGPT_REQUIRE(a->K % 8 == 0, "upgpt_gemm: K (=%d) must be a multiple of 8 (16-byte TMA rows)", a->K);
  const bool conv = a->mode == UPGPT_GEMM_CONV3X3 || a->mode == UPGPT_GEMM_CONV3X3_S2PHASE || a->mode == UPGPT_GEMM_CONV1X1;
  GemmParams p{};
  p.flags = a->flags;
  p.batch = conv ? 1 : (a->batch > 0 ? a->batch : 1);
  p.N_total = a->N;
  p.taps = (a->mode == UPGPT_GEMM_CONV3X3 || a->mode == UPGPT_GEMM_CONV3X3_S2PHASE) ? 9 : 1;
  p.kblocks_per_tap = (a->K + 63) / 64;
  p.out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;

  CUtensorMap tmA, tmB;
  int m_rows_total;
  if (conv) {
    UPGPT_REQUIRE(a->H > 0 && a->W > 0 && a->n_imgs > 0, "upgpt_gemm(conv): bad geometry");
    UPGPT_REQUIRE(a->W <= 128 || a->W % 128 == 0, "upgpt_gemm(conv): W=%d > 128 must be a multiple of 128", a->W);
    p.flags |= GEMM_CONV;
    p.H = a->H; p.W = a->W; p.HW = a->H * a->W; p.n_imgs = a->n_imgs;
    uint32_t box[4];
    p.tiles_per_row = 1;
    if (a->W > 128) {
      p.tile_imgs = 1; p.tile_rows = 1; p.tiles_per_row = a->W / 128;
      p.tiles_per_img = p.tiles_per_row * a->H;
      p.num_m_tiles = p.tiles_per_img * a->n_imgs;
      box[0] = 64; box[1] = 128; box[2] = 1; box[3] = 1;
    } else if (p.HW <= 64) {
      p.tile_imgs = 128 / p.HW; p.tile_rows = a->H; p.tiles_per_img = 1;
      if (p.tile_imgs > a->n_imgs) p.tile_imgs = a->n_imgs;
      if (p.tile_imgs < 1) p.tile_imgs = 1;
      p.num_m_tiles = (a->n_imgs + p.tile_imgs - 1) / p.tile_imgs;
      box[0] = 64; box[1] = a->W; box[2] = a->H; box[3] = p.tile_imgs;
    } else {
      p.tile_imgs = 1; p.tile_rows = 128 / a->W; if (p.tile_rows > a->H) p.tile_rows = a->H;
      p.tiles_per_img = (a->H + p.tile_rows - 1) / p.tile_rows;
      p.num_m_tiles = p.tiles_per_img * a->n_imgs;
      box[0] = 64; box[1] = a->W; box[2] = p.tile_rows; box[3] = 1;
    }
    p.a_bytes = box[0] * box[1] * box[2] * box[3] * 2;
    const int a_imgs = (a->mode == UPGPT_GEMM_CONV3X3_S2PHASE) ? 4 * a->n_imgs : a->n_imgs;
    uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a_imgs};
    const uint64_t lda = a->lda > 0 ? a->lda : a->K;
    uint64_t strides[3] = {lda * 2, lda * 2 * a->W, lda * 2 * a->W * a->H};
    if (make_tmap_f16(&tmA, a->a, 4, dims, strides, box, true)) return -3;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = p.tap_dx[t] = p.tap_dn[t] = 0; }
    if (a->mode == UPGPT_GEMM_CONV3X3) {
      for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3 - 1; p.tap_dx[t] = t % 3 - 1; }
    } else if (a->mode == UPGPT_GEMM_CONV3X3_S2PHASE) {
      // out(y,x) tap (r,s) reads in(2y+r-1, 2x+s-1) = phase[(r+1)&1][(s+1)&1] at (y + (r==0 ? -1 : 0), x + (s==0 ? -1 : 0))
      for (int t = 0; t < 9; ++t) {
        const int rr = t / 3, ss = t % 3;
        const int ph = ((rr + 1) & 1) * 2 + ((ss + 1) & 1);
        p.tap_dy[t] = rr == 0 ? -1 : 0; p.tap_dx[t] = ss == 0 ? -1 : 0; p.tap_dn[t] = ph * a->n_imgs;
      }
    }
    m_rows_total = a->n_imgs * p.HW;
    p.M_total = m_rows_total;
    p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : p.HW;
  } else {
    UPGPT_REQUIRE(a->M > 0, "upgpt_gemm: bad M");
    p.M_total = a->M;
    p.num_m_tiles = (a->M + 127) / 128;
    p.a_bytes = kABytes;
    const uint64_t lda = a->lda > 0 ? a->lda : a->K;
    const uint64_t abs_ = a->a_batch_stride > 0 ? (uint64_t)a->a_batch_stride : lda * (uint64_t)a->M;
    uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)a->M, 1, (uint64_t)p.batch};
    uint64_t strides[3] = {lda * 2, abs_ * 2, abs_ * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    if (make_tmap_f16(&tmA, a->a, 4, dims, strides, box, true)) return -3;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = p.tap_dx[t] = p.tap_dn[t] = 0; }
    m_rows_total = a->M;
    p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : a->M;
  }

  // ---- tile shape / split-K ----
  int bn = a->block_n;
  if (bn <= 0) bn = pick_block_n(a->N, p.num_m_tiles * p.batch, g_num_sms);
  UPGPT_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= 256, "upgpt_gemm: block_n=%d illegal", bn);
  if (p.flags & GEMM_GEGLU) UPGPT_REQUIRE(bn % 32 == 0 && a->N % bn == 0 && a->out16, "upgpt_gemm: GEGLU needs block_n%%32==0, N%%block_n==0, out16");
  p.block_n = bn;
  p.num_n_tiles = (a->N + bn - 1) / bn;
  const int k_iters = p.taps * p.kblocks_per_tap;
  int splits = a->splits;
  if (splits <= 0) {
    splits = 1;
    const int base_tiles = p.num_m_tiles * p.num_n_tiles * p.batch;
    if (!(p.flags & (GEMM_GEGLU | GEMM_CHW)) && base_tiles * 2 <= g_num_sms && k_iters >= 8) {
      splits = g_num_sms / base_tiles;
      if (splits > k_iters / 4) splits = k_iters / 4;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > k_iters) splits = k_iters;
  // every split must own at least one k iteration
  while (splits > 1 && (splits - 1) * ((k_iters + splits - 1) / splits) >= k_iters) --splits;
  p.num_splits = splits;
  if ((p.flags & GEMM_CHW) || a->N % 4 != 0) splits = 1;
  p.num_splits = splits;
  if (splits > 1) {
    UPGPT_REQUIRE(!(p.flags & GEMM_GEGLU), "upgpt_gemm: split-K is not available with the GEGLU epilogue");
    // deterministic split-K workspace: [splits][rows][ws_ld] fp32 partial tiles + one arrival counter per output tile
    p.ws_ld = p.num_n_tiles * bn;
    p.ws_rows = m_rows_total * p.batch;
    const size_t need = (size_t)splits * p.ws_rows * p.ws_ld * sizeof(float);
    const int n_ctr = p.num_m_tiles * p.num_n_tiles * p.batch;
    if (need > g_ws_bytes || n_ctr > kMaxCounters) {
      // shrink the split factor to fit the fixed workspace (its address must stay stable for captured graphs)
      while (splits > 1 && ((size_t)splits * p.ws_rows * p.ws_ld * sizeof(float) > g_ws_bytes || n_ctr > kMaxCounters)) --splits;
      while (splits > 1 && (splits - 1) * ((k_iters + splits - 1) / splits) >= k_iters) --splits;
      p.num_splits = splits;
    }
    p.ws = g_ws;
    p.counters = g_counters;
  }
  {
    const uint64_t ldw = a->ldw > 0 ? a->ldw : a->K;  // elements between taps
    const uint64_t n_stride = ldw * p.taps;            // elements between output channels
    const uint64_t wbs = a->w_batch_stride > 0 ? (uint64_t)a->w_batch_stride : n_stride * (uint64_t)a->N;
    uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)p.taps, (uint64_t)a->N, (uint64_t)p.batch};
    uint64_t strides[3] = {ldw * 2, n_stride * 2, wbs * 2};
    uint32_t box[4] = {64, 1, (uint32_t)bn, 1};
    if (make_tmap_f16(&tmB, a->w, 4, dims, strides, box, true)) return -3;
  }

  p.out32 = a->out32; p.ld32 = a->ld32 > 0 ? a->ld32 : a->N;
  p.out16 = (__half*)a->out16; p.ld16 = a->ld16 > 0 ? a->ld16 : ((p.flags & GEMM_GEGLU) ? a->N / 2 : a->N);
  p.bias = a->bias; p.rowvec = a->rowvec; p.ld_rowvec = a->ld_rowvec > 0 ? a->ld_rowvec : a->N;
  p.res32 = a->res32; p.ldres = a->ldres > 0 ? a->ldres : a->N;
  p.ldT = a->ldT > 0 ? a->ldT : p.rows_per_group;
  if (!(p.flags & GEMM_CHW)) {
    UPGPT_REQUIRE(!p.out32 || (p.ld32 % 4 == 0 && ((uintptr_t)p.out32 & 15) == 0), "upgpt_gemm: out32 must be 16-byte aligned with ld%%4==0");
    UPGPT_REQUIRE(!p.out16 || (p.ld16 % 8 == 0 && ((uintptr_t)p.out16 & 15) == 0), "upgpt_gemm: out16 must be 16-byte aligned with ld%%8==0");
  }

  // ---- pipeline depth from the smem budget ----
  const size_t stage_bytes = (size_t)kABytes + (size_t)bn * 128;
  int stages = (int)(((size_t)g_smem_optin - 1024 - 256) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages > k_iters / splits + 1) stages = k_iters / splits + 1;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + (2 * stages + 4) * 8 + 16;
  UPGPT_REQUIRE(smem <= (size_t)g_smem_optin, "upgpt_gemm: smem %zu > %d", smem, g_smem_optin);

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.num_splits * p.batch;
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  tc_gemm_kernel<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, p);
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}
