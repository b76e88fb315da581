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
    const bool atomic = (p.flags & GEMM_ATOMIC) != 0;
    const bool chw = (p.flags & GEMM_CHW) != 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int bidx = tile / tiles_per_batch;
      int rem = tile - bidx * tiles_per_batch;
      const int split = rem / tiles_mn;
      rem -= split * tiles_mn;
      const int nt = rem / p.num_m_tiles;
      const int mt = rem - nt * p.num_m_tiles;
      // ---- row bookkeeping ----
      bool valid;
      long long grow;  // global output row
      if (conv) {
        if (p.tile_imgs > 1) {
          const int img = mt * p.tile_imgs + r / p.HW;
          valid = (r < p.tile_imgs * p.HW) && (img < p.n_imgs);
          grow = (long long)mt * p.tile_imgs * p.HW + r;
        } else if (p.tiles_per_row > 1) {
          valid = true;                       // W % 128 == 0: every lane is a pixel
          grow = (long long)mt * 128 + r;     // tiles enumerate the image in raster order
        } else {
          const int img = mt / p.tiles_per_img;
          const int y0 = (mt - img * p.tiles_per_img) * p.tile_rows;
          const int y = y0 + r / p.W;
          valid = (r < p.tile_rows * p.W) && (y < p.H);
          grow = (long long)img * p.HW + (long long)y0 * p.W + r;
        }
      } else {
        const int m = mt * 128 + r;
        valid = m < p.M_total;
        grow = (long long)bidx * p.M_total + m;
      }
      const int group = (int)(grow / p.rows_per_group);
      const int rig = (int)(grow - (long long)group * p.rows_per_group);
      const bool extras = !atomic || split == 0;
      const float* rv = (p.rowvec && extras) ? p.rowvec + (size_t)group * p.ld_rowvec : nullptr;
      const float* rs = (p.res32 && extras) ? p.res32 + (size_t)grow * p.ldres : nullptr;
      const float* bs = (p.bias && extras) ? p.bias : nullptr;

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
        for (int j0 = 0; j0 < p.block_n; j0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_acc + (uint32_t)j0, v);
          tmem_ld_wait();
          if (valid) {
            const int n0 = nt * p.block_n + j0;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float x = __uint_as_float(v[i]) * p.out_scale;
              const int n = n0 + i;
              if (n < p.N_total) {
                if (bs) x += bs[n];
                if (rv) x += rv[n];
                if (rs) x += rs[n];
              }
              f[i] = x;
            }
            if (chw) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = n0 + i;
                if (n < p.N_total) {
                  const size_t o = ((size_t)group * p.N_total + n) * (size_t)p.ldT + rig;
                  if (p.out32) {
                    if (atomic) atomicAdd(p.out32 + o, f[i]); else p.out32[o] = f[i];
                  }
                  if (p.out16) p.out16[o] = __float2half_rn(f[i]);
                }
              }
            } else if (n0 + 16 <= p.N_total) {
              if (p.out32) {
                float* dst = p.out32 + (size_t)grow * p.ld32 + n0;
                if (atomic) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) atomicAdd(dst + i, f[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) ((float4*)dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                }
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
                  if (p.out32) {
                    float* dst = p.out32 + (size_t)grow * p.ld32 + n;
                    if (atomic) atomicAdd(dst, f[i]); else *dst = f[i];
                  }
                  if (p.out16) p.out16[(size_t)grow * p.ld16 + n] = __float2half_rn(f[i]);
                }
              }
            }
          }
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

static int gemm_device_setup() {
  if (g_attr_set) return 0;
  int dev = 0;
  UPGPT_CHECK_CUDA(cudaGetDevice(&dev));
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  UPGPT_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
  g_attr_set = true;
  return 0;
}

static int pick_block_n(int N, int want_ctas_per_mtile_hint) {
  // largest tile that divides N and is a legal UMMA N (multiple of 16, <= 256); else pad with the smallest waste.
  static const int cands[] = {256, 224, 192, 160, 128, 112, 96, 80, 64, 48, 32, 16};
  if (N <= 256 && N % 16 == 0 && want_ctas_per_mtile_hint <= 1) return N;
  int best = 0;
  for (int c : cands) {
    if (N % c == 0 && N / c >= want_ctas_per_mtile_hint) { best = c; break; }
  }
  if (best) return best;
  for (int c : cands) if (N % c == 0) return c;   // could not reach the hint; take the largest divisor
  int n = ((N + 15) / 16) * 16;
  return n <= 256 ? n : 128;
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
  if (bn <= 0) {
    int hint = 1;
    if (p.num_m_tiles * ((a->N + 255) / 256) * p.batch < g_num_sms / 2) hint = 2;
    if (p.num_m_tiles * p.batch <= 8) hint = 4;
    bn = pick_block_n(a->N, hint);
  }
  UPGPT_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= 256, "upgpt_gemm: block_n=%d illegal", bn);
  if (p.flags & GEMM_GEGLU) UPGPT_REQUIRE(bn % 32 == 0 && a->N % bn == 0 && a->out16, "upgpt_gemm: GEGLU needs block_n%%32==0, N%%block_n==0, out16");
  p.block_n = bn;
  p.num_n_tiles = (a->N + bn - 1) / bn;
  const int k_iters = p.taps * p.kblocks_per_tap;
  int splits = a->splits;
  if (splits <= 0) {
    splits = 1;
    const int base_tiles = p.num_m_tiles * p.num_n_tiles * p.batch;
    if (a->out32 && !a->out16 && !(p.flags & GEMM_GEGLU) && base_tiles * 2 <= g_num_sms && k_iters >= 8) {
      splits = g_num_sms / base_tiles;
      if (splits > k_iters / 4) splits = k_iters / 4;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > k_iters) splits = k_iters;
  // every split must own at least one k iteration
  while (splits > 1 && (splits - 1) * ((k_iters + splits - 1) / splits) >= k_iters) --splits;
  p.num_splits = splits;
  if (splits > 1 && a->res32 == a->out32) splits = 1;  // in-place residual cannot be combined with the zero-fill
  p.num_splits = splits;
  if (splits > 1) {
    UPGPT_REQUIRE(a->out32 && !a->out16 && !(p.flags & GEMM_GEGLU), "upgpt_gemm: split-K needs an fp32-only output");
    p.flags |= GEMM_ATOMIC;
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

  if (p.flags & GEMM_ATOMIC) {
    // split-K partial sums are reduced with red.add.f32: zero the destination first (a memset node under capture)
    const size_t rows = (size_t)m_rows_total * p.batch;
    if (p.flags & GEMM_CHW) {
      const size_t groups = (rows + p.rows_per_group - 1) / p.rows_per_group;
      UPGPT_CHECK_CUDA(cudaMemsetAsync(p.out32, 0, groups * p.N_total * (size_t)p.ldT * 4, stream));
    } else {
      UPGPT_CHECK_CUDA(cudaMemset2DAsync(p.out32, (size_t)p.ld32 * 4, 0, (size_t)p.N_total * 4, rows, stream));
    }
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.num_splits * p.batch;
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  tc_gemm_kernel<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, p);
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}
