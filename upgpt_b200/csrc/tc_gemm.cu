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
//   warps2-9 = two epilogue groups (tcgen05.ld -> bias / timestep-embedding row vector / residual / GEGLU -> global).
// * Persistent over tiles with a 2-deep TMEM accumulator ring so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Error-compensated mode (GEMM_X3, the precision the parity gate runs in): both operands carry [hi | lo] fp16 planes
//   (x = hi + lo to ~22 bits). The planes are a 5th tensor-map dimension; per 64-wide k-block the producer loads the hi
//   pair (Ah, Wh) and the lo pair (Al, Wl) into two consecutive pipeline slots and the issuer forms
//   Ah*Wh + Al*Wh + Ah*Wl from them: 3x the MMA work on 2x (not 3x) the operand bytes.
//
// Replaces (reference call sites): F.conv2d in ResBlock/Downsample/Upsample (openaimodel.py:116-118,151-153,204,230,241),
// nn.Linear / 1x1 conv in SpatialTransformer/CrossAttention/FeedForward (attention.py:37-64,161-168,233-248),
// and the VAE decoder convolutions (model.py:82-141,535-568).
#include "common.cuh"
#include "tc_gemm.cuh"
#include <type_traits>
#include <cstdlib>
#include <mutex>
#include "../../include/upgpt_b200.h"

namespace upgpt {

static constexpr int kGemmThreads = 320;   // TMA producer warp + MMA issuer warp + 2 epilogue groups of 4 warps
static constexpr int kABytes = 128 * 64 * 2;  // smem slot for one A stage

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory"); return t; }
#define TS(slot) do { if (p.debug_ts) p.debug_ts[(size_t)blockIdx.x * 16 + (slot)] = gtime(); } while (0)

// fp16 store of 4 consecutive columns; with `plane` > 0 also the error-compensation plane: [hi | lo] at column
// offsets 0, plane (operand layout of the fp16x3 precision mode)
__device__ __forceinline__ void store_h4(__half* dst, float4 v, int plane) {
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const uint2 hi = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
  *(uint2*)dst = hi;
  if (plane > 0) {
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    *(uint2*)(dst + plane) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
  }
}

// (A two-CTAs-per-SM variant -- <= 102 registers, <= 110 KB shared memory, so that a PDL successor's prologue overlaps its predecessor's
// tail -- was measured in round 2: the spills and the 2-stage pipeline cost more than the overlap gained, 4.27 -> 4.53 ms per U-Net step.)
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmH, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // carve-up by integer offsets from the __shared__ array, so that every access below compiles to LDS/STS (casting
  // through uintptr_t would demote the pointers to the generic address space)
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * kABytes;
  const uint32_t off_bar = (uint32_t)p.stages * (kABytes + b_bytes) + p.pipe_pad;
  uint64_t* bar_full = (uint64_t*)(smem + off_bar);
  uint64_t* bar_empty = bar_full + p.stages;
  uint64_t* bar_tfull = bar_empty + p.stages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_res = bar_tempty + 2;                        // [2] residual chunk landed (TMA epilogue)
  uint32_t* tmem_base_smem = (uint32_t*)(bar_res + 2);
  uint32_t* split_flag = tmem_base_smem + 1;
  const uint32_t off_stage = (off_bar + (2u * p.stages + 6u) * 8u + 8u + 1023u) & ~1023u;
  // per epilogue group (2 groups): one staging chunk, one residual chunk, the row tables and the tile's bias
  float* stage = (float*)(smem + off_stage);                 // 2 x [128 rows][128 B] epilogue staging (swizzle-128B layout)
  float* resbuf = (float*)(smem + off_stage + 2 * 16384);    // 2 x [128 rows][32 fp32] residual chunks (only if res_tma)
  const uint32_t off_tab = off_stage + 2 * 16384 + (p.res_tma ? 2 * 16384 : 0);
  long long* row_tab = (long long*)(smem + off_tab);         // 2 x [128] global output row of each tile row (-1 = not stored)
  int* grp_tab = (int*)(smem + off_tab + 2 * 128 * 8);       // 2 x [128] row group (image / batch entry) of each tile row
  float* bias_s = (float*)(smem + off_tab + 2 * 128 * 8 + 2 * 128 * 4);   // 2 x [256] bias of this tile's columns
  float* ln_tab = bias_s + 2 * 256;            // 2 x {mean[128], rstd[128]} per-row LayerNorm statistics (folded LN consumer)
  float* lns_s = ln_tab + 2 * 256;             // 2 x [256] column sums s[n] of this tile's columns (folded LN consumer)
  float* rsx = lns_s + 2 * 256;                // [2 tile parities][128 rows][2] row-stat exchange between the epilogue groups (producer)
  const bool gn = p.gn_acc[0] != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TS(0);
  pdl_launch_dependents();

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
      mbar_init(&bar_tempty[s], 8);
      mbar_init(&bar_res[s], 1);
    }
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmR);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, tmem_cols);

  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  // image (row group) index of tile row 0 of M tile `mt`, and this CTA's moments -> shared accumulators / shared -> global
  auto tile_img_base = [&](int mt, int bidx) -> int {
    if (p.flags & GEMM_CONV) return p.tile_imgs > 1 ? mt * p.tile_imgs : p.fd_tiles_per_img.div(mt);
    return (int)(((long long)bidx * p.M_total + (long long)mt * 128) / p.rows_per_group);
  };
  // GroupNorm moments (include/upgpt_b200.h: gn_acc). The lanes of a warp hold the {sum, sumsq} of 32 consecutive result columns
  // (lane = column) over some rows of image `img`; columns of one GroupNorm group are contiguous, so a segmented shuffle reduction
  // (fixed order) leaves each group's total in the first lane of its run, which adds it to the global 64-bit fixed-point accumulator
  // with a fire-and-forget integer atomic (associative: the accumulated value does not depend on the arrival order).
  // Must be called by all 32 lanes; n < 0 marks a lane without a column.
  auto gn_commit = [&](int img, int n, float sv, float qv) {
    if (p.gn_dbg & 2) return;
#pragma unroll
    for (int cns = 0; cns < 2; ++cns) {
      if (!p.gn_acc[cns]) continue;
      const int g = n >= 0 ? (p.gn_choff[cns] + n) / p.gn_cpg[cns] : -1;
      float a = sv, b = qv;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float a2 = __shfl_down_sync(0xffffffffu, a, off), b2 = __shfl_down_sync(0xffffffffu, b, off);
        const int g2 = __shfl_down_sync(0xffffffffu, g, off);
        if (lane + off < 32 && g2 == g) { a += a2; b += b2; }
      }
      const int gprev = __shfl_up_sync(0xffffffffu, g, 1);
      if (g >= 0 && (lane == 0 || gprev != g) && !(p.gn_dbg & 1)) {
        unsigned long long* acc = (unsigned long long*)(p.gn_acc[cns] + ((size_t)img * p.gn_groups + g) * 2);
        atomicAdd(acc, (unsigned long long)__float2ll_rn(a * 16777216.f));
        atomicAdd(acc + 1, (unsigned long long)__float2ll_rn(b * 1048576.f));
      }
    }
  };

  const int tiles_per_batch = tiles_mn * p.num_splits;
  const int num_tiles = tiles_per_batch * p.batch;
  const int k_iters_total = p.taps * p.kblocks_per_tap;
  const int k_per_split = (k_iters_total + p.num_splits - 1) / p.num_splits;

  // tile index -> (batch entry, k split, n tile, m tile); in cluster mode the splits of one output tile are the CTAs of one cluster
  auto decode_tile = [&](int tile, int& bidx, int& split, int& nt, int& mt) {
    if (p.cluster_reduce) {
      const int q = p.fd_splits.div(tile);
      split = tile - q * p.num_splits;
      tile = q;
      bidx = p.fd_tiles_mn.div(tile);
    } else {
      bidx = p.fd_tiles_per_batch.div(tile);
      tile -= bidx * tiles_per_batch;
      split = p.fd_tiles_mn.div(tile);
    }
    const int rem = tile - p.fd_tiles_mn.div(tile) * tiles_mn;
    nt = p.fd_m_tiles.div(rem);
    mt = rem - nt * p.num_m_tiles;
  };

  // origin (coords 1..3) of M tile `mt` in the 4-D A / C / R tensor maps
  auto tile_origin = [&](int mt, int bidx, int& c1, int& c2, int& c3) {
    if (p.flags & GEMM_CONV) {
      if (p.tile_imgs > 1) {
        c1 = 0; c2 = 0; c3 = mt * p.tile_imgs;
      } else {
        const int img = p.fd_tiles_per_img.div(mt);
        const int t = mt - img * p.tiles_per_img;
        const int ty = p.fd_tiles_per_row.div(t);
        c1 = (t - ty * p.tiles_per_row) * p.tile_cols; c2 = ty * p.tile_rows; c3 = img;
      }
    } else {
      c1 = mt * 128; c2 = 0; c3 = bidx;
    }
  };

  // the first tile of this CTA is decoded by every thread here, while the TMEM allocation and the predecessor's tail are still
  // in flight (it only depends on blockIdx): the producer lane's first TMA is then issued right after the PDL wait
  int f_bidx, f_split, f_nt, f_mt, f_c1, f_c2, f_c3;
  decode_tile((int)blockIdx.x, f_bidx, f_split, f_nt, f_mt);
  tile_origin(f_mt, f_bidx, f_c1, f_c2, f_c3);
  auto decode_cached = [&](int tile, int& bidx, int& split, int& nt, int& mt) {
    if (tile == (int)blockIdx.x) { bidx = f_bidx; split = f_split; nt = f_nt; mt = f_mt; }
    else decode_tile(tile, bidx, split, nt, mt);
  };

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail. The producer lane goes further when W
  // is static (model weights): it issues the W boxes of its first tile's leading k-blocks into the operand ring -- and L2 prefetches of
  // the following ones -- BEFORE the dependency wait, so the weight stream from HBM overlaps the predecessor too; only the A boxes
  // (activations written by the predecessor) are issued after the wait.
  int pre_slots = 0;
  if (threadIdx.x == 0 && p.w_prefetch && (int)blockIdx.x < num_tiles) {
    const int k_begin = f_split * k_per_split;
    const int k_end = min(k_begin + k_per_split, k_iters_total);
    int tap = k_begin / p.kblocks_per_tap;
    int kb = k_begin - tap * p.kblocks_per_tap;
    int kit = k_begin;
    for (; kit < k_end && pre_slots + p.x3 < p.stages; ++kit, ++kb) {
      if (kb == p.kblocks_per_tap) { kb = 0; ++tap; }
      for (int plane = 0; plane <= p.x3; ++plane) {
        mbar_arrive_expect_tx(&bar_full[pre_slots], p.a_bytes + b_bytes);
        tma_load_5d(sB + (size_t)pre_slots * b_bytes, &tmB, &bar_full[pre_slots], kb * 64, tap, f_nt * p.block_n, f_bidx, plane);
        ++pre_slots;
      }
    }
    if (p.w_prefetch > 1) {
      for (int n = 0; kit < k_end && n < p.w_prefetch; ++kit, ++kb, ++n) {
        if (kb == p.kblocks_per_tap) { kb = 0; ++tap; }
        for (int plane = 0; plane <= p.x3; ++plane) tma_prefetch_l2_5d(&tmB, kb * 64, tap, f_nt * p.block_n, f_bidx, plane);
      }
    }
  }
  pdl_wait();
  if (threadIdx.x == 0) TS(1);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int bidx, split, nt, mt;
        decode_cached(tile, bidx, split, nt, mt);
        int c1, c2, c3;  // A box origin (before tap shift)
        if (tile == (int)blockIdx.x) { c1 = f_c1; c2 = f_c2; c3 = f_c3; }
        else tile_origin(mt, bidx, c1, c2, c3);
        const int k_begin = split * k_per_split;
        const int k_end = min(k_begin + k_per_split, k_iters_total);
        int tap = k_begin / p.kblocks_per_tap;
        int kb = k_begin - tap * p.kblocks_per_tap;
        const int tofs = (p.flags & GEMM_UP2) ? bidx * 4 : 0;   // parity-wise tap offsets (bidx = output parity)
        for (int kit = k_begin; kit < k_end; ++kit, ++kb) {
          if (kb == p.kblocks_per_tap) { kb = 0; ++tap; }
          for (int plane = 0; plane <= p.x3; ++plane) {   // x3: slot pair {(Ah, Wh), (Al, Wl)}
            if (kit == k_begin && plane == 0) TS(2);       // coordinates decoded, about to issue the first TMA
            if (pre_slots > 0) {
              // first pass over this slot: its W box and the barrier's byte count were issued before the dependency wait
              --pre_slots;
              tma_load_5d(sA + (size_t)stage * kABytes, &tmA, &bar_full[stage], kb * 64, c1 + p.tap_dx[tofs + tap],
                          c2 + p.tap_dy[tofs + tap], c3 + p.tap_dn[tofs + tap], plane);
            } else {
              mbar_wait(&bar_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&bar_full[stage], p.a_bytes + b_bytes);
              tma_load_5d(sA + (size_t)stage * kABytes, &tmA, &bar_full[stage], kb * 64, c1 + p.tap_dx[tofs + tap],
                          c2 + p.tap_dy[tofs + tap], c3 + p.tap_dn[tofs + tap], plane);
              tma_load_5d(sB + (size_t)stage * b_bytes, &tmB, &bar_full[stage], kb * 64, tap, nt * p.block_n, bidx, plane);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        TS(3);
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
        int bidx, split, nt, mt;
        decode_cached(tile, bidx, split, nt, mt);
        const int k_begin = split * k_per_split;
        const int k_end = min(k_begin + k_per_split, k_iters_total);
        mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        if (p.x3) {
          for (int kit = k_begin; kit < k_end; ++kit) {
            // slots (stage, stage + 1) hold the hi and lo plane pairs of this k-block (stages is even: same phase)
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            if (kit == k_begin) TS(4);
            if (kit == k_end - 1) TS(6);
            const uint64_t ah = make_desc_kmajor_sw128(smem_u32(sA + (size_t)stage * kABytes));
            const uint64_t wh = make_desc_kmajor_sw128(smem_u32(sB + (size_t)stage * b_bytes));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16_ss(d_tmem, ah + (uint64_t)(2 * k), wh + (uint64_t)(2 * k), idesc, (kit > k_begin || k > 0) ? 1u : 0u);
            mbar_wait(&bar_full[stage + 1], phase);
            tc_fence_after();
            const uint64_t al = make_desc_kmajor_sw128(smem_u32(sA + (size_t)(stage + 1) * kABytes));
            const uint64_t wl = make_desc_kmajor_sw128(smem_u32(sB + (size_t)(stage + 1) * b_bytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_f16_ss(d_tmem, al + (uint64_t)(2 * k), wh + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_f16_ss(d_tmem, ah + (uint64_t)(2 * k), wl + (uint64_t)(2 * k), idesc, 1u);
            tc_commit(&bar_empty[stage]);
            tc_commit(&bar_empty[stage + 1]);
            stage += 2;
            if (stage == p.stages) { stage = 0; phase ^= 1; }
          }
        } else
        for (int kit = k_begin; kit < k_end; ++kit) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          if (kit == k_begin) TS(4);
          if (kit == k_begin + 1) TS(5);
          if (kit == k_end - 1) TS(6);
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
        TS(7);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ================================
    // Two epilogue groups of 4 warps (one warp per TMEM lane quadrant in each): group g drains the accumulator column chunks c with
    // c % 2 == g.  One warp per SM sub-partition leaves every dependent instruction's latency exposed (measured: ~0.6 us per
    // 32-column chunk, 3-10 us for a split-K slice reduction); two warps per scheduler overlap them.
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;       // accumulator row handled by this thread
    const int tid2 = grp * 128 + r;       // index among the 256 epilogue threads
    const uint32_t gbar = 1u + (uint32_t)grp;   // named barrier of this group (128 threads); barrier 3 = both groups (256)
    float* const st = stage + grp * 4096;        // this group's staging chunk [128 rows][128 B], swizzle-128B layout
    float* const rbuf = resbuf + grp * 4096;     // this group's residual chunk (res_tma)
    uint64_t* const rbar = &bar_res[grp];
    long long* const rtab = row_tab + grp * 128;
    int* const gtab = grp_tab + grp * 128;
    float* const bias_g = bias_s + grp * 256;
    const bool stamp = tid2 == 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    int tile_ctr = 0;
    const bool conv = (p.flags & GEMM_CONV) != 0;
    const bool chw = (p.flags & GEMM_CHW) != 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int bidx, split, nt, mt;
      decode_cached(tile, bidx, split, nt, mt);
      // ---- row bookkeeping: tile row -> (valid, global output row) ----
      bool valid;
      long long grow;  // global output row of this thread's accumulator lane
      if (conv) {
        if (p.tile_imgs > 1) {
          const int img = mt * p.tile_imgs + r / p.HW;
          valid = (r < p.tile_imgs * p.HW) && (img < p.n_imgs);
          grow = (long long)mt * p.tile_imgs * p.HW + r;
          if (p.flags & GEMM_UP2) {
            // low-resolution pixel (y, x) of image img -> pixel (2y + a, 2x + b) of the 2H x 2W result, (a, b) = this tile's parity
            const int pix = r - (r / p.HW) * p.HW;
            const int y = pix / p.W;
            grow = (long long)img * 4 * p.HW + (long long)(2 * y + (bidx >> 1)) * (2 * p.W) + 2 * (pix - y * p.W) + (bidx & 1);
          }
        } else {
          // tile = tile_rows x tile_cols pixels of one image (tile_cols divides W; the last row tile may be ragged)
          const int img = mt / p.tiles_per_img;
          const int t = mt - img * p.tiles_per_img;
          const int ty = t / p.tiles_per_row;
          const int x0 = (t - ty * p.tiles_per_row) * p.tile_cols;
          const int ry = r / p.tile_cols;
          const int y = ty * p.tile_rows + ry;
          valid = (r < p.tile_rows * p.tile_cols) && (y < p.H);
          grow = (long long)img * p.HW + (long long)y * p.W + x0 + (r - ry * p.tile_cols);
          if (p.flags & GEMM_UP2)
            grow = (long long)img * 4 * p.HW + (long long)(2 * y + (bidx >> 1)) * (2 * p.W) + 2 * (x0 + (r - ry * p.tile_cols)) + (bidx & 1);
        }
      } else {
        const int m = mt * 128 + r;
        valid = m < p.M_total;
        grow = (long long)bidx * p.M_total + m;
      }
      const int group = (int)(grow / p.rows_per_group);
      const int rig = (int)(grow - (long long)group * p.rows_per_group);
      const float* rv = p.rowvec ? p.rowvec + (size_t)group * p.ld_rowvec : nullptr;
      const float* rs = p.res32 ? p.res32 + (size_t)grow * p.ldres : nullptr;
      const float* bs = p.bias;
      // publish this thread's row bookkeeping for the flat (coalesced) passes: first wait until every thread of the group has
      // finished reading the previous tile's tables / residual chunk; a later group barrier orders the writes before the reads
      // folded LayerNorm (consumer): this row's mean / rstd from the producer's per-tile partial sums, summed in slot order
      const bool ln = p.ln_stats != nullptr;
      float ln_mu = 0.f, ln_rs = 1.f;
      if (ln && valid) {
        const float2* sp = (const float2*)(p.ln_stats + (size_t)grow * p.ln_slots * 2);
        float su = 0.f, sq = 0.f;
        for (int i = 0; i < p.ln_slots; ++i) { const float2 t = sp[i]; su += t.x; sq += t.y; }
        ln_mu = su * p.ln_inv_c;
        const float var = fmaxf(fmaf(-ln_mu, ln_mu, sq * p.ln_inv_c), 0.f);
        ln_rs = rsqrtf(var + p.ln_eps);
      }
      float* const lnmu_t = ln_tab + grp * 256;
      float* const lnrs_t = lnmu_t + 128;
      float* const lns_g = lns_s + grp * 256;
      named_bar_sync(gbar, 128);
      rtab[r] = valid ? grow : -1;
      gtab[r] = group;
      if (ln) { lnmu_t[r] = ln_mu; lnrs_t[r] = ln_rs; }
      float rs_sum = 0.f, rs_sq = 0.f;       // producer: this row's {sum, sum of squares} over the chunks this thread handles
      const bool tma_epi = p.epi_mode != 0 && p.num_splits == 1;
      const bool f16out = p.epi_mode == 2;
      const int cw = f16out ? 64 : 32;                 // chunk width of the TMA-store epilogue
      const int nch = (p.block_n + cw - 1) / cw;
      int oc1 = 0, oc2 = 0, oc3 = 0;
      if (tma_epi) {
        tile_origin(mt, bidx, oc1, oc2, oc3);
        for (int i = r; i < p.block_n; i += 128) {
          const int n = nt * p.block_n + i;
          bias_g[i] = (p.bias && n < p.N_total) ? p.bias[n] : 0.f;
          if (ln) lns_g[i] = n < p.N_total ? p.ln_colsum[n] : 0.f;
        }
        if (p.res_tma && r == 0 && grp < nch) {
          // prefetch the residual of this group's first chunk while the main loop is still running
          mbar_arrive_expect_tx(rbar, p.a_bytes);
          tma_load_4d(rbuf, &tmR, rbar, nt * p.block_n + grp * 32, oc1, oc2, oc3);
        }
      }
      if (p.flags & GEMM_GEGLU) {     // (filled while the main loop is still running; read after the group barrier below)
        for (int i = r; i < p.block_n; i += 128) {
          const int n = nt * p.block_n + i;
          bias_g[i] = bs ? bs[n] : 0.f;
          lns_g[i] = ln ? p.ln_colsum[n] : 0.f;
        }
      }

      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
      if (stamp) TS(8);
      if (tma_epi && p.res_bulk && r == 0 && grp < nch) {
        // every MMA of this CTA's only tile has completed: the operand slots are dead. All residual chunks of this group are fetched
        // into them in one go (the per-chunk fetch exposed one L2 round trip per 32 columns).
        int cnt = 0;
        for (int c = grp; c < nch; c += 2) ++cnt;
        mbar_arrive_expect_tx(rbar, p.a_bytes * (uint32_t)cnt);
        for (int c = grp; c < nch; c += 2)
          tma_load_4d(smem + p.off_resb + (uint32_t)c * 16384u, &tmR, rbar, nt * p.block_n + c * 32, oc1, oc2, oc3);
      }
      bool bulk_waited = false;
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * p.block_n) + ((uint32_t)(quad * 32) << 16);
      // all TMEM reads of this warp for this tile are done: hand the accumulator stage back to the MMA issuer
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[acc]);
      };

      // --------------------------------------------------------------------------------------------------------
      // Epilogue data path: TMEM -> registers (one accumulator row per thread) -> XOR-swizzled smem staging chunk
      // [128 rows][32 fp32] -> either a TMA bulk store of the finished chunk, or a flat coalesced pass (8 consecutive
      // threads cover one row's 128 bytes) that applies bias / row vector / residual and stores fp32 / fp16 / split-K
      // partials with full-line transactions.  Channel-major stores skip the staging: there one row per thread is
      // already the coalesced direction.
      // --------------------------------------------------------------------------------------------------------
      const bool splitk = p.num_splits > 1;
      const size_t ws_split_stride = (size_t)p.ws_rows * p.ws_ld;

      // writes this thread's row (ncols fp32, ncols in {16, 32}) into the group's staging buffer
      auto stage_row = [&](const float* f, int ncols) {
        float* rowp = st + r * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q * 4 < ncols) *(float4*)(rowp + ((q ^ (r & 7)) << 2)) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      };
      // flat pass over the staged chunk; mode 0: final values (+extras) -> out32/out16, 1: split-K partial -> workspace,
      // 2: GEGLU result -> out16 only (n_out0 = first output column of the chunk)
      auto flush_chunk = [&](int n_out0, int ncols, int mode) {
        // thread r owns float4 column q = r % (ncols/4) of rows (r / (ncols/4)) + k * (128 / (ncols/4)), k = 0..ncols/4-1:
        // the column (hence bias) is loop-invariant and the k iterations are independent (loads first, then stores)
        const int sh = ncols == 32 ? 3 : 2;
        const int q = r & ((1 << sh) - 1);
        const int rr0 = r >> sh;
        const int rstep = 128 >> sh;
        const int n = n_out0 + (q << 2);
        const int iters = 1 << sh;
        if (mode == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k >= iters) break;
            const int rr = rr0 + k * rstep;
            const long long gr = rtab[rr];
            if (gr < 0) continue;
            const float4 v = *(const float4*)(st + rr * 32 + ((q ^ (rr & 7)) << 2));
            __stcg((float4*)(p.ws + (size_t)split * ws_split_stride + (size_t)gr * p.ws_ld + n), v);
          }
          return;
        }
        if (mode == 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k >= iters) break;
            const int rr = rr0 + k * rstep;
            const long long gr = rtab[rr];
            if (gr < 0) continue;
            const float4 v = *(const float4*)(st + rr * 32 + ((q ^ (rr & 7)) << 2));
            store_h4(p.out16 + (size_t)gr * p.ld16 + n, v, p.out16_plane);
          }
          return;
        }
        if (n >= p.N_total) return;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = *(const float4*)(p.bias + n);
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ln) s4 = *(const float4*)(p.ln_colsum + n);
        long long gr[8];
        float4 v[8], e1[8], e2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gr[k] = -1;
          if (k >= iters) continue;
          const int rr = rr0 + k * rstep;
          gr[k] = rtab[rr];
          v[k] = *(const float4*)(st + rr * 32 + ((q ^ (rr & 7)) << 2));
          e1[k] = make_float4(0.f, 0.f, 0.f, 0.f); e2[k] = e1[k];
          if (gr[k] >= 0) {
            if (p.rowvec) e1[k] = *(const float4*)(p.rowvec + (size_t)gtab[rr] * p.ld_rowvec + n);
            if (p.res32) e2[k] = *(const float4*)(p.res32 + (size_t)gr[k] * p.ldres + n);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k >= iters || gr[k] < 0) continue;
          float4 o = v[k];
          if (ln) {
            const int rr = rr0 + k * rstep;
            const float mu = lnmu_t[rr], rsd = lnrs_t[rr];
            o.x = rsd * fmaf(-mu, s4.x, o.x); o.y = rsd * fmaf(-mu, s4.y, o.y); o.z = rsd * fmaf(-mu, s4.z, o.z); o.w = rsd * fmaf(-mu, s4.w, o.w);
          }
          o.x += b4.x + e1[k].x + e2[k].x; o.y += b4.y + e1[k].y + e2[k].y; o.z += b4.z + e1[k].z + e2[k].z; o.w += b4.w + e1[k].w + e2[k].w;
          if (p.out32) *(float4*)(p.out32 + (size_t)gr[k] * p.ld32 + n) = o;
          if (p.out16) store_h4(p.out16 + (size_t)gr[k] * p.ld16 + n, o, p.out16_plane);
        }
      };

      if (tma_epi) {
        named_bar_sync(gbar, 128);   // bias_g / tables of this tile are complete before any thread of the group reads them
        // ---- TMA-store epilogue: extras are applied in registers on the thread's own row, the finished chunk is staged in
        //      the swizzle-128B layout and one elected thread hands it to the TMA unit (clipping = tile raggedness) ----
        if (grp >= nch) release_acc();   // this group has no chunk in such a narrow tile
        for (int c = grp; c < nch; c += 2) {
          const int j0 = c * cw;
          const int n0 = nt * p.block_n + j0;
          uint32_t pk[32];   // the 128 staged bytes of this thread's row
          if (!f16out) {
            // the per-group row vector (timestep embedding) is fetched before the TMEM load so both latencies overlap
            float4 e[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              e[q] = (rv && n0 + 4 * q + 3 < p.N_total) ? *(const float4*)(rv + n0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t v[32];
            tmem_ld32(t_acc + (uint32_t)j0, v);
            tmem_ld_wait();
            if (stamp && c == 0) TS(12);
            float f[32];
            {
              // bias (and column sums) of the chunk's 32 columns as 16-byte shared-memory loads (j0 is a multiple of 32)
              const float4* b4p = (const float4*)(bias_g + j0);
              const float4* s4p = (const float4*)(lns_g + j0);
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                const float4 bb = b4p[i4];
                const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
                if (ln) {
                  const float4 ss = s4p[i4];
                  const float sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    f[4 * i4 + e] = fmaf(ln_rs, fmaf(-ln_mu, sv[e], __uint_as_float(v[4 * i4 + e]) * p.out_scale), bv[e]);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) f[4 * i4 + e] = fmaf(__uint_as_float(v[4 * i4 + e]), p.out_scale, bv[e]);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { f[4 * q] += e[q].x; f[4 * q + 1] += e[q].y; f[4 * q + 2] += e[q].z; f[4 * q + 3] += e[q].w; }
            if (p.res_tma) {
              mbar_wait(rbar, res_phase);
              res_phase ^= 1;
              const float* rb = rbuf + r * 32;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 x = *(const float4*)(rb + ((q ^ (r & 7)) << 2));
                f[4 * q] += x.x; f[4 * q + 1] += x.y; f[4 * q + 2] += x.z; f[4 * q + 3] += x.w;
              }
            } else if (p.res_bulk) {
              if (!bulk_waited) { mbar_wait(rbar, 0); bulk_waited = true; }
              const float* rb = (const float*)(smem + p.off_resb + (uint32_t)c * 16384u) + r * 32;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 x = *(const float4*)(rb + ((q ^ (r & 7)) << 2));
                f[4 * q] += x.x; f[4 * q + 1] += x.y; f[4 * q + 2] += x.z; f[4 * q + 3] += x.w;
              }
            } else if (rs) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (valid && n0 + 4 * q + 3 < p.N_total) {
                  const float4 x = *(const float4*)(rs + n0 + 4 * q);
                  f[4 * q] += x.x; f[4 * q + 1] += x.y; f[4 * q + 2] += x.z; f[4 * q + 3] += x.w;
                }
              }
            }
            if (p.rowstats) {
              // columns past N_total carry exact zeros (zero-filled operands, bias, row vector and residual): no mask needed
#pragma unroll
              for (int i = 0; i < 32; ++i) { rs_sum += f[i]; rs_sq = fmaf(f[i], f[i], rs_sq); }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = __float_as_uint(f[i]);
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld32(t_acc + (uint32_t)(j0 + 32 * h), v);
              tmem_ld_wait();
              const float4* b4p = (const float4*)(bias_g + j0 + 32 * h);
              const float4* s4p = (const float4*)(lns_g + j0 + 32 * h);
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                const float4 bb = b4p[i4];
                const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
                float o[4];
                if (ln) {
                  const float4 ss = s4p[i4];
                  const float sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[e] = fmaf(ln_rs, fmaf(-ln_mu, sv[e], __uint_as_float(v[4 * i4 + e]) * p.out_scale), bv[e]);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[e] = fmaf(__uint_as_float(v[4 * i4 + e]), p.out_scale, bv[e]);
                }
                __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
                pk[16 * h + 2 * i4] = *(uint32_t*)&h0;
                pk[16 * h + 2 * i4 + 1] = *(uint32_t*)&h1;
              }
            }
          }
          if (stamp && c == 0) TS(13);
          if (c + 2 >= nch) release_acc();            // last chunk of this group: its TMEM reads are done
          if (r == 0) tma_store_wait_read<0>();       // the group's previous bulk store has finished reading `st`
          named_bar_sync(gbar, 128);                  // ... and every thread has finished reading `rbuf`
          if (stamp && c == 0) TS(14);
          if (p.res_tma && r == 0 && c + 2 < nch) {
            mbar_arrive_expect_tx(rbar, p.a_bytes);
            tma_load_4d(rbuf, &tmR, rbar, n0 + 64, oc1, oc2, oc3);
          }
          {
            uint8_t* rowp = (uint8_t*)st + r * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *(uint4*)(rowp + ((q ^ (r & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          uint8_t* const hst = smem + p.off_hst + (uint32_t)grp * 16384u;
          if (!f16out && p.h_tma) {
            // fp16 copy of this thread's 32 finished columns as [plane][row][32 halves] (64-byte rows, no swizzle): hi, then the
            // error-compensation plane lo = fp16(x - hi)
            uint32_t hh[16], ll[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float a0 = __uint_as_float(pk[i]), a1 = __uint_as_float(pk[i + 1]);
              __half2 h2 = __floats2half2_rn(a0, a1);
              hh[i >> 1] = *(uint32_t*)&h2;
              const float2 fh = __half22float2(h2);
              __half2 l2 = __floats2half2_rn(a0 - fh.x, a1 - fh.y);
              ll[i >> 1] = *(uint32_t*)&l2;
            }
            uint4* hp = (uint4*)(hst + r * 64);
#pragma unroll
            for (int q = 0; q < 4; ++q) hp[q] = make_uint4(hh[4 * q], hh[4 * q + 1], hh[4 * q + 2], hh[4 * q + 3]);
            if (p.h_planes == 2) {
              uint4* lp = (uint4*)(hst + 8192 + r * 64);
#pragma unroll
              for (int q = 0; q < 4; ++q) lp[q] = make_uint4(ll[4 * q], ll[4 * q + 1], ll[4 * q + 2], ll[4 * q + 3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(gbar, 128);
          if (r == 0) {
            tma_store_4d(&tmC, st, n0, oc1, oc2, oc3);
            if (!f16out && p.h_tma) tma_store_4d(&tmH, hst, n0, oc1, 0, oc3);
            tma_store_commit();
          }
          if (stamp && c == 0) TS(15);
          if (gn && !f16out) {
            // GroupNorm moments of this chunk: warp = one 32-row block of the staged fp32 values, lane = column; a lane sums its column
            // over the block's rows of one image (fixed order), gn_commit reduces the columns of a group and adds the total
            const int n = n0 + lane;
            const bool cok = n < p.N_total;
            const int sw = lane >> 2, el = lane & 3;
            float sv = 0.f, qv = 0.f;
            int cur = -1;
            for (int rr = quad * 32; rr < quad * 32 + 32; ++rr) {     // rtab / gtab: the same for every lane -> uniform control flow
              if (rtab[rr] < 0) continue;
              const int img = gtab[rr];
              if (img != cur) {
                if (cur >= 0) gn_commit(cur, cok ? n : -1, sv, qv);
                cur = img; sv = 0.f; qv = 0.f;
              }
              if (cok) {
                const float v = st[rr * 32 + (((sw ^ (rr & 7)) << 2) | el)];
                sv += v; qv = fmaf(v, v, qv);
              }
            }
            if (cur >= 0) gn_commit(cur, cok ? n : -1, sv, qv);
          }
        }
        if (p.rowstats) {
          // the two groups drained alternate chunks of the same rows: group 1 hands its partial over, group 0 adds in a fixed order
          float2* const xs = (float2*)rsx + ((tile_ctr & 1) * 128);
          if (grp == 1) xs[r] = make_float2(rs_sum, rs_sq);
          named_bar_sync(3, 256);
          if (grp == 0 && valid) {
            const float2 o = xs[r];
            *(float2*)(p.rowstats + ((size_t)grow * p.num_n_tiles + nt) * 2) = make_float2(rs_sum + o.x, rs_sq + o.y);
          }
          ++tile_ctr;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        if (stamp) TS(9);
        continue;
      }
      if (p.flags & GEMM_GEGLU) {
        // tile columns are [x (half) | gate (half)]; out16[:, nt*half + j] = (x + bx) * gelu(gate + bg)
        const int half_n = p.block_n >> 1;
        bool first = true;
        // this tile's bias (and, with a folded LayerNorm, column sums) staged once in shared memory: the per-element global loads
        // of the inner loop made this epilogue -- which already bounds the GEGLU projections -- a third slower
        named_bar_sync(gbar, 128);
        const float ln_ms = ln_mu * ln_rs;         // D = rstd acc - (mean rstd) colsum + bias; rstd = 1, mean = 0 without a folded LayerNorm
        for (int j0 = grp * 32; j0 < half_n; j0 += 64) {
          const int ncols = min(32, half_n - j0);
          float f[32];
          for (int h0 = 0; h0 < ncols; h0 += 16) {
            uint32_t xr[16], gr_[16];
            tmem_ld16(t_acc + (uint32_t)(j0 + h0), xr);
            tmem_ld16(t_acc + (uint32_t)(half_n + j0 + h0), gr_);
            tmem_ld_wait();
            const int cx = j0 + h0, cg = cx + half_n;
            // per-column constants as 16-byte shared-memory loads (cx, cg are multiples of 16 columns): one LDS.128 per 4 elements and
            // array instead of 4 scalar broadcasts per element
            const float4* bx4 = (const float4*)(bias_g + cx);
            const float4* bg4 = (const float4*)(bias_g + cg);
            const float4* sx4 = (const float4*)(lns_g + cx);
            const float4* sg4 = (const float4*)(lns_g + cg);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bx = bx4[i4], bg = bg4[i4];
              float cxv[4] = {bx.x, bx.y, bx.z, bx.w}, cgv[4] = {bg.x, bg.y, bg.z, bg.w};
              if (ln) {
                const float4 sx = sx4[i4], sg = sg4[i4];
                cxv[0] = fmaf(-ln_ms, sx.x, cxv[0]); cxv[1] = fmaf(-ln_ms, sx.y, cxv[1]); cxv[2] = fmaf(-ln_ms, sx.z, cxv[2]); cxv[3] = fmaf(-ln_ms, sx.w, cxv[3]);
                cgv[0] = fmaf(-ln_ms, sg.x, cgv[0]); cgv[1] = fmaf(-ln_ms, sg.y, cgv[1]); cgv[2] = fmaf(-ln_ms, sg.z, cgv[2]); cgv[3] = fmaf(-ln_ms, sg.w, cgv[3]);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * i4 + e;
                // rstd = 1 and (mean rstd) = 0 without a folded LayerNorm: the same two FMAs serve both forms
                const float xv = fmaf(ln_rs, __uint_as_float(xr[i]), cxv[e]);
                const float gv = fmaf(ln_rs, __uint_as_float(gr_[i]), cgv[e]);
                f[h0 + i] = xv * gelu_erf_f(gv);
              }
            }
          }
          if (stamp && j0 == 0) TS(12);
          if (p.geglu_tma && ncols == 32) {
            // fp16 [hi | lo] planes of this thread's 32 finished columns -> staging [plane][row][32 halves] -> one TMA store per chunk
            // (the flat coalesced pass cost ~1 us per chunk: as much as two thirds of the GELU arithmetic)
            uint32_t hh[16], ll[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              __half2 h2 = __floats2half2_rn(f[i], f[i + 1]);
              hh[i >> 1] = *(uint32_t*)&h2;
              const float2 fh = __half22float2(h2);
              __half2 l2 = __floats2half2_rn(f[i] - fh.x, f[i + 1] - fh.y);
              ll[i >> 1] = *(uint32_t*)&l2;
            }
            if (r == 0) tma_store_wait_read<0>();      // the group's previous bulk store has finished reading the staging buffer
            named_bar_sync(gbar, 128);
            uint8_t* const hst = (uint8_t*)st;
            uint4* hp = (uint4*)(hst + r * 64);
#pragma unroll
            for (int q = 0; q < 4; ++q) hp[q] = make_uint4(hh[4 * q], hh[4 * q + 1], hh[4 * q + 2], hh[4 * q + 3]);
            if (p.h_planes == 2) {
              uint4* lp = (uint4*)(hst + 8192 + r * 64);
#pragma unroll
              for (int q = 0; q < 4; ++q) lp[q] = make_uint4(ll[4 * q], ll[4 * q + 1], ll[4 * q + 2], ll[4 * q + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(gbar, 128);
            if (r == 0) {
              tma_store_4d(&tmH, hst, nt * half_n + j0, mt * 128, 0, bidx);
              tma_store_commit();
            }
            first = true;     // the staging buffer is guarded by the bulk-store wait above, not by the flat pass's barrier
            if (stamp && j0 == 0) TS(14);
            continue;
          }
          if (p.geglu_tma && r == 0) tma_store_wait_read<0>();
          if (!first || p.geglu_tma) named_bar_sync(gbar, 128);   // the previous chunk's flat pass / bulk store has left the staging buffer
          first = false;
          stage_row(f, ncols);
          named_bar_sync(gbar, 128);
          if (stamp && j0 == 0) TS(13);
          flush_chunk(nt * half_n + j0, ncols, 2);
          if (stamp && j0 == 0) TS(14);
        }
        if (stamp) TS(15);
      } else if (chw) {
        for (int j0 = grp * 16; j0 < p.block_n; j0 += 32) {
          uint32_t v[16];
          tmem_ld16(t_acc + (uint32_t)j0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = nt * p.block_n + j0 + i;
              if (n < p.N_total) {
                float x = __uint_as_float(v[i]) * p.out_scale;
                if (bs) x += bs[n];
                if (rv) x += rv[n];
                if (rs) x += rs[n];
                const size_t o = ((size_t)group * p.N_total + n) * (size_t)p.ldT + rig;
                if (p.out32) p.out32[o] = x;
                if (p.out16) p.out16[o] = __float2half_rn(x);
              }
            }
          }
        }
      } else {
        bool first = true;
        for (int j0 = grp * 32; j0 < p.block_n; j0 += 64) {
          const int ncols = min(32, p.block_n - j0);
          float f[32];
          if (ncols == 32) {
            uint32_t v[32];
            tmem_ld32(t_acc + (uint32_t)j0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
          } else {
            uint32_t v[16];
            tmem_ld16(t_acc + (uint32_t)j0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
          }
          if (p.cluster_reduce) {
            // split-K partial of this chunk -> this CTA's smem partial tile P[chunk][128 rows][32 fp32] (swizzled like the staging
            // chunks), laid over the operand pipeline slots: the main loop of this CTA's only tile has drained them
            float* rowp = (float*)smem + (j0 >> 5) * 4096 + r * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q * 4 < ncols) *(float4*)(rowp + ((q ^ (r & 7)) << 2)) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            continue;
          }
          if (!first) named_bar_sync(gbar, 128);   // the previous chunk's flat pass has left the staging buffer
          first = false;
          stage_row(f, ncols);
          named_bar_sync(gbar, 128);
          flush_chunk(nt * p.block_n + j0, ncols, splitk ? 1 : 0);
        }
      }
      // the accumulator stage can be recycled now
      release_acc();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (stamp) TS(9);

      if (splitk && !p.cluster_reduce) {
        // ---- deterministic split-K reduction through the global workspace (fallback when the splits are not one cluster) ----
        const int tile_id = bidx * tiles_mn + nt * p.num_m_tiles + mt;
        int* ctr = p.counters + 2 * tile_id;
        const int n4 = p.block_n >> 2;
        const size_t stride4 = ws_split_stride >> 2;
        // Sums the elements idx in [lo, hi) of this tile (idx = row * n4 + float4 column) over all splits IN SPLIT ORDER, applies the
        // extras and stores.  Several elements per thread are in flight at once and the split loop is unrolled, so the L2 round
        // trips of the partial-tile loads overlap instead of serialising.
        auto reduce_range = [&](int lo, int hi) {
          constexpr int U = 2;
          for (int base = lo + tid2; base < hi; base += 256 * U) {
            const float4* src[U];
            long long gr[U];
            int nn[U], gg[U];
            float4 a[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int idx = base + u * 256;
              gr[u] = -1; nn[u] = 0; gg[u] = 0; src[u] = nullptr;
              a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (idx < hi) {
                const int rr = idx / n4;
                const int n = nt * p.block_n + ((idx - rr * n4) << 2);
                const long long g = rtab[rr];
                if (g >= 0 && n < p.N_total) {
                  gr[u] = g; nn[u] = n; gg[u] = gtab[rr];
                  src[u] = (const float4*)(p.ws + (size_t)g * p.ws_ld + n);
                }
              }
            }
            float4 e0[U], e1[U], e2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              e0[u] = e1[u] = e2[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (gr[u] < 0) continue;
              if (p.bias) e0[u] = *(const float4*)(p.bias + nn[u]);
              if (p.rowvec) e1[u] = *(const float4*)(p.rowvec + (size_t)gg[u] * p.ld_rowvec + nn[u]);
              if (p.res32) e2[u] = *(const float4*)(p.res32 + (size_t)gr[u] * p.ldres + nn[u]);
            }
#pragma unroll 4
            for (int sp = 0; sp < p.num_splits; ++sp) {
              float4 t[U];
#pragma unroll
              for (int u = 0; u < U; ++u) t[u] = gr[u] >= 0 ? __ldcg(src[u] + (size_t)sp * stride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int u = 0; u < U; ++u) { a[u].x += t[u].x; a[u].y += t[u].y; a[u].z += t[u].z; a[u].w += t[u].w; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (gr[u] < 0) continue;
              float4 o = a[u];
              o.x += e0[u].x; o.y += e0[u].y; o.z += e0[u].z; o.w += e0[u].w;
              o.x += e1[u].x; o.y += e1[u].y; o.z += e1[u].z; o.w += e1[u].w;
              o.x += e2[u].x; o.y += e2[u].y; o.z += e2[u].z; o.w += e2[u].w;
              if (p.out32) *(float4*)(p.out32 + (size_t)gr[u] * p.ld32 + nn[u]) = o;
              if (p.out16) store_h4(p.out16 + (size_t)gr[u] * p.ld16 + nn[u], o, p.out16_plane);
            }
          }
        };
        __threadfence();
        named_bar_sync(3, 256);
        if (stamp) TS(12);
        if (p.coop_reduce) {
          // every split of this tile is resident (cooperative launch, one tile per CTA): all of them wait for the last
          // partial and then each reduces its own 1/num_splits slice of the tile -> the reduction is parallel as well
          if (tid2 == 0) {
            atomicAdd(ctr, 1);
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
              if (seen < p.num_splits) __nanosleep(40);
            } while (seen < p.num_splits);
          }
          named_bar_sync(3, 256);
          if (stamp) TS(13);
          __threadfence();
          const int total = 128 * n4;
          const int per = (total + p.num_splits - 1) / p.num_splits;
          const int lo = split * per, hi = min(lo + per, total);
          reduce_range(lo, hi);
          if (stamp) TS(14);
          named_bar_sync(3, 256);
          if (stamp) TS(15);
          if (tid2 == 0) {
            const int d = atomicAdd(ctr + 1, 1);
            if (d == p.num_splits - 1) { ctr[0] = 0; ctr[1] = 0; }   // last reader re-arms the counters for the next launch
          }
        } else {
          // fallback (several tiles per CTA): the last split to arrive reduces the whole tile
          if (tid2 == 0) {
            const int prev = atomicAdd(ctr, 1);
            const int last = prev == p.num_splits - 1;
            if (last) *ctr = 0;
            *split_flag = (uint32_t)last;
          }
          named_bar_sync(3, 256);
          const bool is_last = *split_flag != 0;
          named_bar_sync(3, 256);   // everyone has read the flag before a later tile may overwrite it
          if (is_last) {
            __threadfence();
            reduce_range(0, 128 * n4);
          }
        }
      }
    }
  }

  if (p.cluster_reduce) {
    // ---- deterministic split-K reduction over distributed shared memory ----
    // The num_splits CTAs of this cluster hold the fp32 partials of ONE output tile in their own shared memory.  After a cluster
    // barrier, CTA `rank` sums slice `rank` of the tile over all CTAs in rank order (bit-reproducible) with ld.shared::cluster,
    // applies the extras and stores: the partials never travel through L2 (measured: the global-workspace round trip cost 5-10 us
    // per launch at ~2.5 TB/s aggregate, more than the main loop of the 4x4 / 8x8 levels).
    cluster_sync_all();
    if (threadIdx.x == 64) TS(12);
    if (warp >= 2) {
      const int grp = (warp - 2) >> 2;
      const int r = (warp & 3) * 32 + lane;
      const int tid2 = grp * 128 + r;
      const long long* rtab = row_tab + grp * 128;
      const int* gtab = grp_tab + grp * 128;
      const int split = f_split, nt = f_nt;   // cluster mode: this CTA's only tile
      const uint32_t p_base = smem_u32(smem);
      // CTA `split` owns rows [row_lo, row_hi) of the tile.  8 consecutive threads cover one row's 128 bytes of a 32-column chunk
      // (coalesced 128-byte stores), 32 rows per pass; all index math is shifts and adds (an earlier flat-index version spent
      // ~0.3 us per element in integer divisions and predicated address arithmetic at two warps per scheduler).
      auto reduce_slice = [&](auto s_c) {
        constexpr int SS = decltype(s_c)::value;   // cluster size as a constant: the SS partial loads of an element are all in flight
        const int rows_per = (128 + SS - 1) / SS;
        const int row_lo = split * rows_per;
        const int row_hi = min(row_lo + rows_per, 128);
        const int q = tid2 & 7;
        const int nchunks = (p.block_n + 31) >> 5;
        uint32_t rbase[SS];
#pragma unroll
        for (int j = 0; j < SS; ++j) rbase[j] = mapa_shared(p_base, (uint32_t)j);
        const bool ln = p.ln_stats != nullptr;
        const float* lnmu_t = ln_tab + grp * 256;
        const float* lnrs_t = lnmu_t + 128;
        float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};   // producer row stats: <= 4 rows of this slice per thread
        // UC column chunks of a row are processed together: all their partial / row-vector / residual loads are issued before the
        // first use, so one DSMEM + L2 round trip (~450 cycles under load) is paid per UC elements instead of per element
        constexpr int UC = SS <= 3 ? 4 : 2;
        for (int c0 = 0; c0 < nchunks; c0 += UC) {
          int n[UC];
          bool ok[UC];
          float4 b4[UC], s4[UC];
#pragma unroll
          for (int u = 0; u < UC; ++u) {
            const int ncol = (c0 + u) * 32 + q * 4;
            n[u] = nt * p.block_n + ncol;
            ok[u] = ncol < p.block_n && n[u] < p.N_total;
            b4[u] = (ok[u] && p.bias) ? *(const float4*)(p.bias + n[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
            s4[u] = (ok[u] && ln) ? *(const float4*)(p.ln_colsum + n[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          int it = 0;
          for (int rr = row_lo + (tid2 >> 3); rr < row_hi; rr += 32, ++it) {
            const long long g = rtab[rr];
            float4 go[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) go[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < 0 && !gn) continue;
            if (g >= 0) {
            const uint32_t off0 = (uint32_t)(c0 * 16384 + rr * 128 + ((q ^ (rr & 7)) << 4));
            float4 t[UC][SS], e1[UC], e2[UC];
            const float* rvp = p.rowvec ? p.rowvec + (size_t)gtab[rr] * p.ld_rowvec : nullptr;
            const float* rsp = p.res32 ? p.res32 + (size_t)g * p.ldres : nullptr;
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const uint32_t off = ok[u] ? off0 + (uint32_t)u * 16384u : off0;   // masked chunks re-read a valid address
#pragma unroll
              for (int j = 0; j < SS; ++j) {
                // own partial through the local port (DSMEM sustains only ~10-20 B/clk per SM)
                if (j == split) t[u][j] = *(const float4*)(smem + off);
                else t[u][j] = ld_cluster_f4(rbase[j] + off);
              }
              e1[u] = (ok[u] && rvp) ? *(const float4*)(rvp + n[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
              e2[u] = (ok[u] && rsp) ? *(const float4*)(rsp + n[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              if (!ok[u]) continue;
              float4 o = t[u][0];   // partials are summed in rank order whoever owns the row => bit-reproducible, batch-order independent
#pragma unroll
              for (int j = 1; j < SS; ++j) { o.x += t[u][j].x; o.y += t[u][j].y; o.z += t[u][j].z; o.w += t[u][j].w; }
              if (ln) {
                const float mu = lnmu_t[rr], rsd = lnrs_t[rr];
                o.x = rsd * fmaf(-mu, s4[u].x, o.x); o.y = rsd * fmaf(-mu, s4[u].y, o.y);
                o.z = rsd * fmaf(-mu, s4[u].z, o.z); o.w = rsd * fmaf(-mu, s4[u].w, o.w);
              }
              o.x += b4[u].x + e1[u].x + e2[u].x; o.y += b4[u].y + e1[u].y + e2[u].y;
              o.z += b4[u].z + e1[u].z + e2[u].z; o.w += b4[u].w + e1[u].w + e2[u].w;
              if (p.out32) *(float4*)(p.out32 + (size_t)g * p.ld32 + n[u]) = o;
              if (p.out16) store_h4(p.out16 + (size_t)g * p.ld16 + n[u], o, p.out16_plane);
              if (p.rowstats && it < 4) {
                st_s[it] += (o.x + o.y) + (o.z + o.w);
                st_q[it] = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, st_q[it]))));
              }
              go[u] = o;
            }
            }
            if (gn) {
              // GroupNorm moments: the 4 rows a warp handles per iteration (4 row lanes x 8 column quads) belong to one image. Butterfly
              // over the row lanes (every lane ends with its column quad's totals), lane (j, q) keeps column 4q + j, one shuffle
              // brings column c to lane c, gn_commit reduces the columns of a group and adds the total.
              const int img = __shfl_sync(0xffffffffu, g >= 0 ? gtab[rr] : -1, 0);      // row lane 0's row: valid if any of the 4 is
#pragma unroll
              for (int u = 0; u < UC; ++u) {
                float sv[4] = {go[u].x, go[u].y, go[u].z, go[u].w}, qv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  qv[k] = sv[k] * sv[k];
                  sv[k] += __shfl_xor_sync(0xffffffffu, sv[k], 8); qv[k] += __shfl_xor_sync(0xffffffffu, qv[k], 8);
                  sv[k] += __shfl_xor_sync(0xffffffffu, sv[k], 16); qv[k] += __shfl_xor_sync(0xffffffffu, qv[k], 16);
                }
                const int j = lane >> 3;
                const float ms = j == 0 ? sv[0] : (j == 1 ? sv[1] : (j == 2 ? sv[2] : sv[3]));
                const float mq = j == 0 ? qv[0] : (j == 1 ? qv[1] : (j == 2 ? qv[2] : qv[3]));
                const int src = ((lane & 3) << 3) | (lane >> 2);          // lane holding column `lane` of this chunk
                const float cs = __shfl_sync(0xffffffffu, ms, src), cq = __shfl_sync(0xffffffffu, mq, src);
                const int ncol = nt * p.block_n + (c0 + u) * 32 + lane;
                const bool cok = img >= 0 && (c0 + u) * 32 + lane < p.block_n && ncol < p.N_total;
                gn_commit(img < 0 ? 0 : img, cok ? ncol : -1, cs, cq);
              }
            }
          }
        }
        if (p.rowstats) {
          // a row's 8 column quads of every chunk sit in 8 consecutive lanes: butterfly over them (fixed order), lane q == 0 stores
          int it = 0;
          for (int rr = row_lo + (tid2 >> 3); rr < row_lo + ((rows_per + 31) & ~31) && it < 4; rr += 32, ++it) {
            float a = st_s[it], b = st_q[it];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (q == 0 && rr < row_hi) {
              const long long g = rtab[rr];
              if (g >= 0) *(float2*)(p.rowstats + ((size_t)g * p.num_n_tiles + nt) * 2) = make_float2(a, b);
            }
          }
        }
      };
      switch (p.num_splits) {
        case 2: reduce_slice(std::integral_constant<int, 2>{}); break;
        case 3: reduce_slice(std::integral_constant<int, 3>{}); break;
        case 4: reduce_slice(std::integral_constant<int, 4>{}); break;
        case 5: reduce_slice(std::integral_constant<int, 5>{}); break;
        case 6: reduce_slice(std::integral_constant<int, 6>{}); break;
        case 7: reduce_slice(std::integral_constant<int, 7>{}); break;
        default: reduce_slice(std::integral_constant<int, 8>{}); break;
      }
      if (tid2 == 0) TS(14);
    }
    cluster_sync_all();   // no CTA may exit (and release its shared memory) while a sibling still reads its partial tile
  }

  if (warp >= 2 && ((warp & 3) * 32 + lane) == 0) tma_store_wait_read<0>();   // smem must outlive the bulk stores' reads
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TS(10);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  if (threadIdx.x == 32) TS(11);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// Per-device state (a process may drive several GPUs: model.to('cuda:1'), one engine per device). Function attributes (the
// > 48 KB dynamic shared memory opt-in) are per device, and the split-K workspace / arrival counters must live on the device
// that runs the kernel. Two workspace slots per device: slot 1 serves launches on the library's auxiliary stream (the
// parallel branch of a forked program, runtime.cu), so that concurrent GEMMs never share partials or counters.
static constexpr int kMaxCounters = 1 << 16;
static constexpr int kMaxDevices = 64;
struct GemmDev {
  bool ready = false;
  int num_sms = 0, smem_optin = 0;
  int max_clusters[9] = {0};         // co-resident clusters of size S (S CTAs must share a GPC), from the occupancy API
  float* ws[kStreamSlots] = {};      // split-K partial-tile workspaces per stream slot (allocated on a slot's first use, then fixed:
                                     // graph-stable addresses; the engines run every program once eagerly before they capture it)
  size_t ws_bytes = 0;
  int* counters[kStreamSlots] = {};
};
static GemmDev g_gdev[kMaxDevices];
static std::mutex g_gemm_mu;
static long long* g_debug_ts = nullptr;

static int gemm_device_setup(GemmDev** out) {
  int dev = 0;
  UPGPT_CHECK_CUDA(cudaGetDevice(&dev));
  UPGPT_REQUIRE(dev >= 0 && dev < kMaxDevices, "upgpt_gemm: device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_gemm_mu);
  GemmDev& d = g_gdev[dev];
  *out = &d;
  if (d.ready) return 0;
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
  UPGPT_CHECK_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  UPGPT_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin));
  for (int S = 1; S <= 8; ++S) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S * d.num_sms); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = d.smem_optin;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_kernel, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = d.num_sms / (2 * S); }
    d.max_clusters[S] = n;
  }
  d.ws_bytes = (size_t)96 << 20;
  d.ready = true;
  return 0;
}

// split-K workspace + arrival counters of one stream slot (global-workspace fallback of the split-K reduction)
static int gemm_slot_workspace(GemmDev* d, int slot) {
  std::lock_guard<std::mutex> lk(g_gemm_mu);
  if (d->ws[slot]) return 0;
  UPGPT_CHECK_CUDA(cudaMalloc(&d->ws[slot], d->ws_bytes));
  UPGPT_CHECK_CUDA(cudaMalloc(&d->counters[slot], kMaxCounters * sizeof(int)));
  UPGPT_CHECK_CUDA(cudaMemset(d->counters[slot], 0, kMaxCounters * sizeof(int)));
  return 0;
}

// Tile-width / split-K choice by a small cost model of THIS kernel (measured on B200 with the in-kernel %globaltimer
// timeline, tools/gpu_gemm_timeline.py): a CTA ingests operands through TMA at ~100 GB/s (one [128][64] fp16 A box plus one
// [bn][64] B box per k-iteration, ~0.3 us for 44 KB), pays ~2.5 us of prologue/teardown, ~0.6 us per 32-column epilogue chunk,
// and a split-K tile additionally writes + re-reads its fp32 partial tile and synchronises (~2.5 us + 0.9 us per chunk).
// Wide tiles minimise A re-reads; split-K supplies the parallelism that small-M layers lack.
struct TileChoice { int bn; int splits; };
static double g_sm_weight = getenv("UPGPT_GEMM_SM_WEIGHT") ? atof(getenv("UPGPT_GEMM_SM_WEIGHT")) : 0.0;
static TileChoice choose_tiling(const int* g_max_clusters, int N, int m_tiles_x_batch, int k_iters, int num_sms, int gran, bool must_divide,
                                bool allow_split, int chunk_cols, bool x3) {
  TileChoice best{gran, 1};
  double best_t = 1e30;
  auto round_up = [](int x, int m) { return (x + m - 1) / m * m; };
  const int n_cap = round_up(N, gran) < 256 ? round_up(N, gran) : 256 / gran * gran;
  for (int bn = n_cap; bn >= (n_cap < 64 ? n_cap : 64); bn -= gran) {
    if (must_divide && N % bn) continue;
    if (x3 && bn > 224 && n_cap > 224) continue;   // a 256-wide x3 tile leaves smem for one {hi, lo} slot pair only: no load / MMA overlap
    const int n_tiles = (N + bn - 1) / bn;
    const int base = m_tiles_x_batch * n_tiles;
    // per k-block: operand ingest at ~100 GB/s/SM vs the MMAs at ~10 TFLOP/s/SM; x3 = 2 plane pairs loaded, 3 products issued
    const double ingest_us = (x3 ? 2.0 : 1.0) * (16384.0 + 128.0 * bn) / 100e3;
    const double mma_us = (x3 ? 3.0 : 1.0) * (2.0 * 128.0 * bn * 64.0) / 10e6;
    const double us_per_iter = ingest_us > mma_us ? ingest_us : mma_us;
    const double epi = 0.6 * ((bn + chunk_cols - 1) / chunk_cols);
    // split-K = the CTAs of one cluster (<= 8): partial tiles stay in shared memory and are reduced over DSMEM
    // UPGPT_GEMM_MAX_SPLITS: cap of the split-K factor (experiments; a cap of 4 was the first form of the throughput mode: 49.0 -> 50.85
    // images/s with 3 lanes, profiles/r02_split_cap_under_lanes.txt; the SM-time weight above does the same job per shape)
    static const int split_cap = getenv("UPGPT_GEMM_MAX_SPLITS") ? atoi(getenv("UPGPT_GEMM_MAX_SPLITS")) : 8;
    int max_splits = allow_split ? (k_iters / 2 < 8 ? k_iters / 2 : 8) : 1;
    if (max_splits > split_cap) max_splits = split_cap < 1 ? 1 : split_cap;
    for (int sp = 1; sp <= (max_splits < 1 ? 1 : max_splits); ++sp) {
      const int ctas = base * sp;
      const int capacity = sp == 1 ? num_sms : g_max_clusters[sp] * sp;   // CTAs of size-sp clusters that fit at once
      if (sp > 1 && ctas > capacity) break;                               // keep split tiles to one wave
      const int waves = (ctas + capacity - 1) / capacity;
      const int iters = (k_iters + sp - 1) / sp;
      double t = waves * (2.5 + iters * us_per_iter + (sp == 1 ? epi : 0.3 * ((bn + 31) / 32)));
      // DSMEM reduction: every CTA pulls (sp-1)/sp of a [128][bn] fp32 tile from its siblings at ~35 GB/s, + 2 cluster barriers
      if (sp > 1) t += 1.2 + 512.0 * bn * (sp - 1) / sp / 35e3;
      // throughput mode (upgpt_gemm_set_sm_weight): SM time in the objective, cost = latency x (1 + w x CTAs / SMs). With one batch in
      // flight a small layer should spread over as many SMs as its latency gains from; with several batches in flight (lanes) every
      // CTA beyond the necessary ones takes an SM from another batch's kernel and pays its own prologue / drain / reduction.
      if (g_sm_weight > 0.0) t *= 1.0 + g_sm_weight * (double)(ctas < num_sms ? ctas : num_sms) / (double)num_sms;
      if (t < best_t - 1e-9) { best_t = t; best = {bn, sp}; }
    }
  }
  return best;
}

}  // namespace upgpt

using namespace upgpt;

static int gemm_run(const upgpt_gemm_args* a, cudaStream_t stream, int* plan_out) {
  const bool dry = plan_out != nullptr;      // upgpt_gemm_plan: same decisions, no tensor maps, no launch
  GemmDev* gd = nullptr;
  { const int rc = gemm_device_setup(&gd); if (rc) return rc; }
  const int g_num_sms = gd->num_sms, g_smem_optin = gd->smem_optin;
  const int* g_max_clusters = gd->max_clusters;
  const size_t g_ws_bytes = gd->ws_bytes;
  const int ws_slot = stream_slot(stream);
  UPGPT_REQUIRE(a && (dry || (a->a && a->w)), "upgpt_gemm: null operand");
  UPGPT_REQUIRE(a->out32 || a->out16, "upgpt_gemm: no output");
  UPGPT_REQUIRE(!((a->flags & UPGPT_GEMM_F_SPLIT3OUT) && (a->flags & UPGPT_GEMM_F_CHW)), "upgpt_gemm: SPLIT3OUT is not available with channel-major stores");
  UPGPT_REQUIRE(a->K > 0 && a->N > 0, "upgpt_gemm: bad K/N");
  UPGPT_REQUIRE(a->K % 8 == 0, "upgpt_gemm: K (=%d) must be a multiple of 8 (16-byte TMA rows)", a->K);
  const bool s2 = a->mode == UPGPT_GEMM_CONV3X3_S2PHASE || a->mode == UPGPT_GEMM_CONV3X3_S2PHASE_ASYM;
  const bool up2 = a->mode == UPGPT_GEMM_CONV3X3_UP2;
  const bool conv = a->mode == UPGPT_GEMM_CONV3X3 || s2 || a->mode == UPGPT_GEMM_CONV1X1 || up2;
  GemmParams p{};
  p.flags = a->flags;
  p.x3 = (a->flags & UPGPT_GEMM_F_X3) ? 1 : 0;
  const uint64_t n_planes = p.x3 ? 2 : 1;          // operand planes [hi | lo] = 5th tensor-map dimension, K elements apart
  const uint64_t ld_default = n_planes * (uint64_t)a->K;
  p.batch = up2 ? 4 : (conv ? 1 : (a->batch > 0 ? a->batch : 1));   // UP2: the "batch" index of a tile is its output parity (a, b)
  p.N_total = a->N;
  p.taps = (a->mode == UPGPT_GEMM_CONV3X3 || s2) ? 9 : (up2 ? 4 : 1);
  p.kblocks_per_tap = (a->K + 63) / 64;
  p.out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;

  CUtensorMap tmA, tmB;
  int m_rows_total;
  if (conv) {
    UPGPT_REQUIRE(a->H > 0 && a->W > 0 && a->n_imgs > 0, "upgpt_gemm(conv): bad geometry");
      p.flags |= GEMM_CONV;
    p.H = a->H; p.W = a->W; p.HW = a->H * a->W; p.n_imgs = a->n_imgs;
    uint32_t box[5];
    box[4] = 1;
    p.tiles_per_row = 1;
    p.tile_cols = a->W;
    if (p.HW <= 64) {
      p.tile_imgs = 128 / p.HW; p.tile_rows = a->H; p.tiles_per_img = 1;
      if (p.tile_imgs > a->n_imgs) p.tile_imgs = a->n_imgs;
      if (p.tile_imgs < 1) p.tile_imgs = 1;
      p.num_m_tiles = (a->n_imgs + p.tile_imgs - 1) / p.tile_imgs;
      box[0] = 64; box[1] = a->W; box[2] = a->H; box[3] = p.tile_imgs;
    } else {
      // tile_cols = the divisor of W (<= 128) that fills most of the 128 accumulator lanes
      int best_tc = 1, best_fill = 0;
      for (int tc = 1; tc <= 128 && tc <= a->W; ++tc) {
        if (a->W % tc) continue;
        int tr = 128 / tc; if (tr > a->H) tr = a->H;
        if (tc * tr >= best_fill) { best_fill = tc * tr; best_tc = tc; }
      }
      p.tile_imgs = 1; p.tile_cols = best_tc;
      p.tile_rows = 128 / best_tc; if (p.tile_rows > a->H) p.tile_rows = a->H;
      p.tiles_per_row = a->W / best_tc;
      p.tiles_per_img = p.tiles_per_row * ((a->H + p.tile_rows - 1) / p.tile_rows);
      p.num_m_tiles = p.tiles_per_img * a->n_imgs;
      box[0] = 64; box[1] = p.tile_cols; box[2] = p.tile_rows; box[3] = 1;
    }
    p.a_bytes = box[0] * box[1] * box[2] * box[3] * 2;
    const int a_imgs = s2 ? 4 * a->n_imgs : a->n_imgs;
    uint64_t dims[5] = {(uint64_t)a->K, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a_imgs, n_planes};
    const uint64_t lda = a->lda > 0 ? a->lda : ld_default;
    uint64_t strides[4] = {lda * 2, lda * 2 * a->W, lda * 2 * a->W * a->H, (uint64_t)a->K * 2};
    if (!dry && make_tmap_f16(&tmA, a->a, 5, dims, strides, box, true)) return -3;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = p.tap_dx[t] = p.tap_dn[t] = 0; }
    if (a->mode == UPGPT_GEMM_CONV3X3) {
      for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3 - 1; p.tap_dx[t] = t % 3 - 1; }
    } else if (up2) {
      // out(2y + a, 2x + b) reads the upsampled rows 2y + a + {-1, 0, 1} = source rows y + {a - 1, a}: tap (u, v) of parity (a, b)
      // sits at source offset (a - 1 + u, b - 1 + v) and carries the sum of the 3x3 weights that land on that source pixel
      p.flags |= GEMM_UP2;
      for (int par = 0; par < 4; ++par)
        for (int t = 0; t < 4; ++t) { p.tap_dy[par * 4 + t] = (par >> 1) - 1 + (t >> 1); p.tap_dx[par * 4 + t] = (par & 1) - 1 + (t & 1); }
    } else if (a->mode == UPGPT_GEMM_CONV3X3_S2PHASE) {
      // out(y,x) tap (r,s) reads in(2y+r-1, 2x+s-1) = phase[(r+1)&1][(s+1)&1] at (y + (r==0 ? -1 : 0), x + (s==0 ? -1 : 0))
      for (int t = 0; t < 9; ++t) {
        const int rr = t / 3, ss = t % 3;
        const int ph = ((rr + 1) & 1) * 2 + ((ss + 1) & 1);
        p.tap_dy[t] = rr == 0 ? -1 : 0; p.tap_dx[t] = ss == 0 ? -1 : 0; p.tap_dn[t] = ph * a->n_imgs;
      }
    } else if (a->mode == UPGPT_GEMM_CONV3X3_S2PHASE_ASYM) {
      // out(y,x) tap (r,s) reads in(2y+r, 2x+s) = phase[r&1][s&1] at (y + (r==2), x + (s==2)); the row/column past the end is TMA zero fill
      for (int t = 0; t < 9; ++t) {
        const int rr = t / 3, ss = t % 3;
        const int ph = (rr & 1) * 2 + (ss & 1);
        p.tap_dy[t] = rr == 2 ? 1 : 0; p.tap_dx[t] = ss == 2 ? 1 : 0; p.tap_dn[t] = ph * a->n_imgs;
      }
    }
    m_rows_total = a->n_imgs * p.HW;
    p.M_total = m_rows_total;
    p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : (up2 ? 4 * p.HW : p.HW);
  } else {
    UPGPT_REQUIRE(a->M > 0, "upgpt_gemm: bad M");
    p.M_total = a->M;
    p.num_m_tiles = (a->M + 127) / 128;
    p.a_bytes = kABytes;
    const uint64_t lda = a->lda > 0 ? a->lda : ld_default;
    const uint64_t abs_ = a->a_batch_stride > 0 ? (uint64_t)a->a_batch_stride : lda * (uint64_t)a->M;
    uint64_t dims[5] = {(uint64_t)a->K, (uint64_t)a->M, 1, (uint64_t)p.batch, n_planes};
    uint64_t strides[4] = {lda * 2, abs_ * 2, abs_ * 2, (uint64_t)a->K * 2};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    if (!dry && make_tmap_f16(&tmA, a->a, 5, dims, strides, box, true)) return -3;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = p.tap_dx[t] = p.tap_dn[t] = 0; }
    m_rows_total = a->M;
    p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : a->M;
  }

  // ---- tile shape / split-K ----
  // epilogue flavour: TMA-store of finished chunks when the layout allows it (row-major output, 16-byte aligned rows)
  const bool chw_out = (p.flags & GEMM_CHW) != 0;
  const bool geglu = (p.flags & GEMM_GEGLU) != 0;
  int epi_mode = 0;
  if (up2) UPGPT_REQUIRE(!chw_out && !geglu && !a->rowstats_out && !a->ln_stats && !a->gn_acc, "upgpt_gemm(UP2): plain row-major epilogue only");
  if (!chw_out && !geglu && !up2) {      // UP2 scatters a tile's rows over the 2H x 2W image: flat row-table stores, no box store
    const int ld32 = a->ld32 > 0 ? a->ld32 : a->N;
    const int ld16 = a->ld16 > 0 ? a->ld16 : a->N;
    if (a->out32 && ld32 % 4 == 0 && ((uintptr_t)a->out32 & 15) == 0) epi_mode = 1;
    else if (!a->out32 && a->out16 && ld16 % 8 == 0 && ((uintptr_t)a->out16 & 15) == 0 && !a->res32 && !a->rowvec &&
             !(a->flags & UPGPT_GEMM_F_SPLIT3OUT)) epi_mode = 2;
  }
  const int k_iters = p.taps * p.kblocks_per_tap;
  int bn = a->block_n;
  int splits = a->splits;
  if (bn > 0 && ((epi_mode == 1 && bn % 32) || (epi_mode == 2 && bn % 64))) epi_mode = 0;   // explicit tile width wins
  if (bn <= 0 || splits <= 0) {
    const int gran = epi_mode == 2 ? 64 : (epi_mode == 1 ? 32 : 16);
    const bool can_split = !chw_out && !geglu && a->N % 4 == 0 && splits <= 0;
    if (bn <= 0 && a->N <= 16) { bn = 16; }
    if (bn <= 0) {
      const TileChoice tc = choose_tiling(g_max_clusters, a->N, p.num_m_tiles * p.batch, k_iters, g_num_sms, gran, /*must_divide=*/epi_mode == 0,
                                          can_split, epi_mode == 2 ? 64 : 32, p.x3 != 0);
      bn = tc.bn;
      if (splits <= 0) splits = tc.splits;
    } else if (splits <= 0) {
      splits = 1;
      const int base = p.num_m_tiles * ((a->N + bn - 1) / bn) * p.batch;
      if (can_split && base * 2 <= g_num_sms && k_iters >= 8) {
        splits = g_num_sms / base;
        if (splits > k_iters / 2) splits = k_iters / 2;
        if (splits > 8) splits = 8;
        {
          static const int split_cap2 = getenv("UPGPT_GEMM_MAX_SPLITS") ? atoi(getenv("UPGPT_GEMM_MAX_SPLITS")) : 8;
          if (splits > split_cap2) splits = split_cap2 < 1 ? 1 : split_cap2;
        }
        while (splits > 1 && base * splits > g_max_clusters[splits] * splits) --splits;
        if (splits < 1) splits = 1;
      }
    }
  }
  UPGPT_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= 256, "upgpt_gemm: block_n=%d illegal", bn);
  if (p.flags & GEMM_GEGLU) UPGPT_REQUIRE(bn % 32 == 0 && a->N % bn == 0 && a->out16, "upgpt_gemm: GEGLU needs block_n%%32==0, N%%block_n==0, out16");
  p.block_n = bn;
  p.num_n_tiles = (a->N + bn - 1) / bn;
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
    const int n_ctr = 2 * p.num_m_tiles * p.num_n_tiles * p.batch;
    if (need > g_ws_bytes || n_ctr > kMaxCounters) {
      // shrink the split factor to fit the fixed workspace (its address must stay stable for captured graphs)
      while (splits > 1 && ((size_t)splits * p.ws_rows * p.ws_ld * sizeof(float) > g_ws_bytes || n_ctr > kMaxCounters)) --splits;
      while (splits > 1 && (splits - 1) * ((k_iters + splits - 1) / splits) >= k_iters) --splits;
      p.num_splits = splits;
    }
    if (!dry) {
      // (only the global-workspace fallback reads these; the cluster reduction keeps its partials in shared memory)
      const int rc = gemm_slot_workspace(gd, ws_slot);
      if (rc) return rc;
    }
    p.ws = gd->ws[ws_slot];
    p.counters = gd->counters[ws_slot];
  }
  {
    const uint64_t ldw = a->ldw > 0 ? a->ldw : ld_default;  // elements between taps
    const uint64_t n_stride = ldw * p.taps;            // elements between output channels
    const uint64_t wbs = a->w_batch_stride > 0 ? (uint64_t)a->w_batch_stride : n_stride * (uint64_t)a->N;
    uint64_t dims[5] = {(uint64_t)a->K, (uint64_t)p.taps, (uint64_t)a->N, (uint64_t)p.batch, n_planes};
    uint64_t strides[4] = {ldw * 2, n_stride * 2, wbs * 2, (uint64_t)a->K * 2};
    uint32_t box[5] = {64, 1, (uint32_t)bn, 1, 1};
    if (!dry && make_tmap_f16(&tmB, a->w, 5, dims, strides, box, true)) return -3;
  }

  p.debug_ts = g_debug_ts;
  {
    // UPGPT_GEMM_WPREFETCH: 0 = off (default), 1 = operand ring only, n > 1 = ring + L2 prefetch of up to n further k-blocks.
    // Measured on the bbox.yaml step at B = 8 (profiles/r02_probe_weight_prefetch.jsonl): 4.32 / 4.29 / 4.36 ms for 0 / 1 / 64 -- inside
    // the run-to-run noise: the chain is not bound by the weight stream (the CTA that starts last has no lead time to use). With
    // three batches in flight (lanes) "off" is 1 % faster (49.1 vs 48.5 images/s, profiles/r02_knobs_under_lanes.txt): an early CTA
    // that sits on filled ring slots keeps a neighbour lane's CTA off the SM.
    static const int wp_env = getenv("UPGPT_GEMM_WPREFETCH") ? atoi(getenv("UPGPT_GEMM_WPREFETCH")) : 0;
    p.w_prefetch = (a->flags & UPGPT_GEMM_F_W_STATIC) ? wp_env : 0;
  }
  p.rowstats = a->rowstats_out;
  p.ln_stats = a->ln_stats; p.ln_slots = a->ln_slots; p.ln_eps = a->ln_eps; p.ln_colsum = a->ln_colsum;
  p.ln_inv_c = 1.f / (float)a->K;
  if (p.ln_stats) {
    UPGPT_REQUIRE(p.ln_slots > 0 && p.ln_colsum && !conv && p.batch == 1, "upgpt_gemm: folded LayerNorm needs ln_slots > 0, ln_colsum and a plain GEMM");
    UPGPT_REQUIRE(!(p.flags & GEMM_CHW), "upgpt_gemm: folded LayerNorm is not available with channel-major stores");
    UPGPT_REQUIRE(((uintptr_t)p.ln_stats & 7) == 0 && ((uintptr_t)p.ln_colsum & 15) == 0, "upgpt_gemm: ln_stats / ln_colsum alignment");
  }
  if (p.rowstats) {
    UPGPT_REQUIRE(((uintptr_t)p.rowstats & 7) == 0 && !(p.flags & (GEMM_CHW | GEMM_GEGLU)) && a->out32, "upgpt_gemm: rowstats_out needs an fp32 row-major result");
  }
  p.gn_acc[0] = a->gn_acc; p.gn_acc[1] = a->gn_acc ? a->gn_acc2 : nullptr;
  { static const int dbg = getenv("UPGPT_GN_DBG") ? atoi(getenv("UPGPT_GN_DBG")) : 0; p.gn_dbg = dbg; }
  p.gn_groups = a->gn_groups; p.gn_cpg[0] = a->gn_cpg; p.gn_cpg[1] = a->gn_cpg2; p.gn_choff[0] = a->gn_choff; p.gn_choff[1] = a->gn_choff2;
  if (p.gn_acc[0]) {
    UPGPT_REQUIRE(a->out32 && !(p.flags & (GEMM_CHW | GEMM_GEGLU)) && !p.ln_stats && !p.rowstats,
                  "upgpt_gemm: gn_acc needs an fp32 row-major result and no folded LayerNorm on the same launch");
    UPGPT_REQUIRE(p.gn_groups > 0 && p.gn_cpg[0] > 0 && (!p.gn_acc[1] || p.gn_cpg[1] > 0) && p.rows_per_group % 4 == 0,
                  "upgpt_gemm: gn_acc needs gn_groups, gn_cpg > 0 and rows_per_group %% 4 == 0");
    const int imgs_per_tile = conv ? (p.tile_imgs > 1 ? p.tile_imgs : 1)
                                   : (p.rows_per_group % 128 == 0 ? 1 : (128 + p.rows_per_group - 1) / p.rows_per_group + 1);
    for (int c = 0; c < 2; ++c) {
      if (!p.gn_acc[c]) continue;
      p.gn_gt[c] = (bn + p.gn_cpg[c] - 1) / p.gn_cpg[c] + 1;
      UPGPT_REQUIRE(imgs_per_tile * p.gn_gt[c] <= 128, "upgpt_gemm: gn_acc: %d images x %d groups per tile exceed the 128 accumulator slots", imgs_per_tile, p.gn_gt[c]);
      UPGPT_REQUIRE((p.gn_choff[c] + a->N + p.gn_cpg[c] - 1) / p.gn_cpg[c] <= p.gn_groups, "upgpt_gemm: gn_acc: channels exceed gn_groups x gn_cpg");
    }
  }
  p.out32 = a->out32; p.ld32 = a->ld32 > 0 ? a->ld32 : a->N;
  const int n_out16 = (p.flags & GEMM_GEGLU) ? a->N / 2 : a->N;
  p.out16_plane = (a->flags & UPGPT_GEMM_F_SPLIT3OUT) ? n_out16 : 0;
  p.out16 = (__half*)a->out16; p.ld16 = a->ld16 > 0 ? a->ld16 : (p.out16_plane ? 2 * n_out16 : n_out16);
  p.bias = a->bias; p.rowvec = a->rowvec; p.ld_rowvec = a->ld_rowvec > 0 ? a->ld_rowvec : a->N;
  p.res32 = a->res32; p.ldres = a->ldres > 0 ? a->ldres : a->N;
  p.ldT = a->ldT > 0 ? a->ldT : p.rows_per_group;
  if (!(p.flags & GEMM_CHW)) {
    UPGPT_REQUIRE(a->N % 4 == 0, "upgpt_gemm: N (=%d) must be a multiple of 4 unless the channel-major epilogue is used", a->N);
    UPGPT_REQUIRE(!p.bias || ((uintptr_t)p.bias & 15) == 0, "upgpt_gemm: bias must be 16-byte aligned");
    UPGPT_REQUIRE(!p.rowvec || (((uintptr_t)p.rowvec & 15) == 0 && p.ld_rowvec % 4 == 0), "upgpt_gemm: rowvec must be 16-byte aligned with ld%%4==0");
    UPGPT_REQUIRE(!p.res32 || (((uintptr_t)p.res32 & 15) == 0 && p.ldres % 4 == 0), "upgpt_gemm: res32 must be 16-byte aligned with ld%%4==0");
    UPGPT_REQUIRE(!p.out32 || (p.ld32 % 4 == 0 && ((uintptr_t)p.out32 & 15) == 0), "upgpt_gemm: out32 must be 16-byte aligned with ld%%4==0");
    UPGPT_REQUIRE(!p.out16 || (p.ld16 % 8 == 0 && ((uintptr_t)p.out16 & 15) == 0), "upgpt_gemm: out16 must be 16-byte aligned with ld%%8==0");
  }

  // ---- epilogue maps (TMA store of C, TMA prefetch of the residual) ----
  p.epi_mode = (splits > 1) ? 0 : epi_mode;
  p.res_tma = 0;
  CUtensorMap tmC = tmA, tmR = tmA;   // placeholders when unused (never dereferenced)
  if (p.epi_mode) {
    const int eb = p.epi_mode == 1 ? 4 : 2;
    const uint32_t cw = p.epi_mode == 1 ? 32 : 64;
    const void* cbase = p.epi_mode == 1 ? (const void*)p.out32 : (const void*)p.out16;
    const uint64_t ldc = p.epi_mode == 1 ? (uint64_t)p.ld32 : (uint64_t)p.ld16;
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    if (conv) {
      dims[0] = a->N; dims[1] = a->W; dims[2] = a->H; dims[3] = a->n_imgs;
      strides[0] = ldc * eb; strides[1] = ldc * eb * a->W; strides[2] = ldc * eb * a->W * a->H;
      box[0] = cw; box[1] = p.tile_imgs > 1 ? a->W : p.tile_cols; box[2] = p.tile_imgs > 1 ? a->H : p.tile_rows;
      box[3] = p.tile_imgs > 1 ? p.tile_imgs : 1;
    } else {
      dims[0] = a->N; dims[1] = a->M; dims[2] = 1; dims[3] = p.batch;
      strides[0] = ldc * eb; strides[1] = ldc * eb * a->M; strides[2] = ldc * eb * a->M;
      box[0] = cw; box[1] = 128; box[2] = 1; box[3] = 1;
    }
    if (!dry && make_tmap(&tmC, eb, cbase, 4, dims, strides, box, true)) return -3;
    if (p.epi_mode == 1 && p.res32 && p.ldres % 4 == 0 && ((uintptr_t)p.res32 & 15) == 0) {
      const uint64_t ldr = p.ldres;
      if (conv) { strides[0] = ldr * 4; strides[1] = ldr * 4 * a->W; strides[2] = ldr * 4 * a->W * a->H; }
      else { strides[0] = ldr * 4; strides[1] = ldr * 4 * a->M; strides[2] = ldr * 4 * a->M; }
      if (!dry && make_tmap(&tmR, 4, p.res32, 4, dims, strides, box, true)) return -3;
      p.res_tma = 1;
    }
  }

  // ---- one tile per CTA: the dead operand pipeline serves the epilogue (bulk residual prefetch, TMA-stored fp16 planes) ----
  const bool tmR_valid = p.res_tma == 1;
  const bool single_tile = p.num_splits == 1 && p.num_m_tiles * p.num_n_tiles * p.batch <= g_num_sms;
  static const bool epi2_env = getenv("UPGPT_GEMM_NO_EPI2") == nullptr;
  const bool want_bulk = epi2_env && single_tile && p.epi_mode == 1 && tmR_valid;
  const bool want_h = epi2_env && single_tile && p.epi_mode == 1 && !conv && p.out16 != nullptr && p.ld16 % 8 == 0;
  if (want_bulk) p.res_tma = 0;      // no dedicated per-chunk residual buffers: more pipeline stages instead

  // ---- pipeline depth from the smem budget ----
  const size_t stage_bytes = (size_t)kABytes + (size_t)bn * 128;
  size_t epi_bytes = 0;
  int stages = 0;
  for (int pass = 0; pass < 2; ++pass) {
    epi_bytes = 2 * 16384 + (p.res_tma ? 2 * 16384 : 0) + 2 * (128 * 8 + 128 * 4 + 256 * 4) + 3 * 2048 /*LayerNorm tables*/ + 1024 /*alignment slack*/;
    stages = (int)(((size_t)g_smem_optin - 1024 - 256 - epi_bytes) / stage_bytes);
    // x3 needs two slot pairs in flight to overlap loads with MMAs: give up the residual TMA prefetch buffers first
    if (p.x3 && stages < 4 && p.res_tma) { p.res_tma = 0; continue; }
    break;
  }
  const int loads_per_split = (p.x3 ? 2 : 1) * ((k_iters + splits - 1) / splits);
  if (stages > (p.x3 ? 8 : 6)) stages = p.x3 ? 8 : 6;
  if (stages > loads_per_split + 1) stages = loads_per_split + 1;
  if (p.x3) stages &= ~1;
  if (stages < 2) stages = 2;
  p.stages = stages;
  CUtensorMap tmH = tmA;
  {
    const size_t pipe_bytes = (size_t)stages * stage_bytes;
    const int nch32 = (bn + 31) / 32;
    auto pad_for = [&](int units) { const size_t need = (size_t)units * 16384; return need > pipe_bytes ? need - pipe_bytes : (size_t)0; };
    auto fits = [&](int units) { return 1024 + pipe_bytes + pad_for(units) + (2 * stages + 6) * 8 + 16 + epi_bytes <= (size_t)g_smem_optin; };
    int units = 0;
    if (want_bulk && fits(nch32)) { p.res_bulk = 1; p.off_resb = 0; units = nch32; }
    if (want_h && fits(units + 2)) {
      p.h_tma = 1; p.off_hst = (uint32_t)units * 16384u; units += 2;
      p.h_planes = p.out16_plane > 0 ? 2 : 1;
      uint64_t dims[4] = {(uint64_t)a->N, (uint64_t)a->M, (uint64_t)p.h_planes, (uint64_t)p.batch};
      uint64_t strides[3] = {(uint64_t)p.ld16 * 2, (uint64_t)(p.out16_plane > 0 ? p.out16_plane : a->N) * 2, (uint64_t)p.ld16 * 2 * a->M};
      uint32_t box[4] = {32, 128, (uint32_t)p.h_planes, 1};
      if (!dry && make_tmap(&tmH, 2, p.out16, 4, dims, strides, box, false)) return -3;
    }
    p.pipe_pad = (uint32_t)pad_for(units);
    if (epi2_env && geglu && !conv && p.out16 && p.ld16 % 8 == 0 && ((uintptr_t)p.out16 & 15) == 0 && (bn / 2) % 32 == 0) {
      // GEGLU: every chunk of the tile's 'x' half is 32 columns wide -> TMA-stored fp16 planes from the groups' staging buffers
      p.geglu_tma = 1;
      p.h_planes = p.out16_plane > 0 ? 2 : 1;
      const int n_out = a->N / 2;
      uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)a->M, (uint64_t)p.h_planes, (uint64_t)p.batch};
      uint64_t strides[3] = {(uint64_t)p.ld16 * 2, (uint64_t)(p.out16_plane > 0 ? p.out16_plane : n_out) * 2, (uint64_t)p.ld16 * 2 * a->M};
      uint32_t box[4] = {32, 128, (uint32_t)p.h_planes, 1};
      if (!dry && make_tmap(&tmH, 2, p.out16, 4, dims, strides, box, false)) return -3;
    }
  }
  const size_t smem = 1024 + stages * stage_bytes + p.pipe_pad + (2 * stages + 6) * 8 + 16 + epi_bytes;
  UPGPT_REQUIRE(smem <= (size_t)g_smem_optin, "upgpt_gemm: smem %zu > %d", smem, g_smem_optin);

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.num_splits * p.batch;
  p.fd_splits = make_fastdiv(p.num_splits);
  p.fd_tiles_mn = make_fastdiv(p.num_m_tiles * p.num_n_tiles);
  p.fd_tiles_per_batch = make_fastdiv(p.num_m_tiles * p.num_n_tiles * p.num_splits);
  p.fd_m_tiles = make_fastdiv(p.num_m_tiles);
  p.fd_tiles_per_img = make_fastdiv(p.tiles_per_img);
  p.fd_tiles_per_row = make_fastdiv(p.tiles_per_row);
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  // split-K flavour: the splits of a tile as one thread-block cluster reducing over DSMEM (one tile per CTA; the fp32 partial
  // tile [128][bn] is laid over the drained operand slots), else the global-workspace reduction
  p.cluster_reduce = (p.num_splits > 1 && p.num_splits <= 8 && (size_t)stages * stage_bytes >= (size_t)512 * bn &&
                      getenv("UPGPT_NO_CLUSTER_SPLITK") == nullptr) ? 1 : 0;
  p.coop_reduce = (!p.cluster_reduce && p.num_splits > 1 && num_tiles <= grid) ? 1 : 0;
  if (p.ln_stats || p.rowstats)
    UPGPT_REQUIRE(p.num_splits == 1 || p.cluster_reduce, "upgpt_gemm: folded LayerNorm needs the cluster split-K reduction (splits=%d)", p.num_splits);
  if (p.rowstats)
    UPGPT_REQUIRE(p.cluster_reduce || p.epi_mode == 1, "upgpt_gemm: rowstats_out needs the TMA-store epilogue (16-byte aligned fp32 rows)");
  if (p.rowstats && a->rowstats_slots > 0 && !dry)
    UPGPT_REQUIRE(p.num_n_tiles == a->rowstats_slots,
                  "upgpt_gemm: rowstats_out: this launch is tiled into %d N tiles but its consumers were sized for %d partial slots per row "
                  "(the tiling objective / UPGPT_GEMM_* settings changed since the program was recorded: rebuild the engine)",
                  p.num_n_tiles, a->rowstats_slots);
  const int epi_path = (p.num_splits == 1 && p.epi_mode == 1) ? 1 : (p.cluster_reduce ? 2 : 0);
  if (p.gn_acc[0]) {
    UPGPT_REQUIRE(epi_path != 0, "upgpt_gemm: gn_acc needs the TMA-store epilogue or the cluster split-K reduction (block_n=%d splits=%d)", p.block_n, p.num_splits);
    UPGPT_REQUIRE(epi_path != 2 || 128 % p.num_splits == 0, "upgpt_gemm: gn_acc with split-K needs a power-of-two split factor (got %d)", p.num_splits);
  }
  if (dry) {
    plan_out[0] = p.block_n; plan_out[1] = p.num_n_tiles; plan_out[2] = p.num_splits; plan_out[3] = p.stages; plan_out[4] = epi_path;
    plan_out[5] = plan_out[6] = plan_out[7] = 0;
    return 0;
  }
  if (p.cluster_reduce) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_tiles); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.num_splits; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    UPGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel, tmA, tmB, tmC, tmR, tmH, p));
  } else if (p.coop_reduce) {
    // the distributed split-K reduction spins on the arrival of sibling CTAs: a cooperative launch guarantees (or refuses)
    // co-residency instead of risking a deadlock
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    UPGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel, tmA, tmB, tmC, tmR, tmH, p));
  } else {
    UPGPT_CHECK_CUDA(launch_k(tc_gemm_kernel, dim3(grid), dim3(kGemmThreads), smem, stream, tmA, tmB, tmC, tmR, tmH, p));
  }
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_gemm(const upgpt_gemm_args* a, void* stream_) { return gemm_run(a, (cudaStream_t)stream_, nullptr); }

extern "C" int upgpt_gemm_set_sm_weight(double w) {
  UPGPT_REQUIRE(w >= 0.0 && w <= 1000.0, "upgpt_gemm_set_sm_weight: weight %f out of range", w);
  g_sm_weight = w;
  return 0;
}

extern "C" int upgpt_gemm_plan(const upgpt_gemm_args* a, int plan[8]) {
  UPGPT_REQUIRE(plan, "upgpt_gemm_plan: null plan");
  return gemm_run(a, nullptr, plan);
}

// bring-up instrumentation: when set, every CTA of tc_gemm_kernel stamps %globaltimer at 12 points into buf[cta][16]
extern "C" int upgpt_debug_set_gemm_timestamps(long long* buf) {
  g_debug_ts = buf;
  return 0;
}

UPGPT_TRACE_TU(tc_gemm)
