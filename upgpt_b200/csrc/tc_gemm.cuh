// Parameter block of the tcgen05 GEMM / implicit-GEMM convolution kernel (tc_gemm.cu).
#pragma once
#include <stdint.h>
#include <cuda_fp16.h>

namespace upgpt {

enum GemmFlags : uint32_t {
  GEMM_GEGLU = 1u << 1,    // tile columns are [x | gate] halves; out16 gets x * gelu(gate)
  GEMM_CHW = 1u << 2,      // outputs stored channel-major: out[(group * N_total + n) * ldT + row_in_group]
  GEMM_CONV = 1u << 3,
  GEMM_UP2 = 1u << 7,      // conv over the nearest-x2 upsampled image as four parity-wise 2x2 convolutions over the low-resolution operand
  GEMM_W_STATIC = 1u << 6, // W holds model weights (independent of the stream's preceding launches): fetched BEFORE the PDL dependency wait
  GEMM_SPLIT3OUT = 1u << 4,  // (public flag) out16 written as error-compensated [hi | lo] planes
  GEMM_X3 = 1u << 5,         // (public flag) operands carry [hi | lo] planes; D = Ah*Wh + Al*Wh + Ah*Wl
};

// Division by a launch-invariant divisor as multiply-high + shift (valid for dividends < 2^31). The tile decode of the single TMA
// producer lane is a chain of ~8 dependent integer divisions: with hardware-emulated division it cost ~1 us before the first load
// of every launch (in-kernel timeline), i.e. ~0.2 ms per U-Net step.
struct FastDiv {
  uint32_t mul, shift, d;
#ifdef __CUDACC__
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shift); }
#endif
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0, 0, (uint32_t)(d < 1 ? 1 : d)};
  if (f.d > 1) {
    uint32_t l = 0;
    while ((1u << l) < f.d) ++l;                    // ceil(log2(d))
    const uint32_t p = 31 + l;
    f.mul = (uint32_t)((((uint64_t)1 << p) + f.d - 1) / f.d);
    f.shift = p - 32;
  }
  return f;
}

struct GemmParams {
  // ---- problem ----
  int M_total;          // rows per batch entry
  int N_total;          // output columns
  int batch;            // batched GEMM count (A coord3 / B coord3 = batch index), 1 otherwise
  int block_n;          // UMMA N per tile (multiple of 16, <= 256)
  int num_m_tiles, num_n_tiles, num_splits;
  int taps;             // 1 (GEMM / 1x1) or 9 (3x3)
  int kblocks_per_tap;  // ceil(K_per_tap / 64)
  int stages;           // smem pipeline depth (even in x3 mode: a k-block occupies the slot pair {hi planes, lo planes})
  int x3;               // error-compensated product of [hi | lo] operand planes: 3 MMAs per k-step on 2 loaded plane pairs
  uint32_t a_bytes;     // bytes one A TMA box delivers (<= 16384)
  uint32_t flags;
  // ---- conv geometry (GEMM_CONV) ----
  int H, W, HW;
  int tile_rows;        // image rows per M tile (tile_imgs == 1)
  int tile_imgs;        // whole images per M tile (HW * tile_imgs <= 128)
  int tiles_per_img;
  int tiles_per_row;    // column tiles per image row (W / tile_cols)
  int tile_cols;        // pixels per tile row (divides W)
  int n_imgs;           // number of images addressed by the output
  int tap_dy[16], tap_dx[16], tap_dn[16];   // per tap; GEMM_UP2: per (parity, tap) = [parity * 4 + tap]
  // ---- epilogue ----
  float* out32;         // [rows, ld32] (or CHW)
  int ld32;
  __half* out16;        // optional fp16 copy of the result (GEGLU: the only output)
  int ld16;
  int out16_plane;      // > 0: out16 rows are [hi | lo] planes of this many columns each (fp16x3 operand layout)
  const float* bias;    // [N_total] or null
  const float* rowvec;  // [groups, ld_rowvec] added per row-group (timestep-embedding bias), or null
  int ld_rowvec;
  int rows_per_group;   // rows per image / per batch entry (group = row / rows_per_group)
  const float* res32;   // residual [rows, ldres] or null
  int ldres;
  int ldT;              // CHW: stride between channels (>= rows_per_group)
  float out_scale;      // applied to the accumulator before bias (1.0 default)
  // ---- folded LayerNorm (include/upgpt_b200.h) ----
  float* rowstats;          // producer: [rows][num_n_tiles][2] {sum, sumsq} of the final fp32 row over this N tile
  const float* ln_stats;    // consumer: [rows][ln_slots][2]
  int ln_slots;
  float ln_eps, ln_inv_c;   // 1 / K
  const float* ln_colsum;   // [N_total]
  // ---- GroupNorm moments of the result for up to two consumers (include/upgpt_b200.h: gn_acc) ----
  long long* gn_acc[2];
  int gn_groups, gn_cpg[2], gn_choff[2];
  int gn_dbg;               // bring-up (UPGPT_GN_DBG): bit 0 = skip the atomics, bit 1 = skip the whole commit
  int gn_gt[2];             // accumulator slots per image of a tile: the groups an N tile can touch
  int w_prefetch;           // > 0: W boxes of the first tile are issued before the PDL wait (GEMM_W_STATIC); > 1: + L2 prefetch depth
  // ---- deterministic split-K ----
  float* ws;            // [num_splits][ws_rows][ws_ld] partial tiles
  int ws_rows, ws_ld;
  int* counters;        // {arrived, done} counters per (batch, n tile, m tile); zero between launches
  int coop_reduce;      // all splits of a tile are co-resident (cooperative launch): parallel distributed reduction
  int cluster_reduce;   // the splits of a tile form one thread-block cluster: partials stay in shared memory, reduced over DSMEM
  // ---- TMA epilogue ----
  int epi_mode;         // 0 flat stores; 1 fp32 tile chunks [rows][32] by TMA store; 2 fp16 chunks [rows][64] by TMA store
  int res_tma;          // residual tile chunks prefetched by TMA load (epi_mode 1)
  // one tile per CTA: after the main loop the operand pipeline's shared memory is dead and serves the epilogue
  int res_bulk;         // ALL residual chunks of the tile are fetched at once (one TMA round trip instead of one per chunk) into
  uint32_t off_resb;    //   the dead pipeline region at this byte offset, chunk c at + c * 16 KB
  int h_tma;            // the fp16 copy ([hi | lo] planes) of a chunk is staged at off_hst + group * 16 KB as [plane][128 rows][32 halves]
  uint32_t off_hst;     //   and stored by TMA (tmH) instead of the flat pass
  int h_planes;         // 1 or 2
  int geglu_tma;        // GEGLU epilogue: finished 32-column chunks are staged as [plane][128 rows][32 halves] and stored by TMA (tmH)
  uint32_t pipe_pad;    // bytes appended to the operand pipeline region so that the epilogue buffers above fit it
  long long* debug_ts;  // optional [gridDim.x][16] globaltimer stamps (bring-up instrumentation), null in production
  // ---- fast division for the tile decode ----
  FastDiv fd_splits, fd_tiles_mn, fd_tiles_per_batch, fd_m_tiles, fd_tiles_per_img, fd_tiles_per_row;
};

}  // namespace upgpt
