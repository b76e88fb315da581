// Small CUDA-core kernels around the tensor-core path: boundary convolutions with tiny channel counts, the timestep
// embedding MLP (M = batch rows, weight-bandwidth bound GEMV), the fused sampler updates and device-side step state.
//
// Replaces: UNetModel.input_blocks[0] conv 5->224 incl. DiffusionWrapper's hybrid concat (openaimodel.py:519,
// ddpm.py:1567-1570); AutoencoderKL.post_quant_conv + 1/scale_factor (autoencoder.py:303,330-333; ddpm.py:779);
// timestep_embedding (util.py:151-171), time_embed (openaimodel.py:507-511), ResBlock.emb_layers (openaimodel.py:218-224);
// DDIMSampler.p_sample_ddim update (ddim.py:189-203); DDPM p_sample / q_posterior (ddpm.py:224-237,1157-1185).
#include "common.cuh"
#include "../../include/upgpt_b200.h"

namespace upgpt {

// ------------------------------------------------------------------------------------------------------------------
// Direct convolution for tiny Cin (<= 8): input NCHW fp32 taken from up to two tensors (channel concat), weights
// [Cin*k*k][Cout] fp32 (k-major so that threads = output channels read coalesced), output NHWC or NCHW fp32.
// block = (pixel tile of 8) x Cout threads (looped); grid = (HW/8, B)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_small_cin_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, float in_scale, int H,
                      int W, int ksize, const float* __restrict__ wt, const float* __restrict__ bias, int Cout,
                      float* __restrict__ out, int out_nchw) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int TP = 8;
  __shared__ float patch[TP][8 * 9];
  const int Cin = C1 + C2, HW = H * W;
  const int K = Cin * ksize * ksize;
  const int b = blockIdx.y, p0 = blockIdx.x * TP;
  const int pad = ksize / 2;
  for (int i = threadIdx.x; i < TP * K; i += blockDim.x) {
    const int tp = i / K, k = i % K;
    const int ci = k / (ksize * ksize), rs = k % (ksize * ksize);
    const int pp = p0 + tp;
    float v = 0.f;
    if (pp < HW) {
      const int y = pp / W + rs / ksize - pad, x = pp % W + rs % ksize - pad;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        v = ci < C1 ? x1[((size_t)b * C1 + ci) * HW + y * W + x] * in_scale
                    : x2[((size_t)b * C2 + (ci - C1)) * HW + y * W + x];
      }
    }
    patch[tp][k] = v;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc[TP];
    const float bv = bias ? bias[co] : 0.f;
#pragma unroll
    for (int t = 0; t < TP; ++t) acc[t] = bv;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(wt + (size_t)k * Cout + co);
#pragma unroll
      for (int t = 0; t < TP; ++t) acc[t] = fmaf(patch[t][k], w, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      const int pp = p0 + t;
      if (pp < HW) {
        if (out_nchw) out[((size_t)b * Cout + co) * HW + pp] = acc[t];
        else out[((size_t)b * HW + pp) * Cout + co] = acc[t];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sinusoidal timestep embedding: out[b] = [cos(t*f_k) | sin(t*f_k)], f_k = exp(-ln(max_period) * k / half)
// ------------------------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int B, int dim, float max_period,
                                          const float* __restrict__ freqs, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  // freqs (optional): the host-computed fp32 table exp(-ln(max_period) * k / half), so that the table is bit-identical
  // to the reference's torch.exp on the CPU (a 1-ulp difference in a frequency moves cos(t*f) by ~6e-5 at t ~ 1000)
  const float freq = freqs ? freqs[k] : expf(-logf(max_period) * (float)k / (float)half);
  const float arg = (float)t[b] * freq;
  out[(size_t)b * dim + k] = cosf(arg);
  out[(size_t)b * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[(size_t)b * dim + dim - 1] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
// out[b][n] = act_out( sum_k act_in(x[b][k]) * W[n][k] + bias[n] )   for small row counts (B <= 64)
// one warp per output feature n; the W row is read once (float4, coalesced) and reused for all rows.
// ------------------------------------------------------------------------------------------------------------------
template <int RB, bool VEC>
__global__ void __launch_bounds__(256)
linear_small_m_kernel(const float* __restrict__ x, int ldx, int rows, const float* __restrict__ w, const float* __restrict__ bias,
                      int N, int K, int silu_in, int silu_out, float* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* wrow = w + (size_t)n * K;
  for (int r0 = 0; r0 < rows; r0 += RB) {
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.f;
    if (VEC) {
      const float4* wr = (const float4*)wrow;
      const int K4 = K >> 2;
      // the weight row is streamed once from HBM: keep 4 x 512 B of it in flight per warp
      for (int i0 = lane; i0 < K4; i0 += 128) {
        float4 wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = (i0 + 32 * u < K4) ? __ldg(wr + i0 + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 32 * u;
          if (i >= K4) break;
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            if (r0 + r < rows) {
              float4 xv = *(const float4*)(x + (size_t)(r0 + r) * ldx + 4 * i);
              if (silu_in) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
              acc[r] += xv.x * wv[u].x + xv.y * wv[u].y + xv.z * wv[u].z + xv.w * wv[u].w;
            }
          }
        }
      }
    } else {   // K or the row pitches are not multiples of 4 (e.g. the 85-d SMPL vector): scalar loads
      for (int i = lane; i < K; i += 32) {
        const float wv = __ldg(wrow + i);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (r0 + r < rows) {
            float xv = x[(size_t)(r0 + r) * ldx + i];
            if (silu_in) xv = silu_f(xv);
            acc[r] = fmaf(xv, wv, acc[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r0 + r < rows) {
        v += bias ? bias[n] : 0.f;
        if (silu_out) v = silu_f(v);
        out[(size_t)(r0 + r) * ldo + n] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sampler updates. coef rows live in device memory, indexed by *step_ptr (device int) so a captured graph can be replayed.
// DDIM row: {a_t, a_prev, sigma_t, sqrt_one_minus_a_t, temperature}   (all fp32, as the reference builds them)
// ------------------------------------------------------------------------------------------------------------------
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                   const float* __restrict__ coef, const int* __restrict__ step_ptr, int step_imm,
                                   float* __restrict__ x_prev, float* __restrict__ pred_x0, size_t n,
                                   size_t noise_step_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = step_ptr ? *step_ptr : step_imm;
  const float* c = coef + (size_t)step * 5;
  const float a_t = c[0], a_prev = c[1], sigma = c[2], s1m = c[3], temp = c[4];
  const float sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_ap = __fsqrt_rn(a_prev);
  const float dir_c = __fsqrt_rn(__fsub_rn(__fsub_rn(1.f, a_prev), __fmul_rn(sigma, sigma)));
  const float* nz = noise ? noise + (size_t)step * noise_step_stride : nullptr;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[i], e = eps[i];
    const float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(s1m, e)), sqrt_at);
    const float dir = __fmul_rn(dir_c, e);
    float xp = __fadd_rn(__fmul_rn(sqrt_ap, p0), dir);
    const float nv = nz ? __fmul_rn(__fmul_rn(sigma, nz[i]), temp) : 0.f;
    xp = __fadd_rn(xp, nv);
    x_prev[i] = xp;
    if (pred_x0) pred_x0[i] = p0;
  }
}

// DDPM ancestral step (ddpm.py:1157-1185 with clip_denoised=False):
// row: {sqrt_recip_acp, sqrt_recipm1_acp, post_mean_coef1, post_mean_coef2, post_log_var_clipped, nonzero_mask}
__global__ void ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                   const float* __restrict__ coef, const int* __restrict__ step_ptr, int step_imm,
                                   float* __restrict__ x_prev, float* __restrict__ pred_x0, size_t n,
                                   size_t noise_step_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = step_ptr ? *step_ptr : step_imm;
  const float* c = coef + (size_t)step * 6;
  const float sr = c[0], srm1 = c[1], c1 = c[2], c2 = c[3], logvar = c[4], nzmask = c[5];
  const float std = expf(0.5f * logvar);
  const float* nz = noise ? noise + (size_t)step * noise_step_stride : nullptr;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[i], e = eps[i];
    const float x0 = __fsub_rn(__fmul_rn(sr, xv), __fmul_rn(srm1, e));
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
    const float nv = nz ? __fmul_rn(__fmul_rn(nzmask, std), nz[i]) : 0.f;
    x_prev[i] = __fadd_rn(mean, nv);
    if (pred_x0) pred_x0[i] = x0;
  }
}

__global__ void step_state_kernel(int* step_ptr, int op, int value, long long* t_buf, int B, const long long* t_table) {
  pdl_launch_dependents();
  pdl_wait();
  // op 0: set *step = value ; op 1: *step += value ; then (optionally) broadcast t_table[*step] into t_buf[0..B)
  __shared__ int s;
  if (threadIdx.x == 0) {
    int v = *step_ptr;
    if (op == 0) v = value; else if (op == 1) v += value;
    *step_ptr = v;
    s = v;
  }
  __syncthreads();
  if (t_buf && t_table)
    for (int b = threadIdx.x; b < B; b += blockDim.x) t_buf[b] = t_table[s];
}

// out = a * sa + b * sb   (mask blend / q_sample style helpers, ddim.py:144-147; ddpm.py:281-284)
__global__ void axpby_kernel(const float* __restrict__ a, float sa, const float* __restrict__ b, float sb,
                             float* __restrict__ out, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = fmaf(a[i], sa, b ? b[i] * sb : 0.f);
}

// image post-processing of generate_utils.py:165-168: clamp(-1,1) * 0.5 + 0.5, NCHW fp32 -> NHWC uint8 (x255, rounded)
__global__ void to_uint8_nhwc_kernel(const float* __restrict__ x, int B, int C, int HW, uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = (size_t)B * C * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t bp = i / C;
    const size_t b = bp / HW, pp = bp % HW;
    float v = x[(b * C + c) * HW + pp];
    v = fminf(fmaxf(v, -1.f), 1.f) * 0.5f + 0.5f;
    out[i] = (uint8_t)__float2int_rn(v * 255.f);
  }
}

}  // namespace upgpt

using namespace upgpt;

extern "C" int upgpt_conv_small_cin(const float* x1, int C1, const float* x2, int C2, float in_scale, int B, int H, int W,
                                    int ksize, const float* wt_kmajor, const float* bias, int Cout, float* out,
                                    int out_nchw, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x1 && wt_kmajor && out, "conv_small_cin: null");
  UPGPT_REQUIRE(C1 + C2 <= 8 && (ksize == 1 || ksize == 3), "conv_small_cin: Cin<=8, ksize in {1,3}");
  dim3 grid((H * W + 7) / 8, B);
  const int threads = Cout >= 256 ? 256 : ((Cout + 31) / 32) * 32;
  UPGPT_CHECK_CUDA(launch_k(conv_small_cin_kernel, grid, dim3(threads), 0, stream, x1, C1, x2, C2, in_scale == 0.f ? 1.f : in_scale, H, W, ksize,
                            wt_kmajor, bias, Cout, out, out_nchw));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_timestep_embedding(const long long* t, int B, int dim, float max_period, const float* freqs, float* out,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(t && out && dim >= 2, "timestep_embedding: bad args");
  const int n = B * (dim / 2);
  UPGPT_CHECK_CUDA(launch_k(timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, t, B, dim, max_period, freqs, out));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_linear_small_m(const float* x, int ldx, int rows, const float* w, const float* bias, int N, int K,
                                    int silu_in, int silu_out, float* out, int ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && w && out && K > 0 && rows > 0, "linear_small_m: bad args (K=%d)", K);
  if (ldx <= 0) ldx = K;
  if (ldo <= 0) ldo = N;
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)x | (uintptr_t)w) % 16 == 0);
  const int blocks = (N + 7) / 8;
  if (vec) {
    if (rows <= 4) UPGPT_CHECK_CUDA(launch_k(linear_small_m_kernel<4, true>, dim3(blocks), dim3(256), 0, stream, x, ldx, rows, w, bias, N, K, silu_in, silu_out, out, ldo));
    else UPGPT_CHECK_CUDA(launch_k(linear_small_m_kernel<8, true>, dim3(blocks), dim3(256), 0, stream, x, ldx, rows, w, bias, N, K, silu_in, silu_out, out, ldo));
  } else {
    UPGPT_CHECK_CUDA(launch_k(linear_small_m_kernel<4, false>, dim3(blocks), dim3(256), 0, stream, x, ldx, rows, w, bias, N, K, silu_in, silu_out, out, ldo));
  }
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static inline unsigned ew_grid(size_t n) {
  size_t g = (n + 255) / 256;
  return (unsigned)(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

extern "C" int upgpt_ddim_step(const float* x, const float* eps, const float* noise, long long noise_step_stride,
                               const float* coef, const int* step_ptr, int step_imm, float* x_prev, float* pred_x0,
                               long long n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && eps && coef && x_prev && n > 0, "ddim_step: bad args");
  UPGPT_CHECK_CUDA(launch_k(ddim_update_kernel, dim3(ew_grid((size_t)n)), dim3(256), 0, stream, x, eps, noise, coef, step_ptr, step_imm, x_prev, pred_x0,
                            (size_t)n, (size_t)noise_step_stride));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_ddpm_step(const float* x, const float* eps, const float* noise, long long noise_step_stride,
                               const float* coef, const int* step_ptr, int step_imm, float* x_prev, float* pred_x0,
                               long long n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && eps && coef && x_prev && n > 0, "ddpm_step: bad args");
  UPGPT_CHECK_CUDA(launch_k(ddpm_update_kernel, dim3(ew_grid((size_t)n)), dim3(256), 0, stream, x, eps, noise, coef, step_ptr, step_imm, x_prev, pred_x0,
                            (size_t)n, (size_t)noise_step_stride));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_step_state(int* step_ptr, int op, int value, long long* t_buf, int B, const long long* t_table,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(step_ptr, "step_state: null");
  UPGPT_CHECK_CUDA(launch_k(step_state_kernel, dim3(1), dim3(64), 0, stream, step_ptr, op, value, t_buf, B, t_table));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_axpby(const float* a, float sa, const float* b, float sb, float* out, long long n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(a && out && n > 0, "axpby: bad args");
  UPGPT_CHECK_CUDA(launch_k(axpby_kernel, dim3(ew_grid((size_t)n)), dim3(256), 0, stream, a, sa, b, sb, out, (size_t)n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dst[b][i] = table[*step * row_stride + i] for b < B: one row of a per-step table broadcast over the batch (the timestep-embedding
// projections of all denoising steps are computed once per schedule; the step graph only gathers its row)
namespace upgpt {
__global__ void __launch_bounds__(256)
gather_step_row_kernel(const float* __restrict__ table, long long row_stride, const int* __restrict__ step_ptr, float* __restrict__ dst,
                       int B, int n) {
  pdl_launch_dependents();
  pdl_wait();
  const float* src = table + (size_t)(*step_ptr) * row_stride;
  const int n4 = n >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const float4 v = __ldg((const float4*)src + i);
    for (int b = 0; b < B; ++b) *((float4*)(dst + (size_t)b * n) + i) = v;
  }
}
}  // namespace upgpt

extern "C" int upgpt_gather_step_row(const float* table, long long row_stride, const int* step_ptr, float* dst, int B, int n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(table && step_ptr && dst && B > 0 && n > 0 && n % 4 == 0 && row_stride % 4 == 0, "gather_step_row: bad args (n=%d)", n);
  UPGPT_CHECK_CUDA(launch_k(gather_step_row_kernel, dim3((n / 4 + 255) / 256), dim3(256), 0, stream, table, row_stride, step_ptr, dst, B, n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out = (wa*a + wb*b + wc*c + wd*d) * inv_den with the products accumulated left to right (null pointers are skipped):
// the Adams-Bashforth eps combinations of PLMS (plms.py:217-229), e.g. (55 e_t - 59 e_1 + 37 e_2 - 9 e_3) / 24
namespace upgpt {
__global__ void __launch_bounds__(256)
lincomb4_kernel(const float* __restrict__ a, float wa, const float* __restrict__ b, float wb, const float* __restrict__ c, float wc,
                const float* __restrict__ d, float wd, float den, float* __restrict__ out, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // same association as the reference expression: ((wa*a + wb*b) + wc*c) + wd*d, each product rounded, then one division
    float v = wa * a[i];
    if (b) v = v + wb * b[i];
    if (c) v = v + wc * c[i];
    if (d) v = v + wd * d[i];
    out[i] = __fdiv_rn(v, den);
  }
}
}  // namespace upgpt

extern "C" int upgpt_lincomb4(const float* a, float wa, const float* b, float wb, const float* c, float wc, const float* d, float wd,
                              float den, float* out, long long n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(a && out && n > 0 && den != 0.f, "lincomb4: bad args");
  UPGPT_CHECK_CUDA(launch_k(lincomb4_kernel, dim3(ew_grid((size_t)n)), dim3(256), 0, stream, a, wa, b, wb, c, wc, d, wd, den, out, (size_t)n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// q_sample (+ mask blend): out = q * m + (1 - m) * img with q = sqrt(acp[t]) x0 + sqrt(1 - acp[t]) noise (ddpm.py:281-284, the
// known-region blend of ddim.py:144-147 / ddpm.py:1281-1284); mask == NULL: out = q (q_sample, stochastic_encode ddim.py:207-221).
// t per sample (t_per_sample), or one t for the batch read from t_table[*step_ptr] (device-side step counter: graph replays need no
// host data), or t_imm. noise is indexed + step * noise_step_stride (a [S][B][C][HW] table) when the step counter is used.
namespace upgpt {
__global__ void __launch_bounds__(256)
qsample_blend_kernel(const float* __restrict__ x0, const float* __restrict__ noise, long long noise_step_stride, const float* __restrict__ mask,
                     int mask_c, const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ sqrt_acp,
                     const float* __restrict__ sqrt_1m_acp, const long long* __restrict__ t_per_sample, const long long* __restrict__ t_table,
                     const int* __restrict__ step_ptr, int t_imm, int C, int HW, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = step_ptr ? *step_ptr : 0;
  const long long t_all = t_table ? t_table[step] : (long long)t_imm;
  const float* nz = noise + (step_ptr ? (size_t)step * (size_t)noise_step_stride : 0);
  const size_t chw = (size_t)C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw, r = i - b * chw;
    const long long t = t_per_sample ? t_per_sample[b] : t_all;
    const float q = sqrt_acp[t] * x0[i] + sqrt_1m_acp[t] * nz[i];
    if (mask) {
      const float m = mask[mask_c == 1 ? b * HW + (r % HW) : i];
      out[i] = q * m + (1.f - m) * img[i];
    } else {
      out[i] = q;
    }
  }
}
}  // namespace upgpt

extern "C" int upgpt_qsample_blend(const float* x0, const float* noise, long long noise_step_stride, const float* mask, int mask_c,
                                   const float* img, float* out, const float* sqrt_acp, const float* sqrt_1m_acp,
                                   const long long* t_per_sample, const long long* t_table, const int* step_ptr, int t_imm, int B, int C,
                                   int HW, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x0 && noise && out && sqrt_acp && sqrt_1m_acp && B > 0 && C > 0 && HW > 0, "qsample_blend: bad args");
  UPGPT_REQUIRE(!mask || (img && (mask_c == 1 || mask_c == C)), "qsample_blend: mask needs img and 1 or C mask channels");
  const size_t n = (size_t)B * C * HW;
  UPGPT_CHECK_CUDA(launch_k(qsample_blend_kernel, dim3(ew_grid(n)), dim3(256), 0, stream, x0, noise, noise_step_stride, mask, mask_c, img, out,
                            sqrt_acp, sqrt_1m_acp, t_per_sample, t_table, step_ptr, t_imm, C, HW, n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// z = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * out_scale from NCHW moments [B][2C][HW] = {mean | logvar}
// (DiagonalGaussianDistribution.sample / .mode, distributions.py:24-37; x scale_factor of get_first_stage_encoding, ddpm.py:569-576)
namespace upgpt {
__global__ void __launch_bounds__(256)
gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise, float out_scale, float* __restrict__ out,
                       int C, int HW, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t chw = (size_t)C * HW;
    const size_t b = i / chw, r = i - b * chw;
    const float mean = moments[b * 2 * chw + r];
    float v = mean;
    if (noise) {
      float lv = moments[b * 2 * chw + chw + r];
      lv = fminf(fmaxf(lv, -30.f), 20.f);
      v = fmaf(expf(0.5f * lv), noise[i], mean);
    }
    out[i] = v * out_scale;
  }
}
}  // namespace upgpt

extern "C" int upgpt_gaussian_sample(const float* moments, const float* noise, float out_scale, float* out, int B, int C, int HW,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(moments && out && B > 0 && C > 0 && HW > 0, "gaussian_sample: bad args");
  const size_t n = (size_t)B * C * HW;
  UPGPT_CHECK_CUDA(launch_k(gaussian_sample_kernel, dim3(ew_grid(n)), dim3(256), 0, stream, moments, noise, out_scale, out, C, HW, n));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_to_uint8_nhwc(const float* x, int B, int C, int HW, uint8_t* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && out, "to_uint8_nhwc: null");
  UPGPT_CHECK_CUDA(launch_k(to_uint8_nhwc_kernel, dim3(ew_grid((size_t)B * C * HW)), dim3(256), 0, stream, x, B, C, HW, out));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- CUDA graph helpers (stream capture of a sequence of upgpt_* calls) ----
#include <unordered_map>
#include <vector>
#include <mutex>
namespace upgpt { void count_launches(long long n); }
static std::unordered_map<void*, long long> g_graph_kernels;
static std::mutex g_graph_mu;

extern "C" int upgpt_capture_begin(void* stream_) {
  UPGPT_CHECK_CUDA(cudaStreamBeginCapture((cudaStream_t)stream_, cudaStreamCaptureModeThreadLocal));
  return 0;
}
extern "C" int upgpt_capture_end(void* stream_, void** graph_exec_out) {
  cudaGraph_t g = nullptr;
  UPGPT_CHECK_CUDA(cudaStreamEndCapture((cudaStream_t)stream_, &g));
  size_t n_nodes = 0;
  long long kernels = 0;
  if (cudaGraphGetNodes(g, nullptr, &n_nodes) == cudaSuccess && n_nodes > 0) {
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    if (cudaGraphGetNodes(g, nodes.data(), &n_nodes) == cudaSuccess) {
      for (size_t i = 0; i < n_nodes; ++i) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++kernels;
      }
    }
  }
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  UPGPT_CHECK_CUDA(e);
  {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    g_graph_kernels[(void*)ge] = kernels;
  }
  *graph_exec_out = (void*)ge;
  return 0;
}
extern "C" int upgpt_graph_launch(void* graph_exec, void* stream_) {
  UPGPT_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream_));
  {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    auto it = g_graph_kernels.find(graph_exec);
    if (it != g_graph_kernels.end()) count_launches(it->second);
  }
  return 0;
}
extern "C" long long upgpt_graph_kernel_count(void* graph_exec) {
  std::lock_guard<std::mutex> lk(g_graph_mu);
  auto it = g_graph_kernels.find(graph_exec);
  return it == g_graph_kernels.end() ? -1 : it->second;
}
extern "C" int upgpt_graph_destroy(void* graph_exec) {
  if (graph_exec) {
    {
      std::lock_guard<std::mutex> lk(g_graph_mu);
      g_graph_kernels.erase(graph_exec);
    }
    UPGPT_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  }
  return 0;
}

UPGPT_TRACE_TU(misc)
