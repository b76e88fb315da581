// CUDA-core kernels around the tensor-core GEMMs of the CLIP ViT-L/14 conditioning towers.
//
// The reference runs the towers once per request, before the denoising loop:
//   text : FrozenCLIPEmbedder.forward (ldm/modules/encoders/modules.py:137-162) -> transformers.CLIPTextModel.last_hidden_state
//          (token + position embedding, 12 pre-LN layers with CAUSAL self-attention, quick_gelu MLP, final LayerNorm)
//   image: FrozenClipImageEmbedder2.forward (modules.py:234-256) -> clip.model.VisionTransformer (OpenAI CLIP, ViT-L/14):
//          14x14 stride-14 patch conv (no bias), class token, position embedding, ln_pre, 24 pre-LN layers (QuickGELU MLP),
//          ln_post on the class token, projection 1024 -> 768.
// Every Linear of the towers goes through upgpt_gemm (tcgen05); the image tower's attention through upgpt_attention. What is
// left are HBM-bound gathers / elementwise passes and the 77-token causal attention of the text tower, all below.
#include "common.cuh"
#include "../../include/upgpt_b200.h"

namespace upgpt {

static inline unsigned clip_grid(size_t n, int per_block) {
  size_t g = (n + per_block - 1) / per_block;
  return (unsigned)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

__device__ __forceinline__ void store_planes4(__half* hi_row, int plane, int c, float a0, float a1, float a2, float a3) {
  __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
  *(uint2*)(hi_row + c) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
  if (plane > 0) {
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    __half2 l0 = __floats2half2_rn(a0 - f0.x, a1 - f0.y), l1 = __floats2half2_rn(a2 - f1.x, a3 - f1.y);
    *(uint2*)(hi_row + plane + c) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
  }
}

// out[r][:] = tok_emb[ids[r]][:] + pos_emb[r % seq][:]      (CLIPTextEmbeddings.forward)
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ ids, int rows, int seq, int vocab, const float* __restrict__ tok, const float* __restrict__ pos,
                    int C, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C >> 2;
  const size_t n = (size_t)rows * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C4), c = (int)(i % C4);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4 a = ((const float4*)(tok + (size_t)id * C))[c];
    const float4 b = ((const float4*)(pos + (size_t)(r % seq) * C))[c];
    ((float4*)(out + (size_t)r * C))[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// im2col of a P x P stride-P patch convolution: img NCHW fp32 [n][Cin][S][S] -> rows [n * G * G][Kpad] fp16 (G = S / P), row = one patch
// flattened in (c, py, px) order (the order of conv1.weight.reshape(width, -1)), columns >= Cin*P*P zero; [hi | lo] planes when plane > 0.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, int n, int Cin, int S, int P, int Kpad, int plane, __half* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int G = S / P, K = Cin * P * P, K4 = Kpad >> 2;
  const size_t total = (size_t)n * G * G * K4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k4 = (int)(i % K4);
    const size_t row = i / K4;
    const int gx = (int)(row % G), gy = (int)((row / G) % G), b = (int)(row / ((size_t)G * G));
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k4 * 4 + u;
      v[u] = 0.f;
      if (k < K) {
        const int c = k / (P * P), py = (k / P) % P, px = k % P;
        v[u] = img[(((size_t)b * Cin + c) * S + gy * P + py) * S + gx * P + px];
      }
    }
    store_planes4(out + row * ldo, plane, k4 * 4, v[0], v[1], v[2], v[3]);
  }
}

// x[b][0] = cls + pos[0];  x[b][1 + i] = patch[b][i] + pos[1 + i]      (VisionTransformer.forward before ln_pre)
__global__ void __launch_bounds__(256)
vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos, int n, int T, int C,
                    float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C >> 2;
  const size_t total = (size_t)n * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const int t = (int)((i / C4) % T);
    const size_t b = i / ((size_t)C4 * T);
    const float4 a = t == 0 ? ((const float4*)cls)[c] : ((const float4*)(patch + (b * (T - 1) + (t - 1)) * C))[c];
    const float4 p = ((const float4*)(pos + (size_t)t * C))[c];
    ((float4*)(out + (b * T + t) * C))[c] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// LayerNorm with fp32 output (ln_pre writes the residual stream; final_layer_norm writes the text embedding): one warp per row,
// exact two-pass statistics in registers like layernorm_kernel (norm.cu).
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, int ldx, int rows, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ out, int ldo) {
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
  float4* orow = (float4*)(out + (size_t)warp * ldo);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + j * 32;
    if (i < C4) {
      const float4 g = ((const float4*)gamma)[i], be = ((const float4*)beta)[i];
      orow[i] = make_float4((v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y,
                            (v[j].z - mean) * rstd * g.z + be.z, (v[j].w - mean) * rstd * g.w + be.w);
    }
  }
}

// out16 = QuickGELU(x) = x * sigmoid(1.702 x) as a GEMM operand (clip/model.py QuickGELU; transformers quick_gelu), [hi | lo] planes
// when plane > 0.
__global__ void __launch_bounds__(256)
quick_gelu_cast_kernel(const float* __restrict__ x, size_t rows, int C, int plane, __half* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C >> 2;
  const size_t total = rows * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / C4;
    const int c = (int)(i % C4);
    const float4 a = ((const float4*)(x + r * C))[c];
    auto qg = [](float t) { return __fdividef(t, 1.0f + __expf(-1.702f * t)); };
    store_planes4(out + r * ldo, plane, c * 4, qg(a.x), qg(a.y), qg(a.z), qg(a.w));
  }
}

// Softmax attention for short sequences in fp32 on the CUDA cores (the text tower: 77 tokens, causal): one CTA per (batch, head),
// K and V of the head staged in shared memory, one warp per query row (lane = key for the scores, lane = channel pair for the output).
// qkv: fp32 [B][N][ld] with q at column h*d, k at koff + h*d, v at voff + h*d. out: fp16 [B][N][ldo], head h at columns h*d,
// [hi | lo] planes when plane > 0.   Limits: N <= 128, d <= 64 and even.
__global__ void __launch_bounds__(256)
attention_small_kernel(const float* __restrict__ qkv, int ld, int koff, int voff, int N, int d, float scale, int causal, int plane,
                       __half* __restrict__ out, int ldo, int H) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* sK = sm;                    // [N][d + 1]
  float* sV = sK + N * (d + 1);      // [N][d]
  float* sQ = sV + N * d;            // [8 warps][d]
  float* sP = sQ + 8 * d;            // [8 warps][128]
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const float* base = qkv + (size_t)b * N * ld + h * d;
  for (int i = threadIdx.x; i < N * d; i += blockDim.x) {
    const int r = i / d, c = i % d;
    sK[r * (d + 1) + c] = base[(size_t)r * ld + koff + c];
    sV[r * d + c] = base[(size_t)r * ld + voff + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = sQ + warp * d;
  float* pr = sP + warp * 128;
  for (int i = warp; i < N; i += 8) {
    for (int c = lane; c < d; c += 32) q[c] = base[(size_t)i * ld + c] * scale;
    __syncwarp();
    const int nk = causal ? i + 1 : N;
    float s[4], m = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      s[u] = -INFINITY;
      if (j < nk) {
        float acc = 0.f;
        const float* kr = sK + j * (d + 1);
        for (int c = 0; c < d; ++c) acc = fmaf(q[c], kr[c], acc);
        s[u] = acc;
      }
      m = fmaxf(m, s[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + 32 * u;
      const float e = j < nk ? expf(s[u] - m) : 0.f;
      if (j < 128) pr[j] = e;
      l += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __syncwarp();
    const float inv = 1.f / l;
    __half* orow = out + ((size_t)b * N + i) * ldo + h * d;
    for (int c = 2 * lane; c < d; c += 64) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < nk; ++j) {
        const float pj = pr[j];
        a0 = fmaf(pj, sV[j * d + c], a0);
        a1 = fmaf(pj, sV[j * d + c + 1], a1);
      }
      a0 *= inv; a1 *= inv;
      __half2 hh = __floats2half2_rn(a0, a1);
      *(__half2*)(orow + c) = hh;
      if (plane > 0) {
        const float2 f = __half22float2(hh);
        *(__half2*)(orow + plane + c) = __floats2half2_rn(a0 - f.x, a1 - f.y);
      }
    }
    __syncwarp();
  }
}

}  // namespace upgpt

using namespace upgpt;

extern "C" int upgpt_embed_tokens(const long long* ids, int rows, int seq, int vocab, const float* tok_emb, const float* pos_emb, int C,
                                  float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(ids && tok_emb && pos_emb && out && rows > 0 && seq > 0 && vocab > 0 && C > 0 && C % 4 == 0, "embed_tokens: bad args (C=%d)", C);
  UPGPT_CHECK_CUDA(launch_k(embed_tokens_kernel, dim3(clip_grid((size_t)rows * (C / 4), 256)), dim3(256), 0, stream, ids, rows, seq, vocab,
                            tok_emb, pos_emb, C, out));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_patchify(const float* img, int n, int Cin, int S, int P, int Kpad, int split3, void* out16, int ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(img && out16 && n > 0 && Cin > 0 && P > 0 && S > 0 && S % P == 0, "patchify: bad geometry");
  UPGPT_REQUIRE(Kpad % 8 == 0 && Kpad >= Cin * P * P, "patchify: Kpad (=%d) must be a multiple of 8 and >= Cin*P*P", Kpad);
  if (ldo <= 0) ldo = split3 ? 2 * Kpad : Kpad;
  UPGPT_REQUIRE(ldo % 8 == 0 && ldo >= (split3 ? 2 : 1) * Kpad, "patchify: bad ldo");
  const size_t total = (size_t)n * (S / P) * (S / P) * (Kpad / 4);
  UPGPT_CHECK_CUDA(launch_k(patchify_kernel, dim3(clip_grid(total, 256)), dim3(256), 0, stream, img, n, Cin, S, P, Kpad, split3 ? Kpad : 0,
                            (__half*)out16, ldo));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_vit_assemble(const float* patch, const float* cls, const float* pos, int n, int T, int C, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(patch && cls && pos && out && n > 0 && T > 1 && C > 0 && C % 4 == 0, "vit_assemble: bad args");
  UPGPT_CHECK_CUDA(launch_k(vit_assemble_kernel, dim3(clip_grid((size_t)n * T * (C / 4), 256)), dim3(256), 0, stream, patch, cls, pos, n, T, C, out));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_layernorm_f32(const float* x, int ldx, int rows, int C, const float* gamma, const float* beta, float eps, float* out,
                                   int ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && out && gamma && beta && rows > 0 && C % 4 == 0 && C > 0 && C <= 2048, "layernorm_f32: bad args (C=%d)", C);
  if (ldx <= 0) ldx = C;
  if (ldo <= 0) ldo = C;
  UPGPT_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "layernorm_f32: ld must be multiples of 4");
  dim3 grid((rows + 7) / 8);
  if (C <= 256) UPGPT_CHECK_CUDA(launch_k(layernorm_f32_kernel<2>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, out, ldo));
  else if (C <= 512) UPGPT_CHECK_CUDA(launch_k(layernorm_f32_kernel<4>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, out, ldo));
  else if (C <= 1024) UPGPT_CHECK_CUDA(launch_k(layernorm_f32_kernel<8>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, out, ldo));
  else UPGPT_CHECK_CUDA(launch_k(layernorm_f32_kernel<16>, grid, dim3(256), 0, stream, x, ldx, rows, C, gamma, beta, eps, out, ldo));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_quick_gelu_cast(const float* x, long long rows, int C, int split3, void* out16, int ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(x && out16 && rows > 0 && C > 0 && C % 4 == 0, "quick_gelu_cast: bad args");
  if (ldo <= 0) ldo = split3 ? 2 * C : C;
  UPGPT_REQUIRE(ldo % 4 == 0 && ldo >= (split3 ? 2 : 1) * C, "quick_gelu_cast: bad ldo");
  UPGPT_CHECK_CUDA(launch_k(quick_gelu_cast_kernel, dim3(clip_grid((size_t)rows * (C / 4), 256)), dim3(256), 0, stream, x, (size_t)rows, C,
                            split3 ? C : 0, (__half*)out16, ldo));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int upgpt_attention_small(const float* qkv, int ld, int koff, int voff, int B, int H, int N, int d, float scale, int causal,
                                     int split3_out, void* out16, int ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  UPGPT_REQUIRE(qkv && out16 && B > 0 && H > 0, "attention_small: bad args");
  UPGPT_REQUIRE(N > 0 && N <= 128 && d > 0 && d <= 64 && d % 2 == 0, "attention_small: needs N <= 128 and even d <= 64 (N=%d d=%d)", N, d);
  const int plane = split3_out ? H * d : 0;
  if (ldo <= 0) ldo = split3_out ? 2 * H * d : H * d;
  UPGPT_REQUIRE(ldo % 2 == 0 && ldo >= (split3_out ? 2 : 1) * H * d, "attention_small: bad ldo");
  const size_t smem = sizeof(float) * ((size_t)N * (d + 1) + (size_t)N * d + 8 * d + 8 * 128);
  static bool attr_set = false;
  if (!attr_set) {
    UPGPT_CHECK_CUDA(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  UPGPT_REQUIRE(smem <= 96 * 1024, "attention_small: smem %zu too large", smem);
  UPGPT_CHECK_CUDA(launch_k(attention_small_kernel, dim3((unsigned)(B * H)), dim3(256), smem, stream, qkv, ld, koff, voff, N, d, scale, causal,
                            split3_out ? plane : 0, (__half*)out16, ldo, H));
  count_launch();
  UPGPT_CHECK_CUDA(cudaGetLastError());
  return 0;
}

UPGPT_TRACE_TU(clip)
