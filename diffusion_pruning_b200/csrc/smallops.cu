// K9: the two small fp32 operators of the pruning train step that round 1 left on cuBLAS / eager PyTorch:
//   * HyperStructure: y = x W^T + b over the 71 row-concatenated Linears ([B,768] x [768,1620]; hypernet.py:72-79), forward
//     and backward (dW, db, dx). fp32 SIMT on purpose: the architecture logits feed the Gumbel gates and the router, whose
//     assignments must be bit-stable (bf16 tensor-core products would move them); every sum runs in a fixed order.
//   * ContrastiveLoss (pdm/losses/contrastive_loss.py:11-22): row-normalise arch vectors and prompt embeddings of the
//     all-gathered batch, similarity / temperature, row softmax, BCE(softmax_arch^T, softmax_prompt^T) -- fused into a
//     norm kernel, a tiled Gram kernel and one row kernel (softmax + loss), plus the matching backward.
// Sizes are tiny (M = 32..256 rows): these kernels are latency-bound; the point is one deterministic code path with no
// library dependency, not FLOP/s.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

// ---------------------------------------------------------------------------------------------------------------
// fp32 linear
// ---------------------------------------------------------------------------------------------------------------
// one warp per output column n: the weight row is walked once per 4 batch rows (register tile over b)
__global__ void __launch_bounds__(256) linear_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             int B, int K, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const float* wr = w + (size_t)n * K;
  const float bn = bias ? bias[n] : 0.f;
  for (int b0 = blockIdx.y * 4; b0 < B; b0 += gridDim.y * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (b0 + u < B) acc[u] = fmaf(wv, __ldg(x + (size_t)(b0 + u) * K + k), acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && b0 + u < B) y[(size_t)(b0 + u) * N + n] = v + bn;
    }
  }
}

// dW[n,k] = sum_b dy[b,n] x[b,k] (thread per (n,k)); db[n] = sum_b dy[b,n] (first k-block of every n)
__global__ void __launch_bounds__(256) linear_f32_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, float* __restrict__ db, int B, int K,
                                                               int N) {
  const int n = blockIdx.x;
  const int k = blockIdx.y * 256 + threadIdx.x;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = __ldg(dy + (size_t)b * N + n);
    accb += g;
    if (k < K) acc = fmaf(g, __ldg(x + (size_t)b * K + k), acc);
  }
  if (k < K) dw[(size_t)n * K + k] = acc;
  if (db && blockIdx.y == 0 && threadIdx.x == 0) db[n] = accb;
}

// dx[b,k] = sum_n dy[b,n] w[n,k]
__global__ void __launch_bounds__(256) linear_f32_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int B, int K, int N) {
  const int b = blockIdx.x;
  const int k = blockIdx.y * 256 + threadIdx.x;
  if (k >= K) return;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(__ldg(dy + (size_t)b * N + n), __ldg(w + (size_t)n * K + k), acc);
  dx[(size_t)b * K + k] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// contrastive loss
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ a, int M, int D, float* __restrict__ inv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= M) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = __ldg(a + (size_t)i * D + d);
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) inv[i] = rsqrtf(s);
}

// G[i,j] = (a_i . a_j) * inv[i] * inv[j] * inv_temp, 16 x 16 output tile per block, K chunks of 32 through shared memory
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ a, const float* __restrict__ inv, int M, int D,
                                                   float inv_temp, float* __restrict__ G) {
  __shared__ float sa[16][33], sb[16][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  float acc = 0.f;
  for (int k0 = 0; k0 < D; k0 += 32) {
    for (int t = threadIdx.x; t < 16 * 32; t += 256) {
      const int r = t >> 5, c = t & 31;
      sa[r][c] = (i0 + r < M && k0 + c < D) ? __ldg(a + (size_t)(i0 + r) * D + k0 + c) : 0.f;
      sb[r][c] = (j0 + r < M && k0 + c < D) ? __ldg(a + (size_t)(j0 + r) * D + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 32; ++c) acc = fmaf(sa[ty][c], sb[tx][c], acc);
    __syncthreads();
  }
  const int i = i0 + ty, j = j0 + tx;
  if (i < M && j < M) G[(size_t)i * M + j] = acc * inv[i] * inv[j] * inv_temp;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];  // fixed order
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  __syncthreads();
  return t;
}

// one block per row i: Sa[i,:] = softmax(Ga[i,:]), Sp[i,:] = softmax(Gp[i,:]) in place, row_loss[i] = sum_j bce(Sa, Sp)
__global__ void __launch_bounds__(128) contrastive_row_kernel(float* __restrict__ Ga, float* __restrict__ Gp, int M,
                                                              float* __restrict__ row_loss) {
  __shared__ float red[4];
  const int i = blockIdx.x;
  float* ga = Ga + (size_t)i * M;
  float* gp = Gp + (size_t)i * M;
  float ma = -INFINITY, mp = -INFINITY;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    ma = fmaxf(ma, ga[j]);
    mp = fmaxf(mp, gp[j]);
  }
  ma = block_max(ma, red);
  mp = block_max(mp, red);
  float sa = 0.f, sp = 0.f;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    sa += expf(ga[j] - ma);
    sp += expf(gp[j] - mp);
  }
  sa = block_sum(sa, red);
  sp = block_sum(sp, red);
  float l = 0.f;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const float s = expf(ga[j] - ma) / sa, t = expf(gp[j] - mp) / sp;
    ga[j] = s;
    gp[j] = t;
    // F.binary_cross_entropy clamps both logs at -100
    l -= t * fmaxf(logf(s), -100.f) + (1.f - t) * fmaxf(log1pf(-s), -100.f);
  }
  l = block_sum(l, red);
  if (threadIdx.x == 0) row_loss[i] = l;
}

__global__ void __launch_bounds__(256) sum_scale_kernel(const float* __restrict__ v, int n, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// backward, one block per row i: dG[i,:] = Sa (dSa - sum_k Sa dSa), dSa = g (Sa - Sp) / max(Sa (1 - Sa), 1e-12) / M^2
__global__ void __launch_bounds__(128) contrastive_row_bwd_kernel(const float* __restrict__ Sa, const float* __restrict__ Sp,
                                                                  int M, const float* __restrict__ gout,
                                                                  float* __restrict__ dG) {
  __shared__ float red[4];
  const int i = blockIdx.x;
  const float g = gout[0] / ((float)M * (float)M);
  float dot = 0.f;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const float s = Sa[(size_t)i * M + j], t = Sp[(size_t)i * M + j];
    dot += s * (g * (s - t) / fmaxf(s * (1.f - s), 1e-12f));
  }
  dot = block_sum(dot, red);
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const float s = Sa[(size_t)i * M + j], t = Sp[(size_t)i * M + j];
    dG[(size_t)i * M + j] = s * (g * (s - t) / fmaxf(s * (1.f - s), 1e-12f) - dot);
  }
}

// dAhat[i,d] = inv_temp * sum_j (dG[i,j] + dG[j,i]) * a[j,d] * inv[j]     (thread per (i, d))
__global__ void __launch_bounds__(256) contrastive_dhat_kernel(const float* __restrict__ dG, const float* __restrict__ a,
                                                               const float* __restrict__ inv, int M, int D, float inv_temp,
                                                               float* __restrict__ dhat) {
  const int i = blockIdx.x;
  const int d = blockIdx.y * 256 + threadIdx.x;
  if (d >= D) return;
  float acc = 0.f;
  for (int j = 0; j < M; ++j)
    acc = fmaf((dG[(size_t)i * M + j] + dG[(size_t)j * M + i]) * inv[j], __ldg(a + (size_t)j * D + d), acc);
  dhat[(size_t)i * D + d] = acc * inv_temp;
}

// da[i,:] = inv[i] * (dhat[i,:] - ahat[i,:] * (ahat[i,:] . dhat[i,:])),  ahat = a * inv[i]
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ a, const float* __restrict__ inv,
                                                            const float* __restrict__ dhat, int M, int D,
                                                            float* __restrict__ da) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float iv = inv[i];
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) dot = fmaf(a[(size_t)i * D + d] * iv, dhat[(size_t)i * D + d], dot);
  dot = block_sum(dot, red);
  for (int d = threadIdx.x; d < D; d += 256)
    da[(size_t)i * D + d] = iv * (dhat[(size_t)i * D + d] - a[(size_t)i * D + d] * iv * dot);
}

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_linear_f32_fwd(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t K,
                                   int32_t N, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && w && y && B >= 0 && K > 0 && N > 0, "aptp_linear_f32_fwd: bad arguments");
  if (B == 0) return APTP_OK;
  int gy = (B + 3) / 4;
  if (gy > 64) gy = 64;
  linear_f32_fwd_kernel<<<dim3((N + 7) / 8, gy), 256, 0, stream>>>(x, w, bias, y, B, K, N);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_linear_f32_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                                   int32_t B, int32_t K, int32_t N, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && w && dy && B > 0 && K > 0 && N > 0, "aptp_linear_f32_bwd: bad arguments");
  if (dw) linear_f32_wgrad_kernel<<<dim3(N, (K + 255) / 256), 256, 0, stream>>>(x, dy, dw, db, B, K, N);
  if (dx) linear_f32_dgrad_kernel<<<dim3(B, (K + 255) / 256), 256, 0, stream>>>(w, dy, dx, B, K, N);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_contrastive_fwd(const float* arch, int32_t Da, const float* prompt, int32_t Dp, int32_t M,
                                    float arch_temp, float prompt_temp, float* inv_a, float* inv_p, float* Sa, float* Sp,
                                    float* row_loss, float* loss, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(arch && prompt && inv_a && inv_p && Sa && Sp && row_loss && loss && M > 0 && Da > 0 && Dp > 0,
               "aptp_contrastive_fwd: bad arguments");
  APTP_REQUIRE(arch_temp > 0.f && prompt_temp > 0.f, "aptp_contrastive_fwd: temperatures must be positive");
  row_inv_norm_kernel<<<(M + 7) / 8, 256, 0, stream>>>(arch, M, Da, inv_a);
  row_inv_norm_kernel<<<(M + 7) / 8, 256, 0, stream>>>(prompt, M, Dp, inv_p);
  dim3 g((M + 15) / 16, (M + 15) / 16);
  gram_kernel<<<g, 256, 0, stream>>>(arch, inv_a, M, Da, 1.f / arch_temp, Sa);
  gram_kernel<<<g, 256, 0, stream>>>(prompt, inv_p, M, Dp, 1.f / prompt_temp, Sp);
  contrastive_row_kernel<<<M, 128, 0, stream>>>(Sa, Sp, M, row_loss);
  sum_scale_kernel<<<1, 256, 0, stream>>>(row_loss, M, 1.f / ((float)M * (float)M), loss);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_contrastive_bwd(const float* arch, int32_t Da, int32_t M, float arch_temp, const float* inv_a,
                                    const float* Sa, const float* Sp, const float* grad_loss, float* dG, float* dhat,
                                    float* darch, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(arch && inv_a && Sa && Sp && grad_loss && dG && dhat && darch && M > 0 && Da > 0,
               "aptp_contrastive_bwd: bad arguments");
  contrastive_row_bwd_kernel<<<M, 128, 0, stream>>>(Sa, Sp, M, grad_loss, dG);
  contrastive_dhat_kernel<<<dim3(M, (Da + 255) / 256), 256, 0, stream>>>(dG, arch, inv_a, M, Da, 1.f / arch_temp, dhat);
  normalize_bwd_kernel<<<M, 256, 0, stream>>>(arch, inv_a, dhat, M, Da, darch);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
