// Loss front/back end of the pruning train step: the elementwise work either side of the two
// U-Net forwards in Pruner.step (pdm/training/trainer.py:1121-1123, :1181, :1197-1225), HBM-bound, one pass each.
//   aptp_add_noise_velocity   DDIMScheduler.add_noise + get_velocity in one read of (latents, noise)
//   aptp_mse_rows_fwd / _bwd  block-distillation MSE between student and teacher block outputs (bf16 NHWC rows,
//                             the layout the engine produces) and its gradient, written as bf16 NHWC rows again
//   aptp_pred_losses_fwd/_bwd min-SNR-weighted DDPM MSE + distillation MSE on the fp32 predictions
// Reductions are deterministic: every CTA writes one fp64 partial, the caller sums them.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

__device__ __forceinline__ void unpack8l(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// CTA-wide sum of one fp64 per thread (256 threads); result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i];
  }
  __syncthreads();
  return t;
}

// noisy = sqrt(acp[t]) x0 + sqrt(1-acp[t]) n ; target = sqrt(acp[t]) n - sqrt(1-acp[t]) x0 (v-prediction) or n
__global__ void __launch_bounds__(256) add_noise_velocity_kernel(const float4* __restrict__ x0, const float4* __restrict__ nz,
                                                                 const long long* __restrict__ timesteps,
                                                                 const float* __restrict__ sqrt_acp,
                                                                 const float* __restrict__ sqrt_1m_acp,
                                                                 float4* __restrict__ noisy, float4* __restrict__ target,
                                                                 long long total4, int per_sample4, int v_prediction) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = timesteps[i / per_sample4];
    const float sa = sqrt_acp[t], sb = sqrt_1m_acp[t];
    const float4 a = __ldg(x0 + i), n = __ldg(nz + i);
    noisy[i] = make_float4(sa * a.x + sb * n.x, sa * a.y + sb * n.y, sa * a.z + sb * n.z, sa * a.w + sb * n.w);
    if (target) {
      target[i] = v_prediction ? make_float4(sa * n.x - sb * a.x, sa * n.y - sb * a.y, sa * n.z - sb * a.z,
                                             sa * n.w - sb * a.w)
                               : n;
    }
  }
}

// partial[cta] = sum over this CTA's share of (a - b)^2, fp32 differences accumulated in fp64
__global__ void __launch_bounds__(256) mse_rows_fwd_kernel(const __nv_bfloat16* __restrict__ a, int lda,
                                                           const __nv_bfloat16* __restrict__ b, int ldb, long long rows,
                                                           int cv, double* __restrict__ partial) {
  __shared__ double sh[8];
  const long long total = rows * cv;
  const long long T = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * T) {
    uint4 va[2], vb[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * T;
      on[u] = i < total;
      if (on[u]) {
        const long long row = i / cv;
        const int v = (int)(i - row * cv);
        va[u] = __ldg(reinterpret_cast<const uint4*>(a + row * lda + v * 8));
        vb[u] = __ldg(reinterpret_cast<const uint4*>(b + row * ldb + v * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!on[u]) continue;
      float fa[8], fb[8];
      unpack8l(va[u], fa);
      unpack8l(vb[u], fb);
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = fa[e] - fb[e];
        s = fmaf(d, d, s);
      }
      acc += (double)s;
    }
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// da = coef[0] * scale * (a - b), bf16
__global__ void __launch_bounds__(256) mse_rows_bwd_kernel(const __nv_bfloat16* __restrict__ a, int lda,
                                                           const __nv_bfloat16* __restrict__ b, int ldb,
                                                           __nv_bfloat16* __restrict__ da, int ldda, long long rows, int cv,
                                                           const float* __restrict__ coef, float scale) {
  const long long total = rows * cv;
  const float c = coef[0] * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int v = (int)(i - row * cv);
    float fa[8], fb[8];
    unpack8l(__ldg(reinterpret_cast<const uint4*>(a + row * lda + v * 8)), fa);
    unpack8l(__ldg(reinterpret_cast<const uint4*>(b + row * ldb + v * 8)), fb);
    uint4 o;
    o.x = pack_bf16(c * (fa[0] - fb[0]), c * (fa[1] - fb[1]));
    o.y = pack_bf16(c * (fa[2] - fb[2]), c * (fa[3] - fb[3]));
    o.z = pack_bf16(c * (fa[4] - fb[4]), c * (fa[5] - fb[5]));
    o.w = pack_bf16(c * (fa[6] - fb[6]), c * (fa[7] - fb[7]));
    *reinterpret_cast<uint4*>(da + row * ldda + v * 8) = o;
  }
}

// One CTA per (sample, chunk): partial[(b * chunks + c) * 2 + {0,1}] = sum (pred-target)^2 , sum (pred-teacher)^2
__global__ void __launch_bounds__(256) pred_losses_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                              const float* __restrict__ teacher, int per_sample, int chunks,
                                                              double* __restrict__ partial) {
  __shared__ double sh[8];
  const int b = blockIdx.x / chunks, c = blockIdx.x % chunks;
  const long long base = (long long)b * per_sample;
  double a0 = 0.0, a1 = 0.0;
  for (int i = c * 256 + threadIdx.x; i < per_sample; i += chunks * 256) {
    const float pv = pred[base + i];
    const float d0 = pv - target[base + i];
    const float d1 = pv - teacher[base + i];
    a0 += (double)(d0 * d0);
    a1 += (double)(d1 * d1);
  }
  const double t0 = block_sum_256(a0, sh);
  const double t1 = block_sum_256(a1, sh);
  if (threadIdx.x == 0) {
    partial[(size_t)blockIdx.x * 2] = t0;
    partial[(size_t)blockIdx.x * 2 + 1] = t1;
  }
}

// dpred = g_ddpm * w[b] * 2 (pred-target) / (per_sample * batch) + g_distill * 2 (pred-teacher) / (per_sample * batch)
__global__ void __launch_bounds__(256) pred_losses_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                              const float* __restrict__ teacher,
                                                              const float* __restrict__ weight, const float* __restrict__ g,
                                                              float* __restrict__ dpred, long long total, int per_sample,
                                                              float inv_n) {
  const float g0 = g[0] * 2.f * inv_n, g1 = g[1] * 2.f * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float pv = pred[i];
    const float w = weight ? weight[i / per_sample] : 1.f;
    dpred[i] = g0 * w * (pv - target[i]) + g1 * (pv - teacher[i]);
  }
}

static unsigned loss_grid(long long work_items) {
  long long g = (work_items + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_add_noise_velocity(const float* latents, const float* noise, const int64_t* timesteps,
                                       const float* sqrt_acp, const float* sqrt_1m_acp, float* noisy, float* target,
                                       int32_t batch, int32_t per_sample, int32_t v_prediction, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(latents && noise && timesteps && sqrt_acp && sqrt_1m_acp && noisy, "aptp_add_noise_velocity: null pointer");
  APTP_REQUIRE(per_sample > 0 && per_sample % 4 == 0, "aptp_add_noise_velocity: per-sample size must be a multiple of 4");
  if (batch == 0) return APTP_OK;
  const long long total4 = (long long)batch * (per_sample / 4);
  add_noise_velocity_kernel<<<loss_grid(total4), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(latents), reinterpret_cast<const float4*>(noise),
      reinterpret_cast<const long long*>(timesteps), sqrt_acp, sqrt_1m_acp, reinterpret_cast<float4*>(noisy),
      reinterpret_cast<float4*>(target), total4, per_sample / 4, v_prediction);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_mse_rows_fwd(const void* a, int32_t lda, const void* b, int32_t ldb, int64_t rows, int32_t C,
                                 double* partial, int32_t n_partial, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(a && b && partial && n_partial > 0, "aptp_mse_rows_fwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "aptp_mse_rows_fwd: C and pitches must be multiples of 8");
  // exactly n_partial CTAs: every slot of `partial` is written (0 for a CTA without work)
  mse_rows_fwd_kernel<<<(unsigned)n_partial, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a), lda,
                                                              reinterpret_cast<const __nv_bfloat16*>(b), ldb, rows, C / 8,
                                                              partial);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_mse_rows_bwd(const void* a, int32_t lda, const void* b, int32_t ldb, void* da, int32_t ldda,
                                 int64_t rows, int32_t C, const float* coef, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(a && b && da && coef, "aptp_mse_rows_bwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldda % 8 == 0,
               "aptp_mse_rows_bwd: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  mse_rows_bwd_kernel<<<loss_grid(rows * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), lda, reinterpret_cast<const __nv_bfloat16*>(b), ldb,
      reinterpret_cast<__nv_bfloat16*>(da), ldda, rows, C / 8, coef, scale);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_pred_losses_fwd(const float* pred, const float* target, const float* teacher, int32_t batch,
                                    int32_t per_sample, int32_t chunks, double* partial, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(pred && target && teacher && partial, "aptp_pred_losses_fwd: null pointer");
  APTP_REQUIRE(per_sample > 0 && chunks > 0, "aptp_pred_losses_fwd: bad sizes");
  if (batch == 0) return APTP_OK;
  pred_losses_fwd_kernel<<<(unsigned)(batch * chunks), 256, 0, stream>>>(pred, target, teacher, per_sample, chunks, partial);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_pred_losses_bwd(const float* pred, const float* target, const float* teacher, const float* weight,
                                    const float* g, float* dpred, int32_t batch, int32_t per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(pred && target && teacher && g && dpred, "aptp_pred_losses_bwd: null pointer");
  APTP_REQUIRE(per_sample > 0, "aptp_pred_losses_bwd: bad sizes");
  if (batch == 0) return APTP_OK;
  const long long total = (long long)batch * per_sample;
  pred_losses_bwd_kernel<<<loss_grid(total), 256, 0, stream>>>(pred, target, teacher, weight, g, dpred, total, per_sample,
                                                              1.f / (float)total);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
