// K2 (continued): gate-scaled residual / data-movement kernels, all 16-byte vectorised and coalesced.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

__device__ __forceinline__ void unpack8e(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// out = (1-d[b]) * x + d[b] * y
__global__ void __launch_bounds__(256) depth_lerp_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                         const __nv_bfloat16* __restrict__ y, int ldy,
                                                         __nv_bfloat16* __restrict__ out, int ldo, long long rows,
                                                         int cv, const float* __restrict__ d, int rows_per_sample) {
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int v = (int)(i - row * cv);
    const float dd = d[row / rows_per_sample];
    float a[8], b[8];
    unpack8e(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + v * 8)), a);
    unpack8e(__ldg(reinterpret_cast<const uint4*>(y + row * ldy + v * 8)), b);
    uint4 o;
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = (1.f - dd) * a[e] + dd * b[e];
    o.x = pack_bf16(r[0], r[1]);
    o.y = pack_bf16(r[2], r[3]);
    o.z = pack_bf16(r[4], r[5]);
    o.w = pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(out + row * ldo + v * 8) = o;
  }
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const __nv_bfloat16* __restrict__ src, int lds,
                                                        __nv_bfloat16* __restrict__ dst, int ldd, long long rows,
                                                        int cv, const uint8_t* __restrict__ mask,
                                                        int rows_per_sample) {
  const long long total = rows * cv;
  const long long T = (long long)gridDim.x * blockDim.x;
  // 4 independent 16-byte loads in flight per thread before the first store
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * T) {
    uint4 val[4];
    long long off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * T;
      off[u] = -1;
      if (i < total) {
        const long long row = i / cv;
        const int v = (int)(i - row * cv);
        if (!mask || mask[row / rows_per_sample]) {
          val[u] = __ldg(reinterpret_cast<const uint4*>(src + row * lds + v * 8));
          off[u] = row * ldd + v * 8;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (off[u] >= 0) *reinterpret_cast<uint4*>(dst + off[u]) = val[u];
  }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ src,
                                                         __nv_bfloat16* __restrict__ dst, int batch, int H, int W,
                                                         int cv) {
  const long long total = (long long)batch * (2 * H) * (2 * W) * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long pix = i / cv;
    const int ox = (int)(pix % (2 * W));
    pix /= (2 * W);
    const int oy = (int)(pix % (2 * H));
    const int b = (int)(pix / (2 * H));
    const long long sp = ((long long)b * H + (oy >> 1)) * W + (ox >> 1);
    reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + sp * cv + v);
  }
}

// ---- fp32 residual-stream variants (round 2: the stream between blocks is fp32) ----
// 8 channels per work item, read as bf16 (16 B) or fp32 (32 B), written as bf16 or fp32.
template <bool kF32>
__device__ __forceinline__ void load8(const void* base, long long elem_off, float* f) {
  if constexpr (kF32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8e(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off)), f);
  }
}
template <bool kF32>
__device__ __forceinline__ void store8(void* base, long long elem_off, const float* f) {
  if constexpr (kF32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(f[0], f[1], f[2], f[3]);
    p[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = o;
  }
}

template <bool kSrcF32, bool kDstF32>
__global__ void __launch_bounds__(256) copy_rows_cvt_kernel(const void* __restrict__ src, int lds, void* __restrict__ dst,
                                                            int ldd, long long rows, int cv,
                                                            const uint8_t* __restrict__ mask, int rows_per_sample) {
  const long long total = rows * cv;
  const long long T = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * T) {
    float val[2][8];
    long long off[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * T;
      off[u] = -1;
      if (i < total) {
        const long long row = i / cv;
        const int v = (int)(i - row * cv);
        if (!mask || mask[row / rows_per_sample]) {
          load8<kSrcF32>(src, row * lds + v * 8, val[u]);
          off[u] = row * ldd + v * 8;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (off[u] >= 0) store8<kDstF32>(dst, off[u], val[u]);
  }
}

// out = (1-d[b]) * x + d[b] * y, all fp32 rows (DepthGate on the fp32 stream)
__global__ void __launch_bounds__(256) depth_lerp_f32_kernel(const float* x, int ldx, const float* y, int ldy,
                                                             float* out, int ldo, long long rows, int cv,
                                                             const float* __restrict__ d, int rows_per_sample) {
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int v = (int)(i - row * cv);
    const float dd = d[row / rows_per_sample];
    // plain (coherent) loads: `out` may alias `y`
    const float4* xp = reinterpret_cast<const float4*>(x + row * ldx + v * 8);
    const float4* yp = reinterpret_cast<const float4*>(y + row * ldy + v * 8);
    const float4 a0 = xp[0], a1 = xp[1], b0 = yp[0], b1 = yp[1];
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = (1.f - dd) * a[e] + dd * b[e];
    store8<true>(out, row * ldo + v * 8, r);
  }
}

template <bool kSrcF32>
__global__ void __launch_bounds__(256) upsample2x_cvt_kernel(const void* __restrict__ src,
                                                             __nv_bfloat16* __restrict__ dst, int batch, int H, int W,
                                                             int cv) {
  const long long total = (long long)batch * (2 * H) * (2 * W) * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long pix = i / cv;
    const int ox = (int)(pix % (2 * W));
    pix /= (2 * W);
    const int oy = (int)(pix % (2 * H));
    const int b = (int)(pix / (2 * H));
    const long long sp = ((long long)b * H + (oy >> 1)) * W + (ox >> 1);
    float f[8];
    load8<kSrcF32>(src, (sp * cv + v) * 8, f);
    store8<false>(dst, i * 8, f);
  }
}

// conv_in as a GEMM: rows = pixels, K = (tap, cin) zero-padded to 64.
__global__ void __launch_bounds__(256) im2col_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ dst,
                                                           int batch, int Cin, int H, int W) {
  const long long total = (long long)batch * H * W * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i & 63);
    long long pix = i >> 6;
    const int xx = (int)(pix % W);
    pix /= W;
    const int yy = (int)(pix % H);
    const int b = (int)(pix / H);
    float val = 0.f;
    if (k < 9 * Cin) {
      const int tap = k / Cin, c = k - tap * Cin;
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) val = x[(((long long)b * Cin + c) * H + sy) * W + sx];
    }
    dst[i] = __float2bfloat16(val);
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ dst, int batch,
                                          int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i - b * dim;
  const int k = (j < half) ? j : j - half;
  const float freq = expf(-logf(10000.f) * (float)k / (float)half);
  const float arg = t[b] * freq;
  // flip_sin_to_cos=True: first half cos, second half sin
  dst[i] = __float2bfloat16(j < half ? cosf(arg) : sinf(arg));
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// n % 8 == 0 and 16-byte aligned pointers: 32-byte loads, 16-byte stores
__global__ void __launch_bounds__(256) cast_f32_bf16_vec_kernel(const float* __restrict__ src,
                                                                __nv_bfloat16* __restrict__ dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    load8<true>(src, i * 8, f);
    store8<false>(dst, i * 8, f);
  }
}

__global__ void __launch_bounds__(256) silu_bf16_kernel(const __nv_bfloat16* __restrict__ src,
                                                        __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(silu_f(__bfloat162float(src[i])));
}

// Classifier-free-guidance combine + one DDIM step (eta = 0, v-prediction or epsilon), fused:
//   v = v_u + g (v_c - v_u)                                  pruning_pipelines.py:805-807
//   x0 = sa x - sb v ; eps = sa v + sb x   (v-prediction)     diffusers DDIMScheduler.step
//   x_prev = sa_prev x0 + sb_prev eps
// pred holds [uncond batch | cond batch] (pruning_pipelines.py:765, :792), all fp32 NCHW.
__global__ void __launch_bounds__(256) cfg_ddim_kernel(const float* __restrict__ pred, const float* __restrict__ x,
                                                       float* __restrict__ x_out, long long n, float guidance, float sa,
                                                       float sb, float sa_prev, float sb_prev, int v_prediction) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pu = pred[i], pc = pred[n + i];
    const float m = pu + guidance * (pc - pu);
    const float xv = x[i];
    float x0, eps;
    if (v_prediction) {
      x0 = sa * xv - sb * m;
      eps = sa * m + sb * xv;
    } else {
      eps = m;
      x0 = (xv - sb * m) / sa;
    }
    x_out[i] = sa_prev * x0 + sb_prev * eps;
  }
}

static unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_depth_lerp(const void* x, int32_t ldx, const void* y, int32_t ldy, void* out, int32_t ldo,
                               int64_t rows, int32_t C, const float* d, int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && y && out && d, "aptp_depth_lerp: null pointer");
  APTP_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldo % 8 == 0 && rows_per_sample > 0,
               "aptp_depth_lerp: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  depth_lerp_kernel<<<grid_for(rows * (C / 8)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(y), ldy,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, C / 8, d, rows_per_sample);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_copy_rows(const void* src, int32_t lds, void* dst, int32_t ldd, int64_t rows, int32_t C,
                              const uint8_t* sample_mask, int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst, "aptp_copy_rows: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && rows_per_sample > 0,
               "aptp_copy_rows: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  copy_rows_kernel<<<grid_for(rows * (C / 8)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds,
                                                                reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows,
                                                                C / 8, sample_mask, rows_per_sample);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_copy_rows_cvt(const void* src, int32_t src_f32, int32_t lds, void* dst, int32_t dst_f32,
                                  int32_t ldd, int64_t rows, int32_t C, const uint8_t* sample_mask,
                                  int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst, "aptp_copy_rows_cvt: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && rows_per_sample > 0,
               "aptp_copy_rows_cvt: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  const unsigned grid = grid_for(rows * (C / 8));
#define APTP_CVT(SF, DF) \
  copy_rows_cvt_kernel<SF, DF><<<grid, 256, 0, stream>>>(src, lds, dst, ldd, rows, C / 8, sample_mask, rows_per_sample)
  if (src_f32 && dst_f32) APTP_CVT(true, true);
  else if (src_f32) APTP_CVT(true, false);
  else if (dst_f32) APTP_CVT(false, true);
  else APTP_CVT(false, false);
#undef APTP_CVT
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_depth_lerp_f32(const float* x, int32_t ldx, const float* y, int32_t ldy, float* out, int32_t ldo,
                                   int64_t rows, int32_t C, const float* d, int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && y && out && d, "aptp_depth_lerp_f32: null pointer");
  APTP_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldo % 8 == 0 && rows_per_sample > 0,
               "aptp_depth_lerp_f32: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  depth_lerp_f32_kernel<<<grid_for(rows * (C / 8)), 256, 0, stream>>>(x, ldx, y, ldy, out, ldo, rows, C / 8, d,
                                                                     rows_per_sample);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_upsample2x_cvt(const void* src, int32_t src_f32, void* dst, int32_t batch, int32_t H, int32_t W,
                                   int32_t C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst && C % 8 == 0, "aptp_upsample2x_cvt: bad arguments");
  const long long total = (long long)batch * 4 * H * W * (C / 8);
  if (total == 0) return APTP_OK;
  if (src_f32)
    upsample2x_cvt_kernel<true><<<grid_for(total), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), batch,
                                                                    H, W, C / 8);
  else
    upsample2x_cvt_kernel<false><<<grid_for(total), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), batch,
                                                                     H, W, C / 8);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_upsample2x(const void* src, void* dst, int32_t batch, int32_t H, int32_t W, int32_t C,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst && C % 8 == 0, "aptp_upsample2x: bad arguments");
  const long long total = (long long)batch * 4 * H * W * (C / 8);
  if (total == 0) return APTP_OK;
  upsample2x_kernel<<<grid_for(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                        reinterpret_cast<__nv_bfloat16*>(dst), batch, H, W, C / 8);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_im2col_input(const float* sample_nchw, void* dst, int32_t batch, int32_t Cin, int32_t H, int32_t W,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(sample_nchw && dst && 9 * Cin <= 64, "aptp_im2col_input: needs 9*Cin <= 64");
  const long long total = (long long)batch * H * W * 64;
  if (total == 0) return APTP_OK;
  im2col_input_kernel<<<grid_for(total), 256, 0, stream>>>(sample_nchw, reinterpret_cast<__nv_bfloat16*>(dst), batch,
                                                          Cin, H, W);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_timestep_embedding(const float* t, void* dst, int32_t batch, int32_t dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(t && dst && dim % 2 == 0, "aptp_timestep_embedding: bad arguments");
  if (batch == 0) return APTP_OK;
  timestep_embedding_kernel<<<(batch * dim + 255) / 256, 256, 0, stream>>>(t, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                          batch, dim);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst, "aptp_cast_f32_bf16: null pointer");
  if (n == 0) return APTP_OK;
  if (n % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
    cast_f32_bf16_vec_kernel<<<grid_for(n / 8), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n / 8);
  else
    cast_f32_bf16_kernel<<<grid_for(n), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_silu_bf16(const void* src, void* dst, int64_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst, "aptp_silu_bf16: null pointer");
  if (n == 0) return APTP_OK;
  silu_bf16_kernel<<<grid_for(n), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                   reinterpret_cast<__nv_bfloat16*>(dst), n);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_cfg_ddim_step(const float* pred, const float* x, float* x_out, int64_t n, float guidance,
                                  float alpha_t, float alpha_prev, int32_t v_prediction, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(pred && x && x_out, "aptp_cfg_ddim_step: null pointer");
  APTP_REQUIRE(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f,
               "aptp_cfg_ddim_step: alphas_cumprod must be in (0, 1]");
  if (n == 0) return APTP_OK;
  cfg_ddim_kernel<<<grid_for(n), 256, 0, stream>>>(pred, x, x_out, n, guidance, sqrtf(alpha_t), sqrtf(1.f - alpha_t),
                                                  sqrtf(alpha_prev), sqrtf(1.f - alpha_prev), v_prediction);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
