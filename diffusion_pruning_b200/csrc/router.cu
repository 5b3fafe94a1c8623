// K4: the router -- fused Gumbel-sigmoid gate, width/depth normalisation + L2 norm, cosine scores
// with argmax, and Sinkhorn optimal transport with fp64 marginals. All HBM/latency-bound: one pass
// over [B, 1620] per stage, warp-shuffle + shared-memory block reductions, no host synchronisation.
// (The reference issues ~71 x (CPU rand + H2D + ~6 kernels + .all() host sync) per call:
//  pdm/models/vq/quantizer.py:196-215, pdm/utils/estimation_utils.py:5-64.)
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

__device__ __forceinline__ float gumbel_from_uniform(float u) {
  // sample_gumbel: -log(-log(u + eps) + eps), eps = 1e-20 (estimation_utils.py:5-10)
  return -logf(-logf(u + 1e-20f) + 1e-20f);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int GATE_THREADS = 128;
constexpr int MAX_DEPTH = 64;

__global__ void __launch_bounds__(GATE_THREADS)
    gumbel_gate_kernel(const float* __restrict__ z, const float* __restrict__ u, float* __restrict__ out, int n_width,
                       int n_depth, const int* __restrict__ width_starts, int n_gates,
                       const int* __restrict__ depth_order, float inv_t, float base, int non_zero_width) {
  extern __shared__ int sm_i[];
  int* col_gate = sm_i;              // [n_width]
  int* gate_on = sm_i + n_width;     // [n_gates]
  const int row = blockIdx.x;
  const int dim = n_width + n_depth;
  const float* zr = z + (size_t)row * dim;
  const float* ur = u + (size_t)row * dim;
  float* orow = out + (size_t)row * dim;

  for (int g = threadIdx.x; g < n_gates; g += GATE_THREADS) {
    gate_on[g] = 0;
    for (int c = width_starts[g]; c < width_starts[g + 1]; ++c) col_gate[c] = g;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_width; c += GATE_THREADS) {
    const float y = sigmoid_f((zr[c] + gumbel_from_uniform(ur[c]) + base) * inv_t);
    orow[c] = y;
    if (y >= 0.5f) gate_on[col_gate[c]] = 1;  // benign race: every writer stores 1
  }
  if (threadIdx.x == 0 && n_depth > 0) {
    // importance_gumbel_softmax_sample (estimation_utils.py:49-64), sequential like torch.cumsum
    float e[MAX_DEPTH];
    float mx = -INFINITY;
    for (int j = 0; j < n_depth; ++j) mx = fmaxf(mx, zr[n_width + j]);
    float sum = 0.f;
    for (int j = 0; j < n_depth; ++j) {
      e[j] = expf(zr[n_width + j] - mx);
      sum += e[j];
    }
    float run = 0.f;
    for (int j = 0; j < n_depth; ++j) {
      run += e[j] / sum;
      e[j] = run;  // cumsum
    }
    for (int j = 0; j < n_depth; ++j) {
      const float x = e[n_depth - 1 - j];  // flip
      const float lg = logf(x + 1e-6f) - log1pf(-(x - 1e-6f));
      const float y = sigmoid_f((lg + gumbel_from_uniform(ur[n_width + j]) + base) * inv_t);
      orow[n_width + depth_order[j]] = y;
    }
  }
  __syncthreads();
  if (non_zero_width) {
    // rows whose thresholded slice is all-zero get [row, first col] += 0.5 (estimation_utils.py:23-31)
    for (int g = threadIdx.x; g < n_gates; g += GATE_THREADS)
      if (!gate_on[g]) orow[width_starts[g]] += 0.5f;
  }
}

// Backward of gumbel_gate_kernel w.r.t. z (K5): width dz = dy * y(1-y)/T with y recomputed from (z, u)
// (the non-zero-width fix-up adds a constant: estimation_utils.py:30); depth chains through
// sigmoid -> logit -> flip -> cumsum -> softmax (estimation_utils.py:50-64) and the depth_order scatter.
__global__ void __launch_bounds__(GATE_THREADS)
    gumbel_gate_bwd_kernel(const float* __restrict__ z, const float* __restrict__ u, const float* __restrict__ dy,
                           float* __restrict__ dz, int n_width, int n_depth, const int* __restrict__ depth_order,
                           float inv_t, float base) {
  const int row = blockIdx.x;
  const int dim = n_width + n_depth;
  const float* zr = z + (size_t)row * dim;
  const float* ur = u + (size_t)row * dim;
  const float* dyr = dy + (size_t)row * dim;
  float* dzr = dz + (size_t)row * dim;
  for (int c = threadIdx.x; c < n_width; c += GATE_THREADS) {
    const float y = sigmoid_f((zr[c] + gumbel_from_uniform(ur[c]) + base) * inv_t);
    dzr[c] = dyr[c] * y * (1.f - y) * inv_t;
  }
  if (threadIdx.x == 0 && n_depth > 0) {
    float s[MAX_DEPTH], cs[MAX_DEPTH], dc[MAX_DEPTH];
    float mx = -INFINITY;
    for (int j = 0; j < n_depth; ++j) mx = fmaxf(mx, zr[n_width + j]);
    float sum = 0.f;
    for (int j = 0; j < n_depth; ++j) {
      s[j] = expf(zr[n_width + j] - mx);
      sum += s[j];
    }
    float run = 0.f;
    for (int j = 0; j < n_depth; ++j) {
      s[j] = s[j] / sum;
      run += s[j];
      cs[j] = run;
    }
    for (int j = 0; j < n_depth; ++j) {
      const float x = cs[n_depth - 1 - j];
      const float lg = logf(x + 1e-6f) - log1pf(-(x - 1e-6f));
      const float y = sigmoid_f((lg + gumbel_from_uniform(ur[n_width + j]) + base) * inv_t);
      const float dlg = dyr[n_width + depth_order[j]] * y * (1.f - y) * inv_t;
      dc[n_depth - 1 - j] = dlg * (1.f / (x + 1e-6f) + 1.f / (1.f - (x - 1e-6f)));
    }
    // cumsum backward: ds_i = sum_{k >= i} dc_k ; softmax backward: dz_i = s_i (ds_i - sum_k s_k ds_k)
    float acc = 0.f, dot = 0.f;
    for (int i = n_depth - 1; i >= 0; --i) {
      acc += dc[i];
      dc[i] = acc;
      dot += s[i] * acc;
    }
    for (int i = 0; i < n_depth; ++i) dzr[n_width + i] = s[i] * (dc[i] - dot);
  }
}

constexpr int NORMZ_THREADS = 256;

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += scratch[i];
  return t;
}

__global__ void __launch_bounds__(NORMZ_THREADS)
    arch_normalize_kernel(const float* __restrict__ gates, float* __restrict__ out, int dim,
                          const int* __restrict__ col_depth, const float* __restrict__ col_scale, int l2) {
  __shared__ double scratch[32];
  const int row = blockIdx.x;
  const float* g = gates + (size_t)row * dim;
  float* o = out + (size_t)row * dim;
  double ss = 0.0;
  for (int c = threadIdx.x; c < dim; c += NORMZ_THREADS) {
    const int dc = col_depth[c];
    // width_depth_normalize (quantizer.py:239-246): depth-gated blocks keep soft width * depth,
    // everything else is hard_concrete'd (exact 0/1 in the forward pass)
    const float v = (dc >= 0 ? g[c] * g[dc] : (g[c] >= 0.5f ? 1.f : 0.f)) * col_scale[c];
    o[c] = v;
    ss += (double)v * (double)v;
  }
  if (!l2) return;
  const double tot = block_sum_d(ss, scratch);
  const float nrm = (float)sqrt(tot);
  for (int c = threadIdx.x; c < dim; c += NORMZ_THREADS) o[c] = o[c] / nrm;
}

// Backward of width_depth_normalize (no L2): hard_concrete is straight-through (identity), the
// depth-gated slices use the product rule.
__global__ void __launch_bounds__(NORMZ_THREADS)
    arch_normalize_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ dy, float* __restrict__ dx,
                              int dim, const int* __restrict__ col_depth, const float* __restrict__ col_scale) {
  extern __shared__ float dacc[];  // [dim] depth-column accumulators (only depth columns are used)
  const int row = blockIdx.x;
  const float* g = gates + (size_t)row * dim;
  const float* d = dy + (size_t)row * dim;
  float* o = dx + (size_t)row * dim;
  for (int c = threadIdx.x; c < dim; c += NORMZ_THREADS) dacc[c] = 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += NORMZ_THREADS) {
    const int dc = col_depth[c];
    const float gd = d[c] * col_scale[c];
    if (dc >= 0) {
      o[c] = gd * g[dc];
      atomicAdd(&dacc[dc], gd * g[c]);
    } else {
      o[c] = gd;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += NORMZ_THREADS) o[c] += dacc[c];
}

constexpr int MAX_CODES = 32;

// one warp per prompt: scores[b, k] = <a[b], codes[k]>, fp64 accumulation, argmax (first max wins)
__global__ void __launch_bounds__(256)
    route_cosine_kernel(const float* __restrict__ a, const float* __restrict__ codes, float* __restrict__ scores,
                        long long* __restrict__ indices, int batch, int dim, int n_codes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= batch) return;
  const float* ar = a + (size_t)row * dim;
  double acc[MAX_CODES];
#pragma unroll
  for (int k = 0; k < MAX_CODES; ++k) acc[k] = 0.0;
  for (int c = lane; c < dim; c += 32) {
    const double av = (double)ar[c];
#pragma unroll
    for (int k = 0; k < MAX_CODES; ++k)
      if (k < n_codes) acc[k] += av * (double)__ldg(codes + (size_t)k * dim + c);
  }
  float best = -INFINITY;
  int best_k = 0;
#pragma unroll
  for (int k = 0; k < MAX_CODES; ++k) {
    if (k < n_codes) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float f = (float)v;
      if (lane == 0 && scores) scores[(size_t)row * n_codes + k] = f;
      if (f > best) {
        best = f;
        best_k = k;
      }
    }
  }
  if (lane == 0 && indices) indices[row] = best_k;
}

constexpr int SK_THREADS = 1024;

// Single-CTA Sinkhorn phases over Q[b, k] (global memory, L2-resident: B*K floats).
// run_mask bit i = execute phase i; iterations only used when all phases run in one launch.
__global__ void __launch_bounds__(SK_THREADS)
    sinkhorn_kernel(float* __restrict__ Q, const float* __restrict__ scores, double* __restrict__ partial,
                    long long* __restrict__ indices, int batch_local, int batch_global, int K, float epsilon,
                    int phase, int first_iter, int iterations) {
  __shared__ double scratch[32];
  __shared__ float s_row[MAX_CODES];
  const int n = batch_local * K;
  const bool fused = (phase < 0);

  if (fused || phase == 0) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += SK_THREADS) {
      const float q = expf(scores[i] / epsilon);
      Q[i] = q;
      s += (double)q;
    }
    const double tot = block_sum_d(s, scratch);
    if (threadIdx.x == 0) partial[0] = tot;
    if (!fused) return;
    __syncthreads();
  }
  const int iters = fused ? iterations : 1;
  for (int it = 0; it < iters; ++it) {
    if (fused || phase == 1) {
      const bool div_total = fused ? (it == 0) : (first_iter != 0);
      const float total = (float)partial[0];
      // row sums: thread t owns code (t % K) for rows t/K, t/K + SK_THREADS/K, ... (SK_THREADS % K == 0 for K | 1024;
      // otherwise fall back to strided element loop with shared atomics-free per-code pass)
      for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int b = threadIdx.x; b < batch_local; b += SK_THREADS) {
          float q = Q[(size_t)b * K + k];
          if (div_total) {
            q = q / total;
            Q[(size_t)b * K + k] = q;
          }
          s += (double)q;
        }
        const double tot = block_sum_d(s, scratch);
        if (threadIdx.x == 0) partial[1 + k] = tot;
      }
      if (!fused) return;
      __syncthreads();
    }
    if (fused || phase == 2) {
      if (threadIdx.x < K) s_row[threadIdx.x] = (float)partial[1 + threadIdx.x];
      __syncthreads();
      for (int b = threadIdx.x; b < batch_local; b += SK_THREADS) {
        float q[MAX_CODES];
        float cs = 0.f;
        for (int k = 0; k < K; ++k) {
          float v = Q[(size_t)b * K + k];
          v = v / s_row[k];       // Q /= sum_of_rows
          v = v / (float)K;       // Q /= K
          q[k] = v;
          cs += v;
        }
        for (int k = 0; k < K; ++k) {
          float v = q[k] / cs;            // Q /= sum over prototypes
          v = v / (float)batch_global;    // Q /= B
          Q[(size_t)b * K + k] = v;
        }
      }
      if (!fused) return;
      __syncthreads();
    }
  }
  if (fused || phase == 3) {
    for (int b = threadIdx.x; b < batch_local; b += SK_THREADS) {
      float best = -INFINITY;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float v = Q[(size_t)b * K + k] * (float)batch_global;  // Q *= B
        Q[(size_t)b * K + k] = v;
        if (v > best) {
          best = v;
          bk = k;
        }
      }
      indices[b] = bk;
    }
  }
}

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_gumbel_gate_fwd(const float* z, const float* u, float* out, int32_t batch, int32_t n_width,
                                    int32_t n_depth, const int32_t* width_starts, int32_t n_width_gates,
                                    const int32_t* depth_order, float temperature, float base, int32_t non_zero_width,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(z && u && out && width_starts, "aptp_gumbel_gate_fwd: null pointer");
  APTP_REQUIRE(n_depth <= MAX_DEPTH && (n_depth == 0 || depth_order), "aptp_gumbel_gate_fwd: n_depth=%d unsupported", n_depth);
  APTP_REQUIRE(temperature > 0.f, "aptp_gumbel_gate_fwd: temperature must be > 0");
  if (batch == 0) return APTP_OK;
  const size_t smem = (size_t)(n_width + n_width_gates) * sizeof(int);
  APTP_REQUIRE(smem <= 48 * 1024, "aptp_gumbel_gate_fwd: arch vector too wide");
  gumbel_gate_kernel<<<batch, GATE_THREADS, smem, stream>>>(z, u, out, n_width, n_depth, width_starts, n_width_gates,
                                                           depth_order, 1.f / temperature, base, non_zero_width);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_gumbel_gate_bwd(const float* z, const float* u, const float* dy, float* dz, int32_t batch,
                                    int32_t n_width, int32_t n_depth, const int32_t* depth_order, float temperature,
                                    float base, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(z && u && dy && dz, "aptp_gumbel_gate_bwd: null pointer");
  APTP_REQUIRE(n_depth <= MAX_DEPTH && (n_depth == 0 || depth_order), "aptp_gumbel_gate_bwd: n_depth=%d unsupported", n_depth);
  APTP_REQUIRE(temperature > 0.f, "aptp_gumbel_gate_bwd: temperature must be > 0");
  if (batch == 0) return APTP_OK;
  gumbel_gate_bwd_kernel<<<batch, GATE_THREADS, 0, stream>>>(z, u, dy, dz, n_width, n_depth, depth_order,
                                                            1.f / temperature, base);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_arch_normalize_bwd(const float* gates, const float* dy, float* dx, int32_t batch, int32_t dim,
                                       const int32_t* col_depth, const float* col_scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(gates && dy && dx && col_depth && col_scale, "aptp_arch_normalize_bwd: null pointer");
  APTP_REQUIRE((size_t)dim * sizeof(float) <= 48 * 1024, "aptp_arch_normalize_bwd: arch vector too wide");
  if (batch == 0) return APTP_OK;
  arch_normalize_bwd_kernel<<<batch, NORMZ_THREADS, dim * sizeof(float), stream>>>(gates, dy, dx, dim, col_depth,
                                                                                   col_scale);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_arch_normalize(const float* gates, float* out, int32_t batch, int32_t dim, const int32_t* col_depth,
                                   const float* col_scale, int32_t l2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(gates && out && col_depth && col_scale, "aptp_arch_normalize: null pointer");
  if (batch == 0) return APTP_OK;
  arch_normalize_kernel<<<batch, NORMZ_THREADS, 0, stream>>>(gates, out, dim, col_depth, col_scale, l2);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_route_cosine(const float* a_norm, const float* codes_norm, float* scores, int64_t* indices,
                                 int32_t batch, int32_t dim, int32_t n_codes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(a_norm && codes_norm, "aptp_route_cosine: null pointer");
  APTP_REQUIRE(n_codes >= 1 && n_codes <= MAX_CODES, "aptp_route_cosine: n_codes=%d unsupported (max %d)", n_codes, MAX_CODES);
  if (batch == 0) return APTP_OK;
  route_cosine_kernel<<<(batch + 7) / 8, 256, 0, stream>>>(a_norm, codes_norm, scores,
                                                          reinterpret_cast<long long*>(indices), batch, dim, n_codes);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_sinkhorn_phase(int32_t phase, float* Q, const float* scores, double* partial, int64_t* indices,
                                   int32_t batch_local, int32_t batch_global, int32_t n_codes, float epsilon,
                                   int32_t first_iter, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(Q && partial, "aptp_sinkhorn_phase: null pointer");
  APTP_REQUIRE(phase >= 0 && phase <= 3, "aptp_sinkhorn_phase: bad phase %d", phase);
  APTP_REQUIRE(phase != 0 || scores, "aptp_sinkhorn_phase: phase 0 needs scores");
  APTP_REQUIRE(phase != 3 || indices, "aptp_sinkhorn_phase: phase 3 needs indices");
  APTP_REQUIRE(n_codes >= 1 && n_codes <= MAX_CODES, "aptp_sinkhorn_phase: n_codes=%d unsupported", n_codes);
  sinkhorn_kernel<<<1, SK_THREADS, 0, stream>>>(Q, scores, partial, reinterpret_cast<long long*>(indices), batch_local,
                                               batch_global, n_codes, epsilon, phase, first_iter, 1);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_route_sinkhorn(const float* scores, float* Q, double* partial, int64_t* indices, int32_t batch,
                                   int32_t n_codes, float epsilon, int32_t iterations, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(scores && Q && partial && indices, "aptp_route_sinkhorn: null pointer");
  APTP_REQUIRE(n_codes >= 1 && n_codes <= MAX_CODES, "aptp_route_sinkhorn: n_codes=%d unsupported", n_codes);
  sinkhorn_kernel<<<1, SK_THREADS, 0, stream>>>(Q, scores, partial, reinterpret_cast<long long*>(indices), batch, batch,
                                               n_codes, epsilon, -1, 1, iterations);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
