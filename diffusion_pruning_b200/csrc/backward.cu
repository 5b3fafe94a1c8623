// K5: backward kernels of the gated elementwise / normalisation ops. The U-Net weights
// are frozen during pruning (pdm/models/unet/unet_2d_conditional.py:2118-2122), so backward needs only
// activation gradients (bf16, fp32 math inside) and per-(sample, gate) reductions (fp32, atomics into
// caller-zeroed buffers). All kernels are HBM-bound: one read of each operand, one write of each result,
// 16-byte vectors, fixed thread -> channel mapping so per-channel constants and partial sums live in
// registers.
//
// Geometry shared by the [rows, C] kernels: grid = (pixel chunks, 2048-channel column blocks, samples);
// a thread owns one 8-channel vector column and walks down the pixels of its chunk.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int BW_THREADS = 256;
constexpr int BW_COLS = BW_THREADS * 8;  // channels per column block

__device__ __forceinline__ void unpack8b(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8b(const float* o) {
  uint4 r;
  r.x = pack_bf16(o[0], o[1]);
  r.y = pack_bf16(o[2], o[3]);
  r.z = pack_bf16(o[4], o[5]);
  r.w = pack_bf16(o[6], o[7]);
  return r;
}
__device__ __forceinline__ uint4 ldv(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stv(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

struct Geo {
  int b, c0, v_ok, p_first, p_end, p_step;
};
__device__ __forceinline__ Geo make_geo(int C, int hw, int ppc) {
  Geo g;
  g.b = blockIdx.z;
  const int cb0 = blockIdx.y * BW_COLS;
  const int cw = min(BW_COLS, C - cb0);
  const int cvb = cw >> 3;
  const int rpi = BW_THREADS / cvb;
  const int v = threadIdx.x % cvb;
  const int prow = threadIdx.x / cvb;
  g.c0 = cb0 + v * 8;
  g.v_ok = prow < rpi;
  const int p_begin = blockIdx.x * ppc;
  g.p_end = min(hw, p_begin + ppc);
  g.p_first = p_begin + prow;
  g.p_step = rpi;
  return g;
}

// Per-(sample, group) reduction of NQ per-channel partial sums held by each thread: merge runs of equal
// group inside the thread, shared-memory bins per CTA, one global atomic per (group, quantity) per CTA.
template <int NQ>
__device__ __forceinline__ void reduce_groups(float* bins, const float (&acc)[NQ][8], int c0, int C, int gs, bool active,
                                              float* __restrict__ dst, int dst_ld, int b) {
  const int cb0 = blockIdx.y * BW_COLS;
  const int g_first = cb0 / gs;
  const int g_last = (min(cb0 + BW_COLS, C) - 1) / gs;
  const int ng = g_last - g_first + 1;
  for (int i = threadIdx.x; i < ng * NQ; i += BW_THREADS) bins[i] = 0.f;
  __syncthreads();
  if (active) {
    int g_run = -1;
    float run[NQ];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      if (c >= C) break;
      const int g = c / gs - g_first;
      if (g != g_run) {
        if (g_run >= 0) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) atomicAdd(&bins[g_run * NQ + q], run[q]);
        }
        g_run = g;
#pragma unroll
        for (int q = 0; q < NQ; ++q) run[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) run[q] += acc[q][e];
    }
    if (g_run >= 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) atomicAdd(&bins[g_run * NQ + q], run[q]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng * NQ; i += BW_THREADS) {
    const int g = i / NQ, q = i - g * NQ;
    atomicAdd(&dst[((size_t)b * dst_ld + g_first + g) * NQ + q], bins[i]);
  }
}

// exact-erf GELU and its derivative: gelu_erf_fast / gelu_erf_fast_grad (common.cuh), the forms the GEMM epilogue uses

// ------------------------------------------------------------------------------------------------
// y = gate[b, c / group] * u   (WidthGate / LinearWidthGate on q,k,v: gates.py:15-21, blocks.py:250-255)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
    scale_cols_fwd_kernel(const __nv_bfloat16* __restrict__ u, int ldu, __nv_bfloat16* __restrict__ y, int ldy, int C,
                          int hw, const float* __restrict__ gate, int gate_ld, int group, int ppc) {
  const Geo g = make_geo(C, hw, ppc);
  if (!g.v_ok) return;
  float gv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) gv[e] = gate[(size_t)g.b * gate_ld + (g.c0 + e) / group];
  for (int p = g.p_first; p < g.p_end; p += g.p_step) {
    const long long row = (long long)g.b * hw + p;
    float f[8];
    unpack8b(ldv(u + row * ldu + g.c0), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= gv[e];
    stv(y + row * ldy + g.c0, pack8b(f));
  }
}

// du = gate * dy ; dgate[b, k] += sum dy * u
__global__ void __launch_bounds__(BW_THREADS)
    scale_cols_bwd_kernel(const __nv_bfloat16* __restrict__ u, int ldu, const __nv_bfloat16* __restrict__ dy, int lddy,
                          __nv_bfloat16* __restrict__ du, int lddu, int C, int hw, const float* __restrict__ gate,
                          int gate_ld, int group, float* __restrict__ dgate, int ppc) {
  extern __shared__ float bins[];
  const Geo g = make_geo(C, hw, ppc);
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  if (g.v_ok) {
    float gv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) gv[e] = gate[(size_t)g.b * gate_ld + (g.c0 + e) / group];
    for (int p = g.p_first; p < g.p_end; p += g.p_step) {
      const long long row = (long long)g.b * hw + p;
      float fu[8], fd[8], o[8];
      unpack8b(ldv(u + row * ldu + g.c0), fu);
      unpack8b(ldv(dy + row * lddy + g.c0), fd);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[0][e] += fd[e] * fu[e];
        o[e] = fd[e] * gv[e];
      }
      stv(du + row * lddu + g.c0, pack8b(o));
    }
  }
  reduce_groups<1>(bins, acc, g.c0, C, group, g.v_ok, dgate, gate_ld, g.b);
}

// ------------------------------------------------------------------------------------------------
// GEGLUGated (blocks.py:41-50), training form on the un-packed projection hg = [h | gate] (ungated):
//   f = (g h) * gelu(g gate)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
    geglu_fwd_kernel(const __nv_bfloat16* __restrict__ hg, int ld, __nv_bfloat16* __restrict__ out, int ldo, int inner,
                     int hw, const float* __restrict__ gate, int gate_ld, int group, int ppc) {
  const Geo g = make_geo(inner, hw, ppc);
  if (!g.v_ok) return;
  float gv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) gv[e] = gate ? gate[(size_t)g.b * gate_ld + (g.c0 + e) / group] : 1.f;
  for (int p = g.p_first; p < g.p_end; p += g.p_step) {
    const long long row = (long long)g.b * hw + p;
    float h[8], t[8], o[8];
    unpack8b(ldv(hg + row * ld + g.c0), h);
    unpack8b(ldv(hg + row * ld + inner + g.c0), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (gv[e] * h[e]) * gelu_erf_fast(gv[e] * t[e]);
    stv(out + row * ldo + g.c0, pack8b(o));
  }
}

__global__ void __launch_bounds__(BW_THREADS)
    geglu_bwd_kernel(const __nv_bfloat16* __restrict__ hg, int ld, const __nv_bfloat16* __restrict__ df, int lddf,
                     __nv_bfloat16* __restrict__ dhg, int lddhg, int inner, int hw, const float* __restrict__ gate,
                     int gate_ld, int group, float* __restrict__ dgate, int ppc) {
  extern __shared__ float bins[];
  const Geo g = make_geo(inner, hw, ppc);
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  if (g.v_ok) {
    float gv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) gv[e] = gate ? gate[(size_t)g.b * gate_ld + (g.c0 + e) / group] : 1.f;
    for (int p = g.p_first; p < g.p_end; p += g.p_step) {
      const long long row = (long long)g.b * hw + p;
      float h[8], t[8], d[8], oh[8], ot[8];
      unpack8b(ldv(hg + row * ld + g.c0), h);
      unpack8b(ldv(hg + row * ld + inner + g.c0), t);
      unpack8b(ldv(df + row * lddf + g.c0), d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float hh = gv[e] * h[e], tt = gv[e] * t[e];
        float gl, gg;
        gelu_erf_fast_grad(tt, gl, gg);
        const float dhh = d[e] * gl;        // d/d(g h)
        const float dtt = d[e] * hh * gg;   // d/d(g gate)
        oh[e] = gv[e] * dhh;
        ot[e] = gv[e] * dtt;
        acc[0][e] += dhh * h[e] + dtt * t[e];
      }
      stv(dhg + row * lddhg + g.c0, pack8b(oh));
      stv(dhg + row * lddhg + inner + g.c0, pack8b(ot));
    }
  }
  if (dgate) reduce_groups<1>(bins, acc, g.c0, inner, group, g.v_ok, dgate, gate_ld, g.b);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+ fused width gate before it, + SiLU after it) backward.
//   xg = g x ; xh = (xg - g mu) * rstd_g ; y = xh gamma + beta ; a = silu(y)
//   dxh = da silu'(y) gamma ; S1 = sum dxh ; S2 = sum dxh xh   (per sample, group; n = hw * gs)
//   dxg = rstd_g (dxh - S1/n - xh S2/n) ; dx = g dxg ; dg = sum dxg x
// ------------------------------------------------------------------------------------------------
struct GnB {
  const __nv_bfloat16* x;
  int ldx;
  const __nv_bfloat16* da;
  int ldda;
  int C, hw, gs;
  float eps;
  const float* stats;  // forward (sum, sumsq) of the ungated x: [B][stats_groups][2]
  int stats_groups;
  const float* gamma;
  const float* beta;
  const float* gate;   // [B][gate_ld] or null
  int gate_ld;
  int silu;
};

__device__ __forceinline__ void gn_coeffs(const GnB& a, int b, int c0, float* mean_g, float* rstd, float* gm, float* bt,
                                          float* gt) {
  const float inv_n = 1.f / ((float)a.hw * (float)a.gs);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    if (c < a.C) {
      const int g = c / a.gs;
      const float su = a.stats[((size_t)b * a.stats_groups + g) * 2];
      const float ss = a.stats[((size_t)b * a.stats_groups + g) * 2 + 1];
      const float gv = a.gate ? a.gate[(size_t)b * a.gate_ld + g] : 1.f;
      const float mean = su * inv_n;
      const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
      gt[e] = gv;
      mean_g[e] = mean;
      rstd[e] = rsqrtf(gv * gv * var + a.eps);
      gm[e] = a.gamma[c];
      bt[e] = a.beta[c];
    } else {
      gt[e] = 0.f; mean_g[e] = 0.f; rstd[e] = 0.f; gm[e] = 0.f; bt[e] = 0.f;
    }
  }
}
__device__ __forceinline__ float silu_grad(float y) {
  float t;  // sigmoid(y) = 0.5 + 0.5 tanh(y / 2): one XU op (matches the forward's silu_tanh_f)
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  const float s = fmaf(0.5f, t, 0.5f);
  return s * (1.f + y * (1.f - s));
}

// kAffine: also accumulate the affine gradients daffine[c] = (dgamma, dbeta) = (sum dy xh, sum dy) over all samples and
// pixels (dy = dL/d(xh gamma + beta), i.e. after the SiLU derivative) -- the weight-training backward of nn.GroupNorm.
template <bool kAffine>
__global__ void __launch_bounds__(BW_THREADS) gn_bwd_stats_kernel(GnB a, float* __restrict__ bstats,
                                                                  float* __restrict__ daffine, int ppc) {
  extern __shared__ float bins[];
  const Geo g = make_geo(a.C, a.hw, ppc);
  float acc[2][8], aff[2][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = aff[0][e] = aff[1][e] = 0.f;
  if (g.v_ok) {
    float mean[8], rstd[8], gm[8], bt[8], gt[8];
    gn_coeffs(a, g.b, g.c0, mean, rstd, gm, bt, gt);
    // 4 pixels per step: 8 independent 16-byte loads in flight per thread before any math
    for (int p = g.p_first; p < g.p_end; p += 4 * g.p_step) {
      uint4 rx[4], rd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + u * g.p_step;
        const long long row = (long long)g.b * a.hw + pp;
        rx[u] = rd[u] = make_uint4(0u, 0u, 0u, 0u);
        if (pp < g.p_end) {
          rx[u] = ldv(a.x + row * a.ldx + g.c0);
          rd[u] = ldv(a.da + row * a.ldda + g.c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p + u * g.p_step >= g.p_end) break;
        float fx[8], fd[8];
        unpack8b(rx[u], fx);
        unpack8b(rd[u], fd);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = gt[e] * (fx[e] - mean[e]) * rstd[e];
          float dy = fd[e];
          if (a.silu) dy *= silu_grad(xh * gm[e] + bt[e]);
          const float dxh = dy * gm[e];
          acc[0][e] += dxh;
          acc[1][e] += dxh * xh;
          if (kAffine) {
            aff[0][e] += dy * xh;
            aff[1][e] += dy;
          }
        }
      }
    }
  }
  reduce_groups<2>(bins, acc, g.c0, a.C, a.gs, g.v_ok, bstats, a.stats_groups, g.b);
  if (kAffine) {
    __syncthreads();  // the bins are reused: per-channel "groups" of one, destination [C][2], no sample index
    reduce_groups<2>(bins, aff, g.c0, a.C, 1, g.v_ok, daffine, 0, 0);
  }
}

__global__ void __launch_bounds__(BW_THREADS)
    gn_bwd_apply_kernel(GnB a, const float* __restrict__ bstats, __nv_bfloat16* __restrict__ dx, int lddx,
                        int accumulate, float* __restrict__ dgate, int ppc) {
  extern __shared__ float bins[];
  const Geo g = make_geo(a.C, a.hw, ppc);
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  if (g.v_ok) {
    float mean[8], rstd[8], gm[8], bt[8], gt[8], s1[8], s2[8];
    gn_coeffs(a, g.b, g.c0, mean, rstd, gm, bt, gt);
    const float inv_n = 1.f / ((float)a.hw * (float)a.gs);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g.c0 + e;
      const int grp = (c < a.C) ? c / a.gs : 0;
      s1[e] = bstats[((size_t)g.b * a.stats_groups + grp) * 2] * inv_n;
      s2[e] = bstats[((size_t)g.b * a.stats_groups + grp) * 2 + 1] * inv_n;
    }
    for (int p = g.p_first; p < g.p_end; p += 2 * g.p_step) {
      uint4 rx[2], rd[2], ro[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = p + u * g.p_step;
        const long long row = (long long)g.b * a.hw + pp;
        rx[u] = rd[u] = ro[u] = make_uint4(0u, 0u, 0u, 0u);
        if (pp < g.p_end) {
          rx[u] = ldv(a.x + row * a.ldx + g.c0);
          rd[u] = ldv(a.da + row * a.ldda + g.c0);
          if (accumulate) ro[u] = *reinterpret_cast<const uint4*>(dx + row * lddx + g.c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = p + u * g.p_step;
        if (pp >= g.p_end) break;
        const long long row = (long long)g.b * a.hw + pp;
        float fx[8], fd[8], o[8];
        unpack8b(rx[u], fx);
        unpack8b(rd[u], fd);
        unpack8b(ro[u], o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = gt[e] * (fx[e] - mean[e]) * rstd[e];
          float dy = fd[e];
          if (a.silu) dy *= silu_grad(xh * gm[e] + bt[e]);
          const float dxh = dy * gm[e];
          const float dxg = rstd[e] * (dxh - s1[e] - xh * s2[e]);
          acc[0][e] += dxg * fx[e];
          o[e] += gt[e] * dxg;
        }
        stv(dx + row * lddx + g.c0, pack8b(o));
      }
    }
  }
  if (dgate) reduce_groups<1>(bins, acc, g.c0, a.C, a.gs, g.v_ok, dgate, a.gate_ld, g.b);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (one warp per token row): dx (+)= rstd (dyh - mean(dyh) - xh mean(dyh xh)), dyh = dy gamma
// ------------------------------------------------------------------------------------------------
template <int SLOTS>
__global__ void __launch_bounds__(256)
    layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                         __nv_bfloat16* __restrict__ dx, int lddx, long long rows, int C, float eps,
                         const float* __restrict__ gamma, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int cv = C / 8;
  float fx[SLOTS][8], fd[SLOTS][8];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < SLOTS; ++q) {
    const int v = lane + q * 32;
    if (v < cv) {
      unpack8b(ldv(x + row * ldx + v * 8), fx[q]);
      unpack8b(ldv(dy + row * lddy + v * 8), fd[q]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) fx[q][e] = fd[q][e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s += fx[q][e];
  }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < SLOTS; ++q) {
    if (lane + q * 32 < cv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = fx[q][e] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int q = 0; q < SLOTS; ++q) {
    const int v = lane + q * 32;
    if (v < cv) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        fx[q][e] = (fx[q][e] - mean) * rstd;  // xh
        fd[q][e] *= gg[e];                    // dyh
        m1 += fd[q][e];
        m2 += fd[q][e] * fx[q][e];
      }
    }
  }
  m1 = warp_sum(m1) / (float)C;
  m2 = warp_sum(m2) / (float)C;
#pragma unroll
  for (int q = 0; q < SLOTS; ++q) {
    const int v = lane + q * 32;
    if (v < cv) {
      float o[8];
      if (accumulate) unpack8b(*reinterpret_cast<const uint4*>(dx + row * lddx + v * 8), o);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rstd * (fd[q][e] - m1 - fx[q][e] * m2);
      stv(dx + row * lddx + v * 8, pack8b(o));
    }
  }
}

// LayerNorm affine gradients: daffine[c] = (dgamma, dbeta) = (sum_rows dy xh, sum_rows dy). Persistent warps stride over
// the rows with the statistics recomputed per row (exact two-pass, like the forward); per-lane sums meet in shared
// memory and leave the CTA as one atomic per (channel, quantity).
template <int SLOTS>
__global__ void __launch_bounds__(256)
    layernorm_affine_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ dy, int lddy,
                                long long rows, int C, float eps, float* __restrict__ daffine) {
  extern __shared__ float sh_aff[];  // [2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cv = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += 256) sh_aff[i] = 0.f;
  __syncthreads();
  float ag[SLOTS][8], ab[SLOTS][8];
#pragma unroll
  for (int q = 0; q < SLOTS; ++q)
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[q][e] = ab[q][e] = 0.f;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    float fx[SLOTS][8], fd[SLOTS][8];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const int v = lane + q * 32;
      if (v < cv) {
        unpack8b(ldv(x + row * ldx + v * 8), fx[q]);
        unpack8b(ldv(dy + row * lddy + v * 8), fd[q]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) fx[q][e] = fd[q][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += fx[q][e];
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      if (lane + q * 32 < cv) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = fx[q][e] - mean;
          ss += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ag[q][e] += fd[q][e] * (fx[q][e] - mean) * rstd;
        ab[q][e] += fd[q][e];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < SLOTS; ++q) {
    const int v = lane + q * 32;
    if (v < cv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&sh_aff[v * 8 + e], ag[q][e]);
        atomicAdd(&sh_aff[C + v * 8 + e], ab[q][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    atomicAdd(&daffine[2 * i], sh_aff[i]);
    atomicAdd(&daffine[2 * i + 1], sh_aff[C + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// DepthGate backward (gates.py:36-42): out = (1-d) x + d y
//   dy = d dout ; dx (+)= (1-d) dout ; dd[b] += sum dout (y - x)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
    depth_lerp_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int lddo, const __nv_bfloat16* __restrict__ x, int ldx,
                          const __nv_bfloat16* __restrict__ y, int ldy, __nv_bfloat16* __restrict__ dy, int lddy,
                          __nv_bfloat16* __restrict__ dx, int lddx, int accumulate, int C, int hw,
                          const float* __restrict__ d, float* __restrict__ dd, int ppc) {
  __shared__ float red[BW_THREADS / 32];
  const Geo g = make_geo(C, hw, ppc);
  float acc = 0.f;
  if (g.v_ok) {
    const float dv = d[g.b];
    for (int p = g.p_first; p < g.p_end; p += g.p_step) {
      const long long row = (long long)g.b * hw + p;
      float fo[8], fx[8], fy[8], o1[8], o2[8];
      unpack8b(ldv(dout + row * lddo + g.c0), fo);
      unpack8b(ldv(x + row * ldx + g.c0), fx);
      unpack8b(ldv(y + row * ldy + g.c0), fy);
      if (accumulate) unpack8b(*reinterpret_cast<const uint4*>(dx + row * lddx + g.c0), o2);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o2[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc += fo[e] * (fy[e] - fx[e]);
        o1[e] = dv * fo[e];
        o2[e] += (1.f - dv) * fo[e];
      }
      stv(dy + row * lddy + g.c0, pack8b(o1));
      stv(dx + row * lddx + g.c0, pack8b(o2));
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < BW_THREADS / 32; ++i) t += red[i];
    atomicAdd(&dd[g.b], t);
  }
}

// dst (+)= src over [rows, C] (gradient fan-in of skip connections / residual branches)
__global__ void __launch_bounds__(256) add_rows_kernel(const __nv_bfloat16* __restrict__ src, int lds,
                                                       __nv_bfloat16* __restrict__ dst, int ldd, long long rows, int cv) {
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int v = (int)(i - row * cv);
    float a[8], b[8];
    unpack8b(ldv(src + row * lds + v * 8), a);
    unpack8b(*reinterpret_cast<const uint4*>(dst + row * ldd + v * 8), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] += a[e];
    stv(dst + row * ldd + v * 8, pack8b(b));
  }
}

// nearest x2 upsample backward: dX[b,y,x] = sum of the 4 children
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dyp,
                                                             __nv_bfloat16* __restrict__ dxp, int batch, int H, int W,
                                                             int cv) {
  const long long total = (long long)batch * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long pix = i / cv;
    const int x = (int)(pix % W);
    pix /= W;
    const int y = (int)(pix % H);
    const int b = (int)(pix / H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long sp = ((long long)b * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx;
        float f[8];
        unpack8b(ldv(dyp + (sp * cv + v) * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    stv(dxp + i * 8, pack8b(acc));
  }
}

// stride-2 conv dgrad helper: dYu[b, 2oy, 2ox] = dY[b, oy, ox], zero elsewhere (H, W = OUTPUT size of the conv)
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const __nv_bfloat16* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst, int batch, int H, int W,
                                                            int cv) {
  const long long total = (long long)batch * 4 * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long pix = i / cv;
    const int x = (int)(pix % (2 * W));
    pix /= (2 * W);
    const int y = (int)(pix % (2 * H));
    const int b = (int)(pix / (2 * H));
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (((x | y) & 1) == 0) val = ldv(src + ((((long long)b * H + (y >> 1)) * W + (x >> 1)) * cv + v) * 8);
    stv(dst + i * 8, val);
  }
}

static int bw_ppc(int hw, int batch, int C) {
  const int colblocks = (C + BW_COLS - 1) / BW_COLS;
  long long target = 6LL * sm_count();
  long long chunks = (target + (long long)batch * colblocks - 1) / ((long long)batch * colblocks);
  if (chunks < 1) chunks = 1;
  int ppc = (int)((hw + chunks - 1) / chunks);
  if (ppc < 8) ppc = 8;
  if (ppc > hw) ppc = hw;
  return ppc;
}
static unsigned bw_grid1d(long long total) {
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
static size_t bins_bytes(int C, int gs, int nq) {
  const int cols = C < BW_COLS ? C : BW_COLS;
  return (size_t)(cols / gs + 3) * nq * sizeof(float);
}

}  // namespace aptp

using namespace aptp;

#define BW_GRID(C, hw, batch)                           \
  const int ppc = bw_ppc(hw, batch, C);                 \
  dim3 grid((hw + ppc - 1) / ppc, (C + BW_COLS - 1) / BW_COLS, batch)

extern "C" int aptp_scale_cols_fwd(const void* u, int32_t ldu, void* y, int32_t ldy, int32_t batch, int32_t hw, int32_t C,
                                   const float* gate, int32_t gate_ld, int32_t group, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(u && y && gate && group > 0, "aptp_scale_cols_fwd: bad arguments");
  APTP_REQUIRE(C % 8 == 0 && ldu % 8 == 0 && ldy % 8 == 0, "aptp_scale_cols_fwd: C and pitches must be multiples of 8");
  if (batch == 0 || hw == 0) return APTP_OK;
  BW_GRID(C, hw, batch);
  scale_cols_fwd_kernel<<<grid, BW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(u), ldu,
                                                        reinterpret_cast<__nv_bfloat16*>(y), ldy, C, hw, gate, gate_ld,
                                                        group, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_scale_cols_bwd(const void* u, int32_t ldu, const void* dy, int32_t lddy, void* du, int32_t lddu,
                                   int32_t batch, int32_t hw, int32_t C, const float* gate, int32_t gate_ld,
                                   int32_t group, float* dgate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(u && dy && du && gate && dgate && group > 0, "aptp_scale_cols_bwd: bad arguments");
  APTP_REQUIRE(C % 8 == 0 && ldu % 8 == 0 && lddy % 8 == 0 && lddu % 8 == 0,
               "aptp_scale_cols_bwd: C and pitches must be multiples of 8");
  if (batch == 0 || hw == 0) return APTP_OK;
  BW_GRID(C, hw, batch);
  scale_cols_bwd_kernel<<<grid, BW_THREADS, bins_bytes(C, group, 1), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(u), ldu, reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
      reinterpret_cast<__nv_bfloat16*>(du), lddu, C, hw, gate, gate_ld, group, dgate, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_geglu_fwd(const void* hg, int32_t ld, void* out, int32_t ldo, int32_t batch, int32_t hw,
                              int32_t inner, const float* gate, int32_t gate_ld, int32_t group, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(hg && out && (gate == nullptr || group > 0), "aptp_geglu_fwd: bad arguments");
  APTP_REQUIRE(inner % 8 == 0 && ld % 8 == 0 && ldo % 8 == 0, "aptp_geglu_fwd: sizes must be multiples of 8");
  if (batch == 0 || hw == 0) return APTP_OK;
  BW_GRID(inner, hw, batch);
  geglu_fwd_kernel<<<grid, BW_THREADS, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(hg), ld,
                                                   reinterpret_cast<__nv_bfloat16*>(out), ldo, inner, hw, gate, gate_ld,
                                                   group > 0 ? group : 1, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_geglu_bwd(const void* hg, int32_t ld, const void* df, int32_t lddf, void* dhg, int32_t lddhg,
                              int32_t batch, int32_t hw, int32_t inner, const float* gate, int32_t gate_ld,
                              int32_t group, float* dgate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(hg && df && dhg && (gate == nullptr || (group > 0 && dgate)), "aptp_geglu_bwd: bad arguments");
  APTP_REQUIRE(inner % 8 == 0 && ld % 8 == 0 && lddf % 8 == 0 && lddhg % 8 == 0,
               "aptp_geglu_bwd: sizes must be multiples of 8");
  if (batch == 0 || hw == 0) return APTP_OK;
  BW_GRID(inner, hw, batch);
  const int grp = group > 0 ? group : inner;
  geglu_bwd_kernel<<<grid, BW_THREADS, bins_bytes(inner, grp, 1), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(hg), ld, reinterpret_cast<const __nv_bfloat16*>(df), lddf,
      reinterpret_cast<__nv_bfloat16*>(dhg), lddhg, inner, hw, gate, gate_ld, grp, gate ? dgate : nullptr, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

static int groupnorm_bwd_impl(const void* x, int32_t ldx, const void* da, int32_t ldda, void* dx, int32_t lddx,
                                  int32_t accumulate, int32_t batch, int32_t hw, int32_t C, int32_t group_size, float eps,
                                  const float* stats, int32_t stats_groups, const float* gamma, const float* beta,
                                  const float* gate, int32_t gate_ld, int32_t silu, float* bstats, float* dgate,
                                  float* daffine, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && da && dx && stats && gamma && beta && bstats, "aptp_groupnorm_bwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldda % 8 == 0 && lddx % 8 == 0 && group_size > 0,
               "aptp_groupnorm_bwd: C and pitches must be multiples of 8");
  APTP_REQUIRE((C + group_size - 1) / group_size <= stats_groups, "aptp_groupnorm_bwd: stats_groups too small");
  APTP_REQUIRE(gate == nullptr || dgate != nullptr, "aptp_groupnorm_bwd: gate without dgate");
  if (batch == 0 || hw == 0) return APTP_OK;
  GnB a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x);
  a.ldx = ldx;
  a.da = reinterpret_cast<const __nv_bfloat16*>(da);
  a.ldda = ldda;
  a.C = C;
  a.hw = hw;
  a.gs = group_size;
  a.eps = eps;
  a.stats = stats;
  a.stats_groups = stats_groups;
  a.gamma = gamma;
  a.beta = beta;
  a.gate = gate;
  a.gate_ld = gate_ld;
  a.silu = silu;
  BW_GRID(C, hw, batch);
  APTP_CUDA_CHECK(cudaMemsetAsync(bstats, 0, (size_t)batch * stats_groups * 2 * sizeof(float), stream));
  if (daffine) {
    const size_t sm = bins_bytes(C, group_size, 2) > bins_bytes(C, 1, 2) ? bins_bytes(C, group_size, 2) : bins_bytes(C, 1, 2);
    gn_bwd_stats_kernel<true><<<grid, BW_THREADS, sm, stream>>>(a, bstats, daffine, ppc);
  } else {
    gn_bwd_stats_kernel<false><<<grid, BW_THREADS, bins_bytes(C, group_size, 2), stream>>>(a, bstats, nullptr, ppc);
  }
  APTP_CUDA_CHECK(cudaGetLastError());
  gn_bwd_apply_kernel<<<grid, BW_THREADS, bins_bytes(C, group_size, 1), stream>>>(
      a, bstats, reinterpret_cast<__nv_bfloat16*>(dx), lddx, accumulate, gate ? dgate : nullptr, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_groupnorm_bwd(const void* x, int32_t ldx, const void* da, int32_t ldda, void* dx, int32_t lddx,
                                  int32_t accumulate, int32_t batch, int32_t hw, int32_t C, int32_t group_size, float eps,
                                  const float* stats, int32_t stats_groups, const float* gamma, const float* beta,
                                  const float* gate, int32_t gate_ld, int32_t silu, float* bstats, float* dgate,
                                  void* stream_) {
  return groupnorm_bwd_impl(x, ldx, da, ldda, dx, lddx, accumulate, batch, hw, C, group_size, eps, stats, stats_groups, gamma,
                            beta, gate, gate_ld, silu, bstats, dgate, nullptr, stream_);
}

extern "C" int aptp_groupnorm_bwd_affine(const void* x, int32_t ldx, const void* da, int32_t ldda, void* dx, int32_t lddx,
                                         int32_t accumulate, int32_t batch, int32_t hw, int32_t C, int32_t group_size,
                                         float eps, const float* stats, int32_t stats_groups, const float* gamma,
                                         const float* beta, const float* gate, int32_t gate_ld, int32_t silu, float* bstats,
                                         float* dgate, float* daffine, void* stream_) {
  APTP_REQUIRE(daffine != nullptr, "aptp_groupnorm_bwd_affine: null daffine");
  return groupnorm_bwd_impl(x, ldx, da, ldda, dx, lddx, accumulate, batch, hw, C, group_size, eps, stats, stats_groups, gamma,
                            beta, gate, gate_ld, silu, bstats, dgate, daffine, stream_);
}

extern "C" int aptp_layernorm_affine_bwd(const void* x, int32_t ldx, const void* dy, int32_t lddy, int64_t rows, int32_t C,
                                         float eps, float* daffine, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && dy && daffine, "aptp_layernorm_affine_bwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && C <= 32 * 8 * 5 && ldx % 8 == 0 && lddy % 8 == 0,
               "aptp_layernorm_affine_bwd: unsupported C=%d (multiple of 8, at most 1280)", C);
  if (rows == 0) return APTP_OK;
  const int slots = (C / 8 + 31) / 32;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  const size_t sm = (size_t)2 * C * sizeof(float);
#define APTP_LNA_LAUNCH(S)                                                                                       \
  layernorm_affine_bwd_kernel<S><<<(unsigned)blocks, 256, sm, stream>>>(                                         \
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(dy), lddy, rows, C, eps, daffine)
  if (slots <= 1) APTP_LNA_LAUNCH(1);
  else if (slots == 2) APTP_LNA_LAUNCH(2);
  else if (slots == 3) APTP_LNA_LAUNCH(3);
  else APTP_LNA_LAUNCH(5);
#undef APTP_LNA_LAUNCH
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_layernorm_bwd(const void* x, int32_t ldx, const void* dy, int32_t lddy, void* dx, int32_t lddx,
                                  int32_t accumulate, int64_t rows, int32_t C, float eps, const float* gamma,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && dy && dx && gamma, "aptp_layernorm_bwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && C <= 32 * 8 * 8 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0,
               "aptp_layernorm_bwd: unsupported C=%d", C);
  if (rows == 0) return APTP_OK;
  const long long blocks = (rows + 7) / 8;
  const int slots = (C / 8 + 31) / 32;
#define APTP_LNB_LAUNCH(S)                                                                                     \
  layernorm_bwd_kernel<S><<<(unsigned)blocks, 256, 0, stream>>>(                                               \
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(dy), lddy,       \
      reinterpret_cast<__nv_bfloat16*>(dx), lddx, rows, C, eps, gamma, accumulate)
  if (slots <= 1) APTP_LNB_LAUNCH(1);
  else if (slots == 2) APTP_LNB_LAUNCH(2);
  else if (slots == 3) APTP_LNB_LAUNCH(3);
  else if (slots <= 5) APTP_LNB_LAUNCH(5);
  else APTP_LNB_LAUNCH(8);
#undef APTP_LNB_LAUNCH
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_depth_lerp_bwd(const void* dout, int32_t lddo, const void* x, int32_t ldx, const void* y, int32_t ldy,
                                   void* dy, int32_t lddy, void* dx, int32_t lddx, int32_t accumulate, int32_t batch,
                                   int32_t hw, int32_t C, const float* d, float* dd, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(dout && x && y && dy && dx && d && dd, "aptp_depth_lerp_bwd: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lddo % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0,
               "aptp_depth_lerp_bwd: C and pitches must be multiples of 8");
  if (batch == 0 || hw == 0) return APTP_OK;
  BW_GRID(C, hw, batch);
  depth_lerp_bwd_kernel<<<grid, BW_THREADS, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dout), lddo, reinterpret_cast<const __nv_bfloat16*>(x), ldx,
      reinterpret_cast<const __nv_bfloat16*>(y), ldy, reinterpret_cast<__nv_bfloat16*>(dy), lddy,
      reinterpret_cast<__nv_bfloat16*>(dx), lddx, accumulate, C, hw, d, dd, ppc);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_add_rows(const void* src, int32_t lds, void* dst, int32_t ldd, int64_t rows, int32_t C,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst, "aptp_add_rows: null pointer");
  APTP_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "aptp_add_rows: C and pitches must be multiples of 8");
  if (rows == 0) return APTP_OK;
  add_rows_kernel<<<bw_grid1d(rows * (C / 8)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds,
                                                                reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, C / 8);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_upsample2x_bwd(const void* dy, void* dx, int32_t batch, int32_t H, int32_t W, int32_t C,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(dy && dx && C % 8 == 0, "aptp_upsample2x_bwd: bad arguments");
  const long long total = (long long)batch * H * W * (C / 8);
  if (total == 0) return APTP_OK;
  upsample2x_bwd_kernel<<<bw_grid1d(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                             reinterpret_cast<__nv_bfloat16*>(dx), batch, H, W, C / 8);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_zero_insert2x(const void* src, void* dst, int32_t batch, int32_t H, int32_t W, int32_t C,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(src && dst && C % 8 == 0, "aptp_zero_insert2x: bad arguments");
  const long long total = (long long)batch * 4 * H * W * (C / 8);
  if (total == 0) return APTP_OK;
  zero_insert2x_kernel<<<bw_grid1d(total), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                            reinterpret_cast<__nv_bfloat16*>(dst), batch, H, W, C / 8);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
