// K2: HBM-bound normalisation kernels over NHWC / token-major bf16.
//   aptp_groupnorm_stats : per-(sample, group) sum / sum-of-squares, one read of x
//   aptp_groupnorm_apply : y = [silu]((g*x - g*mean) * rstd' * gamma + beta), one read + one write
//   aptp_layernorm       : one warp per token row, values held in registers (exact two-pass)
// Thread -> channel mapping is fixed (each thread owns one 16-byte vector of 8 channels and walks
// down the pixels), so per-channel partial sums and the affine parameters live in registers and every
// global access is a coalesced 16-byte vector.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

#ifndef APTP_GN_SILU
#define APTP_GN_SILU silu_f
#endif
constexpr int NORM_THREADS = 256;
constexpr int MAX_SLOTS = 2;  // 8-channel vectors per thread: supports up to 256*2*8 = 4096 channels

struct Src2 {
  const void* x0;  // bf16 or fp32 (kF32) rows
  const void* x1;
  int c0, ld0, c1, ld1;
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// 8 consecutive channels of one pixel as raw 16-byte vectors: one for bf16 rows, two for fp32 rows (the residual
// stream between blocks is kept in fp32).
template <bool kF32>
struct Vec8 {
  uint4 q[kF32 ? 2 : 1];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < (kF32 ? 2 : 1); ++i) q[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ void to_float(float* f) const {
    if constexpr (kF32) {
      f[0] = __uint_as_float(q[0].x); f[1] = __uint_as_float(q[0].y);
      f[2] = __uint_as_float(q[0].z); f[3] = __uint_as_float(q[0].w);
      f[4] = __uint_as_float(q[1].x); f[5] = __uint_as_float(q[1].y);
      f[6] = __uint_as_float(q[1].z); f[7] = __uint_as_float(q[1].w);
    } else {
      unpack8(q[0], f);
    }
  }
};

template <bool kF32>
__device__ __forceinline__ Vec8<kF32> load_vec(const Src2& s, long long pix, int c) {
  // c is a multiple of 8; c0 is a multiple of 8 so a vector never straddles the two sources
  Vec8<kF32> r;
  if constexpr (kF32) {
    const float* p = (c < s.c0) ? reinterpret_cast<const float*>(s.x0) + pix * s.ld0 + c
                                : reinterpret_cast<const float*>(s.x1) + pix * s.ld1 + (c - s.c0);
    r.q[0] = __ldg(reinterpret_cast<const uint4*>(p));
    r.q[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  } else {
    const __nv_bfloat16* p = (c < s.c0) ? reinterpret_cast<const __nv_bfloat16*>(s.x0) + pix * s.ld0 + c
                                        : reinterpret_cast<const __nv_bfloat16*>(s.x1) + pix * s.ld1 + (c - s.c0);
    r.q[0] = __ldg(reinterpret_cast<const uint4*>(p));
  }
  return r;
}

// Statistics pass. DETERMINISTIC (round 2; round 1 used fp32 atomics, so a forward was not bit-reproducible):
//   1. every thread keeps per-channel partial sums of its pixels in registers and stores them to shared memory,
//   2. one warp per group sums the group's (pixel-row, channel) cells in a fixed order + a fixed shuffle tree,
//   3. the CTA writes its per-group partials to the workspace; the LAST CTA of a sample to arrive (ticket counter)
//      adds the chunks' partials in chunk order, whichever CTA that happens to be, and resets the counter.
template <bool kF32>
__global__ void __launch_bounds__(NORM_THREADS)
    gn_stats_kernel(Src2 s, int hw, int gs, const int* __restrict__ sample_channels, float* __restrict__ stats,
                    int stats_groups, int pix_per_cta, unsigned int* __restrict__ tickets, float* __restrict__ partial) {
  extern __shared__ float cells[];  // [2][rows_per_iter][cv * 8]  (sum plane, then sum-of-squares plane)
  __shared__ int s_last;
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y;
  const int ctot = sample_channels ? sample_channels[b] : (s.c0 + s.c1);
  if (ctot <= 0) return;  // uniform over the sample's CTAs
  const int cv = (ctot + 7) / 8;
  const int groups = (ctot + gs - 1) / gs;
  const int cpad = cv * 8;

  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  float sum[MAX_SLOTS][8], sq[MAX_SLOTS][8];
#pragma unroll
  for (int q = 0; q < MAX_SLOTS; ++q)
#pragma unroll
    for (int e = 0; e < 8; ++e) sum[q][e] = sq[q][e] = 0.f;

  int n_rows;  // pixel rows held in shared memory
  if (cv <= NORM_THREADS) {
    const int rows_per_iter = NORM_THREADS / cv;
    n_rows = rows_per_iter;
    const int v = threadIdx.x % cv;
    const int prow = threadIdx.x / cv;
    if (prow < rows_per_iter) {
      // software pipeline: the 4 loads of the NEXT step are in flight while this step's values are reduced
      const int step = 4 * rows_per_iter;
      Vec8<kF32> cur[4], nxt[4];
      auto load4 = [&](Vec8<kF32>* dst, int p) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = p + u * rows_per_iter;
          if (pp < p_end) dst[u] = load_vec<kF32>(s, (long long)b * hw + pp, v * 8);
          else dst[u].zero();
        }
      };
      load4(cur, p_begin + prow);
      for (int p = p_begin + prow; p < p_end; p += step) {
        load4(nxt, p + step);  // past the end: zeros, no memory access
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          cur[u].to_float(f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sum[0][e] += f[e];
            sq[0][e] += f[e] * f[e];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        cells[prow * cpad + v * 8 + e] = sum[0][e];
        cells[(n_rows + prow) * cpad + v * 8 + e] = sq[0][e];
      }
    }
  } else {
    n_rows = 1;
    for (int p = p_begin; p < p_end; ++p) {
#pragma unroll
      for (int q = 0; q < MAX_SLOTS; ++q) {
        const int v = threadIdx.x + q * NORM_THREADS;
        if (v < cv) {
          float f[8];
          load_vec<kF32>(s, (long long)b * hw + p, v * 8).to_float(f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sum[q][e] += f[e];
            sq[q][e] += f[e] * f[e];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < MAX_SLOTS; ++q) {
      const int v = threadIdx.x + q * NORM_THREADS;
      if (v < cv) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          cells[v * 8 + e] = sum[q][e];
          cells[cpad + v * 8 + e] = sq[q][e];
        }
      }
    }
  }
  __syncthreads();
  // fixed-order reduction: warp w owns groups w, w + 8, ...; lane l adds cells l, l + 32, ... then a shuffle tree
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my_partial = partial + ((size_t)b * gridDim.x + blockIdx.x) * (2 * stats_groups);
  for (int g = warp; g < groups; g += NORM_THREADS / 32) {
    const int c_lo = g * gs, c_hi = min(ctot, c_lo + gs);
    const int width = c_hi - c_lo;
    const int n_cells = n_rows * width;
    float a = 0.f, q2 = 0.f;
    for (int i = lane; i < n_cells; i += 32) {
      const int r = i / width, c = c_lo + (i - r * width);
      a += cells[r * cpad + c];
      q2 += cells[(n_rows + r) * cpad + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      q2 += __shfl_xor_sync(0xffffffffu, q2, o);
    }
    if (lane == 0) {
      my_partial[2 * g] = a;
      my_partial[2 * g + 1] = q2;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[b], 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* pb = partial + (size_t)b * gridDim.x * (2 * stats_groups);
  for (int i = threadIdx.x; i < 2 * groups; i += NORM_THREADS) {
    float acc = 0.f;
    for (unsigned int ch = 0; ch < gridDim.x; ++ch) acc += __ldcg(pb + (size_t)ch * (2 * stats_groups) + i);
    stats[(size_t)b * stats_groups * 2 + i] = acc;
  }
  if (threadIdx.x == 0) tickets[b] = 0u;  // ready for the next launch on this stream
}

template <bool kF32>
__global__ void __launch_bounds__(NORM_THREADS)
    gn_apply_kernel(Src2 s, __nv_bfloat16* __restrict__ y, int ldy, int hw, int gs, float eps,
                    const float* __restrict__ stats, int stats_groups, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int affine_ld, const int* __restrict__ sample_seg,
                    const int* __restrict__ sample_channels, const float* __restrict__ gate, int gate_ld, int silu,
                    int pix_per_cta, __nv_bfloat16* __restrict__ raw, int ld_raw) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y;
  const int ctot = sample_channels ? sample_channels[b] : (s.c0 + s.c1);
  if (ctot <= 0) return;
  int cstore = (ctot + 63) & ~63;
  if (cstore > ldy) cstore = ldy;
  const int cv = (cstore + 7) / 8;
  const int seg = sample_seg ? sample_seg[b] : 0;
  const float* gm = gamma + (size_t)seg * affine_ld;
  const float* bt = beta + (size_t)seg * affine_ld;
  const float inv_n = 1.f / ((float)hw * (float)gs);
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);

  const int slots = (cv + NORM_THREADS - 1) / NORM_THREADS;
  const int rows_per_iter = (cv <= NORM_THREADS) ? NORM_THREADS / cv : 1;
  for (int q = 0; q < slots; ++q) {
    const int v = (cv <= NORM_THREADS) ? (int)(threadIdx.x % cv) : (int)(threadIdx.x + q * NORM_THREADS);
    const int prow = (cv <= NORM_THREADS) ? (int)(threadIdx.x / cv) : 0;
    if (v >= cv || prow >= rows_per_iter) continue;
    // per-channel scale/shift folded from (gate, mean, rstd, gamma, beta)
    float a[8], sft[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      if (c < ctot) {
        const int g = c / gs;
        const float su = stats[((size_t)b * stats_groups + g) * 2];
        const float ss = stats[((size_t)b * stats_groups + g) * 2 + 1];
        const float gt = gate ? gate[(size_t)b * gate_ld + g] : 1.f;
        const float mean = su * inv_n;
        float var = ss * inv_n - mean * mean;
        var = fmaxf(var, 0.f);
        // stats are of the un-gated tensor; GroupNorm(g*x) = (g*x - g*mean) * rsqrt(g^2 var + eps)
        const float rstd = rsqrtf(gt * gt * var + eps);
        const float w = gm[c] * rstd * gt;
        a[e] = w;
        sft[e] = bt[c] - mean * w;
      } else {
        a[e] = 0.f;
        sft[e] = 0.f;
      }
    }
    const bool has_data = v * 8 < ctot;
    const int step = 4 * rows_per_iter;
    Vec8<kF32> cur[4], nxt[4];
    auto load4 = [&](Vec8<kF32>* dst, int p) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + u * rows_per_iter;
        if (has_data && pp < p_end) dst[u] = load_vec<kF32>(s, (long long)b * hw + pp, v * 8);
        else dst[u].zero();
      }
    };
    load4(cur, p_begin + prow);
    for (int p = p_begin + prow; p < p_end; p += step) {
      load4(nxt, p + step);  // next step's loads in flight under this step's math and stores
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + u * rows_per_iter;
        if (pp >= p_end) break;
        float f[8];
        cur[u].to_float(f);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = f[e] * a[e] + sft[e];
          if (silu) t = APTP_GN_SILU(t);
          o[e] = (v * 8 + e < ctot) ? t : 0.f;  // zero K padding even if stale memory holds inf/nan
        }
        uint4 out;
        out.x = pack_bf16(o[0], o[1]);
        out.y = pack_bf16(o[2], o[3]);
        out.z = pack_bf16(o[4], o[5]);
        out.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(y + ((long long)b * hw + pp) * ldy + v * 8) = out;
        if (raw != nullptr) {  // bf16 copy of the UN-normalised input (the 1x1 shortcut's A operand), same pass
          uint4 rw;
          rw.x = pack_bf16(v * 8 + 0 < ctot ? f[0] : 0.f, v * 8 + 1 < ctot ? f[1] : 0.f);
          rw.y = pack_bf16(v * 8 + 2 < ctot ? f[2] : 0.f, v * 8 + 3 < ctot ? f[3] : 0.f);
          rw.z = pack_bf16(v * 8 + 4 < ctot ? f[4] : 0.f, v * 8 + 5 < ctot ? f[5] : 0.f);
          rw.w = pack_bf16(v * 8 + 6 < ctot ? f[6] : 0.f, v * 8 + 7 < ctot ? f[7] : 0.f);
          *reinterpret_cast<uint4*>(raw + ((long long)b * hw + pp) * ld_raw + v * 8) = rw;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    }
  }
}

// One warp per token row at a time, SLOTS 16-byte vectors per lane held in registers (exact two-pass). Warps are
// persistent (grid-stride over rows) and the NEXT row's vectors are already in flight while the current row is
// reduced, normalised and stored: with C = 320 a lane only has 1-2 loads of its own row to keep in flight, far
// too few to cover HBM latency. SLOTS is a template parameter so narrow rows keep the register count low.
constexpr int LN_MAXV = 8;
template <int SLOTS>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                        __nv_bfloat16* __restrict__ y, int ldy, long long rows, int C,
                                                        float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const uint8_t* __restrict__ sample_active,
                                                        int rows_per_sample) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * 8;
  const int cv = C / 8;
  const float inv_c = 1.f / (float)C;
  auto load_row = [&](uint4* dst, long long row) {
    const bool on = row < rows && (!sample_active || sample_active[row / rows_per_sample]);
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const int v = lane + q * 32;
      dst[q] = (on && v < cv) ? __ldg(reinterpret_cast<const uint4*>(x + row * ldx + v * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  // narrow rows: the affine parameters of this lane's channels stay in registers for all rows; wide rows
  // re-read them (L1-resident) to keep the register count, and with it the resident warps, where HBM needs them
  constexpr bool kAffineInRegs = SLOTS <= 2;
  float gg[kAffineInRegs ? SLOTS : 1][8], bb[kAffineInRegs ? SLOTS : 1][8];
  if constexpr (kAffineInRegs) {
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const int v = lane + q * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gg[q][e] = (v < cv) ? __ldg(gamma + v * 8 + e) : 0.f;
        bb[q][e] = (v < cv) ? __ldg(beta + v * 8 + e) : 0.f;
      }
    }
  }
  long long row = (long long)blockIdx.x * 8 + warp;
  uint4 cur[SLOTS], nxt[SLOTS];
  load_row(cur, row);
  for (; row < rows; row += stride) {
    load_row(nxt, row + stride);
    const bool on = !sample_active || sample_active[row / rows_per_sample];
    if (on) {
      float f[SLOTS][8];
      float sm = 0.f;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        unpack8(cur[q], f[q]);
#pragma unroll
        for (int e = 0; e < 8; ++e) sm += f[q][e];  // padded vectors are zero
      }
      const float mean = warp_sum(sm) * inv_c;
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        const int v = lane + q * 32;
        if (v < cv) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = f[q][e] - mean;
            ss += d * d;
          }
        }
      }
      const float rstd = rsqrtf(warp_sum(ss) * inv_c + eps);
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        const int v = lane + q * 32;
        if (v < cv) {
          float o[8];
          if constexpr (kAffineInRegs) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (f[q][e] - mean) * rstd * gg[q][e] + bb[q][e];
          } else {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
            const float gl[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bl[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (f[q][e] - mean) * rstd * gl[e] + bl[e];
          }
          uint4 out;
          out.x = pack_bf16(o[0], o[1]);
          out.y = pack_bf16(o[2], o[3]);
          out.z = pack_bf16(o[4], o[5]);
          out.w = pack_bf16(o[6], o[7]);
          *reinterpret_cast<uint4*>(y + row * ldy + v * 8) = out;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) cur[q] = nxt[q];
  }
}

// LayerNorm for the SD-2.1 widths: C / 8 = 5 * LPR 16-byte vectors per row with LPR in {8, 16, 32}
// (C = 320, 640, 1280): LPR lanes share a row with exactly 5 vectors each (no idle lane slots, unlike the
// 32-lanes-per-row mapping above where C = 320 fills 40 of 64 slots), so a warp works on 32 / LPR rows at once and
// keeps 5 useful loads per lane in flight; warps are persistent and the next row group is prefetched.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm5_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                         __nv_bfloat16* __restrict__ y, int ldy, long long rows, int C,
                                                         float eps, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta,
                                                         const uint8_t* __restrict__ sample_active,
                                                         int rows_per_sample) {
  constexpr int RPW = 32 / LPR;  // rows per warp per step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long stride = (long long)gridDim.x * 8 * RPW;
  const float inv_c = 1.f / (float)C;
  auto load_row = [&](uint4* dst, long long row) {
    const bool on = row < rows && (!sample_active || sample_active[row / rows_per_sample]);
#pragma unroll
    for (int q = 0; q < 5; ++q)
      dst[q] = on ? __ldg(reinterpret_cast<const uint4*>(x + row * ldx + (l + q * LPR) * 8)) : make_uint4(0u, 0u, 0u, 0u);
  };
  long long row = ((long long)blockIdx.x * 8 + warp) * RPW + sub;
  uint4 cur[5], nxt[5];
  load_row(cur, row);
  for (; row - sub < rows; row += stride) {  // warp-uniform trip count (row - sub is the warp's first row)
    load_row(nxt, row + stride);
    const bool on = row < rows && (!sample_active || sample_active[row / rows_per_sample]);
    float f[5][8];
    float sm = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      unpack8(cur[q], f[q]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sm += f[q][e];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    const float mean = sm * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[q][e] - mean;
        ss += d * d;
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
    if (on) {
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int c = (l + q * LPR) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        const float gl[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bl[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (f[q][e] - mean) * rstd * gl[e] + bl[e];
        uint4 out;
        out.x = pack_bf16(o[0], o[1]);
        out.y = pack_bf16(o[2], o[3]);
        out.z = pack_bf16(o[4], o[5]);
        out.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(y + row * ldy + c) = out;
      }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) cur[q] = nxt[q];
  }
}

// GroupNorm statistics from the per-channel partials written by GEMM epilogues (APTP_EPI_GN_STATS): one warp per
// (sample, group) adds the group's (32-row block, channel) cells in a fixed order + a fixed shuffle tree. The partial
// planes are 1/8 (fp32 stream: 1/16 per plane) of the tensor itself, so this replaces a full statistics pass over HBM.
constexpr int GNF_SLICES = 4;  // threads per channel: each walks every 4th 32-row block (more loads in flight)
__global__ void __launch_bounds__(256 * GNF_SLICES)
    gn_partials_finalize_kernel(const float* __restrict__ sum0, const float* __restrict__ sq0, int c0, int ld0,
                                const float* __restrict__ sum1, const float* __restrict__ sq1, int ld1, int ctot_all,
                                int blocks, int gs, int groups_per_cta, const int* __restrict__ sample_channels,
                                float* __restrict__ stats, int stats_groups) {
  // thread (x, y) = channel x, block slice y (coalesced walks down the 32-row blocks y, y + 4, ...: four independent
  // accumulators each, so 16 loads per plane are in flight per channel); then one thread per group adds the slices and
  // its gs channel totals in a fixed order: every sum has a fixed order
  __shared__ float ssum[GNF_SLICES][256], ssq[GNF_SLICES][256];
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x;
  const int ctot = sample_channels ? sample_channels[b] : ctot_all;
  if (ctot <= 0) return;
  const int groups = (ctot + gs - 1) / gs;
  const int g0 = blockIdx.y * groups_per_cta;
  if (g0 >= groups) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = g0 * gs + tx;
  float a = 0.f, q = 0.f;
  if (tx < groups_per_cta * gs && c < ctot) {
    const float* ps;
    const float* pq;
    int ld;
    if (c < c0) {
      ps = sum0 + c; pq = sq0 + c; ld = ld0;
    } else {
      ps = sum1 + (c - c0); pq = sq1 + (c - c0); ld = ld1;
    }
    const size_t row0 = (size_t)b * blocks;
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    int blk = ty;
    for (; blk + 3 * GNF_SLICES < blocks; blk += 4 * GNF_SLICES) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a4[u] += __ldg(ps + (row0 + blk + u * GNF_SLICES) * ld);
        q4[u] += __ldg(pq + (row0 + blk + u * GNF_SLICES) * ld);
      }
    }
    for (; blk < blocks; blk += GNF_SLICES) {
      a4[0] += __ldg(ps + (row0 + blk) * ld);
      q4[0] += __ldg(pq + (row0 + blk) * ld);
    }
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  }
  ssum[ty][tx] = a;
  ssq[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && tx < groups_per_cta && g0 + tx < groups) {
    float ga = 0.f, gq = 0.f;
    for (int e = 0; e < gs; ++e) {
#pragma unroll
      for (int y = 0; y < GNF_SLICES; ++y) {
        ga += ssum[y][tx * gs + e];
        gq += ssq[y][tx * gs + e];
      }
    }
    stats[((size_t)b * stats_groups + g0 + tx) * 2] = ga;
    stats[((size_t)b * stats_groups + g0 + tx) * 2 + 1] = gq;
  }
}

// LayerNorm row statistics for the LN-fold GEMM epilogue: (mean, rstd) of every row from the per-chunk (sum, sumsq)
// partials the producing GEMM wrote, summed in chunk order (deterministic). One thread per row.
__global__ void __launch_bounds__(256) ln_rowstats_kernel(const float2* __restrict__ partial, int chunks, long long rows,
                                                          float inv_c, float eps, float2* __restrict__ out,
                                                          const uint8_t* __restrict__ sample_active, int rows_per_sample) {
  pdl_launch();
  pdl_wait();
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  if (sample_active && !sample_active[row / rows_per_sample]) return;
  const float4* pp = reinterpret_cast<const float4*>(partial + row * chunks);
  float su = 0.f, sq = 0.f;
  for (int i = 0; i < chunks / 2; ++i) {
    const float4 q = __ldg(pp + i);
    su += q.x + q.z;
    sq += q.y + q.w;
  }
  const float mu = su * inv_c;
  out[row] = make_float2(mu, rsqrtf(fmaxf(sq * inv_c - mu * mu, 0.f) + eps));
}

static int pick_pix_per_cta(int hw, int batch) {
  // ~8 CTAs per SM across the grid (two waves at the register-limited occupancy), at least 16 pixels per CTA
  int target_ctas = 8 * sm_count();
  int chunks = (target_ctas + batch - 1) / batch;
  if (chunks < 1) chunks = 1;
  int ppc = (hw + chunks - 1) / chunks;
  if (ppc < 16) ppc = 16;
  if (ppc > hw) ppc = hw;
  return ppc;
}

}  // namespace aptp

using namespace aptp;

static int check_src(const void* x0, int c0, int ld0, const void* x1, int c1, int ld1, const char* who) {
  APTP_REQUIRE(x0 != nullptr && c0 > 0, "%s: x0 is null", who);
  APTP_REQUIRE(c0 % 8 == 0 && ld0 % 8 == 0, "%s: c0/ld0 must be multiples of 8", who);
  APTP_REQUIRE(c1 == 0 || (x1 != nullptr && c1 % 8 == 0 && ld1 % 8 == 0), "%s: bad second source", who);
  APTP_REQUIRE(c0 + c1 <= NORM_THREADS * MAX_SLOTS * 8, "%s: too many channels (%d)", who, c0 + c1);
  return APTP_OK;
}

// workspace of aptp_groupnorm_stats: [batch] ticket counters (zero on entry, left zero) + per-CTA partials
static long long gn_ws_bytes(int batch, int hw, int stats_groups) {
  const int ppc = pick_pix_per_cta(hw, batch);
  const long long chunks = (hw + ppc - 1) / ppc;
  const long long tickets = ((long long)batch * 4 + 255) & ~255LL;
  return tickets + (long long)batch * chunks * 2 * stats_groups * 4;
}

extern "C" int64_t aptp_groupnorm_stats_workspace(int32_t batch, int32_t hw, int32_t stats_groups) {
  if (batch <= 0 || hw <= 0 || stats_groups <= 0) return 0;
  return gn_ws_bytes(batch, hw, stats_groups);
}

extern "C" int aptp_groupnorm_stats(const void* x0, int32_t c0, int32_t ld0, const void* x1, int32_t c1, int32_t ld1,
                                    int32_t x_f32, int32_t batch, int32_t hw, int32_t group_size,
                                    const int32_t* sample_channels, float* stats, int32_t stats_groups,
                                    void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_src(x0, c0, ld0, x1, c1, ld1, "aptp_groupnorm_stats");
  if (rc) return rc;
  APTP_REQUIRE(stats && group_size > 0 && batch > 0 && hw > 0, "aptp_groupnorm_stats: bad arguments");
  APTP_REQUIRE((c0 + c1 + group_size - 1) / group_size <= stats_groups, "aptp_groupnorm_stats: stats_groups too small");
  APTP_REQUIRE(workspace && workspace_bytes >= gn_ws_bytes(batch, hw, stats_groups) &&
                   (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "aptp_groupnorm_stats: workspace too small (%lld bytes needed, see aptp_groupnorm_stats_workspace)",
               gn_ws_bytes(batch, hw, stats_groups));
  Src2 s{x0, x1, c0, ld0, c1, ld1};
  const int ppc = pick_pix_per_cta(hw, batch);
  dim3 grid((hw + ppc - 1) / ppc, batch);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (((long long)batch * 4 + 255) & ~255LL));
  const int C = c0 + c1;
  const size_t smem = (size_t)2 * (C > NORM_THREADS * 8 ? C : NORM_THREADS * 8) * sizeof(float);
  if (x_f32)
    APTP_CUDA_CHECK(launch_pdl_at(2, gn_stats_kernel<true>, grid, dim3(NORM_THREADS), smem, stream, s, hw, group_size,
                               sample_channels, stats, stats_groups, ppc, tickets, partial));
  else
    APTP_CUDA_CHECK(launch_pdl_at(2, gn_stats_kernel<false>, grid, dim3(NORM_THREADS), smem, stream, s, hw, group_size,
                               sample_channels, stats, stats_groups, ppc, tickets, partial));
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_groupnorm_stats_from_partials(const float* sum0, const float* sq0, int32_t c0, int32_t ld0,
                                                  const float* sum1, const float* sq1, int32_t c1, int32_t ld1,
                                                  int32_t blocks, int32_t batch, int32_t group_size,
                                                  const int32_t* sample_channels, float* stats, int32_t stats_groups,
                                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(sum0 && sq0 && c0 > 0 && ld0 >= c0 && stats && blocks > 0 && batch > 0 && group_size > 0,
               "aptp_groupnorm_stats_from_partials: bad arguments");
  APTP_REQUIRE(c1 == 0 || (sum1 && sq1 && ld1 >= c1), "aptp_groupnorm_stats_from_partials: bad second source");
  const int groups = (c0 + c1 + group_size - 1) / group_size;
  APTP_REQUIRE(groups <= stats_groups, "aptp_groupnorm_stats_from_partials: stats_groups too small");
  APTP_REQUIRE(group_size <= 256, "aptp_groupnorm_stats_from_partials: group_size %d > 256", group_size);
  const int gpc = 256 / group_size;  // whole groups per CTA
  APTP_CUDA_CHECK(launch_pdl_at(2, gn_partials_finalize_kernel, dim3(batch, (groups + gpc - 1) / gpc), dim3(256, GNF_SLICES),
                             (size_t)0, stream, sum0, sq0, c0, ld0, sum1, sq1, ld1, c0 + c1, blocks, group_size, gpc,
                             sample_channels, stats, stats_groups));
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_groupnorm_apply(const void* x0, int32_t c0, int32_t ld0, const void* x1, int32_t c1, int32_t ld1,
                                    int32_t x_f32, void* y, int32_t ldy, int32_t batch, int32_t hw,
                                    int32_t group_size, float eps, const float* stats, int32_t stats_groups,
                                    const float* gamma, const float* beta, int32_t affine_ld,
                                    const int32_t* sample_seg, const int32_t* sample_channels, const float* gate,
                                    int32_t gate_ld, int32_t silu, void* stream_) {
  return aptp_groupnorm_apply_raw(x0, c0, ld0, x1, c1, ld1, x_f32, y, ldy, batch, hw, group_size, eps, stats, stats_groups,
                                  gamma, beta, affine_ld, sample_seg, sample_channels, gate, gate_ld, silu, nullptr, 0,
                                  stream_);
}

extern "C" int aptp_groupnorm_apply_raw(const void* x0, int32_t c0, int32_t ld0, const void* x1, int32_t c1, int32_t ld1,
                                        int32_t x_f32, void* y, int32_t ldy, int32_t batch, int32_t hw,
                                        int32_t group_size, float eps, const float* stats, int32_t stats_groups,
                                        const float* gamma, const float* beta, int32_t affine_ld,
                                        const int32_t* sample_seg, const int32_t* sample_channels, const float* gate,
                                        int32_t gate_ld, int32_t silu, void* raw_out, int32_t raw_ld, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(raw_out == nullptr || (raw_ld % 8 == 0 && raw_ld >= c0 + c1), "aptp_groupnorm_apply_raw: bad raw_ld");
  int rc = check_src(x0, c0, ld0, x1, c1, ld1, "aptp_groupnorm_apply");
  if (rc) return rc;
  APTP_REQUIRE(y && stats && gamma && beta && ldy % 8 == 0, "aptp_groupnorm_apply: bad arguments");
  Src2 s{x0, x1, c0, ld0, c1, ld1};
  const int ppc = pick_pix_per_cta(hw, batch);
  dim3 grid((hw + ppc - 1) / ppc, batch);
  if (x_f32)
    APTP_CUDA_CHECK(launch_pdl_at(2, gn_apply_kernel<true>, grid, dim3(NORM_THREADS), (size_t)0, stream, s,
                               reinterpret_cast<__nv_bfloat16*>(y), ldy, hw, group_size, eps, stats, stats_groups, gamma, beta,
                               affine_ld, sample_seg, sample_channels, gate, gate_ld, silu, ppc,
                               reinterpret_cast<__nv_bfloat16*>(raw_out), raw_ld));
  else
    APTP_CUDA_CHECK(launch_pdl_at(2, gn_apply_kernel<false>, grid, dim3(NORM_THREADS), (size_t)0, stream, s,
                               reinterpret_cast<__nv_bfloat16*>(y), ldy, hw, group_size, eps, stats, stats_groups, gamma, beta,
                               affine_ld, sample_seg, sample_channels, gate, gate_ld, silu, ppc,
                               reinterpret_cast<__nv_bfloat16*>(raw_out), raw_ld));
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_ln_rowstats(const float* partial, int32_t chunks, int64_t rows, int32_t C, float eps, float* out,
                                const uint8_t* sample_active, int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(partial && out && chunks > 0 && chunks % 2 == 0 && C > 0 && rows_per_sample > 0 &&
                   (reinterpret_cast<uintptr_t>(partial) & 15) == 0,
               "aptp_ln_rowstats: bad arguments (chunks must be even, partial 16-byte aligned)");
  if (rows == 0) return APTP_OK;
  APTP_CUDA_CHECK(launch_pdl_at(2, ln_rowstats_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), (size_t)0, stream,
                             reinterpret_cast<const float2*>(partial), chunks, (long long)rows, 1.f / (float)C, eps,
                             reinterpret_cast<float2*>(out), sample_active, rows_per_sample));
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_layernorm(const void* x, int32_t ldx, void* y, int32_t ldy, int64_t rows, int32_t C, float eps,
                              const float* gamma, const float* beta, const uint8_t* sample_active,
                              int32_t rows_per_sample, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(x && y && gamma && beta, "aptp_layernorm: null pointer");
  APTP_REQUIRE(C % 8 == 0 && C <= 32 * 8 * LN_MAXV && ldx % 8 == 0 && ldy % 8 == 0, "aptp_layernorm: unsupported C=%d", C);
  APTP_REQUIRE(rows_per_sample > 0, "aptp_layernorm: rows_per_sample must be > 0");
  if (rows == 0) return APTP_OK;
  const long long max_blocks = 8LL * sm_count();  // persistent warps: each walks rows with the next one prefetched
  const int lpr = (C % 40 == 0) ? C / 40 : 0;
  if (lpr == 8) {  // C = 320 (measured: wider rows are as fast on the 32-lanes-per-row kernel)
    const int rpw = 32 / lpr;
    long long blocks = (rows + 8 * rpw - 1) / (8 * rpw);
    if (blocks > max_blocks) blocks = max_blocks;
#define APTP_LN5_LAUNCH(L)                                                                                          \
  layernorm5_kernel<L><<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx,       \
                                                             reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, C, eps, \
                                                             gamma, beta, sample_active, rows_per_sample)
    APTP_LN5_LAUNCH(8);
#undef APTP_LN5_LAUNCH
    APTP_CUDA_CHECK(cudaGetLastError());
    return APTP_OK;
  }
  long long blocks = (rows + 7) / 8;
  if (blocks > max_blocks) blocks = max_blocks;
  const int slots = (C / 8 + 31) / 32;
#define APTP_LN_LAUNCH(S)                                                                                          \
  layernorm_kernel<S><<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx,       \
                                                            reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, C, eps, \
                                                            gamma, beta, sample_active, rows_per_sample)
  if (slots <= 1) APTP_LN_LAUNCH(1);
  else if (slots == 2) APTP_LN_LAUNCH(2);
  else if (slots == 3) APTP_LN_LAUNCH(3);
  else if (slots <= 5) APTP_LN_LAUNCH(5);
  else APTP_LN_LAUNCH(8);
#undef APTP_LN_LAUNCH
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
