// K8: weight gradients of the path's convolutions / linears on tcgen05 (the kernel the fine-tune
// step of a compacted expert needs, pdm/training/trainer.py:1683-1765 -> autograd of F.conv2d / F.linear w.r.t.
// the weight; the pruning stage itself keeps the U-Net frozen, unet_2d_conditional.py:2118-2122).
//
//   dW[n, tap, c] (+)= sum_rows dY[row, n] * A[shift_tap(row), c]        (OHWI layout, taps = 1 or 9)
//
// The reduction runs over rows (= batch x pixels), so BOTH operands are consumed MN-major straight from their natural
// NHWC / token-major layout: a TMA box of [128 rows x 64 columns] (128B-swizzled) is an MN-major operand with K = 128
// rows, atoms of 8 rows x 128 B (SBO 1024), 64-column atoms LBO apart. 3x3 taps are the same shifted 4-D boxes over the
// NHWC activation the forward implicit GEMM uses (zero padding = TMA out-of-bounds fill); dY is read through the
// matching unshifted box so both enumerate the pixels in the same order.
//
//   grid = (n tiles of 128) x (c tiles of 128) x taps, split-K over row stages in gridDim.y
//   warp 0: TMA producer (3-stage ring: dY 2 x 16 KB | A 2 x 16 KB)    warp 1: MMA issuer, M128 x N128 x K16, fp32 in TMEM
//   warps 2-5: epilogue, one accumulator row (= output channel n) per thread, fp32 atomic adds into dW
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int WG_THREADS = 192;
constexpr int WG_ROWS = 128;                   // rows (K of the MMA) per stage
constexpr int WG_ATOM_BYTES = WG_ROWS * 128;   // [128 rows x 64 bf16]
constexpr int WG_STAGE_BYTES = 4 * WG_ATOM_BYTES;
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM_BAR = WG_STAGES * WG_STAGE_BYTES;
constexpr int WG_SMEM_BYTES = WG_SMEM_BAR + 256 + 1024;

struct WgradParams {
  CUtensorMap tmap_dy, tmap_a;
  float* dw;
  long long ld_dw;
  int n_out, k_in, taps;
  int conv;               // 0: linear (2-D maps), 1: 3x3 stride-1 conv (4-D maps), 2: 3x3 stride-2 conv (5-D map of a)
  int ld_a;               // pitch of a (stride-2 view: the two x-parities sit ld_a apart in the fastest dimension)
  int bw, bh, bb;         // pixel box of a stage (bw * bh * bb == 128)
  int Wt, Ht, Bt;         // boxes per image row / column / batch
  int n_stages;           // row stages in total
  int stages_per_split;
  int* abort_flag;
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_SMEM_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* acc_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int n_ct = (p.k_in + 127) / 128;
  const int tap = blockIdx.x % p.taps;
  const int ct = (blockIdx.x / p.taps) % n_ct;
  const int nt = blockIdx.x / (p.taps * n_ct);
  const int n0 = nt * 128, c0 = ct * 128;
  const int s_begin = blockIdx.y * p.stages_per_split;
  const int s_end = min(p.n_stages, s_begin + p.stages_per_split);
  const int n_it = s_end - s_begin;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < WG_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(acc_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (n_it <= 0) {
    // nothing to reduce in this split (uniform per CTA): fall through to the dealloc
  } else if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;  // only used in conv mode (taps == 9)
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_it; ++it) {
      if (!mbar_wait(&empty_bar[stage], phase ^ 1, p.abort_flag)) break;
      uint8_t* st = smem + stage * WG_STAGE_BYTES;
      const int s = s_begin + it;
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
        if (p.conv) {
          const int bx = s % p.Wt, by = (s / p.Wt) % p.Ht, bz = s / (p.Wt * p.Ht);
          const int x0 = bx * p.bw, y0 = by * p.bh, b0 = bz * p.bb;  // box origin in OUTPUT pixels
          tma_load_4d(st, &p.tmap_dy, &full_bar[stage], n0, x0, y0, b0);
          tma_load_4d(st + WG_ATOM_BYTES, &p.tmap_dy, &full_bar[stage], n0 + 64, x0, y0, b0);
          if (p.conv == 1) {
            tma_load_4d(st + 2 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], c0, x0 + dx, y0 + dy, b0);
            tma_load_4d(st + 3 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], c0 + 64, x0 + dx, y0 + dy, b0);
          } else {
            // input y = 2 oy + dy: dy = -1 -> (parity 1, shift -1); 0 -> (0, 0); +1 -> (1, 0)   (same for x)
            const int py = (dy == 0) ? 0 : 1, sy = (dy < 0) ? -1 : 0;
            const int px = (dx == 0) ? 0 : 1, sx = (dx < 0) ? -1 : 0;
            tma_load_5d(st + 2 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], px * p.ld_a + c0, x0 + sx, py, y0 + sy, b0);
            tma_load_5d(st + 3 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], px * p.ld_a + c0 + 64, x0 + sx, py, y0 + sy,
                        b0);
          }
        } else {
          const int r0 = s * WG_ROWS;
          tma_load_2d(st, &p.tmap_dy, &full_bar[stage], n0, r0);
          tma_load_2d(st + WG_ATOM_BYTES, &p.tmap_dy, &full_bar[stage], n0 + 64, r0);
          tma_load_2d(st + 2 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], c0, r0);
          tma_load_2d(st + 3 * WG_ATOM_BYTES, &p.tmap_a, &full_bar[stage], c0 + 64, r0);
        }
      }
      __syncwarp();
      if (++stage == WG_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);  // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int it = 0; it < n_it; ++it) {
      if (!mbar_wait(&full_bar[stage], phase, p.abort_flag)) {
        ok = false;
        break;
      }
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * WG_STAGE_BYTES);
      const uint32_t b_addr = a_addr + 2 * WG_ATOM_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WG_ROWS / 16; ++k)  // 16 rows per step = two 8-row atoms of 1024 B
          umma_bf16_ss(tmem_base, make_desc_mnmajor_sw128(a_addr + k * 2048, WG_ATOM_BYTES),
                       make_desc_mnmajor_sw128(b_addr + k * 2048, WG_ATOM_BYTES), idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == WG_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (ok) umma_commit_e(acc_bar);
  } else {
    // ------------------------------- epilogue ------------------------------------
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // accumulator row = output channel n0 + m
    if (mbar_wait(acc_bar, 0, p.abort_flag)) {
      tc_fence_after();
      const uint32_t t = tmem_base + ((uint32_t)(quad * 32) << 16);
      float* row = p.dw + (long long)(n0 + m) * p.ld_dw + (long long)tap * p.k_in + c0;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t + c * 32, v);
        tmem_ld_wait();
        if (n0 + m < p.n_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + c * 32 + j < p.k_in) atomicAdd(row + c * 32 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// db[n] += sum_rows dy[row, n]: one CTA per (64-column slab, row chunk); 8 lanes x 16 B cover the slab, 32 row lanes
__global__ void __launch_bounds__(256) col_sum_kernel(const __nv_bfloat16* __restrict__ dy, int ld, long long rows, int n_out,
                                                      float* __restrict__ db, int rows_per_cta) {
  __shared__ float red[32][65];
  const int v = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + v * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_cta;
  const long long r_end = min(rows, r_begin + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n_out) {
    for (long long r = r_begin + rl; r < r_end; r += 128) {  // 4 independent 16-byte loads in flight per thread
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        q[u] = (r + 32 * u < r_end) ? __ldg(reinterpret_cast<const uint4*>(dy + (r + 32 * u) * ld + col))
                                    : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16_lo(q[u].x); acc[1] += bf16_hi(q[u].x);
        acc[2] += bf16_lo(q[u].y); acc[3] += bf16_hi(q[u].y);
        acc[4] += bf16_lo(q[u].z); acc[5] += bf16_hi(q[u].z);
        acc[6] += bf16_lo(q[u].w); acc[7] += bf16_hi(q[u].w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][v * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < n_out) atomicAdd(db + c, s);
  }
}

// out[g, n] += sum over the rows of group g (rows_per_group consecutive rows) of dy[row, n]
__global__ void __launch_bounds__(256) col_sum_groups_kernel(const __nv_bfloat16* __restrict__ dy, int ld, int rows_per_group,
                                                             int n_out, float* __restrict__ out, int out_ld, int chunks,
                                                             int rows_per_cta) {
  __shared__ float red[32][65];
  const int v = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + v * 8;
  const int grp = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
  const long long base = (long long)grp * rows_per_group;
  const int r_begin = chunk * rows_per_cta;
  const int r_end = min(rows_per_group, r_begin + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n_out) {
    for (int r = r_begin + rl; r < r_end; r += 128) {  // 4 independent 16-byte loads in flight per thread
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        q[u] = (r + 32 * u < r_end) ? __ldg(reinterpret_cast<const uint4*>(dy + (base + r + 32 * u) * ld + col))
                                    : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16_lo(q[u].x); acc[1] += bf16_hi(q[u].x);
        acc[2] += bf16_lo(q[u].y); acc[3] += bf16_hi(q[u].y);
        acc[4] += bf16_lo(q[u].z); acc[5] += bf16_hi(q[u].z);
        acc[6] += bf16_lo(q[u].w); acc[7] += bf16_hi(q[u].w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][v * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < n_out) atomicAdd(out + (size_t)grp * out_ld + c, s);
  }
}

static int g_wg_smem_set = 0;

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_wgrad(const void* dy, int32_t ld_dy, const void* a, int32_t ld_a, float* dw, int64_t ld_dw,
                          float* dbias, int64_t rows, int32_t n_out, int32_t k_in, int32_t conv3x3, int32_t batch,
                          int32_t H, int32_t W, int32_t bw, int32_t bh, int32_t bb, int32_t splits, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(dy && a && dw, "aptp_wgrad: null pointer");
  APTP_REQUIRE(ld_dy % 8 == 0 && ld_a % 8 == 0 && n_out > 0 && k_in > 0 && rows >= 0 && splits > 0,
               "aptp_wgrad: pitches must be multiples of 8 and sizes positive");
  APTP_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0,
               "aptp_wgrad: operands must be 16-byte aligned");
  const int taps = conv3x3 ? 9 : 1;
  APTP_REQUIRE(ld_dw >= (int64_t)taps * k_in, "aptp_wgrad: ld_dw too small");
  if (rows == 0) return APTP_OK;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  if (conv3x3) {
    // H, W = INPUT size; the reduction runs over the output pixels (Ho x Wo = H x W, or H/2 x W/2 for stride 2)
    APTP_REQUIRE(conv3x3 == 1 || conv3x3 == 2, "aptp_wgrad: conv3x3 must be 0 (linear), 1 (stride 1) or 2 (stride 2)");
    const int Ho = conv3x3 == 2 ? H / 2 : H, Wo = conv3x3 == 2 ? W / 2 : W;
    APTP_REQUIRE(conv3x3 == 1 || (H % 2 == 0 && W % 2 == 0 && k_in == ld_a),
                 "aptp_wgrad: the stride-2 conv needs even H, W and k_in == ld_a");
    APTP_REQUIRE((int64_t)batch * Ho * Wo == rows, "aptp_wgrad: rows != batch * Ho * Wo");
    APTP_REQUIRE(bw * bh * bb == WG_ROWS && Wo % bw == 0 && Ho % bh == 0,
                 "aptp_wgrad: the pixel box must hold 128 pixels and tile the image");  // a partial batch box is zero-filled
    uint64_t dims_y[4] = {(uint64_t)n_out, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)batch};
    uint64_t str_y[3] = {(uint64_t)ld_dy * 2, (uint64_t)Wo * ld_dy * 2, (uint64_t)Ho * Wo * ld_dy * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    int rc = make_tmap_bf16(&p.tmap_dy, dy, 4, dims_y, str_y, box);
    if (conv3x3 == 1) {
      uint64_t dims_a[4] = {(uint64_t)k_in, (uint64_t)W, (uint64_t)H, (uint64_t)batch};
      uint64_t str_a[3] = {(uint64_t)ld_a * 2, (uint64_t)W * ld_a * 2, (uint64_t)H * W * ld_a * 2};
      rc = rc ? rc : make_tmap_bf16(&p.tmap_a, a, 4, dims_a, str_a, box);
    } else {
      // view [b][H/2][2][W/2][2][C] as dims (fastest first): (px*C + c), W/2, py, H/2, b -- as the forward implicit GEMM
      uint64_t dims_a[5] = {(uint64_t)2 * ld_a, (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)batch};
      uint64_t str_a[4] = {(uint64_t)2 * ld_a * 2, (uint64_t)W * ld_a * 2, (uint64_t)2 * W * ld_a * 2,
                           (uint64_t)H * W * ld_a * 2};
      uint32_t box5[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bb};
      rc = rc ? rc : make_tmap_bf16(&p.tmap_a, a, 5, dims_a, str_a, box5);
    }
    if (rc) return rc;
    p.bw = bw;
    p.bh = bh;
    p.bb = bb;
    p.Wt = Wo / bw;
    p.Ht = Ho / bh;
    p.Bt = (batch + bb - 1) / bb;
    p.n_stages = p.Wt * p.Ht * p.Bt;
  } else {
    uint64_t dims_y[2] = {(uint64_t)n_out, (uint64_t)rows};
    uint64_t str_y[1] = {(uint64_t)ld_dy * 2};
    uint64_t dims_a[2] = {(uint64_t)k_in, (uint64_t)rows};
    uint64_t str_a[1] = {(uint64_t)ld_a * 2};
    uint32_t box[2] = {64, WG_ROWS};
    int rc = make_tmap_bf16(&p.tmap_dy, dy, 2, dims_y, str_y, box);
    rc = rc ? rc : make_tmap_bf16(&p.tmap_a, a, 2, dims_a, str_a, box);
    if (rc) return rc;
    p.n_stages = (int)((rows + WG_ROWS - 1) / WG_ROWS);
  }
  p.conv = conv3x3;
  p.ld_a = ld_a;
  p.dw = dw;
  p.ld_dw = ld_dw;
  p.n_out = n_out;
  p.k_in = k_in;
  p.taps = taps;
  if (splits > p.n_stages) splits = p.n_stages;
  p.stages_per_split = (p.n_stages + splits - 1) / splits;
  splits = (p.n_stages + p.stages_per_split - 1) / p.stages_per_split;
  p.abort_flag = device_abort_flag();
  APTP_REQUIRE(p.abort_flag != nullptr, "aptp_wgrad: could not allocate abort flag");
  if (!g_wg_smem_set) {
    APTP_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    g_wg_smem_set = 1;
  }
  const int tiles = ((n_out + 127) / 128) * ((k_in + 127) / 128) * taps;
  APTP_REQUIRE(splits <= 65535, "aptp_wgrad: too many splits");
  wgrad_kernel<<<dim3(tiles, splits), WG_THREADS, WG_SMEM_BYTES, stream>>>(p);
  APTP_CUDA_CHECK(cudaGetLastError());
  if (dbias) {
    const int rows_per_cta = 1024;  // ~4 CTAs per SM at the 64x64 level: enough loads in flight to cover HBM latency
    dim3 grid((n_out + 63) / 64, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
    col_sum_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), ld_dy, rows, n_out, dbias,
                                            rows_per_cta);
    APTP_CUDA_CHECK(cudaGetLastError());
  }
  return APTP_OK;
}

extern "C" int aptp_col_sum_groups(const void* dy, int32_t ld, int32_t groups, int32_t rows_per_group, int32_t n_out,
                                   float* out, int32_t out_ld, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(dy && out, "aptp_col_sum_groups: null pointer");
  APTP_REQUIRE(ld % 8 == 0 && n_out > 0 && n_out % 8 == 0 && rows_per_group > 0 && out_ld >= n_out,
               "aptp_col_sum_groups: bad sizes");
  if (groups == 0) return APTP_OK;
  const int rows_per_cta = 1024;
  const int chunks = (rows_per_group + rows_per_cta - 1) / rows_per_cta;
  APTP_REQUIRE((long long)groups * chunks <= 65535, "aptp_col_sum_groups: grid too large");
  dim3 grid((n_out + 63) / 64, (unsigned)(groups * chunks));
  col_sum_groups_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), ld, rows_per_group, n_out, out,
                                                 out_ld, chunks, rows_per_cta);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
