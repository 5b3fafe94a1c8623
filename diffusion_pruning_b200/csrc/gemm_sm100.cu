// K1: grouped, expert-bucketed GEMM / implicit-GEMM 3x3 conv for sm_100a.
//
//   out[row, n] = epilogue( sum_{tap, c} A[pixel(row) + tap, c] * W[w_row_off + n, tap*pitch + c] )
//
// One persistent CTA per SM walks a host-built tile list. Warp roles:
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, M=128 x N=bn x K=16)
//   warps 2..5  epilogue      (tcgen05.ld 32x32b from the double-buffered TMEM accumulator,
//                              bias / time-embedding / border table / gate / GEGLU / residual, bf16 store)
// The 3x3 conv is an implicit GEMM: for every tap the A tile is a shifted 4-D TMA box over the NHWC
// activation (out-of-bounds rows/cols are zero-filled by the TMA unit = the conv padding); the
// stride-2 down-sampler conv uses a 5-D view that splits H and W into (index, parity).
// Pruned work is skipped, not multiplied by zero: every expert bucket (segment) has its own kept
// column count, kept K-chunk count and compacted weight block; depth-dropped buckets simply have no
// tiles in the list.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;

struct GemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  const aptp_gemm_seg* segs;
  const aptp_gemm_tile* tiles;
  int n_tiles;
  int a_mode;
  int batch, Ho, Wo;  // OUTPUT spatial size (conv modes)
  int bn, bw, bh, bb;
  int stages;
  int k_tap_pitch;
  void* out;
  int out_ld, out_mode;
  const float* bias;
  const float* rowvec;
  int rowvec_ld, rows_per_sample;
  const __nv_bfloat16* residual;
  int res_ld;
  const float* gate;
  int gate_ld, gate_group;
  const float* border_tab;
  int tab_ld;
  int flags;
  int* abort_flag;
};

__device__ __forceinline__ void advance(int& stage, uint32_t& phase, int stages) {
  if (++stage == stages) {
    stage = 0;
    phase ^= 1;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) grouped_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)p.bn * 128u;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.n_tiles && ok; t += gridDim.x) {
        const aptp_gemm_tile tile = p.tiles[t];
        const aptp_gemm_seg seg = p.segs[tile.seg];
        const int taps = (p.a_mode == APTP_A_LINEAR) ? 1 : 9;
        int img = 0, oy0 = 0, ox0 = 0;
        if (p.a_mode != APTP_A_LINEAR) {
          const int hw = p.Ho * p.Wo;
          img = tile.m_base / hw;
          const int rem = tile.m_base - img * hw;
          oy0 = rem / p.Wo;
          ox0 = rem - oy0 * p.Wo;
        }
        for (int tap = 0; tap < taps && ok; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          for (int kc = 0; kc < seg.k_chunks; ++kc) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1, p.abort_flag)) {
              ok = false;
              break;
            }
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            uint8_t* sb = sa + A_STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            if (p.a_mode == APTP_A_LINEAR) {
              tma_load_2d(sa, &p.tmap_a, &full_bar[stage], kc * BK, tile.m_base);
            } else if (p.a_mode == APTP_A_CONV3X3) {
              tma_load_4d(sa, &p.tmap_a, &full_bar[stage], kc * BK, ox0 + dx - 1, oy0 + dy - 1, img);
            } else {
              // input y = 2*oy + dy - 1: dy=0 -> (parity 1, shift -1); dy=1 -> (0, 0); dy=2 -> (1, 0)
              const int py = (dy == 1) ? 0 : 1, sy = (dy == 0) ? -1 : 0;
              const int px = (dx == 1) ? 0 : 1, sx = (dx == 0) ? -1 : 0;
              tma_load_5d(sa, &p.tmap_a, &full_bar[stage], px * p.k_tap_pitch + kc * BK, ox0 + sx, py, oy0 + sy,
                          img);
            }
            tma_load_2d(sb, &p.tmap_b, &full_bar[stage], tap * p.k_tap_pitch + kc * BK, seg.w_row_off + tile.n0);
            advance(stage, phase, stages);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, (uint32_t)p.bn, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.n_tiles && ok; t += gridDim.x) {
        const aptp_gemm_tile tile = p.tiles[t];
        const aptp_gemm_seg seg = p.segs[tile.seg];
        const int kblocks = ((p.a_mode == APTP_A_LINEAR) ? 1 : 9) * seg.k_chunks;
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.bn);
        for (int kb = 0; kb < kblocks; ++kb) {
          if (!mbar_wait(&full_bar[stage], phase, p.abort_flag)) {
            ok = false;
            break;
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_ss(d_tmem, make_desc_kmajor_sw128(a_addr + k * 32), make_desc_kmajor_sw128(b_addr + k * 32),
                         idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          advance(stage, phase, stages);
        }
        if (!ok) break;
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool geglu = (p.flags & APTP_EPI_GEGLU) != 0;
    const int out_cols_per_tile = geglu ? p.bn / 2 : p.bn;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const aptp_gemm_tile tile = p.tiles[t];
      const aptp_gemm_seg seg = p.segs[tile.seg];
      // ---- which output row does this thread own? ----
      long long row;
      bool valid = true;
      int ycls = 1, xcls = 1;
      if (p.a_mode == APTP_A_LINEAR) {
        row = (long long)tile.m_base + r;
      } else {
        const int hw = p.Ho * p.Wo;
        const int img0 = tile.m_base / hw;
        const int rem = tile.m_base - img0 * hw;
        const int oy0 = rem / p.Wo, ox0 = rem - oy0 * p.Wo;
        const int ix = r % p.bw;
        const int iy = (r / p.bw) % p.bh;
        const int ib = r / (p.bw * p.bh);
        const int oy = oy0 + iy, ox = ox0 + ix, img = img0 + ib;
        valid = (img < p.batch) && (oy < p.Ho) && (ox < p.Wo);
        row = ((long long)img * p.Ho + oy) * p.Wo + ox;
        ycls = (oy == 0) ? 0 : ((oy == p.Ho - 1) ? 2 : 1);
        xcls = (ox == 0) ? 0 : ((ox == p.Wo - 1) ? 2 : 1);
      }
      valid = valid && (row < (long long)seg.row_end) && (row >= (long long)seg.row_begin);
      const int sample = valid ? (int)(row / p.rows_per_sample) : 0;
      const int ocol_base = geglu ? tile.n0 / 2 : tile.n0;  // first OUTPUT column of this tile
      const float* tabp =
          p.border_tab ? p.border_tab + seg.tab_off + (size_t)(ycls * 3 + xcls) * p.tab_ld : nullptr;

      const bool got = mbar_wait(&tfull_bar[acc], acc_phase, p.abort_flag);
      if (!got) break;
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.bn);

      for (int c = 0; c * 32 < out_cols_per_tile; ++c) {
        const int col0 = ocol_base + c * 32;
        if (col0 >= seg.n_store) break;  // warp-uniform
        uint32_t ra[32];
        float v[32];
        tmem_ld_32x32(t_addr + c * 32, ra);
        if (geglu) {
          uint32_t rb[32];
          tmem_ld_32x32(t_addr + p.bn / 2 + c * 32, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h = __uint_as_float(ra[j]);
            float g = __uint_as_float(rb[j]);
            const int col = col0 + j;
            if (p.bias && col < seg.n_valid) {
              // packed bias follows the packed (interleaved) weight rows
              h += __ldg(p.bias + seg.vec_off + tile.n0 + c * 32 + j);
              g += __ldg(p.bias + seg.vec_off + tile.n0 + p.bn / 2 + c * 32 + j);
            }
            if (p.gate && valid && col < seg.n_valid) {
              const float gs = __ldg(p.gate + (size_t)sample * p.gate_ld + col / p.gate_group);
              h *= gs;
              g *= gs;
            }
            v[j] = h * gelu_erf_f(g);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < seg.n_valid) v[j] += __ldg(p.bias + seg.vec_off + col0 + j);
          }
          if (p.rowvec && valid) {
            const float* rv = p.rowvec + (size_t)sample * p.rowvec_ld + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < seg.n_valid) v[j] += __ldg(rv + j);
          }
          if (tabp) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < seg.n_valid) v[j] += __ldg(tabp + col0 + j);
          }
          if (p.gate && valid) {
            const float* gp = p.gate + (size_t)sample * p.gate_ld;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < seg.n_valid) v[j] *= __ldg(gp + (col0 + j) / p.gate_group);
          }
          if (p.flags & APTP_EPI_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
          }
        }
        if (!valid) continue;
        if (p.residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.res_ld + seg.out_col_off + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col0 + q * 8 < seg.n_valid) {  // n_valid is a multiple of 8 whenever a residual is used
              const uint4 rr = __ldg(rp + q);
              const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[q * 8 + 2 * e] += bf16_lo(w4[e]);
                v[q * 8 + 2 * e + 1] += bf16_hi(w4[e]);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j >= seg.n_valid) v[j] = 0.f;

        if (p.out_mode == APTP_OUT_BF16) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.out_ld + seg.out_col_off + col0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col0 + q * 8 < seg.n_store) {  // n_store is a multiple of 8
              uint4 o;
              o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
              o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
              o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
              o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
              *reinterpret_cast<uint4*>(op + q * 8) = o;
            }
          }
        } else if (p.out_mode == APTP_OUT_F32) {
          float* op = reinterpret_cast<float*>(p.out) + (size_t)row * p.out_ld + seg.out_col_off + col0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col0 + q * 4 < seg.n_store) {  // n_store multiple of 4
              *reinterpret_cast<float4*>(op + q * 4) =
                  make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
          }
        } else {  // fp32 NCHW: out[(sample*out_ld + col) * rows_per_sample + pixel]
          float* op = reinterpret_cast<float*>(p.out);
          const long long pix = row - (long long)sample * p.rows_per_sample;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < seg.n_valid)
              op[((size_t)sample * p.out_ld + col0 + j) * p.rows_per_sample + pix] = v[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int g_gemm_smem_set = 0;

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_grouped_gemm_fwd(const aptp_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(a != nullptr, "aptp_grouped_gemm_fwd: null args");
  APTP_REQUIRE(a->a && a->w && a->out && a->segs && a->tiles, "aptp_grouped_gemm_fwd: null pointer");
  if (a->n_tiles == 0) return APTP_OK;
  APTP_REQUIRE(a->bn >= 32 && a->bn <= 256 && a->bn % 32 == 0, "aptp_grouped_gemm_fwd: bn=%d must be a multiple of 32 in [32,256]", a->bn);
  APTP_REQUIRE(a->a_ld % 8 == 0 && a->w_ld % 8 == 0, "aptp_grouped_gemm_fwd: a_ld/w_ld must be multiples of 8 (16-byte rows)");
  APTP_REQUIRE(a->out_mode == APTP_OUT_F32_NCHW || a->out_ld % 8 == 0, "aptp_grouped_gemm_fwd: out_ld must be a multiple of 8");
  APTP_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
               "aptp_grouped_gemm_fwd: operands must be 16-byte aligned");
  APTP_REQUIRE(a->rows_per_sample > 0, "aptp_grouped_gemm_fwd: rows_per_sample must be > 0");
  APTP_REQUIRE(!(a->flags & APTP_EPI_GN_STATS), "aptp_grouped_gemm_fwd: APTP_EPI_GN_STATS not implemented yet");
  if (a->flags & APTP_EPI_GEGLU) APTP_REQUIRE(a->bn % 64 == 0, "aptp_grouped_gemm_fwd: GEGLU needs bn %% 64 == 0");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int Ho = 1, Wo = 1;
  if (a->a_mode == APTP_A_LINEAR) {
    uint64_t dims[2] = {(uint64_t)a->a_k, (uint64_t)a->a_rows};
    uint64_t strides[1] = {(uint64_t)a->a_ld * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_bf16(&p.tmap_a, a->a, 2, dims, strides, box);
    if (rc) return rc;
  } else if (a->a_mode == APTP_A_CONV3X3) {
    APTP_REQUIRE(a->bw * a->bh * a->bb == BM, "aptp_grouped_gemm_fwd: conv box %dx%dx%d != 128 pixels", a->bw, a->bh, a->bb);
    Ho = a->H;
    Wo = a->W;
    APTP_REQUIRE(Wo % a->bw == 0 && Ho % a->bh == 0, "aptp_grouped_gemm_fwd: box does not tile the image");
    uint64_t dims[4] = {(uint64_t)a->a_k, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)a->a_ld * 2, (uint64_t)a->W * a->a_ld * 2, (uint64_t)a->H * a->W * a->a_ld * 2};
    uint32_t box[4] = {BK, (uint32_t)a->bw, (uint32_t)a->bh, (uint32_t)a->bb};
    int rc = make_tmap_bf16(&p.tmap_a, a->a, 4, dims, strides, box);
    if (rc) return rc;
  } else if (a->a_mode == APTP_A_CONV3X3_S2) {
    APTP_REQUIRE(a->bw * a->bh * a->bb == BM, "aptp_grouped_gemm_fwd: conv box != 128 pixels");
    APTP_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "aptp_grouped_gemm_fwd: stride-2 conv needs even H, W");
    APTP_REQUIRE(a->a_k == a->a_ld && a->k_tap_pitch == a->a_ld, "aptp_grouped_gemm_fwd: stride-2 conv needs a_k == a_ld == k_tap_pitch");
    Ho = a->H / 2;
    Wo = a->W / 2;
    APTP_REQUIRE(Wo % a->bw == 0 && Ho % a->bh == 0, "aptp_grouped_gemm_fwd: box does not tile the image");
    // view [b][H/2][2][W/2][2][C] as dims (fastest first): (px*C + c), W/2, py, H/2, b
    uint64_t dims[5] = {(uint64_t)2 * a->a_ld, (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)a->batch};
    uint64_t strides[4] = {(uint64_t)2 * a->a_ld * 2, (uint64_t)a->W * a->a_ld * 2,
                           (uint64_t)2 * a->W * a->a_ld * 2, (uint64_t)a->H * a->W * a->a_ld * 2};
    uint32_t box[5] = {BK, (uint32_t)a->bw, 1, (uint32_t)a->bh, (uint32_t)a->bb};
    int rc = make_tmap_bf16(&p.tmap_a, a->a, 5, dims, strides, box);
    if (rc) return rc;
  } else {
    APTP_REQUIRE(false, "aptp_grouped_gemm_fwd: bad a_mode %d", a->a_mode);
  }
  {
    uint64_t dims[2] = {(uint64_t)a->w_ld, (uint64_t)a->w_rows};
    uint64_t strides[1] = {(uint64_t)a->w_ld * 2};
    uint32_t box[2] = {BK, (uint32_t)a->bn};
    int rc = make_tmap_bf16(&p.tmap_b, a->w, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.segs = a->segs;
  p.tiles = a->tiles;
  p.n_tiles = a->n_tiles;
  p.a_mode = a->a_mode;
  p.batch = a->batch;
  p.Ho = Ho;
  p.Wo = Wo;
  p.bn = a->bn;
  p.bw = a->a_mode == APTP_A_LINEAR ? BM : a->bw;
  p.bh = a->a_mode == APTP_A_LINEAR ? 1 : a->bh;
  p.bb = a->a_mode == APTP_A_LINEAR ? 1 : a->bb;
  p.k_tap_pitch = a->k_tap_pitch;
  p.out = a->out;
  p.out_ld = a->out_ld;
  p.out_mode = a->out_mode;
  p.bias = a->bias;
  p.rowvec = a->rowvec;
  p.rowvec_ld = a->rowvec_ld;
  p.rows_per_sample = a->rows_per_sample;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.res_ld = a->res_ld;
  p.gate = a->gate;
  p.gate_ld = a->gate_ld;
  p.gate_group = a->gate_group > 0 ? a->gate_group : 1;
  p.border_tab = a->border_tab;
  p.tab_ld = a->tab_ld;
  p.flags = a->flags;
  p.abort_flag = device_abort_flag();
  APTP_REQUIRE(p.abort_flag != nullptr, "aptp_grouped_gemm_fwd: could not allocate abort flag");

  const int stage_bytes = A_STAGE_BYTES + a->bn * 128;
  const int budget = 225 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  APTP_REQUIRE(stages >= 2, "aptp_grouped_gemm_fwd: tile too large for shared memory");
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256;
  if (!g_gemm_smem_set) {
    APTP_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_gemm_smem_set = 1;
  }
  int grid = a->n_tiles < sm_count() ? a->n_tiles : sm_count();
  grouped_gemm_kernel<<<grid, GEMM_THREADS, smem_bytes, stream>>>(p);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
