// K5: flash-style attention backward on tcgen05 (head_dim 64, no mask), per-sample kept-head lists.
// Two kernels, both built like the forward kernel (TMA producer warp, one MMA-issuing thread, two
// softmax-style warpgroups owning one 128-row TMEM tile each, setmaxnreg register split):
//
//   attn_bwd_dq_kernel   Q-stationary.  Per key tile j (64 keys):
//        S  = Q K_j^T,  dP = dO V_j^T            (M128 x N64 x K64, TMEM)
//        dS = exp2(S c - L) * (dP - delta)        (registers; L = log2-domain LSE of the forward)
//        dQ += dS K_j                             (dS bf16 is the A operand FROM TMEM, K_j MN-major)
//   attn_bwd_dkv_kernel  KV-stationary. Per query tile j (64 queries), transposed problem (lane = key):
//        S^T = K Q_j^T, dP^T = V dO_j^T           (M128 x N64 x K64)
//        P^T = exp2(S^T c - L[q]),  dS^T = P^T * (dP^T - delta[q])
//        dV += P^T dO_j,  dK += dS^T Q_j          (A from TMEM, dO_j / Q_j MN-major)
// Gradients reach the head gates through dq', dk', dv' (blocks.py:250-255).
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int AB_THREADS = 384;
constexpr int AB_D = 64;
constexpr int AB_BIG = 128 * 64 * 2;    // 16 KB: [128 x 64] bf16 tile
constexpr int AB_SMALL = 64 * 64 * 2;   //  8 KB: [ 64 x 64] bf16 tile
constexpr int AB_STAGES = 4;
// smem: 4 big tiles (stationary operands) | (2 small tiles) x stages | barriers
constexpr int AB_SMEM_ST = 0;
constexpr int AB_SMEM_RING = 4 * AB_BIG;
constexpr int AB_SMEM_BAR = AB_SMEM_RING + AB_STAGES * 2 * AB_SMALL;
constexpr int AB_SMEM_BYTES = AB_SMEM_BAR + 256 + 1024;
constexpr int AB_TMEM_COLS = 512;

struct AttnBwdParams {
  CUtensorMap tmap_q128, tmap_do128, tmap_k128, tmap_v128;  // [128 x 64] boxes
  CUtensorMap tmap_q64, tmap_do64, tmap_k64, tmap_v64;      // [ 64 x 64] boxes
  const float* lse2;   // [batch, max_heads, n_q] log2-domain log-sum-exp written by the forward
  const float* delta;  // [batch, max_heads, n_q] rowsum(dO * O)
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int lddq, lddk, lddv;
  int n_q, n_kv, max_heads;
  const int* sample_heads;
  float scale, scale_log2;
  int* abort_flag;
};

struct AbBars {
  uint64_t* st_full;   // stationary tiles landed
  uint64_t* ring_full;
  uint64_t* ring_empty;
  uint64_t* s_full;    // [2]  S / dP (or transposed) accumulators ready
  uint64_t* s_free;    // [2]  softmax holds them in registers
  uint64_t* p_full;    // [2]  bf16 operand(s) written to TMEM
  uint64_t* acc_done;  // [2]  accumulate MMA(s) of the step complete
  uint32_t* tmem_slot;
};

__device__ __forceinline__ AbBars ab_setup(uint8_t* smem, int warp, int lane, int bar_offset = AB_SMEM_BAR) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bar_offset);
  AbBars b;
  b.st_full = bars;
  b.ring_full = bars + 1;
  b.ring_empty = b.ring_full + AB_STAGES;
  b.s_full = b.ring_empty + AB_STAGES;
  b.s_free = b.s_full + 2;
  b.p_full = b.s_free + 2;
  b.acc_done = b.p_full + 2;
  b.tmem_slot = reinterpret_cast<uint32_t*>(b.acc_done + 2);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(b.st_full, 1);
      for (int s = 0; s < AB_STAGES; ++s) {
        mbar_init(&b.ring_full[s], 1);
        mbar_init(&b.ring_empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&b.s_full[i], 1);
        mbar_init(&b.s_free[i], 4);
        mbar_init(&b.p_full[i], 4);
        mbar_init(&b.acc_done[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(b.tmem_slot, AB_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return b;
}

__device__ __forceinline__ void ab_store_rows(__nv_bfloat16* op, const uint32_t* o, float mul) {
#pragma unroll
  for (int q = 0; q < AB_D / 8; ++q) {
    const float* f = reinterpret_cast<const float*>(o) + q * 8;
    *reinterpret_cast<uint4*>(op + q * 8) =
        make_uint4(pack_bf16(f[0] * mul, f[1] * mul), pack_bf16(f[2] * mul, f[3] * mul),
                   pack_bf16(f[4] * mul, f[5] * mul), pack_bf16(f[6] * mul, f[7] * mul));
  }
}

// =====================================================================================================
// dQ: one CTA = 256 queries (two 128-row tiles) of one (sample, kept head); loop over 64-key tiles.
// TMEM per tile w (224 columns): S [0,64) | dP [64,128) | dS bf16 [128,160) | dQ [160,224)
// =====================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
  const int head = blockIdx.y, b = blockIdx.z;
  if (head >= p.sample_heads[b]) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const AbBars B = ab_setup(smem, warp, lane);
  const uint32_t tmem_base = *B.tmem_slot;
  const int n_tiles = (p.n_kv + 63) / 64;
  const int q0 = blockIdx.x * 256;
  const int n_w = (q0 + 128 < p.n_q) ? 2 : 1;
  constexpr int TW = 224;  // TMEM columns per query tile

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // stationary: Q_w at tiles 0,1 ; dO_w at tiles 2,3
      mbar_expect_tx_e(B.st_full, n_w * 2 * AB_BIG);
      for (int w = 0; w < n_w; ++w) {
        tma_load_2d_e(smem + AB_SMEM_ST + w * AB_BIG, &p.tmap_q128, B.st_full, head * AB_D, b * p.n_q + q0 + w * 128);
        tma_load_2d_e(smem + AB_SMEM_ST + (2 + w) * AB_BIG, &p.tmap_do128, B.st_full, head * AB_D,
                    b * p.n_q + q0 + w * 128);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        if (!mbar_wait(&B.ring_empty[stage], phase ^ 1, p.abort_flag)) break;
        uint8_t* sk = smem + AB_SMEM_RING + stage * 2 * AB_SMALL;
        mbar_expect_tx_e(&B.ring_full[stage], 2 * AB_SMALL);
        tma_load_2d_e(sk, &p.tmap_k64, &B.ring_full[stage], head * AB_D, b * p.n_kv + j * 64);
        tma_load_2d_e(sk + AB_SMALL, &p.tmap_v64, &B.ring_full[stage], head * AB_D, b * p.n_kv + j * 64);
        if (++stage == AB_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);  // A from TMEM (K-major), B MN-major
      auto issue_sp = [&](int w, int stg) {
        const uint32_t q_addr = smem_u32(smem + AB_SMEM_ST + w * AB_BIG);
        const uint32_t do_addr = smem_u32(smem + AB_SMEM_ST + (2 + w) * AB_BIG);
        const uint32_t k_addr = smem_u32(smem + AB_SMEM_RING + stg * 2 * AB_SMALL);
        const uint32_t v_addr = k_addr + AB_SMALL;
        const uint32_t t = tmem_base + w * TW;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_e(t, make_desc_kmajor_sw128(q_addr + k * 32), make_desc_kmajor_sw128(k_addr + k * 32), idesc_s,
                       k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_e(t + 64, make_desc_kmajor_sw128(do_addr + k * 32), make_desc_kmajor_sw128(v_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit_e(&B.s_full[w]);
      };
      int stage = 0;
      uint32_t phase = 0;
      bool ok = mbar_wait(B.st_full, 0, p.abort_flag) && mbar_wait(&B.ring_full[0], 0, p.abort_flag);
      if (ok) {
        tc_fence_after();
        for (int w = 0; w < n_w; ++w) issue_sp(w, 0);
      }
      for (int j = 0; j < n_tiles && ok; ++j) {
        int ns = stage + 1;
        uint32_t nphase = phase;
        if (ns == AB_STAGES) {
          ns = 0;
          nphase ^= 1;
        }
        const bool more = j + 1 < n_tiles;
        if (more && !mbar_wait(&B.ring_full[ns], nphase, p.abort_flag)) break;
        for (int w = 0; w < n_w && ok; ++w) {
          ok = mbar_wait(&B.s_free[w], j & 1, p.abort_flag);
          if (ok && more) {
            tc_fence_after();
            issue_sp(w, ns);
          }
        }
        if (!ok) break;
        const uint32_t k_addr = smem_u32(smem + AB_SMEM_RING + stage * 2 * AB_SMALL);
        for (int w = 0; w < n_w; ++w) {
          if (!mbar_wait(&B.p_full[w], j & 1, p.abort_flag)) {
            ok = false;
            break;
          }
          tc_fence_after();
          const uint32_t t = tmem_base + w * TW;
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 keys per step: 8 packed TMEM columns of dS, 2 swizzle atoms of K
            umma_bf16_ts_e(t + 160, t + 128 + k * 8, make_desc_mnmajor_sw128(k_addr + k * 2048, AB_SMALL), idesc_acc,
                         (j | k) != 0);
          umma_commit_e(&B.acc_done[w]);
        }
        if (!ok) break;
        umma_commit_e(&B.ring_empty[stage]);
        stage = ns;
        phase = nphase;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qrow = q0 + w * 128 + r;
    if (w < n_w) {
      const uint32_t t = tmem_base + ((uint32_t)(quad * 32) << 16) + w * TW;
      const size_t vidx = ((size_t)b * p.max_heads + head) * p.n_q + qrow;
      const float L = (qrow < p.n_q) ? p.lse2[vidx] : 0.f;
      const float dl = (qrow < p.n_q) ? p.delta[vidx] : 0.f;
      bool ok = true;
      for (int j = 0; j < n_tiles; ++j) {
        ok = mbar_wait(&B.s_full[w], j & 1, p.abort_flag);
        if (!ok) break;
        tc_fence_after();
        uint32_t s[64], dp[64];
        tmem_ld_32x32(t, s);
        tmem_ld_32x32(t + 64, dp);
        tmem_ld_wait();
        tmem_ld_32x32(t + 32, s + 32);   // keys 32..63 of the tile arrive under the first half's math
        tmem_ld_32x32(t + 96, dp + 32);
        const int kv_left = p.n_kv - j * 64;
        const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const uint64_t nL2 = pack_f32x2(-L, -L), ndl2 = pack_f32x2(-dl, -dl);
        uint32_t pk[32];
        // dS = 2^(S c - L) (dP - delta) for keys [32 h, 32 h + 32), packed fp32x2 math; masked past the last key
        auto half = [&](int h) {
#pragma unroll
          for (int i = 32 * h; i < 32 * h + 32; i += 2) {
            float x0, x1, d0, d1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), scale2, nL2), x0, x1);
            float p0 = ex2_approx_ordered(x0), p1 = ex2_approx_ordered(x1);
            if (i >= kv_left) p0 = 0.f;
            if (i + 1 >= kv_left) p1 = 0.f;
            const uint64_t t2 = add_f32x2(pack_f32x2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), ndl2);
            unpack_f32x2(fma_f32x2(pack_f32x2(p0, p1), t2, 0ull), d0, d1);
            pk[i >> 1] = pack_bf16(d0, d1);
          }
        };
        if (p.n_q > 0) half(0);  // always true; the branch keeps ptxas from sinking this half below the next wait
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.s_free[w]);
        if (p.n_kv > 0) half(1);  // (opaque as above: keeps the early s_free ahead of this half's math)
        if (j > 0) {  // the dQ MMA of the previous step must have consumed dS
          ok = mbar_wait(&B.acc_done[w], (j - 1) & 1, p.abort_flag);
          if (!ok) break;
          tc_fence_after();
        }
        tmem_st_32x32(t + 128, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.p_full[w]);
      }
      if (ok) ok = mbar_wait(&B.acc_done[w], (n_tiles - 1) & 1, p.abort_flag);
      if (ok) {
        tc_fence_after();
        uint32_t o[AB_D];
        tmem_ld_32x32(t + 160, o);
        tmem_ld_32x32(t + 192, o + 32);
        tmem_ld_wait();
        if (qrow < p.n_q) ab_store_rows(p.dq + ((size_t)b * p.n_q + qrow) * p.lddq + head * AB_D, o, p.scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

// =====================================================================================================
// dK, dV: one CTA = 256 keys (two 128-row tiles) of one (sample, kept head); loop over 64-query tiles.
// TMEM per tile w (256 columns): S^T [0,64) | dP^T [64,128) | dV [128,192) | dK [192,256).
// TMEM reads pace this kernel (64 B/clk/SM: S^T + dP^T of both tiles = 2048 clk per query tile), so the
// accumulators are handed back to the tensor pipe as soon as they sit in registers (s_free) and S^T/dP^T of the
// next query tile are computed under this tile's exponentials. P^T and dS^T therefore cannot alias the
// accumulator columns: they go to shared memory as K-major 128B-swizzled A operands (the layout TMA would
// have produced for a [128 x 64] bf16 tile) and the accumulate MMAs are SS. The per-query LSE / delta of the
// NEXT tile are fetched one tile ahead into a per-warp shared-memory slot (coalesced, bounds folded into
// +inf / 0), so the hot loop has no global loads.
// =====================================================================================================
constexpr int AK_SMEM_OP = AB_SMEM_RING + AB_STAGES * 2 * AB_SMALL;  // P^T, dS^T per tile w: 4 x 16 KB
constexpr int AK_SMEM_LD = AK_SMEM_OP + 4 * AB_BIG;                  // 8 warps x 2 slots x (L[64] | D[64]) fp32
constexpr int AK_SMEM_BAR = AK_SMEM_LD + 8 * 2 * 128 * 4;
constexpr int AK_SMEM_BYTES = AK_SMEM_BAR + 256 + 1024;
static_assert(AK_SMEM_BYTES <= 227 * 1024, "dkv kernel shared memory");

__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwdParams p) {
  const int head = blockIdx.y, b = blockIdx.z;
  if (head >= p.sample_heads[b]) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const AbBars B = ab_setup(smem, warp, lane, AK_SMEM_BAR);
  const uint32_t tmem_base = *B.tmem_slot;
  const int n_tiles = (p.n_q + 63) / 64;
  const int k0 = blockIdx.x * 256;
  const int n_w = (k0 + 128 < p.n_kv) ? 2 : 1;
  constexpr int TW = 256;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // stationary: K_w at tiles 0,1 ; V_w at tiles 2,3
      mbar_expect_tx_e(B.st_full, n_w * 2 * AB_BIG);
      for (int w = 0; w < n_w; ++w) {
        tma_load_2d_e(smem + AB_SMEM_ST + w * AB_BIG, &p.tmap_k128, B.st_full, head * AB_D, b * p.n_kv + k0 + w * 128);
        tma_load_2d_e(smem + AB_SMEM_ST + (2 + w) * AB_BIG, &p.tmap_v128, B.st_full, head * AB_D,
                    b * p.n_kv + k0 + w * 128);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        if (!mbar_wait(&B.ring_empty[stage], phase ^ 1, p.abort_flag)) break;
        uint8_t* sq = smem + AB_SMEM_RING + stage * 2 * AB_SMALL;
        mbar_expect_tx_e(&B.ring_full[stage], 2 * AB_SMALL);
        tma_load_2d_e(sq, &p.tmap_q64, &B.ring_full[stage], head * AB_D, b * p.n_q + j * 64);
        tma_load_2d_e(sq + AB_SMALL, &p.tmap_do64, &B.ring_full[stage], head * AB_D, b * p.n_q + j * 64);
        if (++stage == AB_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);  // A K-major (P^T / dS^T in smem), B MN-major
      auto issue_sp = [&](int w, int stg) {
        const uint32_t k_addr = smem_u32(smem + AB_SMEM_ST + w * AB_BIG);
        const uint32_t v_addr = smem_u32(smem + AB_SMEM_ST + (2 + w) * AB_BIG);
        const uint32_t q_addr = smem_u32(smem + AB_SMEM_RING + stg * 2 * AB_SMALL);
        const uint32_t do_addr = q_addr + AB_SMALL;
        const uint32_t t = tmem_base + w * TW;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_e(t, make_desc_kmajor_sw128(k_addr + k * 32), make_desc_kmajor_sw128(q_addr + k * 32), idesc_s,
                       k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_e(t + 64, make_desc_kmajor_sw128(v_addr + k * 32), make_desc_kmajor_sw128(do_addr + k * 32), idesc_s,
                       k != 0);
        umma_commit_e(&B.s_full[w]);
      };
      int stage = 0;
      uint32_t phase = 0;
      bool ok = mbar_wait(B.st_full, 0, p.abort_flag) && mbar_wait(&B.ring_full[0], 0, p.abort_flag);
      if (ok) {
        tc_fence_after();
        for (int w = 0; w < n_w; ++w) issue_sp(w, 0);
      }
      for (int j = 0; j < n_tiles && ok; ++j) {
        int ns = stage + 1;
        uint32_t nphase = phase;
        if (ns == AB_STAGES) {
          ns = 0;
          nphase ^= 1;
        }
        const bool more = j + 1 < n_tiles;
        if (more && !mbar_wait(&B.ring_full[ns], nphase, p.abort_flag)) break;
        for (int w = 0; w < n_w && ok; ++w) {  // S^T / dP^T of tile j are in registers: compute tile j+1 over them
          ok = mbar_wait(&B.s_free[w], j & 1, p.abort_flag);
          if (ok && more) {
            tc_fence_after();
            issue_sp(w, ns);
          }
        }
        if (!ok) break;
        const uint32_t q_addr = smem_u32(smem + AB_SMEM_RING + stage * 2 * AB_SMALL);
        const uint32_t do_addr = q_addr + AB_SMALL;
        for (int w = 0; w < n_w; ++w) {
          if (!mbar_wait(&B.p_full[w], j & 1, p.abort_flag)) {
            ok = false;
            break;
          }
          tc_fence_after();
          const uint32_t t = tmem_base + w * TW;
          const uint32_t p_addr = smem_u32(smem + AK_SMEM_OP + w * 2 * AB_BIG);
          const uint32_t ds_addr = p_addr + AB_BIG;
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dV += P^T dO_j   (16 queries per step)
            umma_bf16_ss_e(t + 128, make_desc_kmajor_sw128(p_addr + k * 32),
                           make_desc_mnmajor_sw128(do_addr + k * 2048, AB_SMALL), idesc_acc, (j | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dK += dS^T Q_j
            umma_bf16_ss_e(t + 192, make_desc_kmajor_sw128(ds_addr + k * 32),
                           make_desc_mnmajor_sw128(q_addr + k * 2048, AB_SMALL), idesc_acc, (j | k) != 0);
          umma_commit_e(&B.acc_done[w]);
        }
        if (!ok) break;
        umma_commit_e(&B.ring_empty[stage]);
        stage = ns;
        phase = nphase;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int krow = k0 + w * 128 + r;
    if (w < n_w) {
      const uint32_t t = tmem_base + ((uint32_t)(quad * 32) << 16) + w * TW;
      const float* Lp = p.lse2 + ((size_t)b * p.max_heads + head) * p.n_q;
      const float* Dp = p.delta + ((size_t)b * p.max_heads + head) * p.n_q;
      const uint32_t ld_slot = smem_u32(smem + AK_SMEM_LD) + (warp - 4) * 1024;  // [2][-L 64 | -D 64] fp32
      const uint32_t op_p = smem_u32(smem + AK_SMEM_OP + w * 2 * AB_BIG) + r * 128;  // this key's 128-byte row of P^T
      const uint32_t op_ds = op_p + AB_BIG;
      const int swz = r & 7;
      // the slots hold -L and -D (operands of the packed fma / add); out-of-range queries: -L = -inf makes
      // P^T = 0 and dS^T = 0 * finite = 0
      auto fetch = [&](int j, float& l0, float& l1, float& d0, float& d1) {
        const int qa = j * 64 + lane, qc = qa + 32;
        l0 = (qa < p.n_q) ? -__ldg(Lp + qa) : -INFINITY;
        l1 = (qc < p.n_q) ? -__ldg(Lp + qc) : -INFINITY;
        d0 = (qa < p.n_q) ? -__ldg(Dp + qa) : 0.f;
        d1 = (qc < p.n_q) ? -__ldg(Dp + qc) : 0.f;
      };
      {
        float l0, l1, d0, d1;
        fetch(0, l0, l1, d0, d1);
        sts_f32(ld_slot + lane * 4, l0);
        sts_f32(ld_slot + (32 + lane) * 4, l1);
        sts_f32(ld_slot + (64 + lane) * 4, d0);
        sts_f32(ld_slot + (96 + lane) * 4, d1);
        __syncwarp();
      }
      bool ok = true;
      for (int j = 0; j < n_tiles; ++j) {
        float nl0 = -INFINITY, nl1 = -INFINITY, nd0 = 0.f, nd1 = 0.f;
        const bool more = j + 1 < n_tiles;
        if (more) fetch(j + 1, nl0, nl1, nd0, nd1);
        ok = mbar_wait(&B.s_full[w], j & 1, p.abort_flag);
        if (!ok) break;
        tc_fence_after();
        uint32_t s[64], dp[64];
        tmem_ld_32x32(t, s);
        tmem_ld_32x32(t + 64, dp);
        tmem_ld_wait();
        tmem_ld_32x32(t + 32, s + 32);   // second half of the tile arrives under the first half's math
        tmem_ld_32x32(t + 96, dp + 32);
        const uint32_t Ls = ld_slot + (j & 1) * 512;  // -L[64] | -D[64] of this query tile
        const uint32_t Ds = Ls + 256;
        const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
        uint32_t pp[32], ds[32];
        // P^T = 2^(S^T c - L), dS^T = P^T (dP^T - D) for queries [32 h, 32 h + 32): packed fp32x2 math
        auto half = [&](int h) {
#pragma unroll
          for (int i = 32 * h; i < 32 * h + 32; i += 4) {
            uint64_t nL0, nL1, nD0, nD1;
            lds_f32x2x2(Ls + i * 4, nL0, nL1);
            lds_f32x2x2(Ds + i * 4, nD0, nD1);
            float x0, x1, x2, x3;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), scale2, nL0), x0, x1);
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), scale2, nL1), x2, x3);
            // ordered: the exponentials of half 0 stay ahead of the wait for half 1's TMEM loads
            const float p0 = ex2_approx_ordered(x0), p1 = ex2_approx_ordered(x1), p2 = ex2_approx_ordered(x2),
                        p3 = ex2_approx_ordered(x3);
            pp[(i >> 1)] = pack_bf16(p0, p1);
            pp[(i >> 1) + 1] = pack_bf16(p2, p3);
            const uint64_t t01 = add_f32x2(pack_f32x2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), nD0);
            const uint64_t t23 = add_f32x2(pack_f32x2(__uint_as_float(dp[i + 2]), __uint_as_float(dp[i + 3])), nD1);
            float d0, d1, d2, d3;
            unpack_f32x2(fma_f32x2(pack_f32x2(p0, p1), t01, 0ull), d0, d1);
            unpack_f32x2(fma_f32x2(pack_f32x2(p2, p3), t23, 0ull), d2, d3);
            ds[(i >> 1)] = pack_bf16(d0, d1);
            ds[(i >> 1) + 1] = pack_bf16(d2, d3);
          }
        };
        // always true, but opaque to the compiler: the branch makes half 0 its own basic block, which keeps ptxas from
        // sinking its math below the wait for half 1's loads (register-only instructions may otherwise move freely
        // across tcgen05.wait / mbarrier.arrive)
        if (p.n_q > 0) half(0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.s_free[w]);  // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
        half(1);
        if (j > 0) {  // the accumulate MMAs of the previous tile must have finished reading P^T / dS^T
          ok = mbar_wait(&B.acc_done[w], (j - 1) & 1, p.abort_flag);
          if (!ok) break;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 16-byte chunk c of row r sits at chunk (c ^ (r & 7)): the TMA 128B swizzle
          const uint32_t off = (uint32_t)((c ^ swz) << 4);
          sts_u32x4(op_p + off, pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
          sts_u32x4(op_ds + off, ds[4 * c], ds[4 * c + 1], ds[4 * c + 2], ds[4 * c + 3]);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor pipe's async-proxy reads
        if (more) {
          const uint32_t nx = ld_slot + ((j + 1) & 1) * 512;
          sts_f32(nx + lane * 4, nl0);
          sts_f32(nx + (32 + lane) * 4, nl1);
          sts_f32(nx + (64 + lane) * 4, nd0);
          sts_f32(nx + (96 + lane) * 4, nd1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.p_full[w]);
      }
      if (ok) ok = mbar_wait(&B.acc_done[w], (n_tiles - 1) & 1, p.abort_flag);
      if (ok) {
        tc_fence_after();
        uint32_t o[AB_D];
        tmem_ld_32x32(t + 128, o);
        tmem_ld_32x32(t + 160, o + 32);
        tmem_ld_wait();
        if (krow < p.n_kv) ab_store_rows(p.dv + ((size_t)b * p.n_kv + krow) * p.lddv + head * AB_D, o, 1.f);
        tmem_ld_32x32(t + 192, o);
        tmem_ld_32x32(t + 224, o + 32);
        tmem_ld_wait();
        if (krow < p.n_kv) ab_store_rows(p.dk + ((size_t)b * p.n_kv + krow) * p.lddk + head * AB_D, o, p.scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

// delta[b, h, q] = sum_d dO[q, h*64 + d] * O[q, h*64 + d]   (one warp per (row, head): 2 elements per lane)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int ldo,
                                                         const __nv_bfloat16* __restrict__ dout, int lddo,
                                                         float* __restrict__ delta, int batch, int n_q, int max_heads,
                                                         const int* __restrict__ sample_heads) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long total = (long long)batch * n_q * max_heads;
  if (wid >= total) return;
  const int h = (int)(wid % max_heads);
  const long long row = wid / max_heads;
  const int b = (int)(row / n_q);
  if (h >= sample_heads[b]) return;
  const uint32_t a = *reinterpret_cast<const uint32_t*>(o + row * ldo + h * 64 + lane * 2);
  const uint32_t d = *reinterpret_cast<const uint32_t*>(dout + row * lddo + h * 64 + lane * 2);
  float v = bf16_lo(a) * bf16_lo(d) + bf16_hi(a) * bf16_hi(d);
  v = warp_sum(v);
  if (lane == 0) delta[((size_t)b * max_heads + h) * n_q + (row - (long long)b * n_q)] = v;
}

static int g_ab_smem_set = 0;

}  // namespace aptp

using namespace aptp;

static int ab_tmap(CUtensorMap* m, const void* base, int ld, int cols, long long rows, int box_rows) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t strides[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  return make_tmap_bf16(m, base, 2, dims, strides, box);
}

extern "C" int aptp_attention_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                  const void* o, int32_t ldo, const void* dout, int32_t lddo, const float* lse2,
                                  float* delta, void* dq, int32_t lddq, void* dk, int32_t lddk, void* dv, int32_t lddv,
                                  int32_t batch, int32_t n_q, int32_t n_kv, const int32_t* sample_heads,
                                  int32_t max_heads, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(q && k && v && o && dout && lse2 && delta && dq && dk && dv && sample_heads,
               "aptp_attention_bwd: null pointer");
  APTP_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 &&
                   lddk % 8 == 0 && lddv % 8 == 0,
               "aptp_attention_bwd: pitches must be multiples of 8");
  APTP_REQUIRE(n_q > 0 && n_kv > 0 && scale > 0.f, "aptp_attention_bwd: bad sizes");
  if (batch == 0 || max_heads == 0) return APTP_OK;
  APTP_REQUIRE(max_heads <= 65535 && batch <= 65535, "aptp_attention_bwd: grid too large");
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  const int cols = max_heads * 64;
  int rc = 0;
  rc = rc ? rc : ab_tmap(&p.tmap_q128, q, ldq, cols, (long long)batch * n_q, 128);
  rc = rc ? rc : ab_tmap(&p.tmap_do128, dout, lddo, cols, (long long)batch * n_q, 128);
  rc = rc ? rc : ab_tmap(&p.tmap_k128, k, ldk, cols, (long long)batch * n_kv, 128);
  rc = rc ? rc : ab_tmap(&p.tmap_v128, v, ldv, cols, (long long)batch * n_kv, 128);
  rc = rc ? rc : ab_tmap(&p.tmap_q64, q, ldq, cols, (long long)batch * n_q, 64);
  rc = rc ? rc : ab_tmap(&p.tmap_do64, dout, lddo, cols, (long long)batch * n_q, 64);
  rc = rc ? rc : ab_tmap(&p.tmap_k64, k, ldk, cols, (long long)batch * n_kv, 64);
  rc = rc ? rc : ab_tmap(&p.tmap_v64, v, ldv, cols, (long long)batch * n_kv, 64);
  if (rc) return rc;
  p.lse2 = lse2;
  p.delta = delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq;
  p.lddk = lddk;
  p.lddv = lddv;
  p.n_q = n_q;
  p.n_kv = n_kv;
  p.max_heads = max_heads;
  p.sample_heads = sample_heads;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.abort_flag = device_abort_flag();
  APTP_REQUIRE(p.abort_flag != nullptr, "aptp_attention_bwd: could not allocate abort flag");
  if (!g_ab_smem_set) {
    APTP_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    APTP_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AK_SMEM_BYTES));
    g_ab_smem_set = 1;
  }
  const long long nd = (long long)batch * n_q * max_heads;
  attn_delta_kernel<<<(unsigned)((nd + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(o), ldo,
                                                                 reinterpret_cast<const __nv_bfloat16*>(dout), lddo,
                                                                 delta, batch, n_q, max_heads, sample_heads);
  APTP_CUDA_CHECK(cudaGetLastError());
  dim3 gq((n_q + 255) / 256, max_heads, batch);
  attn_bwd_dq_kernel<<<gq, AB_THREADS, AB_SMEM_BYTES, stream>>>(p);
  APTP_CUDA_CHECK(cudaGetLastError());
  dim3 gk((n_kv + 255) / 256, max_heads, batch);
  attn_bwd_dkv_kernel<<<gk, AB_THREADS, AK_SMEM_BYTES, stream>>>(p);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
