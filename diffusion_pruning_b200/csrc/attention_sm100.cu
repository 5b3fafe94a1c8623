// K3: flash-style attention (head_dim 64, no mask, non-causal) on tcgen05 with per-sample kept-head
// lists. One CTA = 256 queries (two 128-row tiles) of one (sample, kept head); pruned heads and
// depth-dropped samples launch no work (the reference multiplies q, k, v of a pruned head by zero and
// still runs SDPA over it: pdm/models/unet/blocks.py:250-260).
//
//   warp 0      TMA producer: the two Q tiles, then a ring of (K, V) tiles of 128 keys
//   warp 1      MMA issuer:   S_w = Q_w K_j^T   (M128 x N128 x K64, both operands K-major, S in TMEM)
//                             O_w += P_w V_j    (M128 x N64 x K128; P is the A operand *from TMEM*, V is
//                                                consumed MN-major straight from its [keys, d] layout)
//   warps 4..7  softmax of Q tile 0, warps 8..11 softmax of Q tile 1 (setmaxnreg moves the registers of
//               the producer warpgroup to them): one thread per query row; the
//               whole S row is read once (tcgen05.ld) into registers and S is handed straight back to
//               the tensor pipe (s_free), so QK^T of the NEXT key tile runs underneath this tile's
//               exponentials; exp2 runs on ex2.approx with the softmax scale folded into one FFMA; P is
//               written bf16 to its own TMEM columns (tcgen05.st). O stays in TMEM for the whole KV
//               loop and is only rescaled when the running max grows by more than 2^8 (lazy
//               rescaling), so the steady state is: TMEM read S, 128 x (FFMA + EX2), TMEM write P.
// With short KV (cross-attention, 77 keys) a CTA loops over several query pairs to amortise TMEM
// allocation and barrier setup.
#include <type_traits>
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int ATT_THREADS = 384;  // warpgroup 0: TMA / MMA / 2 idle warps; warpgroups 1, 2: softmax of Q tile 0, 1
constexpr int ATT_BM = 128;   // queries per tile (two tiles per CTA)
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_KV_STAGES = 4;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
// smem map (from 1024-aligned base): 2 slots x (Q0 | Q1) | (K,V) x stages | barriers
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_Q_SLOT_BYTES = 2 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_KV = 2 * ATT_Q_SLOT_BYTES;
constexpr int ATT_SMEM_BAR = ATT_SMEM_KV + ATT_KV_STAGES * 2 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 256 + 1024;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_TMEM_S = 0;      // S0 at cols [0,128), S1 at [128,256)
constexpr int ATT_TMEM_P = 256;    // P0 at [256,320), P1 at [320,384): 128 keys x bf16 = 64 packed columns
constexpr int ATT_TMEM_O = 384;    // O0 at [384,448), O1 at [448,512)
constexpr float ATT_RESCALE_LOG2 = 8.f;
// share of the exponentials computed on the FMA/ALU pipes instead of the XU pipe: ATT_POLY_PAIRS of every
// ATT_POLY_PERIOD column pairs
#ifndef ATT_POLY_PAIRS
#define ATT_POLY_PAIRS 0
#endif
#ifndef ATT_POLY_PERIOD
#define ATT_POLY_PERIOD 4
#endif

struct AttnParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;
  __nv_bfloat16* out;
  int ldo;
  int n_q, n_kv;
  int n_qpairs;
  const int* sample_heads;
  float scale_log2;  // softmax scale * log2(e)
  float* lse2;       // optional [batch, max_heads, n_q]: log2-domain log-sum-exp (for the backward)
  int max_heads;
  int* abort_flag;
};

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  const int head = blockIdx.y, b = blockIdx.z;
  if (head >= p.sample_heads[b]) {  // pruned head / depth-dropped sample: no work at all
    pdl_wait();  // (a grid none of whose CTAs waited would "complete" before its predecessors and unchain its successors)
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars;                       // 2 (Q is double-buffered: the next query pair loads under this one)
  uint64_t* q_empty = bars + 2;                  // 2
  uint64_t* kv_full = bars + 4;                  // stages
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;  // stages
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;   // 2
  uint64_t* s_free = s_full + 2;                 // 2: softmax holds S in registers, S columns reusable
  uint64_t* p_full = s_free + 2;                 // 2
  uint64_t* pv_done = p_full + 2;                // 2: PV MMA of a tile complete (P reusable, O current)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.n_kv + ATT_BN - 1) / ATT_BN;
  // Short KV (cross-attention over 77 text tokens, or the 8x8 level): ONE key tile. K / V stay resident in stage 0 for
  // every query pair of the CTA, the MMA warp looks one query pair ahead (S of the next pair is issued as soon as the
  // softmax warps hold the current S in registers), the softmax skips fully masked 32-key chunks and PV runs over
  // ceil(n_kv / 16) key steps only.
  const bool single = (n_tiles == 1);
  const int pv_steps_last = (p.n_kv - (n_tiles - 1) * ATT_BN + 15) / 16;  // K steps of PV for the last key tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_k);
    tma_prefetch_desc(&p.tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 1);
      }
      for (int s = 0; s < ATT_KV_STAGES; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch();
  pdl_wait();  // q / k / v come from the projection GEMMs before us

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    // Whole warp in uniform control flow; one elected lane issues (a loop nested under `lane == 0` makes
    // ptxas wrap every UTMALDG / UTCHMMA in an R2UR waterfall loop, ~140 cycles per instruction).
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      bool ok = true;
      for (int qp = blockIdx.x; qp < p.n_qpairs && ok; qp += gridDim.x, ++it) {
        const int q0 = qp * 2 * ATT_BM;
        const bool act1 = q0 + ATT_BM < p.n_q;
        const int slot = it & 1;
        if (!mbar_wait(&q_empty[slot], ((it >> 1) & 1) ^ 1, p.abort_flag)) break;
        if (elect_one()) {
          uint8_t* sq = smem + ATT_SMEM_Q + slot * ATT_Q_SLOT_BYTES;
          mbar_expect_tx(&q_full[slot], act1 ? 2 * ATT_TILE_BYTES : ATT_TILE_BYTES);
          tma_load_2d(sq, &p.tmap_q, &q_full[slot], head * ATT_D, b * p.n_q + q0);
          if (act1) tma_load_2d(sq + ATT_TILE_BYTES, &p.tmap_q, &q_full[slot], head * ATT_D, b * p.n_q + q0 + ATT_BM);
        }
        __syncwarp();
        if (single && it > 0) continue;  // K / V of this (sample, head) are already resident in stage 0
        for (int j = 0; j < n_tiles; ++j) {
          if (!mbar_wait(&kv_empty[stage], phase ^ 1, p.abort_flag)) {
            ok = false;
            break;
          }
          if (elect_one()) {
            uint8_t* sk = smem + ATT_SMEM_KV + stage * 2 * ATT_TILE_BYTES;
            mbar_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
            tma_load_2d(sk, &p.tmap_k, &kv_full[stage], head * ATT_D, b * p.n_kv + j * ATT_BN);
            tma_load_2d(sk + ATT_TILE_BYTES, &p.tmap_v, &kv_full[stage], head * ATT_D, b * p.n_kv + j * ATT_BN);
          }
          __syncwarp();
          if (++stage == ATT_KV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    {
      const uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // A (=P) K-major from TMEM, B (=V) MN-major
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      uint32_t gp[2] = {0, 0};  // tiles consumed per Q tile (phase of s_free / p_full)
      bool ok = true;
      auto issue_s = [&](int w, int kv_stage, int q_slot) {
        const uint64_t dq =
            make_desc_kmajor_sw128(smem_base + ATT_SMEM_Q + q_slot * ATT_Q_SLOT_BYTES + w * ATT_TILE_BYTES);
        const uint64_t dk = make_desc_kmajor_sw128(smem_base + ATT_SMEM_KV + kv_stage * 2 * ATT_TILE_BYTES);
        const uint32_t s_tmem = tmem_base + ATT_TMEM_S + w * ATT_BN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)  // +32 B per K=16 step = +2 in the descriptor address field
            umma_bf16_ss(s_tmem, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
          umma_commit(&s_full[w]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, int kv_stage, int j, int ksteps) {
        // O_w (+)= P_w V_j. V: 16 keys = 2 swizzle atoms (2048 B) per step = +128 in the descriptor address field
        const uint64_t dv = make_desc_mnmajor_sw128(
            smem_base + ATT_SMEM_KV + kv_stage * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES, ATT_TILE_BYTES);
        const uint32_t p_tmem = tmem_base + ATT_TMEM_P + w * (ATT_BN / 2);
        const uint32_t o_tmem = tmem_base + ATT_TMEM_O + w * ATT_D;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)  // P: 16 keys = 8 packed 32-bit TMEM columns per step
            umma_bf16_ts(o_tmem, p_tmem + k * 8, dv + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
          umma_commit(&pv_done[w]);
        }
        __syncwarp();
      };
      if (single) {
        // one key tile per query pair, K / V resident: pipeline ACROSS query pairs
        if (blockIdx.x < p.n_qpairs) {
          ok = mbar_wait(&kv_full[0], 0, p.abort_flag) && mbar_wait(&q_full[0], 0, p.abort_flag);
          if (ok) {
            tc_fence_after();
            const int n_w0 = (blockIdx.x * 2 * ATT_BM + ATT_BM < p.n_q) ? 2 : 1;
            for (int w = 0; w < n_w0; ++w) issue_s(w, 0, 0);
          }
        }
        for (int qp = blockIdx.x; qp < p.n_qpairs && ok; qp += gridDim.x, ++it) {
          const int slot = it & 1;
          const int n_w = (qp * 2 * ATT_BM + ATT_BM < p.n_q) ? 2 : 1;
          const int qn = qp + gridDim.x;
          const bool has_next = qn < p.n_qpairs;
          const int n_wn = has_next ? ((qn * 2 * ATT_BM + ATT_BM < p.n_q) ? 2 : 1) : 0;
          if (has_next && !mbar_wait(&q_full[slot ^ 1], ((it + 1) >> 1) & 1, p.abort_flag)) {
            ok = false;
            break;
          }
          for (int w = 0; w < n_w && ok; ++w) {
            ok = mbar_wait(&s_free[w], gp[w] & 1, p.abort_flag);
            if (ok && w < n_wn) {
              tc_fence_after();
              issue_s(w, 0, slot ^ 1);
            }
          }
          if (!ok) break;
          for (int w = 0; w < n_w; ++w) {
            if (!mbar_wait(&p_full[w], gp[w] & 1, p.abort_flag)) {
              ok = false;
              break;
            }
            ++gp[w];
            tc_fence_after();
            issue_pv(w, 0, 0, pv_steps_last);
          }
          if (!ok) break;
          if (elect_one()) umma_commit(&q_empty[slot]);
          __syncwarp();
        }
      } else
      for (int qp = blockIdx.x; qp < p.n_qpairs && ok; qp += gridDim.x, ++it) {
        const int q0 = qp * 2 * ATT_BM;
        const int n_w = (q0 + ATT_BM < p.n_q) ? 2 : 1;
        const int slot = it & 1;
        if (!mbar_wait(&q_full[slot], (it >> 1) & 1, p.abort_flag)) break;
        if (!mbar_wait(&kv_full[stage], phase, p.abort_flag)) break;
        tc_fence_after();
        for (int w = 0; w < n_w; ++w) issue_s(w, stage, slot);
        for (int j = 0; j < n_tiles && ok; ++j) {
          int ns = stage + 1;
          uint32_t nphase = phase;
          if (ns == ATT_KV_STAGES) {
            ns = 0;
            nphase ^= 1;
          }
          const bool more = j + 1 < n_tiles;
          if (more && !mbar_wait(&kv_full[ns], nphase, p.abort_flag)) {
            ok = false;
            break;
          }
          // S_w of the next key tile as soon as the softmax warps hold the current S in registers
          for (int w = 0; w < n_w && ok; ++w) {
            ok = mbar_wait(&s_free[w], gp[w] & 1, p.abort_flag);
            if (ok && more) {
              tc_fence_after();
              issue_s(w, ns, slot);
            }
          }
          if (!ok) break;
          // O_w (+)= P_w V_j once P_w is in tensor memory
          for (int w = 0; w < n_w; ++w) {
            if (!mbar_wait(&p_full[w], gp[w] & 1, p.abort_flag)) {
              ok = false;
              break;
            }
            ++gp[w];
            tc_fence_after();
            issue_pv(w, stage, j, more ? ATT_BN / 16 : pv_steps_last);
          }
          if (!ok) break;
          if (elect_one()) umma_commit(&kv_empty[stage]);
          __syncwarp();
          stage = ns;
          phase = nphase;
        }
        if (elect_one()) umma_commit(&q_empty[slot]);
        __syncwarp();
      }
    }
  }
  } else {
    // ------------------------------- softmax ------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int w = (warp - 4) >> 2;   // Q tile of this warpgroup
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + ATT_TMEM_S + w * ATT_BN;
    const uint32_t p_addr = lane_addr + ATT_TMEM_P + w * (ATT_BN / 2);
    const uint32_t o_addr = lane_addr + ATT_TMEM_O + w * ATT_D;
    uint32_t g = 0;   // key tiles processed (phase of s_full)
    uint32_t gd = 0;  // pv_done phases consumed
    bool ok = true;
    for (int qp = blockIdx.x; qp < p.n_qpairs && ok; qp += gridDim.x) {
      const int q0 = qp * 2 * ATT_BM + w * ATT_BM;
      if (q0 >= p.n_q) continue;  // second tile of a short sequence: nothing to do (warpgroup-uniform)
      float m_used = -INFINITY, l_run = 0.f;
      // One key tile. `ragged` is a compile-time tag: only the last tile of a key count that is not a multiple of 128
      // masks its tail; full tiles carry no per-element compare / select (they were 36 % of the instructions of this
      // issue-bound loop when the mask was a run-time condition that ptxas if-converted).
      auto key_tile = [&](int j, auto ragged) -> bool {
        if (!mbar_wait(&s_full[w], g & 1, p.abort_flag)) return false;
        tc_fence_after();
        uint32_t s[ATT_BN];
        tmem_ld_32x32(s_addr, s);
        tmem_ld_wait();
        tmem_ld_32x32(s_addr + 32, s + 32);  // the rest of the row arrives under the first chunk's exponentials
        tmem_ld_32x32(s_addr + 64, s + 64);
        tmem_ld_32x32(s_addr + 96, s + 96);
        const int kv_left = p.n_kv - j * ATT_BN;  // keys valid in this tile (>= 1)
        // 32-key chunks that hold at least one valid key: fully masked chunks of a ragged tile are skipped (their P
        // columns stay unwritten; PV of this tile only runs over ceil(kv_left / 16) key steps)
        int nch = ATT_BN / 32;
        if constexpr (decltype(ragged)::value) nch = (kv_left + 31) >> 5;
        const bool optimistic = (j > 0);          // exponentials against the STALE max (checked afterwards)
        float mx0 = -INFINITY, mx1 = -INFINITY;
        uint64_t lsum = 0ull;                     // packed (l0, l1) partial row sums of this tile
        const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
        uint64_t negm2 = pack_f32x2(-m_used, -m_used);
        // x = s * scale - m  ->  p = 2^x for 32 columns, packed to bf16. ATT_POLY_PAIRS of every ATT_POLY_PERIOD
        // column pairs may run 2^x on the FMA/ALU pipes (Cody-Waite + degree-3 minimax) instead of MUFU.EX2;
        // measured on B200 (profiles/r01_attention_notes.md) the kernel is not XU-bound at head_dim 64 with two
        // softmax warps per scheduler, so the default keeps every exponential on the XU pipe.
        auto exp_chunk = [&](int c, uint32_t* pk) {
          const float lo_clamp = (m_used - 125.f) / p.scale_log2;  // s below this gives p < 2^-125
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int pair = (c * 32 + i) >> 1;
            float p0, p1;
            if ((pair % ATT_POLY_PERIOD) < ATT_POLY_PAIRS) {
              const float s0 = fmaxf(__uint_as_float(s[c * 32 + i]), lo_clamp);
              const float s1 = fmaxf(__uint_as_float(s[c * 32 + i + 1]), lo_clamp);
              exp2_poly_x2(pack_f32x2(s0, s1), scale2, negm2, p0, p1);
            } else {
              float x0, x1;
              unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1])),
                                     scale2, negm2), x0, x1);
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            lsum = add_f32x2(lsum, pack_f32x2(p0, p1));
            pk[i >> 1] = pack_bf16(p0, p1);
          }
        };
        // After the first key tile the row max itself is NOT tracked per tile: the exponentials run against the stale max
        // and the tile's own row sum tells whether that was safe -- sum(p) <= 2^8 implies every p <= 2^8, the same bound
        // the explicit max test gave. Only when the sum exceeds it (or is not finite) is the max computed (S is still in
        // registers) and the tile redone. This takes 64 FMNMX3 per row and tile out of a loop whose XU pipe idles
        // whenever both softmax warps of a scheduler are outside their exponentials.
        auto mask_chunk = [&](int c) {
          if constexpr (decltype(ragged)::value) {  // ragged last key tile: -inf -> p = 0
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= kv_left) s[c * 32 + i] = 0xff800000u;
          }
        };
        auto max_only = [&](int c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(s[c * 32 + i]));
            mx1 = fmaxf(mx1, __uint_as_float(s[c * 32 + i + 1]));
          }
        };
        auto max_chunk = [&](int c) {
          mask_chunk(c);
#ifdef ATT_TRACK_MAX
          max_only(c);
#else
          if (!optimistic) max_only(c);
#endif
        };
        uint32_t pk0[16], pk1[16];
        max_chunk(0);
        if (optimistic) exp_chunk(0, pk0);
        tmem_ld_wait();
        // the whole S row is in registers: hand S back so QK^T of the next key tile runs under the rest
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[w]);
        if (nch > 1) {
          max_chunk(1);
          if (optimistic) exp_chunk(1, pk1);
        }
        if (j > 0) {  // PV of the previous tile must be complete before P is overwritten / O rescaled
          const bool done = mbar_wait(&pv_done[w], gd & 1, p.abort_flag);
          ++gd;
          if (!done) return false;
          tc_fence_after();
        }
        if (optimistic) {
          tmem_st_32x16(p_addr, pk0);
          if (nch > 1) tmem_st_32x16(p_addr + 16, pk1);
        }
#pragma unroll
        for (int c = 2; c < ATT_BN / 32; ++c) {
          if (c < nch) {
            max_chunk(c);
            if (optimistic) {
              exp_chunk(c, pk0);
              tmem_st_32x16(p_addr + c * 16, pk0);
            }
          }
        }
#ifdef ATT_TRACK_MAX
        const bool need = fmaxf(mx0, mx1) * p.scale_log2 > m_used + ATT_RESCALE_LOG2;  // first tile: m_used = -inf
#else
        bool need = true;  // first tile: the max is known, nothing has been exponentiated yet
        if (optimistic) {
          float l0, l1;
          unpack_f32x2(lsum, l0, l1);
          need = !(l0 + l1 <= 256.f);  // (also true for inf / nan)
        }
#endif
        if (__any_sync(0xffffffffu, need)) {
          // rare after the first tile: the running max grew by more than 2^8. Rescale O and redo this tile's
          // exponentials against the new max (S is still in registers, P has not been handed to the MMA yet).
#ifndef ATT_TRACK_MAX
          if (optimistic) {
#pragma unroll
            for (int c = 0; c < ATT_BN / 32; ++c)
              if (c < nch) max_only(c);
          }
#endif
          const float m_cand = fmaxf(mx0, mx1) * p.scale_log2;
          const float m_new = need ? fmaxf(m_cand, m_used) : m_used;
          if (j > 0) {
            const float factor = need ? ex2_approx(m_used - m_new) : 1.f;
            uint32_t o[ATT_D];
            tmem_ld_32x32(o_addr, o);
            tmem_ld_32x32(o_addr + 32, o + 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < ATT_D; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32(o_addr, o);
            tmem_st_32x32(o_addr + 32, o + 32);
            l_run *= factor;
          }
          m_used = m_new;
          negm2 = pack_f32x2(-m_used, -m_used);
          lsum = 0ull;
#pragma unroll
          for (int c = 0; c < ATT_BN / 32; ++c) {
            if (c < nch) {
              exp_chunk(c, pk0);
              tmem_st_32x16(p_addr + c * 16, pk0);
            }
          }
        }
        {
          float l0, l1;
          unpack_f32x2(lsum, l0, l1);
          l_run += l0 + l1;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w]);
        return true;
      };
      for (int j = 0; j < n_tiles && ok; ++j, ++g) {
        if (p.n_kv - j * ATT_BN >= ATT_BN) {
          ok = key_tile(j, std::false_type{});
        } else {
          ok = key_tile(j, std::true_type{});
        }
      }
      if (!ok) break;
      ok = mbar_wait(&pv_done[w], gd & 1, p.abort_flag);
      ++gd;
      if (!ok) break;
      tc_fence_after();
      {
        const float inv_l = 1.f / l_run;
        const int qrow = q0 + r;
        if (p.lse2 && qrow < p.n_q) p.lse2[((size_t)b * p.max_heads + head) * p.n_q + qrow] = m_used + log2f(l_run);
        __nv_bfloat16* op = p.out + ((size_t)b * p.n_q + qrow) * p.ldo + head * ATT_D;
        uint32_t o[ATT_D];
        tmem_ld_32x32(o_addr, o);
        tmem_ld_32x32(o_addr + 32, o + 32);
        tmem_ld_wait();
        if (qrow < p.n_q) {
#pragma unroll
          for (int q = 0; q < ATT_D / 8; ++q) {
            const float* f = reinterpret_cast<const float*>(o) + q * 8;
            *reinterpret_cast<uint4*>(op + q * 8) =
                make_uint4(pack_bf16(f[0] * inv_l, f[1] * inv_l), pack_bf16(f[2] * inv_l, f[3] * inv_l),
                           pack_bf16(f[4] * inv_l, f[5] * inv_l), pack_bf16(f[6] * inv_l, f[7] * inv_l));
          }
        }
      }
      tc_fence_before();  // O reads are complete before the next item's p_full arrive lets PV overwrite O
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

static int g_attn_smem_set = 0;

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_attention_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                  void* out, int32_t ldo, int32_t batch, int32_t n_q, int32_t n_kv,
                                  const int32_t* sample_heads, int32_t max_heads, float scale, float* lse2,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(q && k && v && out && sample_heads, "aptp_attention_fwd: null pointer");
  APTP_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "aptp_attention_fwd: pitches must be multiples of 8");
  APTP_REQUIRE(n_q > 0 && n_kv > 0, "aptp_attention_fwd: empty sequence");
  APTP_REQUIRE(scale > 0.f, "aptp_attention_fwd: scale must be positive");
  if (batch == 0 || max_heads == 0) return APTP_OK;
  APTP_REQUIRE(max_heads <= 65535 && batch <= 65535, "aptp_attention_fwd: grid too large");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {(uint64_t)max_heads * ATT_D, (uint64_t)batch * n_q};
    uint64_t strides[1] = {(uint64_t)ldq * 2};
    uint32_t box[2] = {ATT_D, ATT_BM};
    int rc = make_tmap_bf16(&p.tmap_q, q, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)max_heads * ATT_D, (uint64_t)batch * n_kv};
    uint64_t strides[1] = {(uint64_t)ldk * 2};
    uint32_t box[2] = {ATT_D, ATT_BN};
    int rc = make_tmap_bf16(&p.tmap_k, k, 2, dims, strides, box);
    if (rc) return rc;
    strides[0] = (uint64_t)ldv * 2;
    rc = make_tmap_bf16(&p.tmap_v, v, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.n_q = n_q;
  p.n_kv = n_kv;
  p.n_qpairs = (n_q + 2 * ATT_BM - 1) / (2 * ATT_BM);
  p.sample_heads = sample_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.max_heads = max_heads;
  p.abort_flag = device_abort_flag();
  APTP_REQUIRE(p.abort_flag != nullptr, "aptp_attention_fwd: could not allocate abort flag");
  if (!g_attn_smem_set) {
    APTP_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    g_attn_smem_set = 1;
  }
  // long KV: one query pair per CTA (32+ KV tiles amortise the setup); short KV (cross-attention): each CTA
  // walks several query pairs, keeping ~8 CTAs per SM worth of work items in the grid
  int gx = p.n_qpairs;
  if (n_kv <= 2 * ATT_BN) {
    const long long ctas_other = (long long)max_heads * batch;
    long long want = (8LL * sm_count() + ctas_other - 1) / ctas_other;
    if (want < 1) want = 1;
    if (want < gx) gx = (int)want;
  }
  dim3 grid(gx, max_heads, batch);
  APTP_CUDA_CHECK(launch_pdl(attention_kernel, grid, dim3(ATT_THREADS), (size_t)ATT_SMEM_BYTES, stream, p));
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
