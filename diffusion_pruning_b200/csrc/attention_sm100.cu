// K3: flash-style attention (head_dim 64, no mask, non-causal) on tcgen05 with per-sample kept-head
// lists. One CTA = 128 queries of one (sample, kept head); pruned heads and depth-dropped samples
// launch no work (the reference multiplies q, k, v of a pruned head by zero and still runs SDPA over
// it: pdm/models/unet/blocks.py:250-260).
//
//   warp 0      TMA producer: Q once, then a ring of (K, V) tiles of 128 keys
//   warp 1      MMA issuer:   S_j = Q K_j^T  (M128 x N128 x K64, both operands K-major)
//                             O_j = P_j V_j  (M128 x N64 x K128, V consumed MN-major straight from
//                                             its [keys, d] layout -- no transpose pass)
//   warps 2..5  softmax:      one thread per query row; S from TMEM (tcgen05.ld), online max/sum in
//                             registers, P written bf16 into 128B-swizzled smem for the PV MMA,
//                             partial O_j read back from TMEM and accumulated in fp32 registers.
// S, P and partial-O are double buffered so QK^T of tile j+1 and PV of tile j overlap the softmax.
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

constexpr int ATT_THREADS = 192;
constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
// smem map (from 1024-aligned base): Q | P0a P0b | P1a P1b | (K,V) x stages | barriers
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_P = ATT_TILE_BYTES;
constexpr int ATT_SMEM_KV = ATT_SMEM_P + 4 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BAR = ATT_SMEM_KV + ATT_KV_STAGES * 2 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 256 + 1024;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_TMEM_S = 0;      // S0 at cols [0,128), S1 at [128,256)
constexpr int ATT_TMEM_O = 256;    // O0 at [256,320), O1 at [320,384)

struct AttnParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;
  __nv_bfloat16* out;
  int ldo;
  int n_q, n_kv;
  const int* sample_heads;
  float scale_log2;  // softmax scale * log2(e)
  int* abort_flag;
};

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  if (head >= p.sample_heads[b]) return;  // pruned head / depth-dropped sample: no work at all

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
  uint64_t* q_full = bars;                       // 1
  uint64_t* kv_full = bars + 1;                  // stages
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;  // stages
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;   // 2
  uint64_t* p_full = s_full + 2;                 // 2
  uint64_t* o_full = p_full + 2;                 // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.n_kv + ATT_BN - 1) / ATT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_k);
    tma_prefetch_desc(&p.tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < ATT_KV_STAGES; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&o_full[i], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(smem + ATT_SMEM_Q, &p.tmap_q, q_full, head * ATT_D, b * p.n_q + qt * ATT_BM);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        if (!mbar_wait(&kv_empty[stage], phase ^ 1, p.abort_flag)) break;
        uint8_t* sk = smem + ATT_SMEM_KV + stage * 2 * ATT_TILE_BYTES;
        mbar_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
        tma_load_2d(sk, &p.tmap_k, &kv_full[stage], head * ATT_D, b * p.n_kv + j * ATT_BN);
        tma_load_2d(sk + ATT_TILE_BYTES, &p.tmap_v, &kv_full[stage], head * ATT_D, b * p.n_kv + j * ATT_BN);
        if (++stage == ATT_KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // B (=V) is MN-major
      const uint32_t q_addr = smem_u32(smem + ATT_SMEM_Q);
      bool ok = mbar_wait(q_full, 0, p.abort_flag);
      // S_0
      if (ok) ok = mbar_wait(&kv_full[0], 0, p.abort_flag);
      if (ok) {
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + ATT_SMEM_KV);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_bf16_ss(tmem_base + ATT_TMEM_S, make_desc_kmajor_sw128(q_addr + k * 32),
                       make_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0);
        umma_commit(&s_full[0]);
      }
      for (int j = 0; j < n_tiles && ok; ++j) {
        const int stage = j % ATT_KV_STAGES;
        if (j + 1 < n_tiles) {
          const int ns = (j + 1) % ATT_KV_STAGES;
          const uint32_t nphase = ((j + 1) / ATT_KV_STAGES) & 1;
          ok = mbar_wait(&kv_full[ns], nphase, p.abort_flag);
          if (!ok) break;
          tc_fence_after();
          const uint32_t k_addr = smem_u32(smem + ATT_SMEM_KV + ns * 2 * ATT_TILE_BYTES);
          const uint32_t s_tmem = tmem_base + ATT_TMEM_S + ((j + 1) & 1) * ATT_BN;
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16_ss(s_tmem, make_desc_kmajor_sw128(q_addr + k * 32), make_desc_kmajor_sw128(k_addr + k * 32),
                         idesc_s, k != 0);
          umma_commit(&s_full[(j + 1) & 1]);
        }
        // O_j = P_j V_j
        ok = mbar_wait(&p_full[j & 1], (j >> 1) & 1, p.abort_flag);
        if (!ok) break;
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + ATT_SMEM_P + (j & 1) * 2 * ATT_TILE_BYTES);
        const uint32_t v_addr = smem_u32(smem + ATT_SMEM_KV + stage * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
        const uint32_t o_tmem = tmem_base + ATT_TMEM_O + (j & 1) * ATT_D;
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          // P: two K-major [128 x 64] chunks; V: 16 keys = 2 swizzle atoms of 8 rows (2048 B) per step
          const uint32_t pa = p_addr + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32;
          umma_bf16_ss(o_tmem, make_desc_kmajor_sw128(pa), make_desc_mnmajor_sw128(v_addr + k * 2048, ATT_TILE_BYTES),
                       idesc_o, k != 0);
        }
        umma_commit(&kv_empty[stage]);
        umma_commit(&o_full[j & 1]);
      }
    }
  } else {
    // ------------------------------- softmax / accumulate -------------------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float o_acc[ATT_D];
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    bool ok = true;
    for (int j = 0; j < n_tiles && ok; ++j) {
      ok = mbar_wait(&s_full[j & 1], (j >> 1) & 1, p.abort_flag);
      if (!ok) break;
      tc_fence_after();
      const uint32_t s_addr = lane_addr + ATT_TMEM_S + (j & 1) * ATT_BN;
      const int kv_left = p.n_kv - j * ATT_BN;  // keys valid in this tile (>= 1)
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(s_addr + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(s[i]));
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);  // first tile: exp2(-inf) = 0
      // pass 2: p = exp2(s*scale - m), row sum, bf16 P into swizzled smem
      uint8_t* pbuf = smem + ATT_SMEM_P + (j & 1) * 2 * ATT_TILE_BYTES;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(s_addr + c * 32, s);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c * 32 + i < kv_left) ? exp2f(__uint_as_float(s[i]) * p.scale_log2 - m_new) : 0.f;
          float p1 = (c * 32 + i + 1 < kv_left) ? exp2f(__uint_as_float(s[i + 1]) * p.scale_log2 - m_new) : 0.f;
          // sum what the tensor core will actually see (bf16-rounded), like flash-attention does not:
          // keep fp32 sum of unrounded p (matches SDPA's fp32 softmax more closely)
          lsum += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        // keys [c*32, c*32+32) -> chunk (c>>1), 16-byte units u = (c&1)*4 + q, swizzled with row%8
        uint8_t* rowp = pbuf + (c >> 1) * ATT_TILE_BYTES + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = ((c & 1) * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(rowp + u * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      // make P visible to the async proxy (tensor core reads smem), then release S and publish P
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      // accumulate the previous tile's partial O (its PV ran while we did this softmax)
      if (j > 0) {
        ok = mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1, p.abort_flag);
        if (!ok) break;
        tc_fence_after();
        const uint32_t o_addr = lane_addr + ATT_TMEM_O + ((j - 1) & 1) * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(o_addr + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(o[i]);
        }
      }
      alpha_prev = alpha;
    }
    if (ok) {
      const int j = n_tiles - 1;
      ok = mbar_wait(&o_full[j & 1], (j >> 1) & 1, p.abort_flag);
      if (ok) {
        tc_fence_after();
        const uint32_t o_addr = lane_addr + ATT_TMEM_O + (j & 1) * ATT_D;
        const float inv_l = 1.f / l_run;
        const int qrow = qt * ATT_BM + r;
        __nv_bfloat16* op = p.out + ((size_t)b * p.n_q + qrow) * p.ldo + head * ATT_D;
#pragma unroll
        for (int c = 0; c < ATT_D / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(o_addr + c * 32, o);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = (o_acc[c * 32 + i] * alpha_prev + __uint_as_float(o[i])) * inv_l;
          if (qrow < p.n_q) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(op + c * 32 + q * 8) =
                  make_uint4(pack_bf16(f[q * 8], f[q * 8 + 1]), pack_bf16(f[q * 8 + 2], f[q * 8 + 3]),
                             pack_bf16(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16(f[q * 8 + 6], f[q * 8 + 7]));
          }
        }
      }
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
                                  const int32_t* sample_heads, int32_t max_heads, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(q && k && v && out && sample_heads, "aptp_attention_fwd: null pointer");
  APTP_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "aptp_attention_fwd: pitches must be multiples of 8");
  APTP_REQUIRE(n_q > 0 && n_kv > 0, "aptp_attention_fwd: empty sequence");
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
  p.sample_heads = sample_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.abort_flag = device_abort_flag();
  APTP_REQUIRE(p.abort_flag != nullptr, "aptp_attention_fwd: could not allocate abort flag");
  if (!g_attn_smem_set) {
    APTP_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    g_attn_smem_set = 1;
  }
  dim3 grid((n_q + ATT_BM - 1) / ATT_BM, max_heads, batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(p);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
