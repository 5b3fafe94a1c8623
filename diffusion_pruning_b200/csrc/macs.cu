// Closed-form MAC accounting of the gated U-Net as one kernel over the [B, 1620] gate matrix (SURVEY 8f rank 3,
// Appendix F). Replaces the module-tree walk of UNet2DConditionModelGated.calc_macs
// (pdm/models/unet/unet_2d_conditional.py:2124-2163; per-block calc_macs at blocks.py:103-119, :144-151, :384-416,
// :598-633, :879-917, :1373-1413) and its ~170 hard_concrete calls (pdm/utils/estimation_utils.py:67-75).
//
// Per gated sub-block s (38 of them) with width gates g in s:   A_s = sum_g ratio_g * P_g,
//   ratio_g = mean_j [gate[g, j] >= 0.5]   (hard_concrete, straight-through gradient d ratio / d gate = 1 / width)
//   no depth gate:  cur_prunable += A_s              cur_total += A_s + F_s
//   depth gate d:   cur_prunable += (A_s + F_s) * d   cur_total += (A_s + F_s) * d     (blocks.py:626-633, :1400-1411)
// with d = [depth >= 0.5] (straight-through), F_s = the sub-block's non-prunable MACs; cur_total carries no gradient.
// One warp per prompt row; lane s handles sub-blocks s and s + 32 (fixed order => deterministic sums, fp64).
#include "common.cuh"
#include "../../include/aptp_sm100.h"

namespace aptp {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double sub_block_A(const float* __restrict__ row, const aptp_macs_gate* __restrict__ gates,
                                              int first, int count) {
  double A = 0.0;
  for (int g = first; g < first + count; ++g) {
    const aptp_macs_gate G = gates[g];
    int kept = 0;
    for (int j = 0; j < G.width; ++j) kept += row[G.col + j] >= 0.5f;
    A += ((double)kept / (double)G.width) * G.macs;
  }
  return A;
}

__global__ void __launch_bounds__(128) macs_ratio_fwd_kernel(const float* __restrict__ arch, int ld, int batch,
                                                             const aptp_macs_gate* __restrict__ gates,
                                                             const aptp_macs_sub* __restrict__ subs, int n_subs,
                                                             double fixed_total, float* __restrict__ cur_prunable,
                                                             float* __restrict__ cur_total) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= batch) return;
  const float* row = arch + (size_t)b * ld;
  double cp = 0.0, ct = 0.0;
  for (int s = lane; s < n_subs; s += 32) {
    const aptp_macs_sub S = subs[s];
    const double A = sub_block_A(row, gates, S.first_gate, S.n_gates);
    if (S.depth_col >= 0) {
      const double d = row[S.depth_col] >= 0.5f ? 1.0 : 0.0;
      cp += (A + S.fixed) * d;
      ct += (A + S.fixed) * d;
    } else {
      cp += A;
      ct += A + S.fixed;
    }
  }
  cp = warp_sum_f64(cp);
  ct = warp_sum_f64(ct);
  if (lane == 0) {
    cur_prunable[b] = (float)cp;
    cur_total[b] = (float)(ct + fixed_total);
  }
}

// darch[b, :] = dcur_prunable[b] * d cur_prunable[b] / d arch[b, :]  (every column of the row is written)
__global__ void __launch_bounds__(128) macs_ratio_bwd_kernel(const float* __restrict__ arch, int ld, int batch,
                                                             const aptp_macs_gate* __restrict__ gates,
                                                             const aptp_macs_sub* __restrict__ subs, int n_subs,
                                                             const float* __restrict__ dcur, float* __restrict__ darch,
                                                             int ldd, int dim) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= batch) return;
  const float* row = arch + (size_t)b * ld;
  float* drow = darch + (size_t)b * ldd;
  for (int c = lane; c < dim; c += 32) drow[c] = 0.f;  // columns not covered by a gate (none in practice)
  __syncwarp();
  const double g = (double)dcur[b];
  for (int s = lane; s < n_subs; s += 32) {
    const aptp_macs_sub S = subs[s];
    double mult = 1.0;
    if (S.depth_col >= 0) {
      mult = row[S.depth_col] >= 0.5f ? 1.0 : 0.0;
      const double A = sub_block_A(row, gates, S.first_gate, S.n_gates);
      drow[S.depth_col] = (float)(g * (A + S.fixed));
    }
    for (int gi = S.first_gate; gi < S.first_gate + S.n_gates; ++gi) {
      const aptp_macs_gate G = gates[gi];
      const float v = (float)(g * mult * G.macs / (double)G.width);
      for (int j = 0; j < G.width; ++j) drow[G.col + j] = v;
    }
  }
}

}  // namespace aptp

using namespace aptp;

extern "C" int aptp_macs_ratio_fwd(const float* arch, int32_t ld, int32_t batch, const aptp_macs_gate* gates,
                                   int32_t n_gates, const aptp_macs_sub* subs, int32_t n_subs, double fixed_total,
                                   float* cur_prunable, float* cur_total, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(arch && gates && subs && cur_prunable && cur_total, "aptp_macs_ratio_fwd: null pointer");
  APTP_REQUIRE(n_gates > 0 && n_subs > 0 && ld > 0, "aptp_macs_ratio_fwd: bad sizes");
  if (batch == 0) return APTP_OK;
  macs_ratio_fwd_kernel<<<(batch + 3) / 4, 128, 0, stream>>>(arch, ld, batch, gates, subs, n_subs, fixed_total,
                                                            cur_prunable, cur_total);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}

extern "C" int aptp_macs_ratio_bwd(const float* arch, int32_t ld, int32_t batch, const aptp_macs_gate* gates,
                                   int32_t n_gates, const aptp_macs_sub* subs, int32_t n_subs, const float* dcur_prunable,
                                   float* darch, int32_t ldd, int32_t dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  APTP_REQUIRE(arch && gates && subs && dcur_prunable && darch, "aptp_macs_ratio_bwd: null pointer");
  APTP_REQUIRE(n_gates > 0 && n_subs > 0 && ld > 0 && ldd >= dim && dim > 0, "aptp_macs_ratio_bwd: bad sizes");
  if (batch == 0) return APTP_OK;
  macs_ratio_bwd_kernel<<<(batch + 3) / 4, 128, 0, stream>>>(arch, ld, batch, gates, subs, n_subs, dcur_prunable, darch,
                                                            ldd, dim);
  APTP_CUDA_CHECK(cudaGetLastError());
  return APTP_OK;
}
