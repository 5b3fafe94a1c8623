// Microbenchmark: TMEM read (tcgen05.ld) and write (tcgen05.st) bandwidth per SM on sm_100a, and MUFU.EX2 rate, the two
// ceilings of the attention kernels' softmax stage (profiles/README.md). One CTA per SM; W warps each read / write
// their own lane quadrant (warp % 4) over 128 fp32 columns per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I diffusion_pruning_b200/csrc -o gpurun_out/tmem_bw tools/microbench/tmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace aptp;

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: tcgen05.ld, 1: tcgen05.st, 2: ex2.approx
__global__ void __launch_bounds__(512, 1) bw_kernel(int mode, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  float acc = 0.f, x = 0.001f * threadIdx.x, x1 = x + 0.1f, x2 = x + 0.2f, x3 = x + 0.3f;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(base + c * 32, r);
      }
      tmem_ld_wait();
      acc += __uint_as_float(r[it & 31]);
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) st32(base + c * 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {  // four independent chains
        x = ex2_approx_ordered(x);
        x1 = ex2_approx_ordered(x1);
        x2 = ex2_approx_ordered(x2);
        x3 = ex2_approx_ordered(x3);
      }
    }
    acc = x + x1 + x2 + x3;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(slot, 512);
  }
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  const char* names[3] = {"tcgen05.ld 32x32b.x32", "tcgen05.st 32x32b.x32", "ex2.approx (4 chains / thread)"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps = 4; warps <= 16; warps *= 2) {
      bw_kernel<<<148, warps * 32, 0>>>(mode, 10, cyc, sink);  // warm-up
      bw_kernel<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148;
      if (mode < 2) {
        const double bytes = (double)iters * warps * 32 * 128 * 4;  // per SM
        printf("%-38s %2d warps/SM: %8.0f cycles, %7.1f B/clk/SM\n", names[mode], warps, avg, bytes / avg);
      } else {
        const double ops = (double)iters * warps * 32 * 128;
        printf("%-38s %2d warps/SM: %8.0f cycles, %7.2f ex2/clk/SM\n", names[mode], warps, avg, ops / avg);
      }
    }
  }
  return 0;
}
