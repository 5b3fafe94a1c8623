// Shared device/host helpers for the sm_100a kernels of the APTP gated U-Net hot path.
// Thin inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA/TMEM),
// plus the process-wide error slot behind aptp_last_error().
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define APTP_OK 0
#define APTP_ERR_INVALID (-1)
#define APTP_ERR_CUDA (-2)
#define APTP_ERR_UNSUPPORTED (-3)

namespace aptp {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
int* device_abort_flag();  // lazily-allocated device int the pipelined kernels set on mbarrier timeout

#define APTP_CUDA_CHECK(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return aptp::cuda_fail(_e, #expr);   \
  } while (0)

#define APTP_REQUIRE(cond, ...)                                 \
  do {                                                          \
    if (!(cond)) {                                              \
      aptp::set_error(__VA_ARGS__);                             \
      return APTP_ERR_INVALID;                                  \
    }                                                           \
  } while (0)

// ---- driver entry point for tensor-map encoding (no link-time libcuda dependency) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// Build a bf16 tiled tensor map with 128B swizzle. dims/strides fastest-first; strides_bytes[0] is
// implicit (2 bytes) and strides_bytes[i] (i>=1) is the byte stride of dim i.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128 = true);

int make_tmap_store64(CUtensorMap* out, const void* base, bool f32, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes);

// ---- programmatic dependent launch ----
// Kernels of the forward are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs may become
// resident (and run their prologue: barrier init, tensor-memory allocation, descriptor prefetch) while the previous kernel
// on the stream is still draining; `pdl_wait()` -- executed before the first access to anything a predecessor wrote --
// blocks until that kernel has completed and its writes are visible. `pdl_launch()` lets the NEXT kernel do the same. In a
// captured CUDA graph the attribute becomes a programmatic edge; without the attribute the two instructions are no-ops.
// Measured on B200 inside the captured graph of the forward (same box): chaining the GEMM and attention kernels this way
// (level 1, the default) 46.5 vs 46.9 ms per step without; chaining the normalisation / statistics kernels as well (level
// 2) is SLOWER (48.5 ms: their many small CTAs become resident early and sit on the SMs the draining GEMM still uses).
// APTP_PDL=0 / 1 / 2 selects the level.
int pdl_level();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_at(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_level() >= level ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  return launch_pdl_at(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must never hang the GPU box. On timeout (~2 s) the abort flag is
// raised and the caller unwinds; the host code polls the flag once per forward / step without synchronising
// (aptp_poll_abort, called from UNet2DConditionModelGated.forward) and at its own sync points (aptp_check_abort).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* abort_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  // The poll loop itself touches nothing but the barrier (try_wait suspends the thread for a hardware-bounded time):
  // a global read of the abort flag on EVERY failed poll made each blocked hand-off cost a ~1 us memory round trip
  // (ncu, round 2: 21 polls per tile in the MMA issuer of the K = 320 layers). Clock and flag are looked at every 64 polls.
  const long long t0 = clock64();
#ifndef APTP_MBAR_POLLS
#define APTP_MBAR_POLLS 64
#endif
#pragma unroll 1
  for (uint32_t spin = 1;; ++spin) {
    if (mbar_try_wait(bar, parity)) return true;
#ifdef APTP_MBAR_SLEEP
    __nanosleep(APTP_MBAR_SLEEP);
#endif
    if ((spin & (uint32_t)(APTP_MBAR_POLLS - 1)) == 0u) {
      if (clock64() - t0 > 4000000000LL || *((volatile int*)abort_flag) != 0) {
        atomicExch(abort_flag, 1);
        return false;
      }
    }
  }
}

// ---- TMA loads (tile mode), completion signalled on an mbarrier ----
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- TMA stores (bulk async-group completion): shared -> global, issued by one thread ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (the buffer may be rewritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- thread-block clusters ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster (also orders shared-memory / mbarrier initialisation cluster-wide)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset (and mbarrier) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA-pair ("2-SM") forms: one tcgen05.mma.cta_group::2 of M = 256 spans both CTAs of a cluster of two ----
// Each CTA stages its own 128 rows of A and HALF of the B tile; the tensor cores of both SMs read both halves, so the
// shared-memory operand traffic per SM drops from (A + B) to (A + B/2) per MMA and B is never duplicated.
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (peer bit 24 of the shared::cluster
// address cleared), data into the executing CTA's own shared memory.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {  // one warp of EACH CTA, same slot
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N columns: N/2 per CTA]; issued by ONE thread of the leader.
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same offset in every CTA of cta_mask once all prior MMAs of the pair complete
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (used by attention: P stays in tensor memory).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane_base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns store (registers -> TMEM); used to park bf16-packed P for the PV MMA.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 32 lanes x 32 columns store (registers -> TMEM); attention parks bf16-packed P here for the PV MMA.
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Same instruction, but pinned in program order relative to the other volatile asm statements (tcgen05.ld /
// tcgen05.wait / mbarrier ops): software-pipelined loops use it so the compiler cannot sink the math of chunk k
// below the wait for chunk k+1.
__device__ __forceinline__ float ex2_approx_ordered(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Explicit shared-space 16-byte accesses (a pointer derived from the aligned dynamic-smem base by integer
// arithmetic is generic to the compiler, which then emits LD.E / ST.E instead of LDS / STS).
__device__ __forceinline__ void lds_f32x2x2(uint32_t addr, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_u32x4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128B swizzle: rows of 64 bf16 (128 B), 8-row
// atoms of 1024 B (SBO), LBO field = 1 (unused for swizzled K-major), version = 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same with an explicit stride between 8-row groups and a start that is NOT aligned to the 1024-byte swizzle
// pattern (rows of a larger tile picked out by moving the start address: the conv halo tile). Measured on B200: the
// tensor core applies the 128B swizzle to the ABSOLUTE shared-memory address (bits 4-6 ^= bits 7-9), exactly as TMA wrote
// the tile, so the descriptor's base-offset field stays 0 for any 128-byte-aligned start (setting it to the row phase
// (start >> 7) & 7 gives wrong products).
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128_ex(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand, 128B swizzle: 64 contiguous bf16 along MN per row (128 B), 8 k-rows per 1024 B
// atom (SBO between k-groups of 8), LBO = byte stride between 64-element MN atoms.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor: bf16 A/B, fp32 accumulate, dense; majors: 0 = K-major, 1 = MN-major.
__device__ __host__ __forceinline__ uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                             uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // c_format = F32
  d |= 1u << 7;   // a_format = BF16
  d |= 1u << 10;  // b_format = BF16
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((N >> 3) & 0x3F) << 17;
  d |= ((M >> 4) & 0x1F) << 24;
  return d;
}

// ---- warp-collective ("elected") issue ----
// Called by ALL 32 lanes of a converged warp; one elected lane issues the instruction. Keeping the issuing
// warp in uniform control flow lets ptxas hold descriptors / barrier addresses in uniform registers; a loop
// nested under `if (lane == 0)` instead wraps every UTCHMMA / UTMALDG in an R2UR waterfall loop (measured:
// ~140 cycles per MMA instruction, the limiter of the round-1 GEMM and attention kernels).
__device__ __forceinline__ void umma_bf16_ss_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  if (elect_one()) umma_bf16_ss(tmem_d, desc_a, desc_b, idesc, accumulate);
}
__device__ __forceinline__ void umma_bf16_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  if (elect_one()) umma_bf16_ts(tmem_d, tmem_a, desc_b, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
}
__device__ __forceinline__ void tma_load_2d_e(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  if (elect_one()) tma_load_2d(smem, map, bar, c0, c1);
}
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
  if (elect_one()) mbar_expect_tx(bar, bytes);
}

// ---- packed fp32x2 math (FFMA2 / FADD2 on sm_100: two lanes per issue slot) ----
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^(s*scale + negm) for two values without the XU pipe. x is formed twice: once plain, once with the
// 1.5*2^23 magic constant added so the rounded integer part n lands in the low mantissa bits; r = x - n is in
// [-0.5, 0.5]; 2^r by a degree-3 minimax polynomial (relative error 7.5e-5, far below bf16 rounding 2^-9 = 2e-3);
// 2^n by adding n to the exponent field. The caller clamps s so that x >= -125.
__device__ __forceinline__ void exp2_poly_x2(uint64_t s2, uint64_t scale2, uint64_t negm2, float& p0, float& p1) {
  const uint64_t magic2 = pack_f32x2(12582912.f, 12582912.f);
  const uint64_t nmagic2 = pack_f32x2(-12582912.f, -12582912.f);
  const uint64_t x2 = fma_f32x2(s2, scale2, negm2);
  const uint64_t xf2 = add_f32x2(x2, magic2);           // mantissa low bits = round(x)
  const uint64_t nr2 = add_f32x2(xf2, nmagic2);         // round(x) as float
  const uint64_t r2 = fma_f32x2(nr2, pack_f32x2(-1.f, -1.f), x2);  // r = x - round(x)
  uint64_t q2 = fma_f32x2(r2, pack_f32x2(0.0551716685f, 0.0551716685f), pack_f32x2(0.242611125f, 0.242611125f));
  q2 = fma_f32x2(q2, r2, pack_f32x2(0.693260968f, 0.693260968f));
  q2 = fma_f32x2(q2, r2, pack_f32x2(0.999928057f, 0.999928057f));
  float q0, q1, f0, f1;
  unpack_f32x2(q2, q0, q1);
  unpack_f32x2(xf2, f0, f1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(f0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(f1) << 23));
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// silu(x) = x * sigmoid(x) = 0.5 x (1 + tanh(x/2)): ONE XU op (tanh.approx, |rel err| ~ 2^-11, below bf16
// rounding) instead of ex2 + rcp -- the GroupNorm+SiLU pass would otherwise need ~70 % of the XU pipe at HBM speed
__device__ __forceinline__ float silu_tanh_f(float x) {
  float t;
  const float h = 0.5f * x;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// Exact-erf GELU (blocks.py:50) with erf from Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16
// resolution): two MUFU ops (rcp, ex2) + 9 FMA-pipe ops, branch-free. The same exp(-x^2/2) is the density term
// of the derivative, so value and gradient share both MUFU ops:
//   Phi(x) = 0.5 (1 + erf(x / sqrt 2)),  gelu = x Phi,  gelu' = Phi + x exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ void gelu_erf_fast_parts(float x, float& Phi, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float half_erf = fmaf(-0.5f * poly, e, 0.5f);  // 0.5 * erf(|x|/sqrt2)
  Phi = 0.5f + copysignf(half_erf, x);
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  float Phi, e;
  gelu_erf_fast_parts(x, Phi, e);
  return x * Phi;
}
// GEGLU value for TWO columns, h * gelu(g), with ONE MUFU op per element and packed FFMA2 for the rest (the K = 320
// GEGLU epilogue is issue- and XU-bound: 12 288 outputs per tile against ~1900 tensor-pipe cycles). Uses
//   gelu(x) = x Phi(x) = relu(x) - |x| (1 - Phi(|x|)),   1 - Phi(a) = 2^q(a),
// q = degree-6 fit of log2 of the Gaussian tail on [0, 6], weighted for the ABSOLUTE error of a 2^q(a): <= 9e-8 in
// fp32 including ex2.approx (tools/fit_gelu_tail.py; the Abramowitz-Stegun form above has 1.5e-7 |x|). Beyond a = 6
// the tail is < 1e-9 and q is clamped there.
__device__ __forceinline__ void geglu_pair(float& h0, float& h1, float g0, float g1) {
  // n = -min(|g|, 6): one FMNMX with source modifiers; the polynomial is written in n (odd coefficients negated) so the
  // last step is a single FFMA2, relu(g) + n * tail
  const uint64_t n2 = pack_f32x2(fmaxf(-fabsf(g0), -6.f), fmaxf(-fabsf(g1), -6.f));
#define APTP_C2(c) pack_f32x2(c, c)
  uint64_t q = fma_f32x2(n2, APTP_C2(3.309331805212423e-05f), APTP_C2(0.0007692242506891489f));
  q = fma_f32x2(q, n2, APTP_C2(0.008080732077360153f));
  q = fma_f32x2(q, n2, APTP_C2(0.05341212823987007f));
  q = fma_f32x2(q, n2, APTP_C2(-0.4587709605693817f));
  q = fma_f32x2(q, n2, APTP_C2(1.1512017250061035f));
  q = fma_f32x2(q, n2, APTP_C2(-0.999993085861206f));
#undef APTP_C2
  float q0, q1, e0, e1;
  unpack_f32x2(q, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const uint64_t gl = fma_f32x2(n2, pack_f32x2(e0, e1), pack_f32x2(fmaxf(g0, 0.f), fmaxf(g1, 0.f)));
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(gl), "l"(pack_f32x2(h0, h1)));
  unpack_f32x2(r, h0, h1);
}
__device__ __forceinline__ void gelu_erf_fast_grad(float x, float& value, float& grad) {
  float Phi, e;
  gelu_erf_fast_parts(x, Phi, e);
  value = x * Phi;
  grad = fmaf(x * 0.3989422804014327f, e, Phi);
}
#endif  // __CUDACC__

}  // namespace aptp
