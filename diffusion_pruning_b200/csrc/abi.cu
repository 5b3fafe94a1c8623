// Library plumbing for libaptp_sm100.so: error slot, device abort flag, tensor-map encoding.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/aptp_sm100.h"
#include <stdarg.h>
#include <string.h>

namespace aptp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return APTP_ERR_CUDA;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// One flag per device (rank processes own one device each, but tests may touch several).
static int* g_abort[64] = {nullptr};
int* device_abort_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_abort[dev]) {
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, sizeof(int)) != cudaSuccess) return nullptr;
    g_abort[dev] = p;
  }
  return g_abort[dev];
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (cuda error %d)", (int)e);
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return APTP_ERR_CUDA;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("tensor map: box[%d]=%u out of range", i, box[i]);
      return APTP_ERR_INVALID;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) {
      set_error("tensor map: stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)strides_bytes[i]);
      return APTP_ERR_INVALID;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides,
                   gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return APTP_ERR_CUDA;
  }
  return APTP_OK;
}

// Output tensor map of the GEMM epilogue's bulk stores: rows x cols row-major, box = 32 rows x 64 bytes (32 bf16 or 16
// fp32 columns), 64-byte swizzle -- the layout the epilogue warps already write their staging tile in (16-byte units
// XORed with (row >> 1) & 3).
int make_tmap_store64(CUtensorMap* out, const void* base, bool f32, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return APTP_ERR_CUDA;
  cuuint64_t gdims[2] = {cols, rows};
  cuuint64_t gstrides[1] = {row_stride_bytes};
  cuuint32_t gbox[2] = {f32 ? 16u : 32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  if (row_stride_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("store tensor map: base / row pitch not 16-byte aligned");
    return APTP_ERR_INVALID;
  }
  CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                   gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (store map) failed with CUresult %d (cols %llu rows %llu pitch %llu)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes);
    return APTP_ERR_CUDA;
  }
  return APTP_OK;
}

int pdl_level() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("APTP_PDL");
    v = e ? atoi(e) : 1;
  }
  return v;
}

}  // namespace aptp

extern "C" int aptp_version(void) { return APTP_ABI_VERSION; }

extern "C" const char* aptp_last_error(void) { return aptp::g_err; }

// Non-blocking variant for the hot path: every call (1) looks at the copy of the flag that the PREVIOUS call enqueued,
// if that copy has landed, and (2) enqueues a fresh 4-byte D2H copy into pinned host memory behind the work already
// on `stream`. No host synchronisation; a timeout is reported one call late at the latest by the next sync point.
struct AbortPoll {
  int* h_flag = nullptr;      // pinned
  cudaEvent_t ev = nullptr;
  bool pending = false;
};
static AbortPoll g_poll[64];

extern "C" int aptp_poll_abort(void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return APTP_OK;
  int* flag = aptp::device_abort_flag();
  if (!flag) return APTP_OK;
  AbortPoll& p = g_poll[dev];
  if (!p.h_flag) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&p.h_flag), sizeof(int), cudaHostAllocDefault) != cudaSuccess) return APTP_OK;
    *p.h_flag = 0;
    if (cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming) != cudaSuccess) return APTP_OK;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return APTP_OK;
  if (p.pending && cudaEventQuery(p.ev) == cudaSuccess) {
    p.pending = false;
    if (*p.h_flag != 0) {
      *p.h_flag = 0;
      cudaMemsetAsync(flag, 0, sizeof(int), stream);
      aptp::set_error("a pipelined kernel timed out on an mbarrier during an earlier launch (pipeline bug, bad tensor "
                      "map, or a GPU time-sliced by a profiler / debugger); results since then are invalid");
      return 1;
    }
  }
  if (!p.pending) {
    if (cudaMemcpyAsync(p.h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
        cudaEventRecord(p.ev, stream) == cudaSuccess)
      p.pending = true;
  }
  return APTP_OK;
}

extern "C" int aptp_check_abort(void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int* flag = aptp::device_abort_flag();
  if (!flag) {
    aptp::set_error("aptp_check_abort: no device abort flag");
    return APTP_ERR_CUDA;
  }
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return aptp::cuda_fail(e, "aptp_check_abort");
  if (h != 0) {
    cudaMemsetAsync(flag, 0, sizeof(int), stream);
    cudaStreamSynchronize(stream);
    aptp::set_error("a pipelined kernel timed out on an mbarrier (pipeline bug or bad tensor map)");
    return 1;
  }
  return 0;
}
