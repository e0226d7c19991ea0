// common.cuh — error plumbing and small device helpers shared by the srgnn_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/srgnn_b200.h"

namespace srg {

// thread-local last-error text (srg_last_error)
char *err_buf();
void set_err(const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_err("%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  if (e == cudaErrorMemoryAllocation) return SRG_ERR_NOMEM;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SRG_ERR_NODEV;
  return SRG_ERR_CUDA;
}

#define SRG_CUDA(call)                                                   \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return srg::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define SRG_LAUNCHED()                                                   \
  do {                                                                   \
    srg::g_launches.fetch_add(1, std::memory_order_relaxed);             \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) return srg::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define SRG_REQUIRE(cond, ...)                                           \
  do {                                                                   \
    if (!(cond)) {                                                       \
      srg::set_err(__VA_ARGS__);                                         \
      return SRG_ERR_INVALID;                                            \
    }                                                                    \
  } while (0)

int require_device();  // SRG_OK or SRG_ERR_NODEV (with message)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Stream-ordered scratch of ONE stream: every block is handed back with cudaFreeAsync on that stream when the
// scope ends, on the success path and on every early error return alike (the frees are ordered after the work
// already queued on the stream, so kernels in flight keep their memory).
struct StreamScratch {
  cudaStream_t s;
  void *ptrs[16];
  int count = 0;
  explicit StreamScratch(cudaStream_t s_) : s(s_) {}
  StreamScratch(const StreamScratch &) = delete;
  StreamScratch &operator=(const StreamScratch &) = delete;
  template <typename T> int alloc(T **p, size_t n_elems) {
    *p = nullptr;
    if (count >= 16) {
      set_err("StreamScratch: too many blocks");
      return SRG_ERR_INVALID;
    }
    void *q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, (n_elems ? n_elems : 1) * sizeof(T), s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    ptrs[count++] = q;
    *p = static_cast<T *>(q);
    return SRG_OK;
  }
  ~StreamScratch() {
    for (int i = count - 1; i >= 0; --i) cudaFreeAsync(ptrs[i], s);
  }
};

// ---- device helpers --------------------------------------------------------------------
// streamed-once data (CSR arrays): read-only path, do not allocate in L1
__device__ __forceinline__ int ld_stream_i32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// gathered feature rows: read-only path, 128-bit, 64-byte DRAM fetch granule on an L2 miss
// (the default fetches the whole 128-byte line: tools/micro/fetch_gran.cu)
__device__ __forceinline__ float4 ld_gather_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace srg
