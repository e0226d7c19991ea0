// fetch_gran.cu — microbenchmark: DRAM bytes fetched per random 16-byte load, per load flavour.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fetch_gran fetch_gran.cu
// Run under ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

template <int MODE>
__global__ void gather(const float4 *__restrict__ a, uint64_t nlines, uint64_t nreq, float4 *out, int span16) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= nreq) return;
  float4 acc = make_float4(0, 0, 0, 0);
  // each thread reads `span16` consecutive 16-byte pieces starting at a random 128-byte line
  uint64_t line = mix(t * 0x9E3779B97F4A7C15ULL + MODE) % nlines;
  const float4 *p = a + line * 8;
  for (int i = 0; i < span16; ++i) {
    float4 v;
    if (MODE == 0) v = p[i];
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    else if (MODE == 3) asm volatile("ld.global.nc.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    else if (MODE == 4) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    else if (MODE == 5) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
    else if (MODE == 6) { float4 w; asm volatile("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w), "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(p + 2 * i)); v.x += w.x; }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (acc.x == 12345.f) out[0] = acc;
}

int main(int argc, char **argv) {
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("set L2 fetch granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), got);
  } else { size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", got); }
  const uint64_t bytes = 8ULL << 30, nlines = bytes / 128, nreq = 16ULL << 20;
  float4 *a, *out;
  cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes); cudaMalloc(&out, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int span = 1; span <= 2; ++span) {
#define RUN(M) { cudaEventRecord(e0); gather<M><<<(unsigned)(nreq / 256), 256>>>(a, nlines, nreq, out, span); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("mode %d span %d B: %.3f ms  %.1f Mreq/ms\n", M, span * 16, ms, nreq / ms / 1e3); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
