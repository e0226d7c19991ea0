// Micro-benchmark: PCIe rate of 2-D (strided) pinned-host <-> device copies against contiguous ones.
// Question: can the host pipeline move FEATURE-COLUMN chunks of the row-major N x F host matrix
// (width = F_chunk * 4 bytes, host pitch = F * 4) at full PCIe speed?  If so, chunk c's hops and its
// device->host copies can start while chunk c+1 is still uploading (full-duplex PCIe).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o memcpy2d memcpy2d.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main() {
  const size_t n = 2449029, F = 100;
  float *h, *d;
  CK(cudaMallocHost(&h, n * F * 4));
  CK(cudaMalloc(&d, n * 104 * 4));
  for (size_t i = 0; i < n * F; i += 1024) h[i] = 1.f;
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  auto time = [&](const char *what, size_t bytes, auto fn) {
    fn();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    for (int r = 0; r < 3; ++r) fn();
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-58s %8.2f ms  %6.1f GB/s\n", what, ms / 3, bytes / (ms / 3) / 1e6);
  };
  time("H2D contiguous n*F*4", n * F * 4, [&] { CK(cudaMemcpyAsync(d, h, n * F * 4, cudaMemcpyHostToDevice, s)); });
  time("D2H contiguous n*F*4", n * F * 4, [&] { CK(cudaMemcpyAsync(h, d, n * F * 4, cudaMemcpyDeviceToHost, s)); });
  time("H2D 2D width 400 (host pitch 400 -> device pitch 416)", n * F * 4,
       [&] { CK(cudaMemcpy2DAsync(d, 416, h, 400, 400, n, cudaMemcpyHostToDevice, s)); });
  time("D2H 2D width 400 (device pitch 416 -> host pitch 400)", n * F * 4,
       [&] { CK(cudaMemcpy2DAsync(h, 400, d, 416, 400, n, cudaMemcpyDeviceToHost, s)); });
  for (int w : {200, 100, 52}) {
    char buf[128];
    snprintf(buf, sizeof buf, "H2D 2D width %d (host pitch 400 -> device pitch %d)", w, (w + 31) / 32 * 32);
    time(buf, n * w, [&] { CK(cudaMemcpy2DAsync(d, (w + 31) / 32 * 32, h, 400, w, n, cudaMemcpyHostToDevice, s)); });
    snprintf(buf, sizeof buf, "D2H 2D width %d (device pitch %d -> host pitch 400)", w, (w + 31) / 32 * 32);
    time(buf, n * w, [&] { CK(cudaMemcpy2DAsync(h, 400, d, (w + 31) / 32 * 32, w, n, cudaMemcpyDeviceToHost, s)); });
  }
  // full duplex: both directions at once on two streams
  cudaStream_t s2;
  CK(cudaStreamCreate(&s2));
  float *h2;
  CK(cudaMallocHost(&h2, n * F * 4));
  CK(cudaEventRecord(a, s));
  CK(cudaMemcpyAsync(d, h, n * F * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h2, d + n * 2, n * F * 4, cudaMemcpyDeviceToHost, s2));
  CK(cudaStreamSynchronize(s2));
  CK(cudaEventRecord(b, s));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("%-58s %8.2f ms  %6.1f GB/s per direction\n", "H2D + D2H concurrently (contiguous)", ms, n * F * 4 / ms / 1e6);
  return 0;
}
