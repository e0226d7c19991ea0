// scan.cuh — int32 exclusive prefix sum (row lengths -> indptr), three small kernels.
// out has n+1 entries, out[n] = total.  Scratch: ceil(n / kScanTile) + 1 ints.
#pragma once
#include "common.cuh"

namespace srg {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int warp_sums[THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < THREADS / 32) ? warp_sums[lane] : 0;
    int wi = warp_incl_scan(w);
    if (lane < THREADS / 32) warp_sums[lane] = wi - w;
    if (lane == THREADS / 32 - 1) block_total = wi;
  }
  __syncthreads();
  const int excl = incl - v + warp_sums[wid];
  *total = block_total;
  __syncthreads();
  return excl;
}

static __global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const int *__restrict__ in, long long n, int *__restrict__ tile_sums) {
  const long long base = (long long)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long idx = base + (long long)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int total;
  (void)block_excl_scan<kScanThreads>(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of tile_sums[0..nt), writes grand total to tile_sums[nt]
static __global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(int *tile_sums, int nt) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nt; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = (idx < nt) ? tile_sums[idx] : 0;
    int total;
    const int excl = block_excl_scan<1024>(v, &total);
    const int carry = carry_s;
    if (idx < nt) tile_sums[idx] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[nt] = carry_s;
}

static __global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int *__restrict__ in, long long n, const int *__restrict__ tile_sums,
                  int nt, int *__restrict__ out) {
  // thread t owns kScanItems consecutive elements so the tile is scanned in order
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long idx = base + i;
    v[i] = (idx < n) ? in[idx] : 0;
    s += v[i];
  }
  int total;
  int run = block_excl_scan<kScanThreads>(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const long long idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_sums[nt];
}

// scratch must hold scan_scratch_ints(n) ints
static inline int64_t scan_scratch_ints(int64_t n) { return ceil_div64(n, kScanTile) + 1; }

static inline int exclusive_scan_i32(const int *in, int64_t n, int *out, int *scratch,
                                     cudaStream_t s) {
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out, 0, sizeof(int), s));
    return SRG_OK;
  }
  const int64_t nt = ceil_div64(n, kScanTile);
  SRG_REQUIRE(nt <= 2147483647LL, "scan: too many tiles");
  scan_tile_sums_kernel<<<(unsigned)nt, kScanThreads, 0, s>>>(in, n, scratch);
  SRG_LAUNCHED();
  scan_tile_offsets_kernel<<<1, 1024, 0, s>>>(scratch, (int)nt);
  SRG_LAUNCHED();
  scan_apply_kernel<<<(unsigned)nt, kScanThreads, 0, s>>>(in, n, scratch, (int)nt, out);
  SRG_LAUNCHED();
  return SRG_OK;
}

}  // namespace srg
