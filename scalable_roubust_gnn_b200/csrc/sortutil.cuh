// sortutil.cuh — the one CUDA-toolkit primitive the library calls: cub::DeviceRadixSort (stable LSD radix
// sort) for (row, col) key sorts of the set-up stages (coo.cu, magnetic.cu).  Never on the per-hop path.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace srg {

static int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && (1LL << b) < n) ++b;
  return b;
}

// stable sort of 64-bit keys (optionally carrying a payload) on `s`
template <typename V>
static int sort_pairs(uint64_t *keys_in, uint64_t *keys_out, V *vals_in, V *vals_out, int64_t m,
                      int end_bit, cudaStream_t s) {
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, m, 0, end_bit, s);
  void *tmp = nullptr;
  SRG_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, m, 0, end_bit, s);
  g_launches.fetch_add(1);
  cudaFreeAsync(tmp, s);
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
  return SRG_OK;
}
static int sort_keys(uint64_t *keys_in, uint64_t *keys_out, int64_t m, int end_bit, cudaStream_t s) {
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys_in, keys_out, m, 0, end_bit, s);
  void *tmp = nullptr;
  SRG_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s));
  cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_in, keys_out, m, 0, end_bit, s);
  g_launches.fetch_add(1);
  cudaFreeAsync(tmp, s);
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortKeys", __FILE__, __LINE__);
  return SRG_OK;
}

}  // namespace srg
