// sparsify.cu — thresholded dense filter output -> CSR on the device, and L1 row normalisation.
//
// Replaces the host-side tail of the wavelet pipeline:
//   wavelet/src/utils.py:98-103   coeffs[coeffs < tol] = 0; nonzero(); csr_matrix(..., float32)
//   SSRG/models/base_scalable/base_model.py:246-251, :265 (per 1000-column block + sparse.hstack)
//   wavelet/src/utils.py:106-112  sklearn.preprocessing.normalize(norm='l1', axis=1)
// The Chebyshev kernel already wrote the thresholded float32 block; here each block is compacted to
// a block-CSR (count, scan, fill with ballot ranks), the blocks are merged row by row in block
// order (= ascending column order, so the rows come out sorted without a sort), and the rows are
// L1-normalised with sklearn's arithmetic (sequential double sum of |v|, then float32(v / sum)).
#include "common.cuh"
#include "scan.cuh"

namespace srg {

// one warp per row: number of non-zeros among the B columns of a dense row
__global__ void __launch_bounds__(256)
dense_row_nnz_kernel(const float *__restrict__ dense, long long ld, long long n, int B, int *__restrict__ counts) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float *d = dense + row * ld;
  int c = 0;
  for (int j = lane; j < B; j += 32) c += (d[j] != 0.0f) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) counts[row] = c;
}

// one warp per row: (column + col0, value) of the non-zeros, in column order
__global__ void __launch_bounds__(256)
dense_to_csr_kernel(const float *__restrict__ dense, long long ld, long long n, int B, int col0,
                    const int *__restrict__ indptr, int *__restrict__ cols, float *__restrict__ vals) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const float *d = dense + row * ld;
  int base = indptr[row];
  for (int j0 = 0; j0 < B; j0 += 32) {
    const int j = j0 + lane;
    const float v = (j < B) ? d[j] : 0.0f;
    const bool keep = v != 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int p = base + __popc(m & lt);
      cols[p] = col0 + j;
      vals[p] = v;
    }
    base += __popc(m);
  }
}

// append block rows to the merged CSR at the per-row cursor, then advance the cursor
__global__ void __launch_bounds__(256)
csr_block_scatter_kernel(long long n, const int *__restrict__ b_indptr, const int *__restrict__ b_cols,
                         const float *__restrict__ b_vals, int *__restrict__ cursor, int *__restrict__ out_cols,
                         float *__restrict__ out_vals) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int s = b_indptr[row], e = b_indptr[row + 1];
  const int dst = cursor[row];
  for (int j = s + lane; j < e; j += 32) {
    out_cols[dst + (j - s)] = b_cols[j];
    out_vals[dst + (j - s)] = b_vals[j];
  }
  __syncwarp();
  if (lane == 0) cursor[row] = dst + (e - s);
}

__global__ void __launch_bounds__(256)
add_counts_kernel(int *__restrict__ total, const int *__restrict__ counts, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) total[i] += counts[i];
}

// sklearn _inplace_csr_row_normalize_l1: sequential DOUBLE sum of |v| per row, then
// v = float32(double(v) / sum) when sum != 0
__global__ void __launch_bounds__(256)
csr_row_normalize_l1_kernel(long long n, const int *__restrict__ indptr, float *__restrict__ vals) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int s = indptr[row], e = indptr[row + 1];
  double sum = 0.0;
  if (lane == 0)
    for (int j = s; j < e; ++j) sum = __dadd_rn(sum, fabs((double)vals[j]));
  sum = __shfl_sync(0xffffffffu, sum, 0);
  if (sum == 0.0) return;
  for (int j = s + lane; j < e; j += 32) vals[j] = __double2float_rn(__ddiv_rn((double)vals[j], sum));
}

}  // namespace srg

using namespace srg;

extern "C" int srg_dense_block_to_csr_f32(const float *dense, int64_t ld, int64_t n, int32_t B, int32_t col0,
                                          int32_t *out_indptr, int32_t *out_cols, float *out_vals,
                                          int64_t capacity, int32_t *row_totals, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && B >= 0 && ld >= B && col0 >= 0 && capacity >= 0, "dense_block_to_csr: bad sizes");
  SRG_REQUIRE(out_indptr, "dense_block_to_csr: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int), s));
    return SRG_OK;
  }
  SRG_REQUIRE(dense || B == 0, "dense_block_to_csr: dense is NULL");
  const bool fill = out_cols != nullptr;
  SRG_REQUIRE(!fill || out_vals, "dense_block_to_csr: out_vals is NULL");
  (void)capacity;  // the caller sized the outputs from a first, count-only call (out_cols == NULL)
  int *scratch = nullptr;
  SRG_CUDA(cudaMallocAsync(&scratch, (size_t)(n + scan_scratch_ints(n)) * sizeof(int), s));
  int *counts = scratch + scan_scratch_ints(n);
  const unsigned wb = (unsigned)ceil_div64(n * 32, 256);
  dense_row_nnz_kernel<<<wb, 256, 0, s>>>(dense, ld, n, B, counts);
  SRG_LAUNCHED();
  rc = exclusive_scan_i32(counts, n, out_indptr, scratch, s);
  if (!rc) {
    if (fill) {
      dense_to_csr_kernel<<<wb, 256, 0, s>>>(dense, ld, n, B, col0, out_indptr, out_cols, out_vals);
      SRG_LAUNCHED();
    }
    if (row_totals) {
      add_counts_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(row_totals, counts, n);
      SRG_LAUNCHED();
    }
  }
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int srg_csr_block_scatter_f32(int64_t n, const int32_t *block_indptr, const int32_t *block_cols,
                                         const float *block_vals, int32_t *cursor, int32_t *out_cols,
                                         float *out_vals, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "csr_block_scatter: negative n");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(block_indptr && cursor && out_cols && out_vals, "csr_block_scatter: NULL pointer");
  csr_block_scatter_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, as_stream(stream)>>>(
      n, block_indptr, block_cols, block_vals, cursor, out_cols, out_vals);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_csr_row_normalize_l1_f32(int64_t n, const int32_t *indptr, float *vals, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "csr_row_normalize_l1: negative n");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(indptr && vals, "csr_row_normalize_l1: NULL pointer");
  csr_row_normalize_l1_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, as_stream(stream)>>>(n, indptr, vals);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && out && (n == 0 || in), "exclusive_scan: bad arguments");
  cudaStream_t s = as_stream(stream);
  int *scratch = nullptr;
  SRG_CUDA(cudaMallocAsync(&scratch, (size_t)scan_scratch_ints(n) * sizeof(int), s));
  rc = exclusive_scan_i32(in, n, out, scratch, s);
  cudaFreeAsync(scratch, s);
  return rc;
}
