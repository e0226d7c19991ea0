// spgemm.cu — sparse x sparse product C = A B in CSR (float32) by expand / sort / compress, and the
// symmetric degree normalisation of a float32 CSR.
//
// Reference call sites:
//   * adj_to_un_in_out_dir_symmetric_norm, SSRG/operators/utils.py:195-260: in_L = P^T P, out_L = P P^T are formed
//     as DENSE N x N float32 products (:216-219) and turned back into sparse matrices with torch.nonzero
//     (:223-227, :243-247); then row sums, pow, D^(r-1) L D^(-r) in float32 (:230-237, :250-257).
//   * torch_sparse.spspmm(Psi, Psi^-1), SSRG/models/base_scalable/base_model.py:208-214 and
//     wavelet/src/gwnn_layer.py:59-75.
// Here the product never leaves sparse form: every a_ik * b_kj is expanded with key (i, j) in (i, k) order, one
// stable key sort (cub::DeviceRadixSort) brings equal (i, j) together in ascending k, a segmented sequential
// fp32 sum compresses them.  Deterministic; the summation order (ascending k) is one of the orders a dense
// sgemm may use, so results agree with the reference to fp32 rounding, not bit for bit.
#include "common.cuh"
#include "scan.cuh"
#include "sortutil.cuh"

namespace srg {

// products per row of A: sum over its entries of the length of the matching row of B
__global__ void spgemm_count_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                    const int *__restrict__ b_ptr, long long n_a, long long k_dim,
                                    int *__restrict__ cnt, unsigned long long *__restrict__ total,
                                    int *__restrict__ flags) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_a) return;
  const int lane = threadIdx.x & 31;
  unsigned long long c = 0;
  for (int p = a_ptr[i] + lane; p < a_ptr[i + 1]; p += 32) {
    const int kk = a_idx[p];
    if (kk < 0 || kk >= k_dim) {
      atomicOr(flags, SRG_FLAG_BAD_INDEX);
      continue;
    }
    c += (unsigned long long)(b_ptr[kk + 1] - b_ptr[kk]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) {
    cnt[i] = (int)min(c, (unsigned long long)2147483647);
    atomicAdd(total, c);
  }
}

// one warp per row of A; its entries in stored order, the row of B spread over the lanes: products of one
// (i, k) pair are contiguous and pairs follow in k order, which the stable sort preserves inside equal keys
__global__ void spgemm_expand_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                     const float *__restrict__ a_val, const int *__restrict__ b_ptr,
                                     const int *__restrict__ b_idx, const float *__restrict__ b_val, long long n_a,
                                     long long k_dim, const int *__restrict__ off, uint64_t *__restrict__ keys,
                                     float *__restrict__ vals) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_a) return;
  const int lane = threadIdx.x & 31;
  long long w = off[i];
  for (int p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
    const int kk = a_idx[p];
    if (kk < 0 || kk >= k_dim) continue;
    const float av = a_val ? a_val[p] : 1.0f;
    const int b0 = b_ptr[kk], b1 = b_ptr[kk + 1];
    for (int t = b0 + lane; t < b1; t += 32) {
      keys[w + (t - b0)] = ((uint64_t)i << 32) | (uint64_t)(unsigned)b_idx[t];
      vals[w + (t - b0)] = __fmul_rn(av, b_val ? b_val[t] : 1.0f);
    }
    w += b1 - b0;
  }
}

__global__ void spgemm_heads_kernel(const uint64_t *__restrict__ keys, long long m, int *__restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per segment head: c_ij = ((p_1 + p_2) + p_3) + ...  in ascending k; exact zeros optionally dropped
// later by the caller (torch.nonzero in the reference) through the `keep` mask
__global__ void spgemm_compress_kernel(const uint64_t *__restrict__ keys, const float *__restrict__ vals,
                                       const int *__restrict__ head, const int *__restrict__ seg, long long m,
                                       uint64_t *__restrict__ u_key, float *__restrict__ u_val, int *__restrict__ keep,
                                       int drop_zeros) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !head[i]) return;
  float s = vals[i];
  for (long long j = i + 1; j < m && !head[j]; ++j) s = __fadd_rn(s, vals[j]);
  if (s != s) s = 0.f;   // in_L[torch.isnan(in_L)] = 0  (utils.py:221)
  const int slot = seg[i];
  u_key[slot] = keys[i];
  u_val[slot] = s;
  keep[slot] = (drop_zeros && s == 0.f) ? 0 : 1;
}

__global__ void spgemm_emit_kernel(const uint64_t *__restrict__ u_key, const float *__restrict__ u_val,
                                   const int *__restrict__ keep, const int *__restrict__ kseg,
                                   const int *__restrict__ total, uint64_t *__restrict__ f_key,
                                   int *__restrict__ out_idx, float *__restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *total || !keep[i]) return;
  const int slot = kseg[i];
  f_key[slot] = u_key[i];
  out_idx[slot] = (int)(u_key[i] & 0xffffffffu);
  out_val[slot] = u_val[i];
}

__global__ void spgemm_row_lower_bound_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ total,
                                              long long n, int *__restrict__ indptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  int lo = 0, hi = *total;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  indptr[r] = lo;
}

// ---- D^(r-1) L D^(-r) of a float32 CSR with float32 arithmetic (utils.py:229-237, :249-257) ----------------
// deg = scatter_add over the row in stored (column) order; pow in float32 with inf -> 0
__global__ void f32_row_sums_kernel(const int *__restrict__ indptr, const float *__restrict__ vals, long long n,
                                    float r, float *__restrict__ dl, float *__restrict__ dr,
                                    float *__restrict__ deg_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = 0.f;
  for (int p = indptr[i]; p < indptr[i + 1]; ++p) d = __fadd_rn(d, vals ? vals[p] : 1.0f);
  float l = powf(d, __fsub_rn(r, 1.0f)), q = powf(d, -r);
  if (isinf(l)) l = 0.f;
  if (isinf(q)) q = 0.f;
  dl[i] = l;
  dr[i] = q;
  if (deg_out) deg_out[i] = d;
}

__global__ void f32_scale_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                 const float *__restrict__ vals, long long n, const float *__restrict__ dl,
                                 const float *__restrict__ dr, float *__restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const float l = dl[i];
  for (int p = indptr[i] + lane; p < indptr[i + 1]; p += 32)
    out[p] = __fmul_rn(__fmul_rn(l, vals ? vals[p] : 1.0f), dr[indices[p]]);
}

// row i of the output = row i of the input followed by the entry (i, i): add_self_loops appends one loop per
// node whatever the row already holds (utils.py:199-201); duplicates are summed by the canonicalisation after
__global__ void csr_append_diag_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, long long n,
                                       int *__restrict__ out_indptr, int *__restrict__ out_indices) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i > n) return;
  const int lane = threadIdx.x & 31;
  if (i == n) {
    if (lane == 0) out_indptr[n] = indptr[n] + (int)n;
    return;
  }
  const int s0 = indptr[i], e0 = indptr[i + 1], d0 = s0 + (int)i;
  if (lane == 0) out_indptr[i] = d0;
  for (int p = s0 + lane; p < e0; p += 32) out_indices[d0 + (p - s0)] = indices[p];
  if (lane == 0) out_indices[d0 + (e0 - s0)] = (int)i;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_csr_append_diagonal(const int32_t *indptr, const int32_t *indices, int64_t n, int32_t *out_indptr,
                                       int32_t *out_indices, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && indptr && out_indptr, "csr_append_diagonal: bad arguments");
  SRG_REQUIRE(n == 0 || out_indices, "csr_append_diagonal: out_indices is NULL");
  csr_append_diag_kernel<<<(unsigned)ceil_div64((n + 1) * 32, 256), 256, 0, as_stream(stream)>>>(indptr, indices, n,
                                                                                              out_indptr, out_indices);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_spgemm_csr_f32(const int32_t *a_indptr, const int32_t *a_indices, const float *a_vals,
                                  int64_t n_a, int64_t k_dim, const int32_t *b_indptr, const int32_t *b_indices,
                                  const float *b_vals, int64_t n_b, int32_t drop_zeros, int32_t *out_indptr,
                                  int32_t *out_indices, float *out_vals, int64_t cap, int64_t *out_nnz,
                                  int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_a >= 0 && k_dim >= 0 && n_b >= 0 && cap >= 0, "spgemm: negative size");
  SRG_REQUIRE(a_indptr && b_indptr && out_indptr && out_nnz && out_flags, "spgemm: NULL pointer");
  SRG_REQUIRE(n_a < 2147483647LL && n_b <= 2147483647LL, "spgemm: exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  *out_nnz = 0;
  if (n_a == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int32_t), s));
    return SRG_OK;
  }
  int *cnt = nullptr;   // cnt (n_a + 1) | off (n_a + 1) | scan scratch
  unsigned long long *total = nullptr;
  SRG_CUDA(cudaMallocAsync(&cnt, (size_t)(2 * (n_a + 1) + scan_scratch_ints(n_a)) * sizeof(int), s));
  SRG_CUDA(cudaMallocAsync(&total, sizeof(unsigned long long), s));
  SRG_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned long long), s));
  int *off = cnt + (n_a + 1), *scr0 = off + (n_a + 1);
  const unsigned wb = (unsigned)ceil_div64(n_a * 32, 256);
  spgemm_count_kernel<<<wb, 256, 0, s>>>(a_indptr, a_indices, b_indptr, n_a, k_dim, cnt, total, out_flags);
  SRG_LAUNCHED();
  unsigned long long h_total = 0;
  SRG_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(h_total), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));   // set-up path: the product count sizes the sort buffers
  cudaFreeAsync(total, s);
  if (h_total > 2147483647ULL) {
    cudaFreeAsync(cnt, s);
    set_err("spgemm: %llu intermediate products exceed the int32 range of the expand-sort-compress path", h_total);
    return SRG_ERR_RANGE;
  }
  const int64_t m = (int64_t)h_total;
  rc = exclusive_scan_i32(cnt, n_a, off, scr0, s);
  if (rc || m == 0) {
    if (!rc) SRG_CUDA(cudaMemsetAsync(out_indptr, 0, (size_t)(n_a + 1) * sizeof(int32_t), s));
    cudaFreeAsync(cnt, s);
    return rc;
  }
  uint64_t *keys = nullptr;   // 2m sort buffers; the first half is reused for the unique keys, then the final keys
  float *vals = nullptr;      // 2m
  int *ints = nullptr;        // head (m+1) | seg (m+1) | keep (m+1) | kseg (m+1) | scratch
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(2 * m) * sizeof(uint64_t), s));
  SRG_CUDA(cudaMallocAsync(&vals, (size_t)(2 * m) * sizeof(float), s));
  SRG_CUDA(cudaMallocAsync(&ints, (size_t)(4 * (m + 1) + scan_scratch_ints(m + 1)) * sizeof(int), s));
  int *head = ints, *seg = head + (m + 1), *keep = seg + (m + 1), *kseg = keep + (m + 1), *scr = kseg + (m + 1);
  const unsigned mb = (unsigned)ceil_div64(m, 256);
  spgemm_expand_kernel<<<wb, 256, 0, s>>>(a_indptr, a_indices, a_vals, b_indptr, b_indices, b_vals, n_a, k_dim, off, keys,
                                          vals);
  SRG_LAUNCHED();
  rc = sort_pairs<float>(keys, keys + m, vals, vals + m, m, 32 + bits_for(n_a > 1 ? n_a : 2), s);
  if (!rc) {
    spgemm_heads_kernel<<<mb, 256, 0, s>>>(keys + m, m, head);
    SRG_LAUNCHED();
    rc = exclusive_scan_i32(head, m, seg, scr, s);   // seg[m] = unique (i, j) pairs
  }
  if (!rc) {
    SRG_CUDA(cudaMemsetAsync(keep, 0, (size_t)(m + 1) * sizeof(int), s));
    spgemm_compress_kernel<<<mb, 256, 0, s>>>(keys + m, vals + m, head, seg, m, keys, vals, keep, drop_zeros);
    SRG_LAUNCHED();
    rc = exclusive_scan_i32(keep, m, kseg, scr, s);  // kseg[m] = entries kept
  }
  int h_nnz = 0;
  if (!rc) {
    SRG_CUDA(cudaMemcpyAsync(&h_nnz, kseg + m, sizeof(int), cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));
    if (h_nnz > cap) {
      set_err("spgemm: the product holds %d entries, the output capacity is %lld", h_nnz, (long long)cap);
      rc = SRG_ERR_RANGE;
    }
  }
  if (!rc) {
    SRG_REQUIRE(h_nnz == 0 || (out_indices && out_vals), "spgemm: NULL output arrays");
    uint64_t *f_key = keys + m;   // the sorted input keys are no longer needed
    if (h_nnz > 0) {
      spgemm_emit_kernel<<<mb, 256, 0, s>>>(keys, vals, keep, kseg, seg + m, f_key, out_indices, out_vals);
      SRG_LAUNCHED();
    }
    spgemm_row_lower_bound_kernel<<<(unsigned)ceil_div64(n_a + 1, 256), 256, 0, s>>>(f_key, kseg + m, n_a, out_indptr);
    SRG_LAUNCHED();
    SRG_CUDA(cudaStreamSynchronize(s));
    *out_nnz = h_nnz;
  }
  cudaFreeAsync(ints, s);
  cudaFreeAsync(vals, s);
  cudaFreeAsync(keys, s);
  cudaFreeAsync(cnt, s);
  return rc;
}

extern "C" int srg_csr_sym_scale_f32(const int32_t *indptr, const int32_t *indices, const float *vals, int64_t n,
                                     float r, float *out_vals, float *out_degree, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "csr_sym_scale: negative size");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && out_vals, "csr_sym_scale: NULL pointer");
  cudaStream_t s = as_stream(stream);
  float *tabs = nullptr;
  SRG_CUDA(cudaMallocAsync(&tabs, (size_t)(2 * n) * sizeof(float), s));
  f32_row_sums_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(indptr, vals, n, r, tabs, tabs + n, out_degree);
  SRG_LAUNCHED();
  f32_scale_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, s>>>(indptr, indices, vals, n, tabs, tabs + n, out_vals);
  SRG_LAUNCHED();
  cudaFreeAsync(tabs, s);
  return SRG_OK;
}
