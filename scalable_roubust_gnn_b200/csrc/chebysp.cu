// chebysp.cu — heat-kernel wavelets Psi_s = sum_k c_sk T_k(L~) applied to the IDENTITY, evaluated as sparse matrices.
//
// The reference (wavelet/src/utils.py:89-104, SSRG/models/base_scalable/base_model.py:236-265) feeds pygsp's cheby_op
// with dense impulse blocks: N x N (or N x 1000) float64 matrices that are almost entirely zero, because T_k(L) e_j
// is supported on the k-hop neighbourhood of j.  Here the same recurrence (restated in oracle/__init__.py:cheby_op;
// csrc/cheby.cu is the dense-block form)
//     T0 = I;  T1 = (L T0 - a2 T0) / a1;              r_s  = (0.5 c_s0) T0 + c_s1 T1
//     Tk = (2/a1) (L T_{k-1} - a2 T_{k-1}) - T_{k-2};  r_s += c_sk Tk            k = 2..M
//     threshold  r_s[r_s < tol] = 0,  float32,  CSR
// runs on the stored entries only: L T_{k-1} is a sparse product (expand every L_ip * T_pj with key (i, j) in (i, p)
// order, ONE stable key sort, sequential fp64 sum of each segment in ascending p = the order of scipy's csr_matvecs
// chain), and the epilogue looks T_{k-1}, T_{k-2} and r_s up in the (nested) patterns of the previous orders.
// Skipping a structural zero changes no bit of a dense evaluation (x + 0 * y = x; only the sign of an exact zero can
// differ, and zeros do not survive the threshold), so the result is bit-identical to the dense-block path and to the
// oracle.  arxiv shape, M = 3: 76 M products instead of 170 dense N x 1000 blocks.
#include <vector>

#include "common.cuh"
#include "scan.cuh"
#include "sortutil.cuh"

namespace srg {

constexpr int kSpMaxScales = 4;

struct SpCoef {
  double a2, inv_scale;            // a2 = lmax / 2;  k == 1: a1 (division), k >= 2: 2 / a1 (multiplication)
  double c_prev[kSpMaxScales];     // k == 1: 0.5 * c_s0
  double c_cur[kSpMaxScales];      // c_sk
  int n_scales;
};

struct SpRes {
  double *r[kSpMaxScales];
};

// T1 and r_s on the pattern of L (which must hold every diagonal entry: T1_ii = (L_ii - a2) / a1 is non-zero even
// for an isolated node); one warp per row
__global__ void __launch_bounds__(256)
chebysp_first_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ lv,
                     long long n, SpCoef c, double *__restrict__ t1, SpRes res, int *__restrict__ flags) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  bool diag = false;
  for (int p = indptr[i] + lane; p < indptr[i + 1]; p += 32) {
    const int j = indices[p];
    const double x = (j == (int)i) ? 1.0 : 0.0;              // T0 = I
    diag |= (j == (int)i);
    // acc = 0 + L_ij * 1 = L_ij;  w = acc - a2 * x;  T1 = w / a1
    const double w = __dsub_rn(lv[p], __dmul_rn(c.a2, x));
    const double tn = __ddiv_rn(w, c.inv_scale);
    t1[p] = tn;
#pragma unroll
    for (int s = 0; s < kSpMaxScales; ++s)
      if (s < c.n_scales) res.r[s][p] = __dadd_rn(__dmul_rn(c.c_prev[s], x), __dmul_rn(c.c_cur[s], tn));
  }
  if (!__any_sync(0xffffffffu, diag) && lane == 0) atomicOr(flags, SRG_FLAG_BAD_INDEX);
}

__global__ void chebysp_count_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                     const int *__restrict__ b_ptr, long long n, int *__restrict__ cnt,
                                     unsigned long long *__restrict__ total) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  unsigned long long c = 0;
  for (int p = a_ptr[i] + lane; p < a_ptr[i + 1]; p += 32) {
    const int kk = a_idx[p];
    c += (unsigned long long)(b_ptr[kk + 1] - b_ptr[kk]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) {
    cnt[i] = (int)min(c, (unsigned long long)2147483647);
    atomicAdd(total, c);
  }
}

// products of row i in (p ascending, j ascending) order: the stable sort keeps the p order inside equal (i, j)
__global__ void chebysp_expand_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                      const double *__restrict__ a_val, const int *__restrict__ b_ptr,
                                      const int *__restrict__ b_idx, const double *__restrict__ b_val, long long n,
                                      const int *__restrict__ off, uint64_t *__restrict__ keys,
                                      double *__restrict__ vals) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  long long w = off[i];
  for (int p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
    const int kk = a_idx[p];
    const double av = a_val[p];
    const int b0 = b_ptr[kk], b1 = b_ptr[kk + 1];
    for (int t = b0 + lane; t < b1; t += 32) {
      keys[w + (t - b0)] = ((uint64_t)i << 32) | (uint64_t)(unsigned)b_idx[t];
      vals[w + (t - b0)] = __dmul_rn(av, b_val[t]);
    }
    w += b1 - b0;
  }
}

__global__ void chebysp_heads_kernel(const uint64_t *__restrict__ keys, long long m, int *__restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per segment head: (L T)_ij = ((p_1 + p_2) + p_3) + ... in ascending p (separately rounded adds)
__global__ void chebysp_compress_kernel(const uint64_t *__restrict__ keys, const double *__restrict__ vals,
                                        const int *__restrict__ head, const int *__restrict__ seg, long long m,
                                        uint64_t *__restrict__ u_key, int *__restrict__ out_idx,
                                        double *__restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !head[i]) return;
  double s = vals[i];
  for (long long j = i + 1; j < m && !head[j]; ++j) s = __dadd_rn(s, vals[j]);
  const int slot = seg[i];
  u_key[slot] = keys[i];
  out_idx[slot] = (int)(keys[i] & 0xffffffffu);
  out_val[slot] = s;
}

__global__ void chebysp_row_ptr_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ total, long long n,
                                       int *__restrict__ indptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  int lo = 0, hi = *total;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  indptr[r] = lo;
}

__device__ __forceinline__ int sp_find(const int *__restrict__ idx, int lo, int hi, int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = idx[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return -1;
}

// epilogue of order k >= 2 on the pattern of C = L T_{k-1} (in place: c_val becomes T_k):
//   w = C_ij - a2 T_{k-1,ij};  T_k = (2/a1) w - T_{k-2,ij};  r_s,ij = r_s,ij(old pattern, else 0) + c_sk T_k
// T_{k-1} and the old r_s share one pattern (p1), T_{k-2} lives on p2 (k == 2: T0 = I, no arrays)
__global__ void __launch_bounds__(256)
chebysp_step_kernel(const int *__restrict__ c_ptr, const int *__restrict__ c_idx, double *__restrict__ c_val, long long n,
                    const int *__restrict__ p1_ptr, const int *__restrict__ p1_idx, const double *__restrict__ t1,
                    SpRes r_old, const int *__restrict__ p2_ptr, const int *__restrict__ p2_idx,
                    const double *__restrict__ t2, SpCoef c, SpRes r_new) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const int a0 = p1_ptr[i], a1 = p1_ptr[i + 1];
  const int b0 = p2_ptr ? p2_ptr[i] : 0, b1 = p2_ptr ? p2_ptr[i + 1] : 0;
  for (int p = c_ptr[i] + lane; p < c_ptr[i + 1]; p += 32) {
    const int j = c_idx[p];
    const int q1 = sp_find(p1_idx, a0, a1, j);
    const double tc = (q1 >= 0) ? t1[q1] : 0.0;
    double tp;
    if (p2_ptr) {
      const int q2 = sp_find(p2_idx, b0, b1, j);
      tp = (q2 >= 0) ? t2[q2] : 0.0;
    } else {
      tp = (j == (int)i) ? 1.0 : 0.0;
    }
    const double w = __dsub_rn(c_val[p], __dmul_rn(c.a2, tc));
    const double tn = __dsub_rn(__dmul_rn(c.inv_scale, w), tp);
    c_val[p] = tn;
#pragma unroll
    for (int s = 0; s < kSpMaxScales; ++s)
      if (s < c.n_scales) {
        const double ro = (q1 >= 0) ? r_old.r[s][q1] : 0.0;
        r_new.r[s][p] = __dadd_rn(ro, __dmul_rn(c.c_cur[s], tn));
      }
  }
}

// threshold + float32 + compaction (wavelet/src/utils.py:98-103): an entry survives when !(r < tol) and its float32
// value is non-zero (csr_matrix(dense) drops zeros)
__device__ __forceinline__ bool sp_keep(double r, double tol, int use_tol) {
  if (use_tol && r < tol) return false;
  return __double2float_rn(r) != 0.0f;
}
__global__ void __launch_bounds__(256)
chebysp_keep_count_kernel(const int *__restrict__ ptr, const double *__restrict__ r, long long n, double tol, int use_tol,
                          int *__restrict__ cnt) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  int c = 0;
  for (int p = ptr[i] + lane; p < ptr[i + 1]; p += 32) c += sp_keep(r[p], tol, use_tol) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) cnt[i] = c;
}
__global__ void __launch_bounds__(256)
chebysp_keep_emit_kernel(const int *__restrict__ ptr, const int *__restrict__ idx, const double *__restrict__ r,
                         long long n, double tol, int use_tol, const int *__restrict__ out_ptr, int *__restrict__ out_idx,
                         float *__restrict__ out_val) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int base = out_ptr[i];
  for (int p0 = ptr[i]; p0 < ptr[i + 1]; p0 += 32) {
    const int p = p0 + lane;
    const bool valid = p < ptr[i + 1];
    const double v = valid ? r[p] : 0.0;
    const bool keep = valid && sp_keep(v, tol, use_tol);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int q = base + __popc(m & lt);
      out_idx[q] = idx[p];
      out_val[q] = __double2float_rn(v);
    }
    base += __popc(m);
  }
}

// result object behind srg_cheby_sparse_*: the thresholded float32 CSR of every scale, device resident
struct SpResult {
  int device = 0;
  int64_t n = 0;
  int n_scales = 0;
  int32_t *indptr[kSpMaxScales] = {nullptr, nullptr, nullptr, nullptr};
  int32_t *indices[kSpMaxScales] = {nullptr, nullptr, nullptr, nullptr};
  float *vals[kSpMaxScales] = {nullptr, nullptr, nullptr, nullptr};
  int64_t nnz[kSpMaxScales] = {0, 0, 0, 0};
  int64_t products = 0, pattern_nnz = 0;
};

struct SpPattern {
  int *ptr = nullptr, *idx = nullptr;
  double *t = nullptr;
  double *r[kSpMaxScales] = {nullptr, nullptr, nullptr, nullptr};
  int64_t nnz = 0;
  bool owns_structure = true;
};

static void sp_free_pattern(SpPattern &p, int n_scales, cudaStream_t s) {
  if (p.owns_structure) {
    if (p.ptr) cudaFreeAsync(p.ptr, s);
    if (p.idx) cudaFreeAsync(p.idx, s);
  }
  if (p.t) cudaFreeAsync(p.t, s);
  for (int i = 0; i < n_scales; ++i)
    if (p.r[i]) cudaFreeAsync(p.r[i], s);
  p = SpPattern();
}

// C = L T (fp64) into a fresh pattern; host synchronisation inside (the product count sizes the sort buffers)
static int sp_product(const int *l_ptr, const int *l_idx, const double *l_val, int64_t n, const SpPattern &b,
                      SpPattern *out, int64_t *products, cudaStream_t s) {
  int rc;
  StreamScratch scratch(s);
  int *cnt = nullptr;
  unsigned long long *total = nullptr;
  if ((rc = scratch.alloc(&cnt, (size_t)(2 * (n + 1) + scan_scratch_ints(n))))) return rc;
  if ((rc = scratch.alloc(&total, 1))) return rc;
  SRG_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned long long), s));
  int *off = cnt + (n + 1), *scr0 = off + (n + 1);
  const unsigned wb = (unsigned)ceil_div64(n * 32, 256);
  chebysp_count_kernel<<<wb, 256, 0, s>>>(l_ptr, l_idx, b.ptr, n, cnt, total);
  SRG_LAUNCHED();
  unsigned long long h_total = 0;
  SRG_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(h_total), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));
  if (h_total > 2147483647ULL) {
    set_err("cheby_sparse: %llu intermediate products exceed the int32 range of the expand-sort-compress path", h_total);
    return SRG_ERR_RANGE;
  }
  const int64_t m = (int64_t)h_total;
  *products += m;
  if ((rc = exclusive_scan_i32(cnt, n, off, scr0, s))) return rc;
  uint64_t *keys = nullptr;
  double *vals = nullptr;
  int *ints = nullptr;
  if ((rc = scratch.alloc(&keys, (size_t)(2 * std::max<int64_t>(m, 1))))) return rc;
  if ((rc = scratch.alloc(&vals, (size_t)(2 * std::max<int64_t>(m, 1))))) return rc;
  if ((rc = scratch.alloc(&ints, (size_t)(2 * (m + 1) + scan_scratch_ints(m + 1))))) return rc;
  int *head = ints, *seg = head + (m + 1), *scr = seg + (m + 1);
  SRG_CUDA(cudaMallocAsync(&out->ptr, (size_t)(n + 1) * sizeof(int), s));
  if (m == 0) {
    SRG_CUDA(cudaMemsetAsync(out->ptr, 0, (size_t)(n + 1) * sizeof(int), s));
    out->nnz = 0;
    return SRG_OK;
  }
  const unsigned mb = (unsigned)ceil_div64(m, 256);
  chebysp_expand_kernel<<<wb, 256, 0, s>>>(l_ptr, l_idx, l_val, b.ptr, b.idx, b.t, n, off, keys, vals);
  SRG_LAUNCHED();
  if ((rc = sort_pairs<double>(keys, keys + m, vals, vals + m, m, 32 + bits_for(n > 1 ? n : 2), s))) return rc;
  chebysp_heads_kernel<<<mb, 256, 0, s>>>(keys + m, m, head);
  SRG_LAUNCHED();
  if ((rc = exclusive_scan_i32(head, m, seg, scr, s))) return rc;   // seg[m] = unique (i, j) pairs
  int h_nnz = 0;
  SRG_CUDA(cudaMemcpyAsync(&h_nnz, seg + m, sizeof(int), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));
  out->nnz = h_nnz;
  SRG_CUDA(cudaMallocAsync(&out->idx, (size_t)std::max(h_nnz, 1) * sizeof(int), s));
  SRG_CUDA(cudaMallocAsync(&out->t, (size_t)std::max(h_nnz, 1) * sizeof(double), s));
  chebysp_compress_kernel<<<mb, 256, 0, s>>>(keys + m, vals + m, head, seg, m, keys, out->idx, out->t);
  SRG_LAUNCHED();
  chebysp_row_ptr_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(keys, seg + m, n, out->ptr);
  SRG_LAUNCHED();
  return SRG_OK;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_cheby_sparse_run(const int32_t *l_indptr, const int32_t *l_indices, const double *l_vals, int64_t n,
                                    int64_t l_nnz, double lmax, const double *coeffs, int32_t n_scales, int32_t order,
                                    double tol, void **out_handle, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(out_handle != nullptr, "cheby_sparse: out_handle is NULL");
  *out_handle = nullptr;
  SRG_REQUIRE(n >= 1 && l_nnz >= 1 && l_indptr && l_indices && l_vals && coeffs, "cheby_sparse: bad arguments");
  SRG_REQUIRE(n_scales >= 1 && n_scales <= kSpMaxScales, "cheby_sparse: 1..%d scales", kSpMaxScales);
  SRG_REQUIRE(order >= 1, "cheby_sparse: order must be >= 1");
  SRG_REQUIRE(lmax > 0.0, "cheby_sparse: lmax must be positive");
  cudaStream_t s = as_stream(stream);
  const double a1 = lmax / 2.0;
  const bool use_tol = tol == tol;
  const unsigned wb = (unsigned)ceil_div64(n * 32, 256);

  StreamScratch scratch(s);
  int *flags = nullptr;
  if ((rc = scratch.alloc(&flags, 1))) return rc;
  SRG_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), s));

  // order 1 on the pattern of L (structure borrowed from the caller)
  SpPattern cur, prev;   // T_{k-1} (+ r_s), T_{k-2}
  cur.ptr = const_cast<int *>(l_indptr);
  cur.idx = const_cast<int *>(l_indices);
  cur.owns_structure = false;
  cur.nnz = l_nnz;
  SRG_CUDA(cudaMallocAsync(&cur.t, (size_t)l_nnz * sizeof(double), s));
  for (int i = 0; i < n_scales; ++i) SRG_CUDA(cudaMallocAsync(&cur.r[i], (size_t)l_nnz * sizeof(double), s));
  SpCoef c;
  c.a2 = a1;
  c.inv_scale = a1;
  c.n_scales = n_scales;
  for (int i = 0; i < kSpMaxScales; ++i) c.c_prev[i] = c.c_cur[i] = 0.0;
  for (int i = 0; i < n_scales; ++i) {
    c.c_prev[i] = 0.5 * coeffs[(size_t)i * (order + 1)];
    c.c_cur[i] = coeffs[(size_t)i * (order + 1) + 1];
  }
  SpRes res;
  for (int i = 0; i < kSpMaxScales; ++i) res.r[i] = cur.r[i];
  chebysp_first_kernel<<<wb, 256, 0, s>>>(l_indptr, l_indices, l_vals, n, c, cur.t, res, flags);
  SRG_LAUNCHED();
  int h_flags = 0;
  SRG_CUDA(cudaMemcpyAsync(&h_flags, flags, sizeof(int), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));
  if (h_flags) {
    sp_free_pattern(cur, n_scales, s);
    set_err("cheby_sparse: the Laplacian must store every diagonal entry (srg_laplacian_csr keeps explicit zeros out: "
            "isolated nodes are not supported by the sparse path)");
    return SRG_ERR_UNSUPPORTED;
  }

  int64_t products = 0;
  for (int k = 2; k <= order; ++k) {
    SpPattern nxt;
    rc = sp_product(l_indptr, l_indices, l_vals, n, cur, &nxt, &products, s);
    if (!rc)
      for (int i = 0; i < n_scales && !rc; ++i) {
        cudaError_t e = cudaMallocAsync(&nxt.r[i], (size_t)std::max<int64_t>(nxt.nnz, 1) * sizeof(double), s);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
      }
    if (rc) {
      sp_free_pattern(nxt, n_scales, s);
      sp_free_pattern(cur, n_scales, s);
      sp_free_pattern(prev, n_scales, s);
      return rc;
    }
    c.inv_scale = 2.0 / a1;
    for (int i = 0; i < n_scales; ++i) c.c_cur[i] = coeffs[(size_t)i * (order + 1) + k];
    SpRes r_old, r_new;
    for (int i = 0; i < kSpMaxScales; ++i) {
      r_old.r[i] = cur.r[i];
      r_new.r[i] = nxt.r[i];
    }
    if (nxt.nnz > 0) {
      chebysp_step_kernel<<<wb, 256, 0, s>>>(nxt.ptr, nxt.idx, nxt.t, n, cur.ptr, cur.idx, cur.t, r_old,
                                             k == 2 ? nullptr : prev.ptr, prev.idx, prev.t, c, r_new);
      SRG_LAUNCHED();
    }
    // the old r_s are folded into the new ones; T_{k-2} is no longer needed
    for (int i = 0; i < n_scales; ++i) {
      cudaFreeAsync(cur.r[i], s);
      cur.r[i] = nullptr;
    }
    sp_free_pattern(prev, n_scales, s);
    prev = cur;
    cur = nxt;
  }

  // threshold, float32, compaction per scale
  SpResult *out = new SpResult();
  cudaGetDevice(&out->device);
  out->n = n;
  out->n_scales = n_scales;
  out->products = products;
  out->pattern_nnz = cur.nnz;
  int *cnt = nullptr;
  rc = scratch.alloc(&cnt, (size_t)(n + scan_scratch_ints(n)));
  for (int i = 0; i < n_scales && !rc; ++i) {
    int *scr = cnt + n;
    cudaError_t e = cudaMalloc(&out->indptr[i], (size_t)(n + 1) * sizeof(int32_t));
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); break; }
    chebysp_keep_count_kernel<<<wb, 256, 0, s>>>(cur.ptr, cur.r[i], n, tol, use_tol ? 1 : 0, cnt);
    if ((rc = exclusive_scan_i32(cnt, n, out->indptr[i], scr, s))) break;
    int h_nnz = 0;
    cudaMemcpyAsync(&h_nnz, out->indptr[i] + n, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    out->nnz[i] = h_nnz;
    e = cudaMalloc(&out->indices[i], (size_t)std::max(h_nnz, 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&out->vals[i], (size_t)std::max(h_nnz, 1) * sizeof(float));
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); break; }
    chebysp_keep_emit_kernel<<<wb, 256, 0, s>>>(cur.ptr, cur.idx, cur.r[i], n, tol, use_tol ? 1 : 0, out->indptr[i],
                                               out->indices[i], out->vals[i]);
    g_launches.fetch_add(2, std::memory_order_relaxed);
  }
  sp_free_pattern(cur, n_scales, s);
  sp_free_pattern(prev, n_scales, s);
  cudaStreamSynchronize(s);
  if (rc) {
    for (int i = 0; i < kSpMaxScales; ++i) {
      cudaFree(out->indptr[i]);
      cudaFree(out->indices[i]);
      cudaFree(out->vals[i]);
    }
    delete out;
    return rc;
  }
  *out_handle = out;
  return SRG_OK;
}

extern "C" int srg_cheby_sparse_info(void *handle, int32_t scale, int64_t *out_nnz, int64_t *out_products,
                                     int64_t *out_pattern_nnz) {
  SRG_REQUIRE(handle != nullptr, "cheby_sparse_info: NULL handle");
  SpResult *r = static_cast<SpResult *>(handle);
  SRG_REQUIRE(scale >= 0 && scale < r->n_scales, "cheby_sparse_info: bad scale %d", scale);
  if (out_nnz) *out_nnz = r->nnz[scale];
  if (out_products) *out_products = r->products;
  if (out_pattern_nnz) *out_pattern_nnz = r->pattern_nnz;
  return SRG_OK;
}

// copy one scale's CSR into caller-provided device arrays (indptr n + 1, indices / vals >= nnz)
extern "C" int srg_cheby_sparse_fetch(void *handle, int32_t scale, int32_t *indptr, int32_t *indices, float *vals,
                                      void *stream) {
  SRG_REQUIRE(handle != nullptr, "cheby_sparse_fetch: NULL handle");
  SpResult *r = static_cast<SpResult *>(handle);
  SRG_REQUIRE(scale >= 0 && scale < r->n_scales && indptr, "cheby_sparse_fetch: bad arguments");
  cudaStream_t s = as_stream(stream);
  SRG_CUDA(cudaMemcpyAsync(indptr, r->indptr[scale], (size_t)(r->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (r->nnz[scale] > 0) {
    SRG_REQUIRE(indices && vals, "cheby_sparse_fetch: NULL output arrays");
    SRG_CUDA(cudaMemcpyAsync(indices, r->indices[scale], (size_t)r->nnz[scale] * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    SRG_CUDA(cudaMemcpyAsync(vals, r->vals[scale], (size_t)r->nnz[scale] * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return SRG_OK;
}

extern "C" int srg_cheby_sparse_free(void *handle) {
  if (!handle) return SRG_OK;
  SpResult *r = static_cast<SpResult *>(handle);
  {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(r->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kSpMaxScales; ++i) {
      cudaFree(r->indptr[i]);
      cudaFree(r->indices[i]);
      cudaFree(r->vals[i]);
    }
    if (prev >= 0) cudaSetDevice(prev);
  }
  delete r;
  return SRG_OK;
}

// ---- lambda_max of the Laplacian on the device (pygsp Graph.estimate_lmax) ----------------------------------------------
// The reference obtains lmax from ARPACK: 1.01 * eigsh(L, k=1, tol=5e-3, ncv=min(N,10)) (wavelet/src/utils.py:83,
// SSRG/models/base_scalable/base_model.py:184) with a random start vector, so its value is only defined to the 5e-3
// relative tolerance of that call.  Here: Lanczos with full re-orthogonalisation in fp64 (SpMV + dots + updates are
// kernels below; the small tridiagonal eigenvalue is a Sturm bisection on the host), stopped when the largest Ritz
// value has moved by less than tol / 10 over two steps.  Deterministic (hashed start vector, ordered reductions).
namespace srg {

__global__ void __launch_bounds__(256)
spmv_f64_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ vals,
                long long n, const double *__restrict__ x, double *__restrict__ y) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int p = indptr[i] + lane; p < indptr[i + 1]; p += 32) acc = __fma_rn(vals[p], x[indices[p]], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[i] = acc;
}

// out[k] = x . V[k]  for k < count (one block per k, fixed-order tree: deterministic)
__global__ void __launch_bounds__(1024)
dots_f64_kernel(const double *__restrict__ x, const double *__restrict__ V, long long n, double *__restrict__ out) {
  __shared__ double sh[1024];
  const double *v = V + (long long)blockIdx.x * n;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc = __fma_rn(x[i], v[i], acc);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// x -= sum_k c[k] V[k]
__global__ void __launch_bounds__(256)
project_out_kernel(double *__restrict__ x, const double *__restrict__ V, long long n, const double *__restrict__ c, int count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i];
  for (int k = 0; k < count; ++k) v = __fma_rn(-c[k], V[(long long)k * n + i], v);
  x[i] = v;
}

__global__ void __launch_bounds__(256)
scale_into_kernel(const double *__restrict__ x, double s, long long n, double *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * s;
}

__global__ void __launch_bounds__(256) hashed_start_kernel(double *__restrict__ v, long long n, unsigned seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 0x9e3779b1u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  v[i] = 0.5 + (double)x * (1.0 / 4294967296.0);      // positive entries: never orthogonal to the top eigenvector's sign pattern by accident
}

// largest eigenvalue of the symmetric tridiagonal (alpha[0..m), beta[0..m-1)) by Sturm bisection
static double tridiag_lambda_max(const std::vector<double> &a, const std::vector<double> &b) {
  const int m = (int)a.size();
  double lo = a[0], hi = a[0];
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i + 1 < m ? fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
  }
  for (int it = 0; it < 200 && hi - lo > 1e-15 * std::max(fabs(hi), fabs(lo)); ++it) {
    const double x = 0.5 * (lo + hi);
    // number of eigenvalues < x = number of negative pivots of T - x I
    int neg = 0;
    double d = 1.0;
    for (int i = 0; i < m; ++i) {
      const double bb = (i > 0) ? b[i - 1] * b[i - 1] : 0.0;
      d = (a[i] - x) - (i > 0 ? bb / (d == 0.0 ? 1e-300 : d) : 0.0);
      if (d < 0.0) ++neg;
    }
    if (neg >= m) hi = x; else lo = x;   // all eigenvalues below x: move down
  }
  return 0.5 * (lo + hi);
}

}  // namespace srg

extern "C" int srg_lanczos_lambda_max_f64(const int32_t *indptr, const int32_t *indices, const double *vals, int64_t n,
                                          double tol, int32_t max_steps, double *out_lambda, int32_t *out_steps,
                                          void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 1 && indptr && indices && vals && out_lambda, "lanczos: bad arguments");
  SRG_REQUIRE(tol > 0.0 && max_steps >= 1, "lanczos: tol must be positive and max_steps >= 1");
  cudaStream_t s = as_stream(stream);
  const int m_max = (int)std::min<int64_t>(std::min<int64_t>(max_steps, n), 256);
  StreamScratch scratch(s);
  double *V = nullptr, *w = nullptr, *c = nullptr;
  if ((rc = scratch.alloc(&V, (size_t)(m_max + 1) * n))) return rc;
  if ((rc = scratch.alloc(&w, (size_t)n))) return rc;
  if ((rc = scratch.alloc(&c, (size_t)(m_max + 2)))) return rc;
  const unsigned nb = (unsigned)ceil_div64(n, 256), wb = (unsigned)ceil_div64(n * 32, 256);
  std::vector<double> alpha, beta, hc(m_max + 2);
  // v_0 = start / |start|
  hashed_start_kernel<<<nb, 256, 0, s>>>(w, n, 2023u);
  SRG_LAUNCHED();
  dots_f64_kernel<<<1, 1024, 0, s>>>(w, w, n, c);
  SRG_LAUNCHED();
  SRG_CUDA(cudaMemcpyAsync(hc.data(), c, sizeof(double), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));
  scale_into_kernel<<<nb, 256, 0, s>>>(w, 1.0 / sqrt(hc[0]), n, V);
  SRG_LAUNCHED();
  double theta = 0.0, theta_1 = 0.0, theta_2 = 0.0;
  int j = 0;
  for (; j < m_max; ++j) {
    spmv_f64_kernel<<<wb, 256, 0, s>>>(indptr, indices, vals, n, V + (size_t)j * n, w);
    SRG_LAUNCHED();
    // full re-orthogonalisation against v_0..v_j (twice is enough); the coefficient on v_j is alpha_j
    double a_j = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      dots_f64_kernel<<<j + 1, 1024, 0, s>>>(w, V, n, c);
      SRG_LAUNCHED();
      project_out_kernel<<<nb, 256, 0, s>>>(w, V, n, c, j + 1);
      SRG_LAUNCHED();
      SRG_CUDA(cudaMemcpyAsync(hc.data(), c + j, sizeof(double), cudaMemcpyDeviceToHost, s));
      SRG_CUDA(cudaStreamSynchronize(s));
      a_j += hc[0];
    }
    alpha.push_back(a_j);
    dots_f64_kernel<<<1, 1024, 0, s>>>(w, w, n, c);
    SRG_LAUNCHED();
    SRG_CUDA(cudaMemcpyAsync(hc.data(), c, sizeof(double), cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));
    const double b_j = sqrt(std::max(hc[0], 0.0));
    theta_2 = theta_1;
    theta_1 = theta;
    theta = tridiag_lambda_max(alpha, beta);
    const bool converged = j >= 4 && fabs(theta - theta_2) <= 0.1 * tol * fabs(theta);
    if (converged || b_j <= 1e-14 * std::max(1.0, fabs(theta)) || j + 1 == m_max) {
      ++j;
      break;
    }
    beta.push_back(b_j);
    scale_into_kernel<<<nb, 256, 0, s>>>(w, 1.0 / b_j, n, V + (size_t)(j + 1) * n);
    SRG_LAUNCHED();
  }
  *out_lambda = theta;
  if (out_steps) *out_steps = j;
  return SRG_OK;
}
