// norm.cu — adjacency normalisation over CSR on the device.
//
// Replaces SSRG/operators/utils.py:81-93 (adj_to_symmetric_norm) as it is driven by
// SSRG/operators/graph_operator/symmetrical_simgraph_laplacian_operator.py:12-15 and
// .../symmetrical_simgraph_ppr_operator.py:13-21:
//
//     A~ = A + I                      (csr_plus_csr: union pattern, sums, exact zeros dropped)
//     d  = A~.sum(1)                  (fp64; numpy add.reduceat order: a[0] + pairwise(a[1:]))
//     dl = d^(r-1), dr = d^(-r)       (inf -> 0)
//     R  = (A~ * diag(dl))^T * diag(dr)   =>  R[a,b] = (A~[b,a] * dl[a]) * dr[b]
//     PPR: (1-alpha) * R + alpha * I
//
// Integer outputs (row pointer, column indices) are bit-exact; values are fp64 products in the
// reference's multiply order.  The only non-reproducible step of the reference is np.power,
// whose last bit depends on the host's libm / SVML build; here the common exponents
// (0, +-0.5, +-1) are correctly rounded and the rest use CUDA pow (<= 2 ulp).
#include "common.cuh"
#include "scan.cuh"

namespace srg {

template <int DT> struct ValLoad;
template <> struct ValLoad<SRG_VAL_ONES> {
  __device__ static __forceinline__ double at(const void *, long long) { return 1.0; }
};
template <> struct ValLoad<SRG_VAL_F32> {
  __device__ static __forceinline__ double at(const void *p, long long j) {
    return (double)static_cast<const float *>(p)[j];
  }
};
template <> struct ValLoad<SRG_VAL_F64> {
  __device__ static __forceinline__ double at(const void *p, long long j) {
    return static_cast<const double *>(p)[j];
  }
};

// ---- stage 1: row lengths of A~ --------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256)
selfloop_rowlen_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                       const void *__restrict__ data, long long n, int *__restrict__ rowlen,
                       int *__restrict__ flags) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int s = indptr[a], e = indptr[a + 1];
  int cnt = 0, prev = -1, fl = 0;
  double diag = 0.0;
  for (int j = s; j < e; ++j) {
    const int b = indices[j];
    if (b <= prev) fl |= SRG_FLAG_UNSORTED;
    if (b < 0 || b >= n) fl |= SRG_FLAG_BAD_INDEX;
    prev = b;
    const double v = ValLoad<DT>::at(data, j);
    if (b == a)
      diag = v;
    else if (v != 0.0)
      ++cnt;
  }
  if (__dadd_rn(diag, 1.0) != 0.0) ++cnt;
  rowlen[a] = cnt;
  if (fl) atomicOr(flags, fl);
}

// ---- numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum) ------
__device__ double np_pairwise_sum(const double *a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  } else if (n <= 128) {
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
      r0 = __dadd_rn(r0, a[i + 0]);
      r1 = __dadd_rn(r1, a[i + 1]);
      r2 = __dadd_rn(r2, a[i + 2]);
      r3 = __dadd_rn(r3, a[i + 3]);
      r4 = __dadd_rn(r4, a[i + 4]);
      r5 = __dadd_rn(r5, a[i + 5]);
      r6 = __dadd_rn(r6, a[i + 6]);
      r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
  }
}

// x^e with the exponents that occur for r in {0, 0.5, 1} correctly rounded
__device__ double pow_tab(double x, double e) {
  double y;
  if (e == 0.0) {
    y = 1.0;
  } else if (e == 1.0) {
    y = x;
  } else if (e == -1.0) {
    y = __ddiv_rn(1.0, x);
  } else if (e == 0.5) {
    y = __dsqrt_rn(x);
  } else if (e == -0.5) {
    // 1/sqrt(x): two-rounding estimate, then one residual-corrected Newton step
    const double y0 = __ddiv_rn(1.0, __dsqrt_rn(x));
    if (isfinite(y0) && y0 > 0.0) {
      const double t = __dmul_rn(x, y0);
      const double terr = __fma_rn(x, y0, -t);
      const double res = __dsub_rn(__fma_rn(-t, y0, 1.0), __dmul_rn(terr, y0));
      y = __fma_rn(__dmul_rn(0.5, y0), res, y0);
    } else {
      y = y0;
    }
  } else {
    y = pow(x, e);
  }
  if (isinf(y)) y = 0.0;  // r_inv_sqrt[np.isinf(...)] = 0  (utils.py:85,89)
  return y;
}

// ---- stage 2a: write A~ (indices, values), degree and the two power tables ---------------------
template <int DT>
__global__ void __launch_bounds__(256)
selfloop_fill_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                     const void *__restrict__ data, long long n, const int *__restrict__ out_indptr,
                     int *__restrict__ out_indices, double *__restrict__ at_val,
                     double *__restrict__ degree, double *__restrict__ dl, double *__restrict__ dr,
                     double e_left, double e_right) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int s = indptr[a], e = indptr[a + 1];
  const int p0 = out_indptr[a];
  int p = p0;
  bool placed = false;
  double diag_add = 1.0;  // value contributed by I
  // pass 1: find the diagonal value of A (rows are sorted, so this is a short scan)
  double diag_a = 0.0;
  for (int j = s; j < e; ++j)
    if (indices[j] == (int)a) diag_a = ValLoad<DT>::at(data, j);
  const double diag = __dadd_rn(diag_a, diag_add);
  for (int j = s; j < e; ++j) {
    const int b = indices[j];
    if (b == (int)a) continue;
    if (!placed && b > (int)a) {
      placed = true;
      if (diag != 0.0) {
        out_indices[p] = (int)a;
        if (DT != SRG_VAL_ONES) at_val[p] = diag;
        ++p;
      }
    }
    const double v = ValLoad<DT>::at(data, j);
    if (v != 0.0) {
      out_indices[p] = b;
      if (DT != SRG_VAL_ONES) at_val[p] = v;
      ++p;
    }
  }
  if (!placed && diag != 0.0) {
    out_indices[p] = (int)a;
    if (DT != SRG_VAL_ONES) at_val[p] = diag;
    ++p;
  }
  const int len = p - p0;
  double d;
  if (DT == SRG_VAL_ONES) {
    // all off-diagonal entries are 1.0, the diagonal is 1.0 or 2.0: exact in any order
    d = (double)(len - 1) + diag;  // diag is 1.0 or 2.0, never dropped
  } else if (len == 0) {
    d = 0.0;
  } else {
    d = __dadd_rn(at_val[p0], np_pairwise_sum(at_val + p0 + 1, len - 1));
    if (len == 1) d = at_val[p0];
  }
  degree[a] = d;
  dl[a] = pow_tab(d, e_left);
  dr[a] = pow_tab(d, e_right);
}

// ---- stage 2b: R[a,b] = (A~[b,a] * dl[a]) * dr[b], symmetric-pattern path ---------------------
// one warp per row; each lane owns entries p = p0+lane, p0+lane+32, ...
template <int DT>
__global__ void __launch_bounds__(256)
sym_norm_values_kernel(long long n, const int *__restrict__ out_indptr,
                       const int *__restrict__ out_indices, const double *__restrict__ at_val,
                       const double *__restrict__ degree, const double *__restrict__ dl,
                       const double *__restrict__ dr, double one_minus_alpha, double alpha,
                       int use_ppr,
                       double *__restrict__ val64, float *__restrict__ val32,
                       int *__restrict__ flags) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= n) return;
  const int lane = threadIdx.x & 31;
  const int p0 = out_indptr[a], p1 = out_indptr[a + 1];
  const double dla = dl[a];
  int fl = 0;
  for (int p = p0 + lane; p < p1; p += 32) {
    const int b = out_indices[p];
    double vt;  // A~[b,a]
    if (b == (int)a) {
      vt = (DT == SRG_VAL_ONES) ? 0.0 : at_val[p];
    } else {
      // binary search for column a in row b
      int lo = out_indptr[b], hi = out_indptr[b + 1];
      int q = -1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int c = out_indices[mid];
        if (c == (int)a) {
          q = mid;
          break;
        }
        if (c < (int)a)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (q < 0) {
        fl |= SRG_FLAG_ASYMMETRIC;
        vt = 0.0;
      } else {
        vt = (DT == SRG_VAL_ONES) ? 1.0 : at_val[q];
      }
    }
    if (DT == SRG_VAL_ONES && b == (int)a) {
      // diagonal of an unweighted graph: 1.0 from I, 2.0 when A already had the loop.
      // degree = (#off-diagonal) + diag  =>  diag = degree - (len - 1), exact small integers.
      vt = degree[a] - (double)(p1 - p0 - 1);
    }
    double v = __dmul_rn(__dmul_rn(vt, dla), dr[b]);
    if (use_ppr) {
      v = __dmul_rn(one_minus_alpha, v);
      if (b == (int)a) v = __dadd_rn(v, alpha);
    }
    if (v == 0.0) fl |= SRG_FLAG_ZERO_PRODUCT;
    if (val64) val64[p] = v;
    if (val32) val32[p] = __double2float_rn(v);
  }
  if (fl) atomicOr(flags, fl);
}

// used by the general (transpose) path in coo.cu
int selfloop_fill_dispatch(const int32_t *indptr, const int32_t *indices, const void *data, int val_dtype,
                           int64_t n, const int32_t *at_indptr, int32_t *at_indices, double *at_val,
                           double *degree, double *dl, double *dr, double r, cudaStream_t s) {
  const unsigned blocks = (unsigned)ceil_div64(n, 256);
  if (val_dtype == SRG_VAL_ONES)
    selfloop_fill_kernel<SRG_VAL_ONES><<<blocks, 256, 0, s>>>(indptr, indices, data, n, at_indptr, at_indices, at_val, degree, dl, dr, r - 1.0, -r);
  else if (val_dtype == SRG_VAL_F32)
    selfloop_fill_kernel<SRG_VAL_F32><<<blocks, 256, 0, s>>>(indptr, indices, data, n, at_indptr, at_indices, at_val, degree, dl, dr, r - 1.0, -r);
  else
    selfloop_fill_kernel<SRG_VAL_F64><<<blocks, 256, 0, s>>>(indptr, indices, data, n, at_indptr, at_indices, at_val, degree, dl, dr, r - 1.0, -r);
  SRG_LAUNCHED();
  return SRG_OK;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_degree_selfloop_csr(const int32_t *indptr, const int32_t *indices,
                                       const void *data, int val_dtype, int64_t n,
                                       int32_t *out_indptr, int32_t *out_count,
                                       int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "degree_selfloop: negative n");
  SRG_REQUIRE(indptr && out_indptr && out_flags, "degree_selfloop: NULL pointer");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "degree_selfloop: bad val_dtype %d", val_dtype);
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || data != nullptr,
              "degree_selfloop: data is NULL but val_dtype says values");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int), s));
    return SRG_OK;
  }
  SRG_REQUIRE(indices != nullptr, "degree_selfloop: indices is NULL");
  int *scratch = nullptr;
  const int64_t scratch_ints = (out_count ? 0 : n) + scan_scratch_ints(n);
  SRG_CUDA(cudaMallocAsync(&scratch, scratch_ints * sizeof(int), s));
  int *rowlen = out_count ? out_count : scratch + scan_scratch_ints(n);
  const int64_t blocks = ceil_div64(n, 256);
  switch (val_dtype) {
    case SRG_VAL_ONES:
      selfloop_rowlen_kernel<SRG_VAL_ONES><<<(unsigned)blocks, 256, 0, s>>>(indptr, indices, data, n, rowlen, out_flags);
      break;
    case SRG_VAL_F32:
      selfloop_rowlen_kernel<SRG_VAL_F32><<<(unsigned)blocks, 256, 0, s>>>(indptr, indices, data, n, rowlen, out_flags);
      break;
    default:
      selfloop_rowlen_kernel<SRG_VAL_F64><<<(unsigned)blocks, 256, 0, s>>>(indptr, indices, data, n, rowlen, out_flags);
      break;
  }
  SRG_LAUNCHED();
  rc = exclusive_scan_i32(rowlen, n, out_indptr, scratch, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

template <int DT>
static int sym_norm_typed(const int32_t *indptr, const int32_t *indices, const void *data, int64_t n,
                          int64_t nnz, const int32_t *out_indptr, double r, double ppr_alpha,
                          int32_t *out_indices, double *out_degree, double *out_val_f64,
                          float *out_val_f32, int32_t *out_flags, cudaStream_t s) {
  // scratch: dl, dr (n doubles each) [+ degree if the caller does not want it] [+ A~ values]
  const int64_t cap = nnz + n;
  const int64_t n_doubles = 2 * n + (out_degree ? 0 : n) + (DT == SRG_VAL_ONES ? 0 : cap);
  double *scratch = nullptr;
  SRG_CUDA(cudaMallocAsync(&scratch, (size_t)n_doubles * sizeof(double), s));
  double *dl = scratch, *dr = scratch + n;
  double *deg = out_degree ? out_degree : scratch + 2 * n;
  double *at_val = (DT == SRG_VAL_ONES) ? nullptr : scratch + 2 * n + (out_degree ? 0 : n);
  const int64_t blocks = ceil_div64(n, 256);
  selfloop_fill_kernel<DT><<<(unsigned)blocks, 256, 0, s>>>(indptr, indices, data, n, out_indptr,
                                                            out_indices, at_val, deg, dl, dr,
                                                            r - 1.0, -r);
  SRG_LAUNCHED();
  const int use_ppr = ppr_alpha >= 0.0 ? 1 : 0;
  const int64_t wblocks = ceil_div64(n * 32, 256);
  SRG_REQUIRE(wblocks <= 2147483647LL, "sym_norm: too many rows");
  sym_norm_values_kernel<DT><<<(unsigned)wblocks, 256, 0, s>>>(
      n, out_indptr, out_indices, at_val, deg, dl, dr, 1.0 - ppr_alpha, ppr_alpha, use_ppr,
      out_val_f64, out_val_f32, out_flags);
  SRG_LAUNCHED();
  cudaFreeAsync(scratch, s);
  return SRG_OK;
}

extern "C" int srg_sym_norm_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                                int val_dtype, int64_t n, int64_t nnz, const int32_t *out_indptr,
                                double r, double ppr_alpha, int32_t *out_indices,
                                double *out_degree, double *out_val_f64, float *out_val_f32,
                                int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "sym_norm: negative size");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && out_indptr && out_indices && out_flags, "sym_norm: NULL pointer");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "sym_norm: bad val_dtype %d", val_dtype);
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || data != nullptr, "sym_norm: data is NULL but val_dtype says values");
  SRG_REQUIRE(nnz + n <= 2147483647LL, "sym_norm: nnz + n exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  switch (val_dtype) {
    case SRG_VAL_ONES:
      return sym_norm_typed<SRG_VAL_ONES>(indptr, indices, data, n, nnz, out_indptr, r, ppr_alpha, out_indices, out_degree, out_val_f64, out_val_f32, out_flags, s);
    case SRG_VAL_F32:
      return sym_norm_typed<SRG_VAL_F32>(indptr, indices, data, n, nnz, out_indptr, r, ppr_alpha, out_indices, out_degree, out_val_f64, out_val_f32, out_flags, s);
    default:
      return sym_norm_typed<SRG_VAL_F64>(indptr, indices, data, n, nnz, out_indptr, r, ppr_alpha, out_indices, out_degree, out_val_f64, out_val_f32, out_flags, s);
  }
}
