// magnetic.cu — magnetic-Laplacian normalisation of a directed graph on the device.
//
// Reference: adj_to_directed_symmetric_mag_norm, SSRG/operators/utils.py:95-138 (torch + torch_sparse.coalesce
// + torch_scatter.scatter_add on the CPU), used by SymDirMagLaplacianGraphOp / SymDirMagComPprGraphOp
// (SSRG/operators/graph_operator/symmetrical_directed_magnetic_{laplacian,comppr}_operator.py):
//   pairs   = [(u, v, w, +w)] ++ [(v, u, w, -w)]                                   (:100-104)
//   (sym, theta)[u, v] = coalesce(pairs, "add");  sym /= 2                          (:105-108)
//   append the loops (i, i, 1.0, 0.0) as SEPARATE entries                           (:109-119)
//   deg[u]  = scatter_add(sym) over the row, coalesced entries in column order, the loop last (:122)
//   x       = (deg[u]^(r-1) * sym) * deg[v]^(-r);  val = x * exp(i * 2 pi q theta)   (:124-130)
//   real / imag CSR = csr_matrix((val.real / val.imag, (row, col)))  — the loop is summed into an
//   existing (i, i) entry here                                                       (:133-136)
// All of it is "sort (row, col) keys, then segment": one stable key sort (cub::DeviceRadixSort), head flags,
// scans and per-entry fp64 arithmetic in the reference's operation order.  Set-up path, not per hop.
#include "common.cuh"
#include "scan.cuh"
#include "sortutil.cuh"

namespace srg {

// key layout: row << 33 | col << 1 | is_appended_loop
__device__ __forceinline__ uint64_t mag_key(long long row, long long col, int loop) {
  return ((uint64_t)row << 33) | ((uint64_t)col << 1) | (uint64_t)loop;
}

// positions: [0, nnz) the entries (u, v), [nnz, 2 nnz) their mirrors (v, u), [2 nnz, 2 nnz + n) the loops —
// the concatenation order of the reference, which the stable sort preserves inside equal keys
template <typename T>
__global__ void __launch_bounds__(256)
mag_pairs_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const T *__restrict__ data,
                 long long n, long long nnz, uint64_t *__restrict__ keys, unsigned *__restrict__ pos,
                 double *__restrict__ sym, double *__restrict__ th, int *__restrict__ flags) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= n) return;
  const int lane = threadIdx.x & 31;
  for (long long j = (long long)indptr[a] + lane; j < indptr[a + 1]; j += 32) {
    int b = indices[j];
    if (b < 0 || b >= n) {
      atomicOr(flags, SRG_FLAG_BAD_INDEX);
      b = 0;
    }
    const double w = data ? (double)data[j] : 1.0;
    keys[j] = mag_key(a, b, 0);
    keys[nnz + j] = mag_key(b, a, 0);
    pos[j] = (unsigned)j;
    pos[nnz + j] = (unsigned)(nnz + j);
    sym[j] = w;
    sym[nnz + j] = w;
    th[j] = w;
    th[nnz + j] = -w;
  }
  if (lane == 0) {
    const long long p = 2 * nnz + a;
    keys[p] = mag_key(a, a, 1);
    pos[p] = (unsigned)p;
    sym[p] = 1.0;
    th[p] = 0.0;
  }
}

// head[i] = 1 when sorted key i starts a new segment; `shift` = 0 compares whole keys (coalesce, loops kept
// apart), 1 drops the loop bit (the final csr_matrix duplicate sum)
__global__ void mag_heads_kernel(const uint64_t *__restrict__ keys, long long m, const int *__restrict__ m_dev,
                                 int shift, int *__restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long live = m_dev ? (long long)*m_dev : m;
  head[i] = (i < live && (i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift))) ? 1 : 0;
}

// one thread per segment head: sums in sorted (= reference concatenation) order
__global__ void mag_coalesce_kernel(const uint64_t *__restrict__ keys, const unsigned *__restrict__ pos,
                                    const int *__restrict__ head, const int *__restrict__ seg, long long m,
                                    const double *__restrict__ sym, const double *__restrict__ th,
                                    uint64_t *__restrict__ e_key, double *__restrict__ e_sym,
                                    double *__restrict__ e_th) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !head[i]) return;
  double s = sym[pos[i]], t = th[pos[i]];
  for (long long j = i + 1; j < m && !head[j]; ++j) {
    s = __dadd_rn(s, sym[pos[j]]);
    t = __dadd_rn(t, th[pos[j]]);
  }
  const int slot = seg[i];
  const bool loop = keys[i] & 1ull;
  e_key[slot] = keys[i];
  e_sym[slot] = loop ? 1.0 : __ddiv_rn(s, 2.0);
  e_th[slot] = loop ? 0.0 : t;
}

// rowptr[r] = first entry whose row is >= r
__global__ void mag_row_lower_bound_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ total,
                                           long long n, int *__restrict__ rowptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  int lo = 0, hi = *total;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> 33) < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = lo;
}

// deg[u] = sum of the coalesced entries of the row in column order, then the appended loop (scatter_add order)
__global__ void mag_degree_kernel(const int *__restrict__ rowptr, const uint64_t *__restrict__ e_key,
                                  const double *__restrict__ e_sym, long long n, double *__restrict__ deg) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (int p = rowptr[r]; p < rowptr[r + 1]; ++p)
    if (!(e_key[p] & 1ull)) d = __dadd_rn(d, e_sym[p]);
  deg[r] = __dadd_rn(d, 1.0);
}

// x = (dl[row] * sym) * dr[col];  (real, imag) = x * (cos, sin)(c * theta)
__global__ void mag_values_kernel(const uint64_t *__restrict__ e_key, const double *__restrict__ e_sym,
                                  const double *__restrict__ e_th, const int *__restrict__ total,
                                  const double *__restrict__ dl, const double *__restrict__ dr, double c,
                                  double *__restrict__ re, double *__restrict__ im) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *total) return;
  const uint64_t k = e_key[i];
  const long long row = (long long)(k >> 33), col = (long long)((k >> 1) & 0xffffffffull);
  const double x = __dmul_rn(__dmul_rn(dl[row], e_sym[i]), dr[col]);
  double sn, cs;
  sincos(__dmul_rn(c, e_th[i]), &sn, &cs);
  re[i] = __dmul_rn(x, cs);
  im[i] = __dmul_rn(x, sn);
}

// final duplicate sum (coalesced (i, i) entry + appended loop) and the optional PPR blend
// real = (1 - alpha) * real + alpha * I,  imag = (1 - alpha) * imag  (comppr_operator.py:33-38)
__global__ void mag_finalize_kernel(const uint64_t *__restrict__ e_key, const int *__restrict__ head,
                                    const int *__restrict__ seg, const int *__restrict__ total,
                                    const double *__restrict__ re, const double *__restrict__ im, double alpha,
                                    int *__restrict__ out_indices, double *__restrict__ re64,
                                    double *__restrict__ im64, float *__restrict__ re32, float *__restrict__ im32,
                                    uint64_t *__restrict__ f_key, int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = *total;
  if (i >= m || !head[i]) return;
  double r = re[i], q = im[i];
  for (long long j = i + 1; j < m && !head[j]; ++j) {
    r = __dadd_rn(r, re[j]);
    q = __dadd_rn(q, im[j]);
  }
  const uint64_t k = e_key[i];
  const long long row = (long long)(k >> 33), col = (long long)((k >> 1) & 0xffffffffull);
  if (alpha >= 0.0) {
    r = __dmul_rn(1.0 - alpha, r);
    q = __dmul_rn(1.0 - alpha, q);
    if (row == col) r = __dadd_rn(r, alpha);
    if (r == 0.0) atomicOr(flags, SRG_FLAG_ZERO_PRODUCT);   // scipy's `+` would drop the entry
  }
  const int slot = seg[i];
  f_key[slot] = k;
  out_indices[slot] = (int)col;
  if (re64) re64[slot] = r;
  if (im64) im64[slot] = q;
  if (re32) re32[slot] = __double2float_rn(r);
  if (im32) im32[slot] = __double2float_rn(q);
}

}  // namespace srg

using namespace srg;

extern "C" int srg_mag_norm_csr(const int32_t *indptr, const int32_t *indices, const void *data, int val_dtype,
                                int64_t n, int64_t nnz, double r, double q_angle, double ppr_alpha,
                                int32_t *out_indptr, int32_t *out_indices, double *out_degree, double *out_real_f64,
                                double *out_imag_f64, float *out_real_f32, float *out_imag_f32,
                                int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "mag_norm: negative size");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "mag_norm: bad val_dtype");
  SRG_REQUIRE(indptr && out_indptr && out_flags, "mag_norm: NULL pointer");
  SRG_REQUIRE(nnz == 0 || indices, "mag_norm: indices is NULL");
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || nnz == 0 || data, "mag_norm: data is NULL");
  const int64_t m = 2 * nnz + n;
  SRG_REQUIRE(m <= 2147483647LL, "mag_norm: 2 nnz + n exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int32_t), s));
    return SRG_OK;
  }
  SRG_REQUIRE(out_indices, "mag_norm: out_indices is NULL");
  uint64_t *keys = nullptr;   // 2m sort buffers + m coalesced keys + m final keys
  unsigned *pos = nullptr;    // 2m
  double *dbl = nullptr;      // sym, th (m each) | e_sym, e_th, re, im (m each) | deg, dl, dr (n each)
  int *ints = nullptr;        // head (m+1), seg (m+1), rowptr (n+1), scan scratch
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(4 * m) * sizeof(uint64_t), s));
  SRG_CUDA(cudaMallocAsync(&pos, (size_t)(2 * m) * sizeof(unsigned), s));
  SRG_CUDA(cudaMallocAsync(&dbl, (size_t)(6 * m + 3 * n) * sizeof(double), s));
  SRG_CUDA(cudaMallocAsync(&ints, (size_t)(2 * (m + 1) + (n + 1) + scan_scratch_ints(m)) * sizeof(int), s));
  double *sym = dbl, *th = dbl + m, *e_sym = dbl + 2 * m, *e_th = dbl + 3 * m, *re = dbl + 4 * m, *im = dbl + 5 * m;
  double *deg = out_degree ? out_degree : dbl + 6 * m, *dl = dbl + 6 * m + n, *dr = dbl + 6 * m + 2 * n;
  int *head = ints, *seg = ints + (m + 1), *rowptr = seg + (m + 1), *scratch = rowptr + (n + 1);
  uint64_t *e_key = keys + 2 * m, *f_key = keys + 3 * m;
  const unsigned wb = (unsigned)ceil_div64(n * 32, 256), mb = (unsigned)ceil_div64(m, 256);

  if (val_dtype == SRG_VAL_F32)
    mag_pairs_kernel<float><<<wb, 256, 0, s>>>(indptr, indices, static_cast<const float *>(data), n, nnz, keys, pos, sym, th, out_flags);
  else if (val_dtype == SRG_VAL_F64)
    mag_pairs_kernel<double><<<wb, 256, 0, s>>>(indptr, indices, static_cast<const double *>(data), n, nnz, keys, pos, sym, th, out_flags);
  else
    mag_pairs_kernel<double><<<wb, 256, 0, s>>>(indptr, indices, nullptr, n, nnz, keys, pos, sym, th, out_flags);
  SRG_LAUNCHED();
  rc = sort_pairs<unsigned>(keys, keys + m, pos, pos + m, m, 33 + bits_for(n), s);
  if (!rc) {
    // coalesce: equal whole keys (appended loops stay separate entries)
    mag_heads_kernel<<<mb, 256, 0, s>>>(keys + m, m, nullptr, 0, head);
    SRG_LAUNCHED();
    rc = exclusive_scan_i32(head, m, seg, scratch, s);   // seg[m] = number of coalesced entries
  }
  if (!rc) {
    mag_coalesce_kernel<<<mb, 256, 0, s>>>(keys + m, pos + m, head, seg, m, sym, th, e_key, e_sym, e_th);
    SRG_LAUNCHED();
    mag_row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(e_key, seg + m, n, rowptr);
    SRG_LAUNCHED();
    mag_degree_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(rowptr, e_key, e_sym, n, deg);
    SRG_LAUNCHED();
    rc = srg_pow_tables_f64(deg, n, r, dl, dr, s);
  }
  if (!rc) {
    // the live entry count stays on the device: park it where the second scan does not overwrite it
    int *total1 = rowptr + n;   // rowptr[n] == seg[m] == number of coalesced entries
    mag_values_kernel<<<mb, 256, 0, s>>>(e_key, e_sym, e_th, total1, dl, dr, q_angle, re, im);
    SRG_LAUNCHED();
    // csr_matrix((vals, (row, col))): duplicates (coalesced (i, i) + appended loop) are summed
    mag_heads_kernel<<<mb, 256, 0, s>>>(e_key, m, total1, 1, head);
    SRG_LAUNCHED();
    rc = exclusive_scan_i32(head, m, seg, scratch, s);   // seg[m] = final entry count
    if (!rc) {
      mag_finalize_kernel<<<mb, 256, 0, s>>>(e_key, head, seg, total1, re, im, ppr_alpha, out_indices, out_real_f64,
                                             out_imag_f64, out_real_f32, out_imag_f32, f_key, out_flags);
      SRG_LAUNCHED();
      mag_row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(f_key, seg + m, n, out_indptr);
      SRG_LAUNCHED();
    }
  }
  cudaFreeAsync(ints, s);
  cudaFreeAsync(dbl, s);
  cudaFreeAsync(pos, s);
  cudaFreeAsync(keys, s);
  return rc;
}
