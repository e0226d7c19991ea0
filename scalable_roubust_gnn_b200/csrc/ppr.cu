// ppr.cu — device stages of the fast PPR-approximation normaliser of directed graphs.
//
// Reference: adj_to_fast_ppr_approx_symmetric_norm, SSRG/operators/utils.py:262-335 (scipy on the CPU):
//   A1 = pattern(A) + one appended loop per node, duplicates summed; r = A1 1; D1 = diag(1 / r)          (:263-277)
//   fixed point  x <- (1 - a) A1^T D1 x + s (z . x),  s = 1 / ((1 + a) n),  until |x - x_old|_2 <= 1e-6
//   or 100 iterations; pi = x / sum(x)                                                                      (:278-296)
//   P = D1 A1;  L = (Pi^1/2 P Pi^-1/2 + Pi^-1/2 P^T Pi^1/2) / 2,  NaN -> 0                                  (:297-301)
//   values -> float32, then D^(r-1) L D^(-r) with D = row sums (float32)                                    (:304-334)
// Stages here (the host loop, the convergence test and the float32 normalisation reuse existing entry points):
//   srg_ppr_iterate_f64   one sweep  y = (1 - a) A1^T (x / r) + s (z . x)  and  |y - x|^2, sum(y)
//   srg_ppr_symmetrize    the union pattern of A1 and A1^T with the two pi-weighted terms summed, / 2
// fp64 throughout; the sums run in a fixed (deterministic) order, not scipy's internal one, so the stationary
// vector agrees to rounding (1e-15) and the final float32 values to float32 rounding.
// Scratch is stream-ordered and scoped (StreamScratch): every return path hands its blocks back.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"
#include "sortutil.cuh"

namespace srg {

// z . x with z_i = c_nz if r_i != 0 else c_z  (utils.py:281-283); one block, fixed-order tree: deterministic
__global__ void __launch_bounds__(1024)
ppr_dot_kernel(const double *__restrict__ x, const double *__restrict__ deg, long long n, double c_nz, double c_z,
               double *__restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc = __dadd_rn(acc, __dmul_rn(deg[i] != 0.0 ? c_nz : c_z, x[i]));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// y_i = (1 - a) * sum_j A1[j, i] * (x_j / r_j) + s * zx      over row i of A1^T (CSR, float32 counts), stored order
__global__ void __launch_bounds__(256)
ppr_sweep_kernel(const int *__restrict__ t_ptr, const int *__restrict__ t_idx, const float *__restrict__ t_cnt,
                 const double *__restrict__ x, const double *__restrict__ deg, long long n, double one_minus_a,
                 double s, const double *__restrict__ zx, double *__restrict__ y, double *__restrict__ part) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0, sy = 0.0;
  if (i < n) {
    double acc = 0.0;
    for (int p = t_ptr[i]; p < t_ptr[i + 1]; ++p) {
      const int j = t_idx[p];
      const double w = __dmul_rn(__dmul_rn(one_minus_a, (double)t_cnt[p]), __ddiv_rn(1.0, deg[j]));
      acc = __dadd_rn(acc, __dmul_rn(w, x[j]));
    }
    const double yi = __dadd_rn(acc, __dmul_rn(s, zx[0]));
    y[i] = yi;
    const double d = __dsub_rn(yi, x[i]);
    d2 = __dmul_rn(d, d);
    sy = yi;
  }
  // per-block partials (fixed-order tree), summed by the finishing kernel: deterministic
  __shared__ double sh_d[256], sh_s[256];
  sh_d[threadIdx.x] = d2;
  sh_s[threadIdx.x] = sy;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh_d[threadIdx.x] = __dadd_rn(sh_d[threadIdx.x], sh_d[threadIdx.x + o]);
      sh_s[threadIdx.x] = __dadd_rn(sh_s[threadIdx.x], sh_s[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * (long long)blockIdx.x] = sh_d[0];
    part[2 * (long long)blockIdx.x + 1] = sh_s[0];
  }
}

__global__ void __launch_bounds__(1024)
ppr_finish_kernel(const double *__restrict__ part, long long blocks, double *__restrict__ out) {
  __shared__ double sh_d[1024], sh_s[1024];
  double d = 0.0, s = 0.0;
  for (long long b = threadIdx.x; b < blocks; b += 1024) {
    d = __dadd_rn(d, part[2 * b]);
    s = __dadd_rn(s, part[2 * b + 1]);
  }
  sh_d[threadIdx.x] = d;
  sh_s[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh_d[threadIdx.x] = __dadd_rn(sh_d[threadIdx.x], sh_d[threadIdx.x + o]);
      sh_s[threadIdx.x] = __dadd_rn(sh_s[threadIdx.x], sh_s[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[1] = sh_d[0];   // |y - x|^2
    out[2] = sh_s[0];   // sum(y)
  }
}

// entry (u, v, c) of A1 gives  t1 = (sqrt(pi_u) * p) * pi_v^-1/2  at key (u, v)  and
//                              t2 = (pi_v^-1/2 * p) * sqrt(pi_u)  at key (v, u),   p = c / r_u   (utils.py:297-300)
__global__ void __launch_bounds__(256)
ppr_pairs_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const float *__restrict__ cnt,
                 const double *__restrict__ deg, const double *__restrict__ x, const double *__restrict__ stats,
                 long long n, long long nnz, uint64_t *__restrict__ keys, unsigned *__restrict__ pos,
                 double *__restrict__ val) {
  const long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n) return;
  const int lane = threadIdx.x & 31;
  const double total = stats[2];                      // sum(x): pi = x / sum(x)
  const double pu = __ddiv_rn(x[u], total);
  const double su = sqrt(pu);
  const double inv_r = __ddiv_rn(1.0, deg[u]);
  for (long long j = (long long)indptr[u] + lane; j < indptr[u + 1]; j += 32) {
    const int v = indices[j];
    const double pv = __ddiv_rn(x[v], total);
    const double iv = pow(pv, -0.5);
    const double p = __dmul_rn(inv_r, (double)cnt[j]);
    keys[j] = ((uint64_t)u << 32) | (uint64_t)(unsigned)v;
    keys[nnz + j] = ((uint64_t)(unsigned)v << 32) | (uint64_t)u;
    pos[j] = (unsigned)j;
    pos[nnz + j] = (unsigned)(nnz + j);
    val[j] = __dmul_rn(__dmul_rn(su, p), iv);
    val[nnz + j] = __dmul_rn(__dmul_rn(iv, p), su);
  }
}

__global__ void ppr_heads_kernel(const uint64_t *__restrict__ keys, long long m, int *__restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void ppr_compress_kernel(const uint64_t *__restrict__ keys, const unsigned *__restrict__ pos,
                                    const int *__restrict__ head, const int *__restrict__ seg, long long m,
                                    const double *__restrict__ val, uint64_t *__restrict__ u_key,
                                    int *__restrict__ out_idx, float *__restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !head[i]) return;
  double s = val[pos[i]];
  for (long long j = i + 1; j < m && !head[j]; ++j) s = __dadd_rn(s, val[pos[j]]);
  s = __ddiv_rn(s, 2.0);
  if (s != s) s = 0.0;                                 // L.data[np.isnan(L.data)] = 0.0  (:301)
  const int slot = seg[i];
  u_key[slot] = keys[i];
  out_idx[slot] = (int)(keys[i] & 0xffffffffu);
  out_val[slot] = __double2float_rn(s);                // torch.FloatTensor(values)  (:309)
}

__global__ void ppr_row_lower_bound_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ total,
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

// ---- two-order PPR approximation (SSRG/operators/utils.py:337-424) ------------------------------------------------
// left eigenvector of the (n + 1) x (n + 1) teleport matrix [[(1 - a) P, a], [1/n, 0]] (the reference calls LAPACK on
// the dense float32 matrix, :353-369): one sweep of the power iteration over the CSR of P^T,
//   y_j = (1 - a) sum_i P_ij x_i + x_n / n   (j < n),      y_n = a sum_{i<n} x_i
__global__ void __launch_bounds__(256)
teleport_sweep_kernel(const int *__restrict__ t_ptr, const int *__restrict__ t_idx, const float *__restrict__ t_val,
                      const double *__restrict__ x, long long n, double one_minus_a, double a,
                      const double *__restrict__ sum_x, double *__restrict__ y, double *__restrict__ part) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d1 = 0.0, sy = 0.0;
  if (j < n) {
    double acc = 0.0;
    for (int p = t_ptr[j]; p < t_ptr[j + 1]; ++p) acc = __dadd_rn(acc, __dmul_rn((double)t_val[p], x[t_idx[p]]));
    const double yj = __dadd_rn(__dmul_rn(one_minus_a, acc), __ddiv_rn(x[n], (double)n));
    y[j] = yj;
    d1 = fabs(__dsub_rn(yj, x[j]));
    sy = yj;
  } else if (j == n) {
    y[n] = __dmul_rn(a, sum_x[0]);
  }
  __shared__ double sh_d[256], sh_s[256];
  sh_d[threadIdx.x] = d1;
  sh_s[threadIdx.x] = sy;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh_d[threadIdx.x] = __dadd_rn(sh_d[threadIdx.x], sh_d[threadIdx.x + o]);
      sh_s[threadIdx.x] = __dadd_rn(sh_s[threadIdx.x], sh_s[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * (long long)blockIdx.x] = sh_d[0];
    part[2 * (long long)blockIdx.x + 1] = sh_s[0];
  }
}

// L = (L_in + L_out) / 2 on the entries where BOTH are non-zero: the in-place masking of :405-410
// (L_in[L_out == 0] = 0, then L_out[L_in == 0] = 0 on the already masked L_in).  Rows of both inputs are sorted.
__global__ void intersect_count_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                       const float *__restrict__ a_val, const int *__restrict__ b_ptr,
                                       const int *__restrict__ b_idx, const float *__restrict__ b_val, long long n,
                                       int *__restrict__ cnt, int *__restrict__ match) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  int c = 0;
  const int b0 = b_ptr[i], b1 = b_ptr[i + 1];
  for (int p = a_ptr[i] + lane; p < a_ptr[i + 1]; p += 32) {
    const int col = a_idx[p];
    int lo = b0, hi = b1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (b_idx[mid] < col) lo = mid + 1; else hi = mid;
    }
    const bool hit = lo < b1 && b_idx[lo] == col && a_val[p] != 0.f && b_val[lo] != 0.f;
    match[p] = hit ? lo : -1;
    c += hit ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[i] = c;
}

__global__ void intersect_fill_kernel(const int *__restrict__ a_ptr, const int *__restrict__ a_idx,
                                      const float *__restrict__ a_val, const float *__restrict__ b_val,
                                      const int *__restrict__ match, const int *__restrict__ out_ptr, long long n,
                                      int *__restrict__ out_idx, float *__restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per row keeps the order
  if (i >= n) return;
  int w = out_ptr[i];
  for (int p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
    const int q = match[p];
    if (q < 0) continue;
    out_idx[w] = a_idx[p];
    out_val[w] = __fdiv_rn(__fadd_rn(a_val[p], b_val[q]), 2.0f);
    ++w;
  }
}

}  // namespace srg

using namespace srg;

extern "C" int srg_teleport_iterate_f64(const int32_t *t_indptr, const int32_t *t_indices, const float *t_vals, int64_t n,
                                        double ppr_alpha, const double *x, double *y, double *stats3, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 1 && t_indptr && t_indices && t_vals && x && y && stats3 && x != y, "teleport_iterate: bad arguments");
  cudaStream_t s = as_stream(stream);
  const long long blocks = ceil_div64(n + 1, 256);
  StreamScratch scratch(s);
  double *part = nullptr;
  if ((rc = scratch.alloc(&part, (size_t)(2 * blocks)))) return rc;
  ppr_dot_kernel<<<1, 1024, 0, s>>>(x, x, n, 1.0, 1.0, stats3);   // stats3[0] = sum_{i<n} x_i
  SRG_LAUNCHED();
  teleport_sweep_kernel<<<(unsigned)blocks, 256, 0, s>>>(t_indptr, t_indices, t_vals, x, n, 1 - ppr_alpha, ppr_alpha, stats3,
                                                        y, part);
  SRG_LAUNCHED();
  ppr_finish_kernel<<<1, 1024, 0, s>>>(part, blocks, stats3);     // stats3[1] = |y - x|_1 (first n), [2] = sum_{j<n} y_j
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_csr_intersect_mean_f32(const int32_t *a_indptr, const int32_t *a_indices, const float *a_vals,
                                          const int32_t *b_indptr, const int32_t *b_indices, const float *b_vals,
                                          int64_t n, int64_t a_nnz, int32_t *out_indptr, int32_t *out_indices,
                                          float *out_vals, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && a_nnz >= 0 && a_indptr && b_indptr && out_indptr, "csr_intersect_mean: bad arguments");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int32_t), s));
    return SRG_OK;
  }
  SRG_REQUIRE(a_nnz == 0 || (a_indices && a_vals && b_indices && b_vals && out_indices && out_vals),
              "csr_intersect_mean: NULL pointer");
  StreamScratch pool(s);
  int *ints = nullptr;   // cnt (n) | match (a_nnz) | scratch
  if ((rc = pool.alloc(&ints, (size_t)(n + std::max<int64_t>(a_nnz, 1) + scan_scratch_ints(n))))) return rc;
  int *cnt = ints, *match = ints + n, *scratch = match + std::max<int64_t>(a_nnz, 1);
  intersect_count_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, s>>>(a_indptr, a_indices, a_vals, b_indptr, b_indices,
                                                                        b_vals, n, cnt, match);
  SRG_LAUNCHED();
  rc = exclusive_scan_i32(cnt, n, out_indptr, scratch, s);
  if (!rc && a_nnz > 0) {
    intersect_fill_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(a_indptr, a_indices, a_vals, b_vals, match, out_indptr,
                                                                     n, out_indices, out_vals);
    SRG_LAUNCHED();
  }
  return rc;
}

extern "C" int srg_ppr_iterate_f64(const int32_t *t_indptr, const int32_t *t_indices, const float *t_counts,
                                   const double *degree, int64_t n, double ppr_alpha, const double *x, double *y,
                                   double *stats3, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 1 && t_indptr && t_indices && t_counts && degree && x && y && stats3 && x != y,
              "ppr_iterate: bad arguments");
  cudaStream_t s = as_stream(stream);
  const double a = ppr_alpha;
  const double c_nz = a * (1 + a), c_z = (1 - a) / (1 + a) + a * (1 + a);
  const double sv = 1 / (1 + a) / (double)n;
  const long long blocks = ceil_div64(n, 256);
  StreamScratch scratch(s);
  double *part = nullptr;
  if ((rc = scratch.alloc(&part, (size_t)(2 * blocks)))) return rc;
  ppr_dot_kernel<<<1, 1024, 0, s>>>(x, degree, n, c_nz, c_z, stats3);
  SRG_LAUNCHED();
  ppr_sweep_kernel<<<(unsigned)blocks, 256, 0, s>>>(t_indptr, t_indices, t_counts, x, degree, n, 1 - a, sv, stats3, y, part);
  SRG_LAUNCHED();
  ppr_finish_kernel<<<1, 1024, 0, s>>>(part, blocks, stats3);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_ppr_symmetrize(const int32_t *indptr, const int32_t *indices, const float *counts,
                                  const double *degree, const double *x, const double *stats3, int64_t n, int64_t nnz,
                                  int32_t *out_indptr, int32_t *out_indices, float *out_vals, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 1 && nnz >= 1 && indptr && indices && counts && degree && x && stats3 && out_indptr && out_indices &&
                  out_vals,
              "ppr_symmetrize: bad arguments");
  const int64_t m = 2 * nnz;
  SRG_REQUIRE(m <= 2147483647LL, "ppr_symmetrize: 2 nnz exceeds the int32 range");
  cudaStream_t s = as_stream(stream);
  uint64_t *keys = nullptr;   // 2m sort buffers + m unique keys
  unsigned *pos = nullptr;    // 2m
  double *val = nullptr;      // m
  int *ints = nullptr;        // head (m+1) | seg (m+1) | scratch
  StreamScratch pool(s);
  if ((rc = pool.alloc(&keys, (size_t)(3 * m)))) return rc;
  if ((rc = pool.alloc(&pos, (size_t)(2 * m)))) return rc;
  if ((rc = pool.alloc(&val, (size_t)m))) return rc;
  if ((rc = pool.alloc(&ints, (size_t)(2 * (m + 1) + scan_scratch_ints(m))))) return rc;
  int *head = ints, *seg = ints + (m + 1), *scratch = seg + (m + 1);
  ppr_pairs_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, s>>>(indptr, indices, counts, degree, x, stats3, n, nnz, keys,
                                                                  pos, val);
  SRG_LAUNCHED();
  rc = sort_pairs<unsigned>(keys, keys + m, pos, pos + m, m, 32 + bits_for(n > 1 ? n : 2), s);
  if (!rc) {
    ppr_heads_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, s>>>(keys + m, m, head);
    SRG_LAUNCHED();
    rc = exclusive_scan_i32(head, m, seg, scratch, s);   // seg[m] = entries of L
  }
  if (!rc) {
    ppr_compress_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, s>>>(keys + m, pos + m, head, seg, m, val, keys + 2 * m,
                                                                   out_indices, out_vals);
    SRG_LAUNCHED();
    ppr_row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(keys + 2 * m, seg + m, n, out_indptr);
    SRG_LAUNCHED();
  }
  return rc;
}
