// cheby.cu — Chebyshev polynomial graph filter (heat-kernel wavelets) with the recurrence fused into
// the SpMM epilogue, fp64.
//
// Replaces the arithmetic the reference delegates to pygsp (absent, un-pinned):
//   wavelet/src/utils.py:83,95,131-133 (estimate_lmax / cheby_op / Heat / compute_cheby_coeff) and
//   SSRG/models/base_scalable/base_model.py:184-189,243.  Restated in oracle/__init__.py (cheby_op):
//
//     T0 = X;  T1 = (L X - a2 X) / a1;          r_s  = (0.5 c_s0) T0 + c_s1 T1
//     Tk = (2/a1) (L T_{k-1} - a2 T_{k-1}) - T_{k-2};   r_s += c_sk Tk          k = 2..M
//     threshold (wavelet/src/utils.py:98):  r_s[r_s < tol] = 0, cast to float32
//
// One kernel launch per order k: the gather L T_{k-1} is a CSR SpMM (sequential multiply-add chain in
// CSR order, no FMA contraction: the arithmetic of scipy's csr_matvecs), and the epilogue forms Tk
// and updates every scale's r_s in the same pass, so each T_k is written once and never re-read
// from HBM for the accumulation.  Both scales share the T_k (the reference recomputes them per
// scale).  Every operation is a separately rounded fp64 op in the oracle's order => bit-exact.
#include "common.cuh"
#include "scan.cuh"

namespace srg {

constexpr int kMaxScales = 4;
struct ChebyStep {
  double a2;          // lmax / 2
  double inv_scale;   // k == 1: a1 (division)   k >= 2: 2 / a1 (multiplication)
  double c_prev[kMaxScales];  // k == 1: 0.5 * c_s0
  double c_cur[kMaxScales];   // c_sk
  double tol;         // threshold applied when `last`
  int n_scales;
  int first;          // k == 1
  int last;           // k == M
  int use_tol;
};

__device__ __forceinline__ double2 ld_gather_d2(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// one warp per (row, chunk of 64 columns); lane owns 2 consecutive columns
template <int U>
__global__ void __launch_bounds__(256)
cheby_step_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                  const double *__restrict__ lvals, long long n, const double2 *__restrict__ Tcur,
                  const double2 *Tprev, double2 *Tnew, long long ld2, int nvec, int chunks,
                  double2 *const r0, double2 *const r1, double2 *const r2, double2 *const r3,
                  float2 *const q0, float2 *const q1, float2 *const q2, float2 *const q3,
                  long long ldq2, ChebyStep st) {
  const int lane = threadIdx.x & 31;
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long row = item / chunks;
  const int chunk = (int)(item - row * chunks);
  if (row >= n) return;
  const int col = chunk * 32 + lane;
  const bool active = col < nvec;
  const int s0 = __ldg(indptr + row), e0 = __ldg(indptr + row + 1);
  double2 acc = make_double2(0.0, 0.0);
  const double2 *Tc = Tcur + col;
  for (int base = s0; base < e0; base += 32) {
    int my_c = 0;
    double my_v = 0.0;
    if (base + lane < e0) {
      my_c = ld_stream_i32(indices + base + lane);
      my_v = __ldg(lvals + base + lane);
    }
    const int cnt = min(32, e0 - base);
    for (int t = 0; t < cnt; t += U) {
      double2 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_c, t + u);
        if (t + u < cnt && active) x[u] = ld_gather_d2(Tc + (long long)c * ld2);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double v = __shfl_sync(0xffffffffu, my_v, t + u);
        if (t + u < cnt && active) {
          acc.x = __dadd_rn(acc.x, __dmul_rn(v, x[u].x));
          acc.y = __dadd_rn(acc.y, __dmul_rn(v, x[u].y));
        }
      }
    }
  }
  if (!active) return;
  const long long off = row * ld2 + col;
  const double2 tc = Tcur[off];
  double2 tn;
  // w = L t - a2 t
  const double wx = __dsub_rn(acc.x, __dmul_rn(st.a2, tc.x));
  const double wy = __dsub_rn(acc.y, __dmul_rn(st.a2, tc.y));
  if (st.first) {
    tn.x = __ddiv_rn(wx, st.inv_scale);
    tn.y = __ddiv_rn(wy, st.inv_scale);
  } else {
    const double2 tp = Tprev[off];
    tn.x = __dsub_rn(__dmul_rn(st.inv_scale, wx), tp.x);
    tn.y = __dsub_rn(__dmul_rn(st.inv_scale, wy), tp.y);
  }
  if (!st.last) Tnew[off] = tn;  // T_M is only needed for the accumulation
  double2 *const rr[kMaxScales] = {r0, r1, r2, r3};
  float2 *const qq[kMaxScales] = {q0, q1, q2, q3};
#pragma unroll
  for (int s = 0; s < kMaxScales; ++s) {
    if (s < st.n_scales) {
      double2 r;
      if (st.first) {
        r.x = __dadd_rn(__dmul_rn(st.c_prev[s], tc.x), __dmul_rn(st.c_cur[s], tn.x));
        r.y = __dadd_rn(__dmul_rn(st.c_prev[s], tc.y), __dmul_rn(st.c_cur[s], tn.y));
      } else {
        r = rr[s][off];
        r.x = __dadd_rn(r.x, __dmul_rn(st.c_cur[s], tn.x));
        r.y = __dadd_rn(r.y, __dmul_rn(st.c_cur[s], tn.y));
      }
      if (st.last && st.use_tol) {
        if (r.x < st.tol) r.x = 0.0;
        if (r.y < st.tol) r.y = 0.0;
      }
      rr[s][off] = r;
      if (st.last && qq[s]) qq[s][row * ldq2 + col] = make_float2(__double2float_rn(r.x), __double2float_rn(r.y));
    }
  }
}

// ---- combinatorial Laplacian L = diag(W 1) - W over a canonical CSR -------------------------------
template <int DT> struct LVal;
template <> struct LVal<SRG_VAL_ONES> { __device__ static double at(const void *, long long) { return 1.0; } };
template <> struct LVal<SRG_VAL_F32> { __device__ static double at(const void *p, long long j) { return (double)static_cast<const float *>(p)[j]; } };
template <> struct LVal<SRG_VAL_F64> { __device__ static double at(const void *p, long long j) { return static_cast<const double *>(p)[j]; } };

template <int DT>
__device__ double pw_leaf(const void *data, long long b, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, LVal<DT>::at(data, b + i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = LVal<DT>::at(data, b + j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], LVal<DT>::at(data, b + i + j));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, LVal<DT>::at(data, b + i));
  return res;
}
template <int DT>
__device__ double pw_sum(const void *data, long long b, int n) {
  if (n <= 128) return pw_leaf<DT>(data, b, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(pw_sum<DT>(data, b, n2), pw_sum<DT>(data, b + n2, n - n2));
}

template <int DT, bool FILL>
__global__ void __launch_bounds__(256)
laplacian_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                 const void *__restrict__ data, long long n, int *__restrict__ rowlen,
                 const int *__restrict__ out_indptr, int *__restrict__ out_indices,
                 double *__restrict__ out_vals, double *__restrict__ degree, int *__restrict__ flags) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int s = indptr[a], e = indptr[a + 1];
  const int len = e - s;
  double d = 0.0;
  if (len == 1) d = LVal<DT>::at(data, s);
  else if (len > 1) d = __dadd_rn(LVal<DT>::at(data, s), pw_sum<DT>(data, s + 1, len - 1));
  double wdiag = 0.0;
  int fl = 0, prev = -1;
  for (int j = s; j < e; ++j) {
    const int b = indices[j];
    if (b <= prev) fl |= SRG_FLAG_UNSORTED;
    if (b < 0 || b >= n) fl |= SRG_FLAG_BAD_INDEX;
    prev = b;
    if (b == (int)a) wdiag = LVal<DT>::at(data, j);
  }
  const double diag = __dsub_rn(d, wdiag);
  if (!FILL) {
    int cnt = (diag != 0.0) ? 1 : 0;
    for (int j = s; j < e; ++j)
      if (indices[j] != (int)a && LVal<DT>::at(data, j) != 0.0) ++cnt;
    rowlen[a] = cnt;
    if (fl) atomicOr(flags, fl);
    return;
  }
  int p = out_indptr[a];
  bool placed = false;
  for (int j = s; j < e; ++j) {
    const int b = indices[j];
    if (b == (int)a) continue;
    if (!placed && b > (int)a) {
      placed = true;
      if (diag != 0.0) { out_indices[p] = (int)a; out_vals[p] = diag; ++p; }
    }
    const double v = LVal<DT>::at(data, j);
    if (v != 0.0) { out_indices[p] = b; out_vals[p] = __dsub_rn(0.0, v); ++p; }
  }
  if (!placed && diag != 0.0) { out_indices[p] = (int)a; out_vals[p] = diag; ++p; }
  if (degree) degree[a] = d;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_laplacian_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                                 int val_dtype, int64_t n, int32_t *out_indptr, int32_t *out_indices,
                                 double *out_vals, double *out_degree, int32_t *out_flags,
                                 void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "laplacian: negative n");
  SRG_REQUIRE(indptr && out_indptr && out_flags, "laplacian: NULL pointer");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "laplacian: bad val_dtype");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int), s));
    return SRG_OK;
  }
  SRG_REQUIRE(indices && out_indices && out_vals, "laplacian: NULL pointer");
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || data, "laplacian: data is NULL");
  int *scratch = nullptr;
  SRG_CUDA(cudaMallocAsync(&scratch, (size_t)(n + scan_scratch_ints(n)) * sizeof(int), s));
  int *rowlen = scratch + scan_scratch_ints(n);
  const unsigned blocks = (unsigned)ceil_div64(n, 256);
#define SRG_LAP(DT)                                                                                   \
  laplacian_kernel<DT, false><<<blocks, 256, 0, s>>>(indptr, indices, data, n, rowlen, nullptr, nullptr, nullptr, nullptr, out_flags); \
  SRG_LAUNCHED();                                                                                     \
  rc = exclusive_scan_i32(rowlen, n, out_indptr, scratch, s);                                         \
  if (!rc) {                                                                                          \
    laplacian_kernel<DT, true><<<blocks, 256, 0, s>>>(indptr, indices, data, n, nullptr, out_indptr, out_indices, out_vals, out_degree, out_flags); \
    SRG_LAUNCHED();                                                                                   \
  }
  if (val_dtype == SRG_VAL_ONES) { SRG_LAP(SRG_VAL_ONES) }
  else if (val_dtype == SRG_VAL_F32) { SRG_LAP(SRG_VAL_F32) }
  else { SRG_LAP(SRG_VAL_F64) }
#undef SRG_LAP
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int srg_cheby_filter_f64(const int32_t *lap_indptr, const int32_t *lap_indices,
                                    const double *lap_vals, int64_t n, const double *X, int64_t ld,
                                    int32_t B, double lmax, const double *coeffs, int32_t n_scales,
                                    int32_t order, double tol, double *const *out_r,
                                    float *const *out_r32, int64_t ld32, double *work0,
                                    double *work1, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && B >= 0, "cheby: negative size");
  SRG_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "cheby: n_scales must be 1..%d", kMaxScales);
  SRG_REQUIRE(order >= 1 && order <= 1024, "cheby: order must be >= 1 (got %d)", order);
  SRG_REQUIRE(coeffs && out_r, "cheby: NULL pointer");
  if (n == 0 || B == 0) return SRG_OK;
  SRG_REQUIRE(lap_indptr && lap_indices && lap_vals && X && work0 && work1, "cheby: NULL pointer");
  SRG_REQUIRE(ld >= B && ld % 2 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)work0 % 16 == 0 && (uintptr_t)work1 % 16 == 0,
              "cheby: fp64 matrices need an even leading dimension >= B and 16-byte alignment");
  SRG_REQUIRE(lmax > 0.0, "cheby: lmax must be positive");
  for (int s = 0; s < n_scales; ++s) {
    SRG_REQUIRE(out_r[s] && (uintptr_t)out_r[s] % 16 == 0, "cheby: out_r[%d] NULL or unaligned", s);
    if (out_r32 && out_r32[s]) SRG_REQUIRE(ld32 >= B && ld32 % 2 == 0 && (uintptr_t)out_r32[s] % 8 == 0, "cheby: bad float32 output layout");
  }
  cudaStream_t st = as_stream(stream);
  const int nvec = (B + 1) / 2;
  const int chunks = (nvec + 31) / 32;
  const int64_t warps = n * (int64_t)chunks;
  const int64_t blocks = ceil_div64(warps * 32, 256);
  SRG_REQUIRE(blocks <= 2147483647LL, "cheby: grid too large");
  const double a1 = lmax / 2.0, a2 = lmax / 2.0;
  const double *tprev = nullptr, *tcur = X;
  double *bufs[2] = {work0, work1};
  for (int k = 1; k <= order; ++k) {
    ChebyStep cs;
    cs.a2 = a2;
    cs.first = (k == 1);
    cs.last = (k == order);
    cs.inv_scale = cs.first ? a1 : 2.0 / a1;
    cs.n_scales = n_scales;
    cs.tol = tol;
    cs.use_tol = (tol == tol) ? 1 : 0;  // NaN = no threshold
    for (int s = 0; s < kMaxScales; ++s) {
      cs.c_prev[s] = (s < n_scales) ? 0.5 * coeffs[(size_t)s * (order + 1)] : 0.0;
      cs.c_cur[s] = (s < n_scales) ? coeffs[(size_t)s * (order + 1) + k] : 0.0;
    }
    double *tnew = bufs[(k - 1) & 1];
    double2 *r[kMaxScales] = {nullptr, nullptr, nullptr, nullptr};
    float2 *q[kMaxScales] = {nullptr, nullptr, nullptr, nullptr};
    for (int s = 0; s < n_scales; ++s) {
      r[s] = reinterpret_cast<double2 *>(out_r[s]);
      if (out_r32 && out_r32[s]) q[s] = reinterpret_cast<float2 *>(out_r32[s]);
    }
    cheby_step_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(
        lap_indptr, lap_indices, lap_vals, n, reinterpret_cast<const double2 *>(tcur),
        reinterpret_cast<const double2 *>(tprev), reinterpret_cast<double2 *>(tnew), ld / 2, nvec,
        chunks, r[0], r[1], r[2], r[3], q[0], q[1], q[2], q[3], ld32 / 2, cs);
    SRG_LAUNCHED();
    tprev = tcur;
    tcur = tnew;
  }
  return SRG_OK;
}
