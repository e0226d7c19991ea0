// spmm.cu — fp32 CSR x dense propagation hop for sm_100a (B200).
//
// Replaces SSRG/operators/csrc/matmul.c:23-40 (FloatCSRMulDenseOMP): for every row i and feature k
//   answer[i,k] = fma(data[j], mat[indices[j],k], answer[i,k])   for j = indptr[i] .. indptr[i+1]-1
// in CSR order, starting from 0.  Every output element here is the same sequential fp32 FMA chain,
// so a hop is bit-identical to the reference; only the mapping to the machine differs:
//
//   * a "group" of G lanes (G = 1..32, power of two) owns one (row, 4*G-float feature chunk);
//     each lane keeps ONE float4 accumulator => 128-bit gathers, 4 FMA chains per lane.
//   * the group loads G (index, value) pairs of its row with one coalesced read and broadcasts
//     them with shuffles, so the CSR arrays are read exactly once per chunk.
//   * U gathers are issued back to back before the first FMA consumes them (memory-level
//     parallelism: U*16 B per lane in flight); the FMAs are then applied in CSR order.
//   * HBM-bound (0.5 flop/B): no shared memory tile reuse exists for X on a random graph, so the
//     kernel is sized for occupancy and bytes in flight, not for tensor cores.
//
// Long rows: handled by the split kernels further down (fixed-order partial sums, deterministic).
#include "common.cuh"

namespace srg {

// ---- vector abstraction: float4 fast path, float scalar path ---------------------------------
template <typename VT> struct VecOps;
template <> struct VecOps<float4> {
  static constexpr int W = 4;
  __device__ static __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static __forceinline__ float4 gather(const float4 *p) { return ld_gather_f4(p); }
  __device__ static __forceinline__ void fma(float a, const float4 &x, float4 &acc) {
    acc.x = fmaf(a, x.x, acc.x);
    acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z);
    acc.w = fmaf(a, x.w, acc.w);
  }
  __device__ static __forceinline__ void store(float4 *p, const float4 &v) { *p = v; }
};
template <> struct VecOps<float> {
  static constexpr int W = 1;
  __device__ static __forceinline__ float zero() { return 0.f; }
  __device__ static __forceinline__ float gather(const float *p) { return __ldg(p); }
  __device__ static __forceinline__ void fma(float a, const float &x, float &acc) {
    acc = fmaf(a, x, acc);
  }
  __device__ static __forceinline__ void store(float *p, const float &v) { *p = v; }
};

constexpr int kSpmmThreads = 256;

template <typename VT, int G, int U, bool ACCUM>
__global__ void __launch_bounds__(kSpmmThreads)
spmm_group_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                  const float *__restrict__ vals, long long n_rows, const VT *__restrict__ X,
                  long long ldx, VT *__restrict__ Y, long long ldy, int nvec, int chunks) {
  using Ops = VecOps<VT>;
  constexpr int GROUPS = kSpmmThreads / G;
  const int g = threadIdx.x % G;
  const long long item = (long long)blockIdx.x * GROUPS + threadIdx.x / G;
  const long long row = item / chunks;
  const int chunk = (int)(item - row * chunks);
  if (row >= n_rows) return;  // group-uniform
  const unsigned gmask =
      (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  const int col = chunk * G + g;  // in units of VT
  const bool active = col < nvec;

  const int st = __ldg(indptr + row);
  const int ed = __ldg(indptr + row + 1);
  VT acc = Ops::zero();
  if (ACCUM && active) acc = Y[row * ldy + col];

  const VT *Xc = X + col;
  for (int base = st; base < ed; base += G) {
    int my_c = 0;
    float my_v = 0.f;
    if (base + g < ed) {
      my_c = ld_stream_i32(indices + base + g);
      my_v = ld_stream_f32(vals + base + g);
    }
    const int cnt = min(G, ed - base);
    for (int t = 0; t < cnt; t += U) {
      VT x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(gmask, my_c, t + u, G);
        if (t + u < cnt && active) x[u] = Ops::gather(Xc + (long long)c * ldx);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float v = __shfl_sync(gmask, my_v, t + u, G);
        if (t + u < cnt && active) Ops::fma(v, x[u], acc);
      }
    }
  }
  if (active) Ops::store(Y + row * ldy + col, acc);
}

template <typename VT, int G, bool ACCUM>
static int launch_group(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                        const VT *X, int64_t ldx, VT *Y, int64_t ldy, int nvec, cudaStream_t s) {
  constexpr int U = (G >= 8) ? 8 : G;
  constexpr int GROUPS = kSpmmThreads / G;
  const int chunks = (nvec + G - 1) / G;
  const int64_t items = n_rows * (int64_t)chunks;
  const int64_t blocks = ceil_div64(items, GROUPS);
  if (blocks > 2147483647LL) {
    set_err("spmm: grid too large (%lld blocks)", (long long)blocks);
    return SRG_ERR_RANGE;
  }
  if (blocks == 0) return SRG_OK;
  spmm_group_kernel<VT, G, U, ACCUM><<<(unsigned)blocks, kSpmmThreads, 0, s>>>(
      indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, chunks);
  SRG_LAUNCHED();
  return SRG_OK;
}

template <typename VT, bool ACCUM>
static int dispatch_group(const int *indptr, const int *indices, const float *vals,
                          int64_t n_rows, const VT *X, int64_t ldx, VT *Y, int64_t ldy, int nvec,
                          cudaStream_t s) {
#define SRG_CASE(GG) \
  return launch_group<VT, GG, ACCUM>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s)
  if (nvec <= 1) SRG_CASE(1);
  if (nvec <= 2) SRG_CASE(2);
  if (nvec <= 4) SRG_CASE(4);
  if (nvec <= 8) SRG_CASE(8);
  if (nvec <= 16) SRG_CASE(16);
  SRG_CASE(32);
#undef SRG_CASE
}

int spmm_csr_f32_impl(const int32_t *indptr, const int32_t *indices, const float *vals,
                      int64_t n_rows, const float *X, int64_t ldx, float *Y, int64_t ldy,
                      int32_t F, bool accumulate, cudaStream_t s) {
  SRG_REQUIRE(n_rows >= 0 && F >= 0, "spmm: negative size (n_rows=%lld, F=%d)", (long long)n_rows, F);
  if (n_rows == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && vals && X && Y, "spmm: NULL pointer argument");
  SRG_REQUIRE(ldx >= F && ldy >= F, "spmm: leading dimension smaller than F (ldx=%lld ldy=%lld F=%d)",
              (long long)ldx, (long long)ldy, F);
  SRG_REQUIRE(X != Y, "spmm: in-place hop (X == Y) is not supported");
  const bool vec_ok = (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)X % 16 == 0) &&
                      ((uintptr_t)Y % 16 == 0);
  if (vec_ok) {
    const int nvec = (F + 3) / 4;
    const float4 *X4 = reinterpret_cast<const float4 *>(X);
    float4 *Y4 = reinterpret_cast<float4 *>(Y);
    return accumulate
               ? dispatch_group<float4, true>(indptr, indices, vals, n_rows, X4, ldx / 4, Y4, ldy / 4, nvec, s)
               : dispatch_group<float4, false>(indptr, indices, vals, n_rows, X4, ldx / 4, Y4, ldy / 4, nvec, s);
  }
  return accumulate ? dispatch_group<float, true>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, F, s)
                    : dispatch_group<float, false>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, F, s);
}

// ---- layout helpers ----------------------------------------------------------------------------
// one warp per row; lanes stride over the destination columns.
__global__ void __launch_bounds__(256)
pack_features_kernel(const float *__restrict__ src, long long ld_src, float *__restrict__ dst,
                     long long ld_dst, long long n, int F, const int *__restrict__ mask) {
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n) return;
  const float *s = src + row * ld_src;
  float *d = dst + row * ld_dst;
  const int *m = mask ? mask + row * (long long)F : nullptr;
  for (int c = threadIdx.x; c < ld_dst; c += 32) {
    float v = 0.f;
    if (c < F) {
      v = s[c];
      if (m) v = __fmul_rn(v, (float)m[c]);  // x * feature_mask (SSRG/data_augument.py:28)
    }
    d[c] = v;
  }
}

__global__ void __launch_bounds__(256)
unpack_features_kernel(const float *__restrict__ src, long long ld_src, float *__restrict__ dst,
                       long long ld_dst, long long n, int F) {
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n) return;
  const float *s = src + row * ld_src;
  float *d = dst + row * ld_dst;
  for (int c = threadIdx.x; c < F; c += 32) d[c] = s[c];
}

}  // namespace srg

using namespace srg;

extern "C" int srg_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                                int64_t n_rows, const float *X, int64_t ldx, float *Y, int64_t ldy,
                                int32_t F, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  return spmm_csr_f32_impl(indptr, indices, vals, n_rows, X, ldx, Y, ldy, F, false, as_stream(stream));
}

extern "C" int srg_propagate_khop_f32(const int32_t *indptr, const int32_t *indices,
                                      const float *vals, int64_t n, float *const *hops, int64_t ld,
                                      int32_t F, int32_t K, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(K >= 0, "propagate: K must be >= 0 (got %d)", K);
  SRG_REQUIRE(hops != nullptr, "propagate: hops is NULL");
  for (int k = 1; k <= K; ++k) {
    rc = spmm_csr_f32_impl(indptr, indices, vals, n, hops[k - 1], ld, hops[k], ld, F, false,
                           as_stream(stream));
    if (rc) return rc;
  }
  return SRG_OK;
}

extern "C" int srg_pack_features_f32(const float *src, int64_t ld_src, float *dst, int64_t ld_dst,
                                     int64_t n, int32_t F, const int32_t *mask, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && F >= 0 && ld_src >= F && ld_dst >= F, "pack: bad sizes");
  if (n == 0 || ld_dst == 0) return SRG_OK;
  SRG_REQUIRE(src && dst, "pack: NULL pointer");
  const int64_t blocks = ceil_div64(n, 8);
  SRG_REQUIRE(blocks <= 2147483647LL, "pack: too many rows");
  pack_features_kernel<<<(unsigned)blocks, dim3(32, 8), 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, n, F, mask);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_unpack_features_f32(const float *src, int64_t ld_src, float *dst, int64_t ld_dst,
                                       int64_t n, int32_t F, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && F >= 0 && ld_src >= F && ld_dst >= F, "unpack: bad sizes");
  if (n == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(src && dst, "unpack: NULL pointer");
  const int64_t blocks = ceil_div64(n, 8);
  SRG_REQUIRE(blocks <= 2147483647LL, "unpack: too many rows");
  unpack_features_kernel<<<(unsigned)blocks, dim3(32, 8), 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, n, F);
  SRG_LAUNCHED();
  return SRG_OK;
}
