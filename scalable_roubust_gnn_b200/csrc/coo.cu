// coo.cu — edge-list / COO stages around the propagation path (integer work, bit-exact):
//
//   srg_edge_gather_i64      edge_index[:, keep]                      SSRG/data_process.py:65-66
//   srg_edges_to_sym_csr     symmetrise + unique -> CSR of ones       SSRG/data_augument.py:99-102
//   srg_csr_canonicalize     sort rows + sum duplicates               what scipy's `adj.tocoo() + eye`
//                                                                     does to non-canonical input
//                                                                     (SSRG/operators/utils.py:82)
//   srg_sym_norm_csr_general R = D^(r-1) A~^T D^(-r) when the pattern of A~ is NOT symmetric: the
//                            explicit transpose of SSRG/operators/utils.py:92
//
// All four are "sort (row, col) keys, then segment" problems.  The key sort is the one place the
// library calls a CUDA-toolkit primitive (cub::DeviceRadixSort, stable LSD radix sort); everything
// else (key construction, duplicate segmentation, row pointer search, value arithmetic) is local
// kernels.  None of this is on the per-hop path.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"
#include "sortutil.cuh"

namespace srg {

__global__ void edge_gather_kernel(const long long *__restrict__ ei, long long E,
                                   const long long *__restrict__ keep, long long Ek,
                                   long long *__restrict__ out, int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ek) return;
  const long long k = keep[i];
  if (k < 0 || k >= E) {
    atomicOr(flags, SRG_FLAG_BAD_INDEX);
    return;
  }
  out[i] = ei[k];
  out[Ek + i] = ei[E + k];
}

// keys of both directions of every edge: (u<<32|v) and (v<<32|u)
__global__ void sym_keys_kernel(const long long *__restrict__ ei, long long E, long long n,
                                uint64_t *__restrict__ keys, int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  const long long u = ei[i], v = ei[E + i];
  if (u < 0 || u >= n || v < 0 || v >= n) {
    atomicOr(flags, SRG_FLAG_BAD_INDEX);
    keys[2 * i] = keys[2 * i + 1] = 0;
    return;
  }
  keys[2 * i] = ((uint64_t)u << 32) | (uint64_t)v;
  keys[2 * i + 1] = ((uint64_t)v << 32) | (uint64_t)u;
}

__global__ void head_flags_kernel(const uint64_t *__restrict__ keys, long long m, int *__restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// compact unique keys: out_col[seg] = low 32 bits; optionally sum the values of a segment in
// stored order (first + rest, sequential)
template <bool WITH_VALS>
__global__ void compact_unique_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ head,
                                      const int *__restrict__ seg, long long m,
                                      const double *__restrict__ vals, int *__restrict__ out_col,
                                      uint64_t *__restrict__ out_key, double *__restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || !head[i]) return;
  const int sidx = seg[i];
  out_col[sidx] = (int)(keys[i] & 0xffffffffu);
  if (out_key) out_key[sidx] = keys[i];
  if (WITH_VALS) {
    double acc = vals[i];
    for (long long j = i + 1; j < m && !head[j]; ++j) acc = __dadd_rn(acc, vals[j]);
    out_val[sidx] = acc;
  }
}

// indptr[r] = first position whose row (key >> 32) is >= r, over `cnt` sorted unique keys
__global__ void row_lower_bound_kernel(const uint64_t *__restrict__ keys, const int *__restrict__ total,
                                       long long n, int *__restrict__ indptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  const int cnt = *total;
  int lo = 0, hi = cnt;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)(keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  indptr[r] = lo;
}

// (row, col, val) keys of a CSR (row-major expansion)
template <typename T>
__global__ void csr_keys_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                const T *__restrict__ data, long long n, uint64_t *__restrict__ keys,
                                double *__restrict__ vals, int *__restrict__ flags, bool transpose) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= n) return;
  const int lane = threadIdx.x & 31;
  for (int j = indptr[a] + lane; j < indptr[a + 1]; j += 32) {
    const int b = indices[j];
    if (b < 0 || b >= n) {
      atomicOr(flags, SRG_FLAG_BAD_INDEX);
      keys[j] = 0;
      vals[j] = 0.0;
      continue;
    }
    keys[j] = transpose ? (((uint64_t)b << 32) | (uint64_t)a) : (((uint64_t)a << 32) | (uint64_t)b);
    vals[j] = data ? (double)data[j] : 1.0;
  }
}

// general path values: entry (a, b) of A~ lands at R[b, a] = (A~[a,b] * dl[b]) * dr[a]
__global__ void general_norm_vals_kernel(const int *__restrict__ at_indptr, const int *__restrict__ at_indices,
                                         const double *__restrict__ at_val, const double *__restrict__ degree,
                                         long long n, const double *__restrict__ dl, const double *__restrict__ dr,
                                         uint64_t *__restrict__ keys, double *__restrict__ vals, int ones) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= n) return;
  const int lane = threadIdx.x & 31;
  const int p0 = at_indptr[a], p1 = at_indptr[a + 1];
  for (int p = p0 + lane; p < p1; p += 32) {
    const int b = at_indices[p];
    double v;
    if (ones) v = (b == (int)a) ? degree[a] - (double)(p1 - p0 - 1) : 1.0;
    else v = at_val[p];
    keys[p] = ((uint64_t)b << 32) | (uint64_t)a;
    vals[p] = __dmul_rn(__dmul_rn(v, dl[b]), dr[a]);
  }
}

__global__ void finalize_general_kernel(const uint64_t *__restrict__ keys, const double *__restrict__ vals,
                                        const int *__restrict__ total, double one_minus_alpha, double alpha,
                                        int use_ppr, int *__restrict__ out_indices, double *__restrict__ v64,
                                        float *__restrict__ v32, int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *total) return;
  const uint64_t k = keys[i];
  const int row = (int)(k >> 32), col = (int)(k & 0xffffffffu);
  double v = vals[i];
  if (use_ppr) {
    v = __dmul_rn(one_minus_alpha, v);
    if (row == col) v = __dadd_rn(v, alpha);
  }
  if (v == 0.0) atomicOr(flags, SRG_FLAG_ZERO_PRODUCT);
  out_indices[i] = col;
  if (v64) v64[i] = v;
  if (v32) v32[i] = __double2float_rn(v);
}

__global__ void set_int_kernel(int *p, const int *src) { *p = *src; }

// ---- synthetic power-law shards (R-MAT, counter-based RNG keyed by (seed, edge id)) -----
// splitmix64: the whole generator is integer arithmetic, restated in numpy by synth.rmat_scrambled_host.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// bijection of [0, 2^scale): hubs of the R-MAT recursion (small ids) are spread over the id range so
// contiguous row blocks carry comparable numbers of entries (Graph500-style vertex scrambling)
__host__ __device__ __forceinline__ uint64_t scramble_id(uint64_t x, int scale, uint64_t seed) {
  const uint64_t mask = (scale >= 64) ? ~0ull : ((1ull << scale) - 1ull);
  const int h = scale / 2 + 1;
  x = (x * 0x9E3779B97F4A7C15ull + splitmix64(seed)) & mask;
  x ^= x >> h;
  x = (x * 0xD6E8FEB86659FD93ull) & mask;
  x ^= x >> h;
  x = (x * 0xCA5A826395121157ull) & mask;
  x ^= x >> h;
  return x;
}

// edge id e -> (u, v): `scale` quadrant draws, 32 random bits each (two per hash)
__device__ __forceinline__ void rmat_edge(uint64_t seed, uint64_t e, int scale, uint64_t ta, uint64_t tab,
                                          uint64_t tabc, uint64_t *u_out, uint64_t *v_out) {
  const uint64_t base = splitmix64(seed ^ (e * 0xA24BAED4963EE407ull));
  uint64_t u = 0, v = 0, bits = 0;
  for (int l = 0; l < scale; ++l) {
    if ((l & 1) == 0) bits = splitmix64(base + (uint64_t)(l >> 1));
    const uint64_t r = (l & 1) ? (bits >> 32) : (bits & 0xffffffffull);
    const uint64_t right = (r >= ta && r < tab) || (r >= tabc);   // quadrant b or d: column bit
    const uint64_t down = (r >= tab);                             // quadrant c or d: row bit
    u = (u << 1) | down;
    v = (v << 1) | right;
  }
  *u_out = scramble_id(u, scale, seed);
  *v_out = scramble_id(v, scale, seed);
}

// every thread walks edge ids e0 + tid, + stride, ...; an accepted pair lands in the shard of each endpoint
// that falls into [row0, row1) as key ((row - row0) << 32 | col).  Emission order is irrelevant (the keys
// are sorted and made unique afterwards), so a global cursor is enough.
__global__ void __launch_bounds__(256)
rmat_shard_keys_kernel(uint64_t seed, long long m_draw, int scale, uint64_t ta, uint64_t tab, uint64_t tabc,
                       long long n, long long row0, long long row1, uint64_t *__restrict__ keys, long long cap,
                       unsigned long long *__restrict__ cursor) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m_draw; e += stride) {
    uint64_t u, v;
    rmat_edge(seed, (uint64_t)e, scale, ta, tab, tabc, &u, &v);
    if (u >= (uint64_t)n || v >= (uint64_t)n || u == v) continue;
    const bool mine_u = (long long)u >= row0 && (long long)u < row1;
    const bool mine_v = (long long)v >= row0 && (long long)v < row1;
    const int cnt = (mine_u ? 1 : 0) + (mine_v ? 1 : 0);
    if (!cnt) continue;
    const unsigned long long at = atomicAdd(cursor, (unsigned long long)cnt);
    if (at + cnt > (unsigned long long)cap) continue;   // overflow: detected by the host from the cursor
    unsigned long long w = at;
    if (mine_u) keys[w++] = ((u - (uint64_t)row0) << 32) | v;
    if (mine_v) keys[w] = ((v - (uint64_t)row0) << 32) | u;
  }
}

// X[i, c] = U[0,1) float32 from hash(seed, global row, global column): shards of any layout agree
__global__ void __launch_bounds__(256)
hash_features_kernel(uint64_t seed, long long row0, long long n_rows, int col0, int F, int f_total,
                     float *__restrict__ out, long long ld) {
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n_rows) return;
  float *o = out + row * ld;
  for (int c = threadIdx.x; c < ld; c += 32) {
    float val = 0.f;
    if (c < F) {
      const uint64_t h = splitmix64(seed ^ (((uint64_t)(row0 + row) * (uint64_t)f_total + (uint64_t)(col0 + c)) * 0x9E3779B97F4A7C15ull));
      val = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
    o[c] = val;
  }
}

// transpose keys of a float32 CSR: entry (a, b) -> key (b << 32 | a), payload = value.  Slots past
// indptr[n] (the arrays may be longer than the matrix) get a key that sorts behind every entry.
__global__ void transpose_keys_f32_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                          const float *__restrict__ vals, long long n, long long cap,
                                          uint64_t *__restrict__ keys, float *__restrict__ pay,
                                          int *__restrict__ flags) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a < n) {
    for (int j = indptr[a] + lane; j < indptr[a + 1]; j += 32) {
      const int b = indices[j];
      const bool bad = (b < 0 || b >= n);
      if (bad) atomicOr(flags, SRG_FLAG_BAD_INDEX);
      keys[j] = bad ? ((uint64_t)n << 32) : (((uint64_t)b << 32) | (uint64_t)a);
      pay[j] = vals ? vals[j] : 1.0f;
    }
  } else if (a == n) {
    for (long long j = (long long)indptr[n] + lane; j < cap; j += 32) {
      keys[j] = (uint64_t)n << 32;
      pay[j] = 0.f;
    }
  }
}

__global__ void split_keys_f32_kernel(const uint64_t *__restrict__ keys, const float *__restrict__ pay,
                                      long long n, long long cap, int *__restrict__ out_indices,
                                      float *__restrict__ out_vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  const bool live = (long long)(keys[i] >> 32) < n;  // padding / bad-index keys carry row n
  out_indices[i] = live ? (int)(keys[i] & 0xffffffffu) : 0;
  out_vals[i] = live ? pay[i] : 0.f;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_edge_gather_i64(const int64_t *edge_index, int64_t E, const int64_t *keep,
                                   int64_t E_keep, int64_t *out, int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(E >= 0 && E_keep >= 0, "edge_gather: negative size");
  if (E_keep == 0) return SRG_OK;
  SRG_REQUIRE(edge_index && keep && out && out_flags, "edge_gather: NULL pointer");
  const int64_t blocks = ceil_div64(E_keep, 256);
  SRG_REQUIRE(blocks <= 2147483647LL, "edge_gather: too many edges");
  edge_gather_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long *>(edge_index), E, reinterpret_cast<const long long *>(keep), E_keep,
      reinterpret_cast<long long *>(out), out_flags);
  SRG_LAUNCHED();
  return SRG_OK;
}

// sorted keys (m of them, possibly with duplicates) -> unique CSR.  vals may be NULL.
static int keys_to_csr(uint64_t *sorted_keys, double *sorted_vals, int64_t m, int64_t n, int32_t *out_indptr,
                       int32_t *out_indices, double *out_vals, int32_t *out_nnz_dev, cudaStream_t s) {
  int *head = nullptr;
  const int64_t ints = 2 * (m + 1) + scan_scratch_ints(m);
  SRG_CUDA(cudaMallocAsync(&head, (size_t)ints * sizeof(int), s));
  int *seg = head + (m + 1);
  int *scratch = seg + (m + 1);
  uint64_t *ukeys = nullptr;
  SRG_CUDA(cudaMallocAsync(&ukeys, (size_t)(m > 0 ? m : 1) * sizeof(uint64_t), s));
  const unsigned blocks = (unsigned)ceil_div64(m, 256);
  int rc = SRG_OK;
  if (m > 0) {
    head_flags_kernel<<<blocks, 256, 0, s>>>(sorted_keys, m, head);
    SRG_LAUNCHED();
  }
  rc = exclusive_scan_i32(head, m, seg, scratch, s);  // seg[m] = number of unique keys
  if (!rc && m > 0) {
    if (sorted_vals)
      compact_unique_kernel<true><<<blocks, 256, 0, s>>>(sorted_keys, head, seg, m, sorted_vals, out_indices, ukeys, out_vals);
    else
      compact_unique_kernel<false><<<blocks, 256, 0, s>>>(sorted_keys, head, seg, m, nullptr, out_indices, ukeys, nullptr);
    SRG_LAUNCHED();
  }
  if (!rc) {
    row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(ukeys, seg + m, n, out_indptr);
    SRG_LAUNCHED();
    if (out_nnz_dev) {
      set_int_kernel<<<1, 1, 0, s>>>(out_nnz_dev, seg + m);
      SRG_LAUNCHED();
    }
  }
  cudaFreeAsync(ukeys, s);
  cudaFreeAsync(head, s);
  return rc;
}

extern "C" int srg_edges_to_sym_csr(const int64_t *edge_index, int64_t E, int64_t n,
                                    int32_t *out_indptr, int32_t *out_indices, int32_t *out_nnz_dev,
                                    int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(E >= 0 && n >= 0, "edges_to_sym_csr: negative size");
  SRG_REQUIRE(out_indptr && out_flags, "edges_to_sym_csr: NULL pointer");
  SRG_REQUIRE(2 * E <= 2147483647LL && n <= 2147483647LL, "edges_to_sym_csr: exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  const int64_t m = 2 * E;
  uint64_t *keys = nullptr;
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(m > 0 ? 2 * m : 1) * sizeof(uint64_t), s));
  if (m > 0) {
    SRG_REQUIRE(edge_index && out_indices, "edges_to_sym_csr: NULL pointer");
    sym_keys_kernel<<<(unsigned)ceil_div64(E, 256), 256, 0, s>>>(reinterpret_cast<const long long *>(edge_index), E, n, keys, out_flags);
    SRG_LAUNCHED();
    rc = sort_keys(keys, keys + m, m, 32 + bits_for(n), s);
  }
  if (!rc) rc = keys_to_csr(keys + m, nullptr, m, n, out_indptr, out_indices, nullptr, out_nnz_dev, s);
  cudaFreeAsync(keys, s);
  return rc;
}

extern "C" int srg_csr_canonicalize(const int32_t *indptr, const int32_t *indices, const void *data,
                                    int val_dtype, int64_t n, int64_t nnz, int32_t *out_indptr,
                                    int32_t *out_indices, double *out_vals, int32_t *out_nnz_dev,
                                    int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "csr_canonicalize: negative size");
  SRG_REQUIRE(indptr && out_indptr && out_flags, "csr_canonicalize: NULL pointer");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "csr_canonicalize: bad val_dtype");
  cudaStream_t s = as_stream(stream);
  uint64_t *keys = nullptr;
  double *vals = nullptr;
  const int64_t m = nnz;
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(m > 0 ? 2 * m : 1) * sizeof(uint64_t), s));
  SRG_CUDA(cudaMallocAsync(&vals, (size_t)(m > 0 ? 2 * m : 1) * sizeof(double), s));
  if (m > 0) {
    SRG_REQUIRE(indices && out_indices && out_vals, "csr_canonicalize: NULL pointer");
    const unsigned wb = (unsigned)ceil_div64(n * 32, 256);
    if (val_dtype == SRG_VAL_F32)
      csr_keys_kernel<float><<<wb, 256, 0, s>>>(indptr, indices, static_cast<const float *>(data), n, keys, vals, out_flags, false);
    else if (val_dtype == SRG_VAL_F64)
      csr_keys_kernel<double><<<wb, 256, 0, s>>>(indptr, indices, static_cast<const double *>(data), n, keys, vals, out_flags, false);
    else
      csr_keys_kernel<double><<<wb, 256, 0, s>>>(indptr, indices, nullptr, n, keys, vals, out_flags, false);
    SRG_LAUNCHED();
    rc = sort_pairs<double>(keys, keys + m, vals, vals + m, m, 32 + bits_for(n), s);
  }
  if (!rc) rc = keys_to_csr(keys + m, vals + m, m, n, out_indptr, out_indices, out_vals, out_nnz_dev, s);
  cudaFreeAsync(vals, s);
  cudaFreeAsync(keys, s);
  return rc;
}

extern "C" int srg_synth_rmat_shard_csr(uint64_t seed, int32_t scale, int64_t m_draw, double a, double b, double c,
                                        int64_t n, int64_t row0, int64_t row1, int64_t cap, int32_t *out_indptr,
                                        int32_t *out_indices, int64_t *out_nnz, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(scale >= 1 && scale <= 31, "rmat_shard: scale must be in 1..31");
  SRG_REQUIRE(n >= 1 && n <= (1LL << scale) && row0 >= 0 && row0 <= row1 && row1 <= n, "rmat_shard: bad row range");
  SRG_REQUIRE(m_draw >= 0 && cap >= 0 && cap <= 2147483647LL, "rmat_shard: bad sizes");
  SRG_REQUIRE(a > 0 && b >= 0 && c >= 0 && a + b + c < 1.0, "rmat_shard: bad quadrant probabilities");
  SRG_REQUIRE(out_indptr && out_nnz && (cap == 0 || out_indices), "rmat_shard: NULL pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t n_loc = row1 - row0;
  const uint64_t ta = (uint64_t)(a * 4294967296.0), tab = (uint64_t)((a + b) * 4294967296.0),
                 tabc = (uint64_t)((a + b + c) * 4294967296.0);
  uint64_t *keys = nullptr;
  unsigned long long *cursor = nullptr;
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(cap > 0 ? 2 * cap : 1) * sizeof(uint64_t), s));
  SRG_CUDA(cudaMallocAsync(&cursor, sizeof(unsigned long long), s));
  SRG_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), s));
  if (m_draw > 0 && cap > 0) {
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(m_draw, 256), 148 * 16);
    rmat_shard_keys_kernel<<<blocks, 256, 0, s>>>(seed, m_draw, scale, ta, tab, tabc, n, row0, row1, keys, cap, cursor);
    SRG_LAUNCHED();
  }
  unsigned long long emitted = 0;
  SRG_CUDA(cudaMemcpyAsync(&emitted, cursor, sizeof(emitted), cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));  // set-up path: the key count sizes the sort
  cudaFreeAsync(cursor, s);
  if (emitted > (unsigned long long)cap) {
    cudaFreeAsync(keys, s);
    set_err("rmat_shard: %llu keys for rows [%lld, %lld) exceed the capacity %lld", emitted, (long long)row0,
            (long long)row1, (long long)cap);
    return SRG_ERR_RANGE;
  }
  const int64_t m = (int64_t)emitted;
  if (m > 0) rc = sort_keys(keys, keys + cap, m, 32 + bits_for(n_loc > 1 ? n_loc : 2), s);
  int32_t *nnz_dev = nullptr;
  if (!rc) {
    SRG_CUDA(cudaMallocAsync(&nnz_dev, sizeof(int32_t), s));
    rc = keys_to_csr(keys + cap, nullptr, m, n_loc, out_indptr, out_indices, nullptr, nnz_dev, s);
  }
  if (!rc) {
    int32_t h = 0;
    SRG_CUDA(cudaMemcpyAsync(&h, nnz_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));
    *out_nnz = h;
  }
  if (nnz_dev) cudaFreeAsync(nnz_dev, s);
  cudaFreeAsync(keys, s);
  return rc;
}

extern "C" int srg_synth_hash_features_f32(uint64_t seed, int64_t row0, int64_t n_rows, int32_t col0, int32_t F,
                                           int32_t f_total, float *out, int64_t ld, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && F >= 0 && col0 >= 0 && f_total >= col0 + F && ld >= F, "hash_features: bad sizes");
  if (n_rows == 0 || ld == 0) return SRG_OK;
  SRG_REQUIRE(out, "hash_features: NULL pointer");
  hash_features_kernel<<<(unsigned)ceil_div64(n_rows, 8), dim3(32, 8), 0, as_stream(stream)>>>(seed, row0, n_rows, col0, F,
                                                                                           f_total, out, ld);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_csr_transpose_f32(const int32_t *indptr, const int32_t *indices, const float *vals, int64_t n,
                                     int64_t nnz, int32_t *out_indptr, int32_t *out_indices, float *out_vals,
                                     int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "csr_transpose: negative size");
  SRG_REQUIRE(indptr && out_indptr && out_flags, "csr_transpose: NULL pointer");
  SRG_REQUIRE(nnz == 0 || (indices && out_indices && out_vals), "csr_transpose: NULL pointer");
  SRG_REQUIRE(n < 2147483647LL && nnz <= 2147483647LL, "csr_transpose: exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  const int64_t m = nnz;  // capacity of the arrays; the live entry count is indptr[n] (device side)
  uint64_t *keys = nullptr;
  float *pay = nullptr;
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(m > 0 ? 2 * m : 1) * sizeof(uint64_t), s));
  SRG_CUDA(cudaMallocAsync(&pay, (size_t)(m > 0 ? 2 * m : 1) * sizeof(float), s));
  if (m > 0) {
    transpose_keys_f32_kernel<<<(unsigned)ceil_div64((n + 1) * 32, 256), 256, 0, s>>>(indptr, indices, vals, n, m, keys,
                                                                                    pay, out_flags);
    SRG_LAUNCHED();
    // stable: equal columns keep their row order, so every transposed row comes out sorted
    rc = sort_pairs<float>(keys, keys + m, pay, pay + m, m, 32 + bits_for(n + 1), s);
  }
  if (!rc) {
    // rows of the transpose = columns of the input; padding keys (row n) sort last and are not counted
    row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(keys + m, indptr + n, n, out_indptr);
    SRG_LAUNCHED();
    if (m > 0) {
      split_keys_f32_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, s>>>(keys + m, pay + m, n, m, out_indices, out_vals);
      SRG_LAUNCHED();
    }
  }
  cudaFreeAsync(pay, s);
  cudaFreeAsync(keys, s);
  return rc;
}

namespace srg {
// defined in norm.cu: fills A~ (pattern, values), degree and the power tables
int selfloop_fill_dispatch(const int32_t *indptr, const int32_t *indices, const void *data, int val_dtype,
                           int64_t n, int64_t nnz, const int32_t *at_indptr, int32_t *at_indices, double *at_val,
                           double *degree, double *dl, double *dr, double r, const int32_t *flags,
                           cudaStream_t s);
}  // namespace srg

extern "C" int srg_sym_norm_csr_general(const int32_t *indptr, const int32_t *indices,
                                        const void *data, int val_dtype, int64_t n, int64_t nnz,
                                        const int32_t *at_indptr, double r, double ppr_alpha,
                                        int32_t *out_indptr, int32_t *out_indices,
                                        double *out_degree, double *out_val_f64, float *out_val_f32,
                                        int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "sym_norm_general: negative size");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && at_indptr && out_indptr && out_indices && out_flags, "sym_norm_general: NULL pointer");
  SRG_REQUIRE((val_dtype & 0xff) <= 2, "sym_norm_general: bad val_dtype");
  SRG_REQUIRE(nnz + n <= 2147483647LL, "sym_norm_general: nnz + n exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  const int64_t cap = nnz + n;
  // scratch: A~ indices, A~ values, degree, dl, dr, keys x2, vals x2
  int *at_indices = nullptr;
  double *dscratch = nullptr;
  uint64_t *keys = nullptr;
  SRG_CUDA(cudaMallocAsync(&at_indices, (size_t)cap * sizeof(int), s));
  SRG_CUDA(cudaMallocAsync(&dscratch, (size_t)(3 * n + 3 * cap) * sizeof(double), s));
  SRG_CUDA(cudaMallocAsync(&keys, (size_t)(2 * cap) * sizeof(uint64_t), s));
  double *deg = out_degree ? out_degree : dscratch;
  double *dl = dscratch + n, *dr = dscratch + 2 * n;
  double *at_val = dscratch + 3 * n, *vals = at_val + cap;
  rc = selfloop_fill_dispatch(indptr, indices, data, val_dtype, n, nnz, at_indptr, at_indices, at_val, deg, dl, dr, r, out_flags, s);
  if (!rc) {
    general_norm_vals_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, s>>>(
        at_indptr, at_indices, at_val, deg, n, dl, dr, keys, vals, (val_dtype & 0xff) == SRG_VAL_ONES ? 1 : 0);
    SRG_LAUNCHED();
  }
  // The number of A~ entries is only known on the device (at_indptr[n]): this directed-graph path
  // reads it back (one 4-byte synchronising copy) to size the sort.
  int h_cnt = 0;
  if (!rc) {
    SRG_CUDA(cudaMemcpyAsync(&h_cnt, at_indptr + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));  // general (directed) path only: one small readback
    rc = sort_pairs<double>(keys, keys + cap, vals, vals + cap, h_cnt, 32 + bits_for(n), s);
  }
  int *total = nullptr;
  if (!rc) {
    SRG_CUDA(cudaMallocAsync(&total, sizeof(int), s));
    SRG_CUDA(cudaMemcpyAsync(total, &h_cnt, sizeof(int), cudaMemcpyHostToDevice, s));
    row_lower_bound_kernel<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, s>>>(keys + cap, total, n, out_indptr);
    SRG_LAUNCHED();
    if (h_cnt > 0) {
      finalize_general_kernel<<<(unsigned)ceil_div64(h_cnt, 256), 256, 0, s>>>(
          keys + cap, vals + cap, total, 1.0 - ppr_alpha, ppr_alpha, ppr_alpha >= 0.0 ? 1 : 0, out_indices,
          out_val_f64, out_val_f32, out_flags);
      SRG_LAUNCHED();
    }
    SRG_CUDA(cudaStreamSynchronize(s));  // h_cnt lives on this stack frame
    cudaFreeAsync(total, s);
  }
  cudaFreeAsync(keys, s);
  cudaFreeAsync(dscratch, s);
  cudaFreeAsync(at_indices, s);
  return rc;
}
