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
#include <algorithm>

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

constexpr int kWeighted = SRG_FLAG_WEIGHTED;
// input defects found by stage 1: later stages must not trust positions / indices derived from it
constexpr int kFatal = SRG_FLAG_UNSORTED | SRG_FLAG_BAD_INDEX | SRG_FLAG_EXPLICIT_ZERO;

// OR bits into the flags word without hammering one address: only threads that would add a bit not
// yet visible issue the atomic (the race is benign, it only costs a redundant atomic).
__device__ __forceinline__ void raise_flags(int *flags, int fl) {
  if (fl && (fl & ~*reinterpret_cast<volatile int *>(flags))) atomicOr(flags, fl);
}

// ---- stage 1: row lengths of A~ for the rows [row0, row0 + n_rows) of an n_cols-column matrix --
// One warp per row (grid-stride), lanes stride over the entries: coalesced for short rows and
// 32-way parallel inside a power-law hub row.
constexpr int kNormBlocks = 148 * 8;

template <int DT>
__global__ void __launch_bounds__(256)
rows_count_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                  const void *__restrict__ data, long long n_rows, long long row0, long long n_cols,
                  int *__restrict__ rowlen, int *__restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  int fl = 0;
  for (long long a = warp0; a < n_rows; a += nwarps) {
    const int ag = (int)(a + row0);
    const int s = indptr[a], e = indptr[a + 1];
    int cnt = 0;
    bool has_diag = false;
    double diag = 0.0;
    for (int j = s + lane; j < e; j += 32) {
      const int b = indices[j];
      const int pb = (j > s) ? indices[j - 1] : -1;
      if (b <= pb) fl |= SRG_FLAG_UNSORTED;
      if (b < 0 || b >= n_cols) fl |= SRG_FLAG_BAD_INDEX;
      const double v = ValLoad<DT>::at(data, j);
      if (DT != SRG_VAL_ONES && v != 1.0) fl |= kWeighted;
      if (b == ag) {
        has_diag = true;
        diag = v;
      } else if (v != 0.0) {
        ++cnt;
      }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    const unsigned dm = __ballot_sync(0xffffffffu, has_diag);
    if (dm) diag = __shfl_sync(0xffffffffu, diag, __ffs(dm) - 1);
    if (__dadd_rn(diag, 1.0) != 0.0) ++cnt;
    if (lane == 0) rowlen[a] = cnt;
  }
  raise_flags(flags, fl);
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

// ---- stage 2a: write A~ (indices, and values when the graph is weighted) and the row degree ------
// One warp per row (grid-stride).  Pass A finds the diagonal value of A and how many kept entries
// precede the diagonal; pass B compacts the kept entries with ballot ranks, leaving the slot of the
// diagonal of A~ = A + I free for lane 0.
template <int DT>
__global__ void __launch_bounds__(256)
rows_fill_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                 const void *__restrict__ data, long long n_rows, long long row0,
                 const int *__restrict__ at_indptr, int *__restrict__ at_indices,
                 double *__restrict__ at_val, double *__restrict__ degree,
                 const int *__restrict__ flags, int force_vals) {
  const bool weighted = (DT != SRG_VAL_ONES) && (force_vals || (*flags & kWeighted));
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long a = warp0; a < n_rows; a += nwarps) {
    const int ag = (int)(a + row0);
    const int s = indptr[a], e = indptr[a + 1];
    const int p0 = at_indptr[a];
    // pass A
    int cnt_lt = 0;
    bool has_diag = false;
    double diag_a = 0.0;
    for (int j = s + lane; j < e; j += 32) {
      const int b = indices[j];
      const double v = ValLoad<DT>::at(data, j);
      if (b == ag) {
        has_diag = true;
        diag_a = v;
      } else if (b < ag && v != 0.0) {
        ++cnt_lt;
      }
    }
    cnt_lt = __reduce_add_sync(0xffffffffu, cnt_lt);
    const unsigned dm = __ballot_sync(0xffffffffu, has_diag);
    if (dm) diag_a = __shfl_sync(0xffffffffu, diag_a, __ffs(dm) - 1);
    const double diag = __dadd_rn(diag_a, 1.0);
    const int keepd = (diag != 0.0) ? 1 : 0;
    // pass B
    int base = p0;
    for (int j0 = s; j0 < e; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < e;
      const int b = valid ? indices[j] : 0;
      const double v = valid ? ValLoad<DT>::at(data, j) : 0.0;
      const bool keep = valid && b != ag && v != 0.0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = base + __popc(m & lt_mask) + ((b > ag) ? keepd : 0);
        at_indices[pos] = b;
        if (weighted) at_val[pos] = v;
      }
      base += __popc(m);
    }
    const int len = (base - p0) + keepd;
    if (lane == 0 && keepd) {
      at_indices[p0 + cnt_lt] = ag;
      if (weighted) at_val[p0 + cnt_lt] = diag;
    }
    __syncwarp();
    if (lane == 0) {
      double d;
      if (!weighted) {
        d = (len == 0) ? 0.0 : (double)(len - 1) + diag;  // entries 1.0, diagonal 1.0 or 2.0: exact
      } else if (len == 0) {
        d = 0.0;
      } else if (len == 1) {
        d = at_val[p0];
      } else {
        d = __dadd_rn(at_val[p0], np_pairwise_sum(at_val + p0 + 1, len - 1));
      }
      degree[a] = d;
    }
  }
}

__global__ void __launch_bounds__(256)
pow_tables_kernel(const double *__restrict__ degree, long long n, double e_left, double e_right,
                  double *__restrict__ dl, double *__restrict__ dr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = degree[i];
  dl[i] = pow_tab(d, e_left);
  dr[i] = pow_tab(d, e_right);
}

__device__ __forceinline__ int ld_idx_64(const int *p) {
  int v;
  asm volatile("ld.global.nc.L2::64B.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- row / segment tasks ---------------------------------------------------------------------------
// The fast kernels below run one warp per TASK.  A task is a whole row, or — for rows longer than
// kSegLen entries (power-law hubs) — one kSegLen-entry segment of it, so a hub is spread over many
// warps.  Segments are listed on the device by seg_plan_kernel (no host round trip); the kernels
// grid-stride over n_rows + *n_seg tasks.
constexpr int kSegLen = 1024;
struct SegPlan {
  int *n_seg;    // device counter
  int *seg_row;  // [cap]
  int *seg_lo;   // [cap]
  int *seg_hi;   // [cap]
  int cap;
};

__global__ void __launch_bounds__(256)
seg_plan_kernel(const int *__restrict__ indptr, long long n_rows, SegPlan p) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int st = indptr[r], ed = indptr[r + 1];
  if (ed - st <= kSegLen) return;
  const int ns = (ed - st + kSegLen - 1) / kSegLen;
  const int s0 = atomicAdd(p.n_seg, ns);
  for (int k = 0; k < ns && s0 + k < p.cap; ++k) {
    p.seg_row[s0 + k] = (int)r;
    p.seg_lo[s0 + k] = st + k * kSegLen;
    p.seg_hi[s0 + k] = min(ed, st + (k + 1) * kSegLen);
  }
}

struct Task {
  long long a;  // local row
  int s, e;     // the row's whole range
  int lo, hi;   // the part this task covers
};
// task t of n_rows + n_seg; returns false when the task is void (a long row's row-task, or t beyond the end)
__device__ __forceinline__ bool get_task(const int *__restrict__ indptr, long long n_rows, const SegPlan &p,
                                         long long t, Task &k) {
  if (t < n_rows) {
    k.a = t;
    k.s = indptr[t];
    k.e = indptr[t + 1];
    k.lo = k.s;
    k.hi = k.e;
    return k.e - k.s <= kSegLen;
  }
  const long long i = t - n_rows;
  if (i >= min(*p.n_seg, p.cap)) return false;
  k.a = p.seg_row[i];
  k.s = indptr[k.a];
  k.e = indptr[k.a + 1];
  k.lo = p.seg_lo[i];
  k.hi = p.seg_hi[i];
  return true;
}

// position of the first entry >= key in the sorted range [lo, hi)
__device__ __forceinline__ int lower_bound_idx(const int *__restrict__ idx, int lo, int hi, int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (idx[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- fast stage 1: canonical input without explicit zeros (anything else raises a flag) ----------------
template <int DT>
__global__ void __launch_bounds__(256)
rows_count_fast_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                       const void *__restrict__ data, long long n_rows, long long row0, long long n_cols,
                       SegPlan plan, int *__restrict__ rowlen, int *__restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n_tasks = n_rows + plan.cap;
  int fl = 0;
  for (long long t = warp0; t < n_tasks; t += nwarps) {
    Task k;
    if (!get_task(indptr, n_rows, plan, t, k)) {
      if (t >= n_rows && t - n_rows >= *plan.n_seg) break;
      continue;
    }
    const int ag = (int)(k.a + row0);
    for (int j = k.lo + lane; j < k.hi; j += 32) {
      const int b = indices[j];
      const int pb = (j > k.s) ? indices[j - 1] : -1;
      if (b <= pb) fl |= SRG_FLAG_UNSORTED;
      if (b < 0 || b >= n_cols) fl |= SRG_FLAG_BAD_INDEX;
      if (DT != SRG_VAL_ONES) {
        const double v = ValLoad<DT>::at(data, j);
        if (v != 1.0) fl |= (v == 0.0) ? (kWeighted | SRG_FLAG_EXPLICIT_ZERO) : kWeighted;
      }
    }
    if (k.lo == k.s && lane == 0) {  // the row's first task also owns the row-level result
      const int q = lower_bound_idx(indices, k.s, k.e, ag);
      const bool hd = q < k.e && indices[q] == ag;
      const double dv = hd ? ValLoad<DT>::at(data, q) : 0.0;
      const int keepd = (__dadd_rn(dv, 1.0) != 0.0) ? 1 : 0;
      rowlen[k.a] = (k.e - k.s) - (hd ? 1 : 0) + keepd;
    }
  }
  raise_flags(flags, fl);
}

// ---- fast stage 2a: every kept entry's output slot follows from its input slot ------------------------
template <int DT>
__global__ void __launch_bounds__(256)
rows_fill_fast_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                      const void *__restrict__ data, long long n_rows, long long row0, SegPlan plan,
                      const int *__restrict__ at_indptr, int *__restrict__ at_indices,
                      double *__restrict__ at_val, double *__restrict__ degree,
                      const int *__restrict__ flags, int force_vals) {
  if (*flags & kFatal) return;  // defective input: the slot arithmetic below would run out of bounds
  const bool weighted = (DT != SRG_VAL_ONES) && (force_vals || (*flags & kWeighted));
  if (!weighted && !force_vals) return;   // all-ones input: tile_fill_unweighted_kernel wrote A~
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n_tasks = n_rows + plan.cap;
  for (long long t = warp0; t < n_tasks; t += nwarps) {
    Task k;
    if (!get_task(indptr, n_rows, plan, t, k)) {
      if (t >= n_rows && t - n_rows >= *plan.n_seg) break;
      continue;
    }
    const int ag = (int)(k.a + row0);
    const int p0 = at_indptr[k.a];
    int q = 0;
    if (lane == 0) q = lower_bound_idx(indices, k.s, k.e, ag);
    q = __shfl_sync(0xffffffffu, q, 0);
    const bool hd = q < k.e && indices[q] == ag;
    const double diag = __dadd_rn(hd ? ValLoad<DT>::at(data, q) : 0.0, 1.0);
    const int keepd = (diag != 0.0) ? 1 : 0;
    const int shift = keepd - (hd ? 1 : 0);  // applied to entries right of the diagonal
    for (int j = k.lo + lane; j < k.hi; j += 32) {
      const int b = indices[j];
      if (b == ag) continue;
      const int pos = p0 + (j - k.s) + ((b > ag) ? shift : 0);
      at_indices[pos] = b;
      if (weighted) at_val[pos] = ValLoad<DT>::at(data, j);
    }
    if (k.lo == k.s && lane == 0) {
      const int len = (k.e - k.s) - (hd ? 1 : 0) + keepd;
      if (keepd) {
        at_indices[p0 + (q - k.s)] = ag;
        if (weighted) at_val[p0 + (q - k.s)] = diag;
      }
      if (!weighted) degree[k.a] = (len == 0) ? 0.0 : (double)(len - 1) + diag;  // 1.0 entries: exact
    }
  }
}

// ---- row-tile kernels: the fast path of canonical input ---------------------------------------------------------------
// A block takes kTileRows consecutive rows; their entries are one contiguous range of the CSR arrays, which the
// block streams with coalesced loads, ONE THREAD PER ENTRY.  The row of an entry comes from a binary search over the
// tile's row pointers in shared memory (8 steps, no global traffic), so there is no per-row overhead, no lane idles
// on short rows, and a power-law hub is simply a longer loop of the block that owns it.  Three passes:
//   tile_count   validation flags (sortedness, index range, weights, explicit zeros) + row length of A~ = A + I
//   tile_fill    unweighted input: every kept entry's slot follows from its input slot; degree = len + 1 (exact)
//   tile_values  R[a,b] = (A~[b,a] * dl[a]) * dr[b] on a symmetric A~, PPR blend, fp32 rounding; symmetry is CHECKED
//                by two 64-bit sums of a hash of (min, max[, value]) over the upper and over the lower entries
//                (equal for every symmetric matrix; a mismatch raises SRG_FLAG_ASYMMETRIC and the caller retries
//                through the general transpose path, which needs no symmetry).  The exact per-entry mirror lookup
//                of round 1 (entry_values_kernel) stays available: srg_set_tuning("exact_sym_check", 1).
constexpr int kTileRows = 256;
constexpr int kHubLen = 8192;      // rows with more entries leave the tile loop and are spread over a whole grid
constexpr int kHubBlocks = 148 * 4;
static int g_exact_sym_check = 0;
void set_exact_sym_check(int v) { g_exact_sym_check = v; }

// rows longer than kHubLen (power-law hubs): registered by the tile kernel that owns them, processed by a second
// launch whose blocks ALL stride over the hub's entries (a 10^6-entry row would otherwise be one block's loop)
struct HubList {
  int *count;   // device counter of SEGMENTS
  int *rows;    // [cap] local row id of the segment
  int *lo;      // [cap] first entry of the segment
  int *hi;      // [cap] one past its last entry
  int cap;
};

// a hub row is cut into kHubLen-entry segments; the hub kernels give one block to each segment
__device__ __forceinline__ void hub_register(const HubList &h, int row, int s, int e) {
  const int ns = (e - s + kHubLen - 1) / kHubLen;
  const int i0 = atomicAdd(h.count, ns);
  for (int k = 0; k < ns && i0 + k < h.cap; ++k) {
    h.rows[i0 + k] = row;
    h.lo[i0 + k] = s + k * kHubLen;
    h.hi[i0 + k] = min(e, s + (k + 1) * kHubLen);
  }
}

// largest t in [0, nr) with cp[t] <= c  (rows may be empty / skipped: the LAST of equal pointers owns the entry)
__device__ __forceinline__ int tile_row_of(const int *cp, int nr, int c) {
  int lo = 0, hi = nr;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cp[mid] <= c) lo = mid; else hi = mid;
  }
  return lo;
}

// Shared prologue of the tile kernels: row pointers of the tile in rp[], compacted pointers (hub rows contribute no
// entries) in cp[]; hub rows are registered when `hubs` is given.  Returns the number of compacted entries.
__device__ __forceinline__ int tile_setup(const int *__restrict__ ptr, long long r0, int nr, int *rp, int *cp,
                                          const HubList *hubs) {
  for (int t = threadIdx.x; t <= nr; t += kTileRows) rp[t] = ptr[r0 + t];
  __syncthreads();
  int len = 0;
  bool hub = false;
  if ((int)threadIdx.x < nr) {
    len = rp[threadIdx.x + 1] - rp[threadIdx.x];
    hub = len > kHubLen;
  }
  if (!__syncthreads_or(hub ? 1 : 0)) {
    // the usual tile: no hub row, the compacted pointers are the row pointers themselves
    const int total0 = rp[nr] - rp[0];
    cp[threadIdx.x] = ((int)threadIdx.x < nr) ? rp[threadIdx.x] - rp[0] : total0;
    if (threadIdx.x == 0) cp[kTileRows] = total0;
    __syncthreads();
    return total0;
  }
  if (hub) {
    if (hubs) hub_register(*hubs, (int)(r0 + threadIdx.x), rp[threadIdx.x], rp[threadIdx.x + 1]);
    len = 0;
  }
  int total;
  const int excl = block_excl_scan<kTileRows>(len, &total);
  cp[threadIdx.x] = excl;
  if (threadIdx.x == 0) cp[kTileRows] = total;
  __syncthreads();
  if ((int)threadIdx.x >= nr) cp[threadIdx.x] = total;   // rows beyond the tile: empty
  __syncthreads();
  return total;
}

// Entry-to-thread mapping of the tile loops: a thread takes kGroup CONSECUTIVE compacted entries, so the row of the
// first comes from ONE binary search and the rows of the others from a short linear advance (rows are ~25 entries
// long).  ncu showed the first version (one search per entry) issue-bound at 152 thread-instructions per entry.
constexpr int kGroup = 4;
// rows / positions of the entries c0 .. c0 + kGroup - 1 (those < total); returns how many are valid
__device__ __forceinline__ int tile_group(const int *rp, const int *cp, int nr, int total, int c0, int *tt, int *jj) {
  if (c0 >= total) return 0;
  int t = tile_row_of(cp, nr, c0);
  const int cnt = min(kGroup, total - c0);
#pragma unroll
  for (int v = 0; v < kGroup; ++v) {
    if (v < cnt) {
      const int c = c0 + v;
      while (c >= cp[t + 1]) ++t;      // cp[nr..] = total: stops inside the tile; skipped (hub / empty) rows are stepped over
      tt[v] = t;
      jj[v] = rp[t] + (c - cp[t]);
    } else {
      tt[v] = 0;
      jj[v] = 0;
    }
  }
  return cnt;
}

// ---- count ------------------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ void count_entry(int j, int b, int pb, double v, int row_start, int ag, long long n_cols,
                                            int &fl, int &diag_bits) {
  if (j > row_start && b <= pb) fl |= SRG_FLAG_UNSORTED;
  if (b < 0 || b >= n_cols) fl |= SRG_FLAG_BAD_INDEX;
  if (DT != SRG_VAL_ONES && v != 1.0) fl |= (v == 0.0) ? (kWeighted | SRG_FLAG_EXPLICIT_ZERO) : kWeighted;
  if (b == ag) diag_bits = 1 | ((__dadd_rn(v, 1.0) == 0.0) ? 2 : 0);
}

template <int DT>
__global__ void __launch_bounds__(kTileRows)
tile_count_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const void *__restrict__ data,
                  long long n_rows, long long row0, long long n_cols, int *__restrict__ rowlen,
                  int *__restrict__ flags, HubList hubs) {
  __shared__ int rp[kTileRows + 1];
  __shared__ int cp[kTileRows + 1];
  __shared__ int hdk[kTileRows];   // bit 0: the row stores its diagonal, bit 1: A~'s diagonal cancels (value -1)
  const long long r0 = (long long)blockIdx.x * kTileRows;
  const int nr = (int)min((long long)kTileRows, n_rows - r0);
  hdk[threadIdx.x] = 0;
  const int total = tile_setup(indptr, r0, nr, rp, cp, &hubs);
  int fl = 0;
  constexpr int G = 2;   // groups per thread and iteration: 2 x kGroup independent loads in flight
  for (int base = 0; base < total; base += G * kTileRows * kGroup) {
    int b[G][kGroup], tt[G][kGroup], jj[G][kGroup], pb0[G], cnt[G];
    double v[G][kGroup];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cnt[g] = tile_group(rp, cp, nr, total, base + (g * kTileRows + (int)threadIdx.x) * kGroup, tt[g], jj[g]);
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        b[g][u] = (u < cnt[g]) ? __ldg(indices + jj[g][u]) : 0;
        v[g][u] = (DT != SRG_VAL_ONES && u < cnt[g]) ? ValLoad<DT>::at(data, jj[g][u]) : 1.0;
      }
      pb0[g] = (cnt[g] > 0 && jj[g][0] > rp[tt[g][0]]) ? __ldg(indices + jj[g][0] - 1) : -1;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        if (u >= cnt[g]) break;
        // previous entry of the same row: the group's previous element unless this one starts its row
        const int pb = (u == 0) ? pb0[g] : ((tt[g][u] == tt[g][u - 1]) ? b[g][u - 1] : -1);
        int bits = 0;
        count_entry<DT>(jj[g][u], b[g][u], pb, v[g][u], rp[tt[g][u]], (int)(r0 + tt[g][u] + row0), n_cols, fl, bits);
        if (bits) hdk[tt[g][u]] = bits;
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < nr) {
    const int t = threadIdx.x, h = hdk[t];
    rowlen[r0 + t] = (rp[t + 1] - rp[t]) - (h & 1) + ((h & 2) ? 0 : 1);   // hub rows: as if no diagonal were stored
  }
  raise_flags(flags, fl);
}

template <int DT>
__global__ void __launch_bounds__(256)
hub_count_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const void *__restrict__ data,
                 long long row0, long long n_cols, int *__restrict__ rowlen, int *__restrict__ flags, HubList hubs) {
  const int nh = min(*hubs.count, hubs.cap);
  int fl = 0;
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int a = hubs.rows[h];
    const int s = indptr[a], e = indptr[a + 1];
    const int ag = (int)(a + row0);
    for (int j = hubs.lo[h] + threadIdx.x; j < hubs.hi[h]; j += blockDim.x) {
      const int b = ld_stream_i32(indices + j);
      const int pb = (j > s) ? __ldg(indices + j - 1) : -1;
      const double v = (DT != SRG_VAL_ONES) ? ValLoad<DT>::at(data, j) : 1.0;
      int bits = 0;
      count_entry<DT>((int)j, b, pb, v, s, ag, n_cols, fl, bits);
      if (bits) rowlen[a] = (e - s) - 1 + ((bits & 2) ? 0 : 1);   // the one stored diagonal fixes the row length up
    }
  }
  raise_flags(flags, fl);
}

// ---- fill (unweighted canonical input: all stored values 1.0) -----------------------------------------------------------
// A~'s row a has len + 1 - hd entries (hd: A stores its diagonal, which becomes 2.0), so hd follows from the two row
// pointers and every kept entry's slot follows from its input slot.
__device__ __forceinline__ void fill_entry(int j, int b, int pb, int s, int e, int ag, int ap_t, int hd,
                                           int *__restrict__ at_indices) {
  const int pos = ap_t + (j - s) + ((b > ag) ? 1 - hd : 0);
  at_indices[pos] = b;                                  // b == ag: the stored diagonal keeps its slot
  if (!hd) {
    // the new diagonal entry sits between the last b < ag and the first b > ag
    if (b > ag && (j == s || pb < ag)) at_indices[pos - 1] = ag;
    else if (b < ag && j == e - 1) at_indices[pos + 1] = ag;
  }
}

__global__ void __launch_bounds__(kTileRows)
tile_fill_unweighted_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, long long n_rows,
                            long long row0, const int *__restrict__ at_indptr, int *__restrict__ at_indices,
                            double *__restrict__ degree, const int *__restrict__ flags, int dt_can_be_weighted,
                            HubList hubs) {
  const int f = *flags;
  if (f & kFatal) return;                                 // defective input: nothing is written
  if (dt_can_be_weighted && (f & kWeighted)) return;      // weighted: rows_fill_fast_kernel does it
  __shared__ int rp[kTileRows + 1];
  __shared__ int cp[kTileRows + 1];
  __shared__ int ap[kTileRows + 1];
  const long long r0 = (long long)blockIdx.x * kTileRows;
  const int nr = (int)min((long long)kTileRows, n_rows - r0);
  for (int t = threadIdx.x; t <= nr; t += kTileRows) ap[t] = at_indptr[r0 + t];
  const int total = tile_setup(indptr, r0, nr, rp, cp, &hubs);
  constexpr int G = 2;
  for (int base = 0; base < total; base += G * kTileRows * kGroup) {
    int bb[G][kGroup], tt[G][kGroup], jj[G][kGroup], pb0[G], cnt[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cnt[g] = tile_group(rp, cp, nr, total, base + (g * kTileRows + (int)threadIdx.x) * kGroup, tt[g], jj[g]);
#pragma unroll
      for (int u = 0; u < kGroup; ++u) bb[g][u] = (u < cnt[g]) ? __ldg(indices + jj[g][u]) : 0;
      pb0[g] = (cnt[g] > 0 && jj[g][0] > rp[tt[g][0]]) ? __ldg(indices + jj[g][0] - 1) : -1;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        if (u >= cnt[g]) break;
        const int t = tt[g][u];
        const int s = rp[t], e = rp[t + 1];
        const int hd = ((ap[t + 1] - ap[t]) == (e - s)) ? 1 : 0;
        const int pb = (u == 0) ? pb0[g] : ((t == tt[g][u - 1]) ? bb[g][u - 1] : -1);
        fill_entry(jj[g][u], bb[g][u], pb, s, e, (int)(r0 + t + row0), ap[t], hd, at_indices);
      }
    }
  }
  if ((int)threadIdx.x < nr) {
    const int t = threadIdx.x;
    const int len = rp[t + 1] - rp[t];
    if (len == 0) at_indices[ap[t]] = (int)(r0 + t + row0);
    degree[r0 + t] = (double)(len + 1);                   // len ones + the added 1.0 (or 2.0 on a stored diagonal)
  }
}

__global__ void __launch_bounds__(256)
hub_fill_unweighted_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, long long row0,
                           const int *__restrict__ at_indptr, int *__restrict__ at_indices,
                           const int *__restrict__ flags, int dt_can_be_weighted, HubList hubs) {
  const int f = *flags;
  if ((f & kFatal) || (dt_can_be_weighted && (f & kWeighted))) return;
  const int nh = min(*hubs.count, hubs.cap);
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int a = hubs.rows[h];
    const int s = indptr[a], e = indptr[a + 1];
    const int ap_t = at_indptr[a];
    const int hd = ((at_indptr[a + 1] - ap_t) == (e - s)) ? 1 : 0;
    const int ag = (int)(a + row0);
    for (int j = hubs.lo[h] + threadIdx.x; j < hubs.hi[h]; j += blockDim.x) {
      const int b = ld_stream_i32(indices + j);
      const int pb = (j > s) ? __ldg(indices + j - 1) : -1;
      fill_entry((int)j, b, pb, s, e, ag, ap_t, hd, at_indices);
    }
  }
}

// ---- values -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

struct ValuesArgs {
  const int *at_indices;
  const double *at_val;
  const double *dr;
  double one_minus_alpha, alpha;
  int use_ppr, check_sym;
  double *val64;
  float *val32;
};

__device__ __forceinline__ void values_entry(const ValuesArgs &g, int p, int b, double vt, double dla, double drb, int ag,
                                             bool weighted, int &fl, unsigned long long &s1, unsigned long long &s2) {
  double v = __dmul_rn(__dmul_rn(vt, dla), drb);
  if (g.use_ppr) {
    v = __dmul_rn(g.one_minus_alpha, v);
    if (b == ag) v = __dadd_rn(v, g.alpha);
  }
  if (v == 0.0) fl |= SRG_FLAG_ZERO_PRODUCT;
  if (g.val64) g.val64[p] = v;
  if (g.val32) g.val32[p] = __double2float_rn(v);
  if (g.check_sym && b != ag) {
    const unsigned lo = (unsigned)min(ag, b), hi = (unsigned)max(ag, b);
    unsigned ha = hash32(lo * 0x9e3779b1u + hash32(hi));
    unsigned hb = hash32(hi * 0x85ebca6bu ^ hash32(lo + 0x27d4eb2fu));
    if (weighted) {
      const unsigned long long vb = (unsigned long long)__double_as_longlong(vt);
      ha = hash32(ha ^ (unsigned)vb);
      hb = hash32(hb ^ (unsigned)(vb >> 32));
    }
    const unsigned long long h1 = ((unsigned long long)ha << 32) | hb;
    const unsigned long long h2 = (unsigned long long)ha * (unsigned long long)(hb | 1u);
    if (b > ag) {
      s1 += h1;
      s2 += h2;
    } else {
      s1 -= h1;
      s2 -= h2;
    }
  }
}

// block-wide sum of the two symmetry accumulators into tri_counts (wrap-around arithmetic)
__device__ __forceinline__ void sym_sums_commit(unsigned long long s1, unsigned long long s2,
                                                unsigned long long *__restrict__ tri_counts) {
  __shared__ unsigned long long red[2][kTileRows / 32];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t1 = 0, t2 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      t1 += red[0][w];
      t2 += red[1][w];
    }
    if (t1) atomicAdd(tri_counts, t1);
    if (t2) atomicAdd(tri_counts + 1, t2);
  }
}

__global__ void __launch_bounds__(kTileRows)
tile_values_kernel(long long n_rows, long long row0, const int *__restrict__ at_indptr, const double *__restrict__ degree,
                   const double *__restrict__ dl, ValuesArgs g, int *__restrict__ flags,
                   unsigned long long *__restrict__ tri_counts, HubList hubs) {
  if (*flags & kFatal) return;  // A~ was not written
  const bool weighted = (*flags & kWeighted) != 0;
  __shared__ int ap[kTileRows + 1];
  __shared__ int cp[kTileRows + 1];
  __shared__ double dla[kTileRows];
  __shared__ double dgv[kTileRows];   // unweighted: value of the row's diagonal entry (1.0 or 2.0)
  const long long r0 = (long long)blockIdx.x * kTileRows;
  const int nr = (int)min((long long)kTileRows, n_rows - r0);
  const int total = tile_setup(at_indptr, r0, nr, ap, cp, &hubs);
  if ((int)threadIdx.x < nr) {
    const int t = threadIdx.x;
    dla[t] = dl[r0 + t + row0];
    dgv[t] = degree[r0 + t] - (double)(ap[t + 1] - ap[t] - 1);   // exact: 1.0 or 2.0
  }
  __syncthreads();
  int fl = 0;
  unsigned long long s1 = 0, s2 = 0;   // upper - lower, two independent sums (wrap-around arithmetic)
  // one entry per thread and iteration here (not the 4-consecutive grouping of count / fill): this pass is bound by
  // the random dr[b] gathers, i.e. by how many of them are in flight, and the grouped form's 88 registers halved the
  // resident warps (measured 683 vs 518 us at the products shape)
  constexpr int U = 4;
  for (int c0 = threadIdx.x; c0 < total; c0 += U * kTileRows) {
    int bb[U], tt[U], pp[U];
    double drb[U], atv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * kTileRows;
      tt[u] = (c < total) ? tile_row_of(cp, nr, c) : 0;
      pp[u] = ap[tt[u]] + (c - cp[tt[u]]);
      bb[u] = (c < total) ? ld_stream_i32(g.at_indices + pp[u]) : 0;
      atv[u] = (weighted && c < total) ? g.at_val[pp[u]] : 1.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) drb[u] = (c0 + u * kTileRows < total) ? __ldg(g.dr + bb[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u * kTileRows >= total) break;
      const int t = tt[u];
      const int ag = (int)(r0 + t + row0);
      const double vt = weighted ? atv[u] : ((bb[u] == ag) ? dgv[t] : 1.0);
      values_entry(g, pp[u], bb[u], vt, dla[t], drb[u], ag, weighted, fl, s1, s2);
    }
  }
  if (g.check_sym) sym_sums_commit(s1, s2, tri_counts);
  raise_flags(flags, fl);
}

__global__ void __launch_bounds__(256)
hub_values_kernel(long long row0, const int *__restrict__ at_indptr, const double *__restrict__ degree,
                  const double *__restrict__ dl, ValuesArgs g, int *__restrict__ flags,
                  unsigned long long *__restrict__ tri_counts, HubList hubs) {
  if (*flags & kFatal) return;
  const bool weighted = (*flags & kWeighted) != 0;
  const int nh = min(*hubs.count, hubs.cap);
  int fl = 0;
  unsigned long long s1 = 0, s2 = 0;
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int a = hubs.rows[h];
    const int s = at_indptr[a], e = at_indptr[a + 1];
    const int ag = (int)(a + row0);
    const double dla = dl[ag];
    const double dgv = degree[a] - (double)(e - s - 1);
    for (int p = hubs.lo[h] + threadIdx.x; p < hubs.hi[h]; p += blockDim.x) {
      const int b = ld_stream_i32(g.at_indices + p);
      const double vt = weighted ? g.at_val[p] : ((b == ag) ? dgv : 1.0);
      values_entry(g, (int)p, b, vt, dla, __ldg(g.dr + b), ag, weighted, fl, s1, s2);
    }
  }
  if (g.check_sym) sym_sums_commit(s1, s2, tri_counts);
  raise_flags(flags, fl);
}

struct HubHolder {
  HubList h;
  int *base = nullptr;
};
static int make_hub_list(int64_t nnz_bound, cudaStream_t s, HubHolder *o) {
  const int64_t cap = 2 * (nnz_bound / kHubLen) + 2;   // segments: <= len / kHubLen + 1 per hub row
  SRG_CUDA(cudaMallocAsync(&o->base, (size_t)(1 + 3 * cap) * sizeof(int), s));
  o->h.count = o->base;
  o->h.rows = o->base + 1;
  o->h.lo = o->h.rows + cap;
  o->h.hi = o->h.lo + cap;
  o->h.cap = (int)cap;
  SRG_CUDA(cudaMemsetAsync(o->base, 0, sizeof(int), s));
  return SRG_OK;
}
static void free_hub_list(HubHolder *o, cudaStream_t s) {
  if (o->base) cudaFreeAsync(o->base, s);
  o->base = nullptr;
}

// weighted graphs only: degree = A~.sum(1) in numpy's add.reduceat order, one thread per row
__global__ void __launch_bounds__(256)
rows_degree_weighted_kernel(long long n_rows, const int *__restrict__ at_indptr, const double *__restrict__ at_val,
                            double *__restrict__ degree, const int *__restrict__ flags, int force) {
  if (*flags & kFatal) return;
  if (!force && !(*flags & kWeighted)) return;
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_rows) return;
  const int p0 = at_indptr[a], len = at_indptr[a + 1] - p0;
  double d = 0.0;
  if (len == 1) d = at_val[p0];
  else if (len > 1) d = __dadd_rn(at_val[p0], np_pairwise_sum(at_val + p0 + 1, len - 1));
  degree[a] = d;
}

// mirror lookup: position of `key` in the sorted range [lo, hi) of idx, or -1.
// What bounds this kernel is the number of per-thread loads (every lane hits a different line, so
// each load is its own L1 wavefront), not latency: bisection costs ~5 dependent loads.  Column ids
// of a row are roughly uniform, so the key sits near the interpolated position: ONE 16-entry window
// around it is fetched with two 256-bit loads (LDG.E.256, 32-byte aligned) and searched in
// registers; only a miss outside the window bisects the remaining side.
// win_max: largest 8-aligned window start that keeps [w0, w0+16) inside the allocation, or -1 when
// the array is too small / not 32-byte aligned (then plain bisection).
__device__ __forceinline__ void ld_idx8(const int *p, int *w) {
  asm volatile("ld.global.nc.L2::64B.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
__device__ __forceinline__ int find_sorted(const int *__restrict__ idx, int lo, int hi, int key, long long n_cols,
                                           int win_max) {
  const int len = hi - lo;
  if (len <= 0) return -1;
  int l = lo, r = hi;
  if (win_max >= 0) {
    const int g = lo + (int)(((long long)len * key) / n_cols);
    const int w0 = min(max(g - 8, 0) & ~7, win_max);
    int w[16];
    ld_idx8(idx + w0, w);
    ld_idx8(idx + w0 + 8, w + 8);
    const int f = max(w0, lo), t = min(w0 + 15, hi - 1);  // valid part of the window
    int q = -1, vf = 0, vt = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int pos = w0 + i;
      if (pos >= lo && pos < hi && w[i] == key) q = pos;
      if (pos == f) vf = w[i];
      if (pos == t) vt = w[i];
    }
    if (q >= 0) return q;
    if (f <= t) {
      if (vf > key) r = f;
      else if (vt < key) l = t + 1;
      else return -1;  // bracketed by the window and absent
    }
  }
  while (l < r) {
    const int mid = (l + r) >> 1;
    const int v = ld_idx_64(idx + mid);
    if (v == key) return mid;
    if (v < key) l = mid + 1; else r = mid;
  }
  return -1;
}

// ---- stage 2b: R[a,b] = (A~[b,a] * dl[a]) * dr[b] on a symmetric A~ -------------------------------
// For a symmetric matrix A~[b,a] = A~[a,b], so every value comes from the entry itself.  Symmetry is
// VERIFIED, not assumed (check_sym): every upper entry (b > a) must find its mirror (b,a) with the
// same value, and the number of upper and lower entries must agree; the mirror map is injective,
// so together these prove pattern and value symmetry.
// The kernel is latency-bound (entry -> row pointer of b -> window of row b), so it is ENTRY
// parallel: one thread per stored entry, every thread an independent chain, hub rows need no special
// case.  The row of each entry comes from expand_rows_kernel (one coalesced pass).
__global__ void __launch_bounds__(256)
expand_rows_kernel(const int *__restrict__ at_indptr, long long n_rows, SegPlan plan, int *__restrict__ at_rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n_tasks = n_rows + plan.cap;
  for (long long t = warp0; t < n_tasks; t += nwarps) {
    Task k;
    if (!get_task(at_indptr, n_rows, plan, t, k)) {
      if (t >= n_rows && t - n_rows >= *plan.n_seg) break;
      continue;
    }
    for (int p = k.lo + lane; p < k.hi; p += 32) at_rows[p] = (int)k.a;
  }
}

__global__ void __launch_bounds__(256, 6)
entry_values_kernel(long long n_rows, long long row0, long long n_cols, const int *__restrict__ at_indptr,
                    const int *__restrict__ at_indices, const int *__restrict__ at_rows,
                    const double *__restrict__ at_val, const double *__restrict__ degree,
                    const double *__restrict__ dl, const double *__restrict__ dr, double one_minus_alpha,
                    double alpha, int use_ppr, int check_sym, double *__restrict__ val64,
                    float *__restrict__ val32, int *__restrict__ flags,
                    unsigned long long *__restrict__ tri_counts, int win_max) {
  if (*flags & kFatal) return;  // A~ was not written
  const bool weighted = (*flags & kWeighted) != 0;
  const long long total = at_indptr[n_rows];
  const long long stride = (long long)gridDim.x * blockDim.x;
  int fl = 0;
  long long diff = 0;  // (#upper - #lower) seen by this thread
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) {
    const int a = at_rows[p];
    const int b = at_indices[p];
    const int ag = (int)(a + row0);
    const double dla = dl[ag];
    const double drb = dr[b];
    double vt;
    if (weighted) vt = at_val[p];
    else if (b == ag) vt = degree[a] - (double)(at_indptr[a + 1] - at_indptr[a] - 1);  // 1.0 or 2.0, exact
    else vt = 1.0;
    if (check_sym && b != ag) {
      if (b > ag) {
        ++diff;
        const int q = find_sorted(at_indices, at_indptr[b], at_indptr[b + 1], ag, n_cols, win_max);
        if (q < 0 || (weighted && at_val[q] != vt)) fl |= SRG_FLAG_ASYMMETRIC;
      } else {
        --diff;
      }
    }
    double v = __dmul_rn(__dmul_rn(vt, dla), drb);
    if (use_ppr) {
      v = __dmul_rn(one_minus_alpha, v);
      if (b == ag) v = __dadd_rn(v, alpha);
    }
    if (v == 0.0) fl |= SRG_FLAG_ZERO_PRODUCT;
    if (val64) val64[p] = v;
    if (val32) val32[p] = __double2float_rn(v);
  }
  if (check_sym) {
    __shared__ long long s_diff[8];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
    if (lane == 0) s_diff[threadIdx.x >> 5] = diff;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long tt = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tt += s_diff[w];
      if (tt) atomicAdd(tri_counts, (unsigned long long)tt);  // wraps: only == 0 matters
    }
  }
  raise_flags(flags, fl);
}

__global__ void tri_compare_kernel(const unsigned long long *tri_counts, int *flags) {
  // exact check: #upper != #lower; hash check: sum over the upper entries != sum over the lower entries
  if (tri_counts[0] != 0ULL || tri_counts[1] != 0ULL) atomicOr(flags, SRG_FLAG_ASYMMETRIC);
}

// defective input: hand back an EMPTY matrix (all row pointers 0) so that a caller who launches
// hops before looking at the flags gathers nothing instead of chasing uninitialised indices
__global__ void __launch_bounds__(256)
void_on_fatal_kernel(const int *__restrict__ flags, int *indptr, long long n_plus_1) {
  if (!(*flags & kFatal)) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_plus_1) indptr[i] = 0;
}

#define SRG_DT_SWITCH(dt, CALL)                       \
  switch (dt) {                                       \
    case SRG_VAL_ONES: { constexpr int DTT = SRG_VAL_ONES; CALL; } break; \
    case SRG_VAL_F32: { constexpr int DTT = SRG_VAL_F32; CALL; } break;   \
    default: { constexpr int DTT = SRG_VAL_F64; CALL; } break;            \
  }

struct PlanHolder {
  SegPlan p;
  int *base = nullptr;
};
// device-side list of the segments of rows longer than kSegLen; nnz_bound >= indptr[n_rows]
static int make_plan(const int *indptr, int64_t n_rows, int64_t nnz_bound, cudaStream_t s, PlanHolder *h) {
  const int64_t cap = 2 * (nnz_bound / kSegLen) + 2;
  SRG_CUDA(cudaMallocAsync(&h->base, (size_t)(1 + 3 * cap) * sizeof(int), s));
  h->p.n_seg = h->base;
  h->p.seg_row = h->base + 1;
  h->p.seg_lo = h->p.seg_row + cap;
  h->p.seg_hi = h->p.seg_lo + cap;
  h->p.cap = (int)cap;
  SRG_CUDA(cudaMemsetAsync(h->base, 0, sizeof(int), s));
  seg_plan_kernel<<<(unsigned)ceil_div64(n_rows, 256), 256, 0, s>>>(indptr, n_rows, h->p);
  SRG_LAUNCHED();
  return SRG_OK;
}
static void free_plan(PlanHolder *h, cudaStream_t s) {
  if (h->base) cudaFreeAsync(h->base, s);
  h->base = nullptr;
}

static inline unsigned norm_grid(int64_t n_rows) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_rows * 32, 256), kNormBlocks));
}

// dt may carry SRG_VAL_HAS_ZEROS: the compacting (slower) kernels that tolerate explicit zeros
int rows_count_launch(const int32_t *indptr, const int32_t *indices, const void *data, int dt, int64_t n_rows,
                      int64_t nnz, int64_t row0, int64_t n_cols, int32_t *rowlen, int32_t *flags, cudaStream_t s) {
  const bool zeros = (dt & SRG_VAL_HAS_ZEROS) != 0;
  dt &= 0xff;
  if (zeros) {
    SRG_DT_SWITCH(dt, (rows_count_kernel<DTT><<<norm_grid(n_rows), 256, 0, s>>>(indptr, indices, data, n_rows, row0, n_cols, rowlen, flags)));
    SRG_LAUNCHED();
    return SRG_OK;
  }
  HubHolder hubs;
  int rc = make_hub_list(nnz, s, &hubs);
  if (rc) return rc;
  SRG_DT_SWITCH(dt, (tile_count_kernel<DTT><<<(unsigned)ceil_div64(n_rows, kTileRows), kTileRows, 0, s>>>(indptr, indices, data, n_rows, row0, n_cols, rowlen, flags, hubs.h)));
  SRG_LAUNCHED();
  if (nnz > kHubLen) {   // a hub row needs more than kHubLen entries
    SRG_DT_SWITCH(dt, (hub_count_kernel<DTT><<<kHubBlocks, 256, 0, s>>>(indptr, indices, data, row0, n_cols, rowlen, flags, hubs.h)));
    SRG_LAUNCHED();
  }
  free_hub_list(&hubs, s);
  return SRG_OK;
}

int rows_fill_launch(const int32_t *indptr, const int32_t *indices, const void *data, int dt, int64_t n_rows,
                     int64_t nnz, int64_t row0, const int32_t *at_indptr, int32_t *at_indices, double *at_val,
                     double *degree, const int32_t *flags, int force_vals, cudaStream_t s) {
  const bool zeros = (dt & SRG_VAL_HAS_ZEROS) != 0;
  dt &= 0xff;
  if (zeros) {
    SRG_DT_SWITCH(dt, (rows_fill_kernel<DTT><<<norm_grid(n_rows), 256, 0, s>>>(indptr, indices, data, n_rows, row0, at_indptr, at_indices, at_val, degree, flags, force_vals)));
    SRG_LAUNCHED();
    return SRG_OK;
  }
  if (!force_vals) {
    // all-ones input (known from the dtype, or found at run time by the count pass): entry-parallel tile kernel
    HubHolder hubs;
    int rch = make_hub_list(nnz, s, &hubs);
    if (rch) return rch;
    tile_fill_unweighted_kernel<<<(unsigned)ceil_div64(n_rows, kTileRows), kTileRows, 0, s>>>(
        indptr, indices, n_rows, row0, at_indptr, at_indices, degree, flags, dt != SRG_VAL_ONES ? 1 : 0, hubs.h);
    SRG_LAUNCHED();
    if (nnz > kHubLen) {
      hub_fill_unweighted_kernel<<<kHubBlocks, 256, 0, s>>>(indptr, indices, row0, at_indptr, at_indices, flags,
                                                           dt != SRG_VAL_ONES ? 1 : 0, hubs.h);
      SRG_LAUNCHED();
    }
    free_hub_list(&hubs, s);
    if (dt == SRG_VAL_ONES) return SRG_OK;
  }
  // weighted (or forced) values: warp-per-task kernels, which leave at once when the flags say "all ones"
  PlanHolder h;
  int rc = make_plan(indptr, n_rows, nnz, s, &h);
  if (rc) return rc;
  SRG_DT_SWITCH(dt, (rows_fill_fast_kernel<DTT><<<norm_grid(n_rows), 256, 0, s>>>(indptr, indices, data, n_rows, row0, h.p, at_indptr, at_indices, at_val, degree, flags, force_vals)));
  SRG_LAUNCHED();
  free_plan(&h, s);
  if (dt != SRG_VAL_ONES) {
    rows_degree_weighted_kernel<<<(unsigned)ceil_div64(n_rows, 256), 256, 0, s>>>(n_rows, at_indptr, at_val, degree, flags, force_vals);
    SRG_LAUNCHED();
  }
  return SRG_OK;
}

// used by the general (transpose) path in coo.cu: A~ with explicit values, degree, power tables
int selfloop_fill_dispatch(const int32_t *indptr, const int32_t *indices, const void *data, int val_dtype,
                           int64_t n, int64_t nnz, const int32_t *at_indptr, int32_t *at_indices, double *at_val,
                           double *degree, double *dl, double *dr, double r, const int32_t *flags,
                           cudaStream_t s) {
  int rc = rows_fill_launch(indptr, indices, data, val_dtype, n, nnz, 0, at_indptr, at_indices, at_val, degree, flags, 1, s);
  if (rc) return rc;
  pow_tables_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(degree, n, r - 1.0, -r, dl, dr);
  SRG_LAUNCHED();
  return SRG_OK;
}

}  // namespace srg

using namespace srg;

static int check_rows_args(const char *who, const int32_t *indptr, const int32_t *indices, const void *data,
                           int val_dtype, int64_t n_rows, int64_t row0, int64_t n_cols) {
  SRG_REQUIRE(n_rows >= 0 && row0 >= 0 && n_cols >= 0 && row0 + n_rows <= n_cols,
              "%s: bad row range (n_rows=%lld row0=%lld n_cols=%lld)", who, (long long)n_rows, (long long)row0, (long long)n_cols);
  SRG_REQUIRE(n_cols <= 2147483647LL, "%s: more than 2^31-1 columns", who);
  SRG_REQUIRE((val_dtype & 0xff) >= 0 && (val_dtype & 0xff) <= 2 && (val_dtype & ~(0xff | SRG_VAL_HAS_ZEROS)) == 0,
              "%s: bad val_dtype %d", who, val_dtype);
  SRG_REQUIRE(n_rows == 0 || indptr, "%s: indptr is NULL", who);
  (void)indices; (void)data;
  return SRG_OK;
}

extern "C" int srg_selfloop_rows_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                                     int val_dtype, int64_t n_rows, int64_t nnz, int64_t row0, int64_t n_cols,
                                     int32_t *out_indptr, int32_t *out_count, int32_t *out_flags,
                                     void *stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_rows_args("selfloop_rows", indptr, indices, data, val_dtype, n_rows, row0, n_cols))) return rc;
  SRG_REQUIRE(out_indptr && out_flags, "selfloop_rows: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (n_rows == 0) {
    SRG_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int), s));
    return SRG_OK;
  }
  SRG_REQUIRE(indices != nullptr, "selfloop_rows: indices is NULL");
  SRG_REQUIRE((val_dtype & 0xff) == SRG_VAL_ONES || data != nullptr, "selfloop_rows: data is NULL but val_dtype says values");
  SRG_REQUIRE(nnz >= 0, "selfloop_rows: negative nnz");
  int *scratch = nullptr;
  const int64_t scratch_ints = (out_count ? 0 : n_rows) + scan_scratch_ints(n_rows);
  SRG_CUDA(cudaMallocAsync(&scratch, scratch_ints * sizeof(int), s));
  int *rowlen = out_count ? out_count : scratch + scan_scratch_ints(n_rows);
  rc = rows_count_launch(indptr, indices, data, val_dtype, n_rows, nnz, row0, n_cols, rowlen, out_flags, s);
  if (!rc) rc = exclusive_scan_i32(rowlen, n_rows, out_indptr, scratch, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

extern "C" int srg_degree_selfloop_csr(const int32_t *indptr, const int32_t *indices,
                                       const void *data, int val_dtype, int64_t n, int64_t nnz,
                                       int32_t *out_indptr, int32_t *out_count,
                                       int32_t *out_flags, void *stream) {
  return srg_selfloop_rows_csr(indptr, indices, data, val_dtype, n, nnz, 0, n, out_indptr, out_count, out_flags, stream);
}

extern "C" int srg_selfloop_fill_rows_csr(const int32_t *indptr, const int32_t *indices,
                                          const void *data, int val_dtype, int64_t n_rows, int64_t nnz,
                                          int64_t row0, int64_t n_cols, const int32_t *at_indptr,
                                          int32_t *at_indices, double *at_val, double *out_degree,
                                          const int32_t *flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_rows_args("selfloop_fill_rows", indptr, indices, data, val_dtype, n_rows, row0, n_cols))) return rc;
  if (n_rows == 0) return SRG_OK;
  SRG_REQUIRE(indices && at_indptr && at_indices && out_degree && flags, "selfloop_fill_rows: NULL pointer");
  SRG_REQUIRE((val_dtype & 0xff) == SRG_VAL_ONES || (data && at_val), "selfloop_fill_rows: weighted input needs data and at_val");
  return rows_fill_launch(indptr, indices, data, val_dtype, n_rows, nnz, row0, at_indptr, at_indices, at_val, out_degree,
                          flags, 0, as_stream(stream));
}

extern "C" int srg_pow_tables_f64(const double *degree, int64_t n, double r, double *out_left,
                                  double *out_right, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0, "pow_tables: negative n");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(degree && out_left && out_right, "pow_tables: NULL pointer");
  pow_tables_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(degree, n, r - 1.0, -r, out_left, out_right);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_norm_values_rows_csr(int32_t *at_indptr, const int32_t *at_indices,
                                        const double *at_val, const double *degree_rows,
                                        int64_t n_rows, int64_t nnz, int64_t row0, int64_t n_cols,
                                        const double *pow_left,
                                        const double *pow_right, double ppr_alpha, int check_symmetry,
                                        double *out_val_f64, float *out_val_f32, int32_t *flags,
                                        void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && row0 >= 0 && nnz >= 0 && n_cols >= row0 + n_rows, "norm_values_rows: bad row range");
  if (n_rows == 0) return SRG_OK;
  SRG_REQUIRE(at_indptr && at_indices && degree_rows && pow_left && pow_right && flags, "norm_values_rows: NULL pointer");
  SRG_REQUIRE(!check_symmetry || row0 == 0, "norm_values_rows: the symmetry check needs every row on this device");
  cudaStream_t s = as_stream(stream);
  unsigned long long *tri = nullptr;
  if (check_symmetry) {
    SRG_CUDA(cudaMallocAsync(&tri, 2 * sizeof(unsigned long long), s));
    SRG_CUDA(cudaMemsetAsync(tri, 0, 2 * sizeof(unsigned long long), s));
  }
  if (!g_exact_sym_check) {
    ValuesArgs g;
    g.at_indices = at_indices;
    g.at_val = at_val;
    g.dr = pow_right;
    g.one_minus_alpha = 1.0 - ppr_alpha;
    g.alpha = ppr_alpha;
    g.use_ppr = ppr_alpha >= 0.0 ? 1 : 0;
    g.check_sym = check_symmetry;
    g.val64 = out_val_f64;
    g.val32 = out_val_f32;
    HubHolder hubs;
    if ((rc = make_hub_list(nnz, s, &hubs))) return rc;
    tile_values_kernel<<<(unsigned)ceil_div64(n_rows, kTileRows), kTileRows, 0, s>>>(n_rows, row0, at_indptr, degree_rows,
                                                                                  pow_left, g, flags, tri, hubs.h);
    SRG_LAUNCHED();
    if (nnz > kHubLen) {
      hub_values_kernel<<<kHubBlocks, 256, 0, s>>>(row0, at_indptr, degree_rows, pow_left, g, flags, tri, hubs.h);
      SRG_LAUNCHED();
    }
    free_hub_list(&hubs, s);
    void_on_fatal_kernel<<<(unsigned)ceil_div64(n_rows + 1, 256), 256, 0, s>>>(flags, at_indptr, n_rows + 1);
    SRG_LAUNCHED();
    if (check_symmetry) {
      tri_compare_kernel<<<1, 1, 0, s>>>(tri, flags);
      SRG_LAUNCHED();
      cudaFreeAsync(tri, s);
    }
    return SRG_OK;
  }
  // 256-bit window loads need a 32-byte aligned index array with >= 16 entries of capacity (nnz)
  const int win_max = ((uintptr_t)at_indices % 32 == 0 && nnz >= 16) ? (int)((nnz - 16) & ~7LL) : -1;
  PlanHolder h;
  if ((rc = make_plan(at_indptr, n_rows, nnz, s, &h))) return rc;
  int *at_rows = nullptr;
  SRG_CUDA(cudaMallocAsync(&at_rows, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int), s));
  expand_rows_kernel<<<norm_grid(n_rows), 256, 0, s>>>(at_indptr, n_rows, h.p, at_rows);
  SRG_LAUNCHED();
  free_plan(&h, s);
  // one thread per stored entry, 4 entries per thread: grid from the capacity bound
  const int64_t eblocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(nnz, 256 * 4), 2147483647LL));
  entry_values_kernel<<<(unsigned)eblocks, 256, 0, s>>>(n_rows, row0, n_cols, at_indptr, at_indices, at_rows, at_val,
                                                        degree_rows, pow_left, pow_right, 1.0 - ppr_alpha, ppr_alpha,
                                                        ppr_alpha >= 0.0 ? 1 : 0, check_symmetry, out_val_f64,
                                                        out_val_f32, flags, tri, win_max);
  SRG_LAUNCHED();
  cudaFreeAsync(at_rows, s);
  void_on_fatal_kernel<<<(unsigned)ceil_div64(n_rows + 1, 256), 256, 0, s>>>(flags, at_indptr, n_rows + 1);
  SRG_LAUNCHED();
  if (check_symmetry) {
    tri_compare_kernel<<<1, 1, 0, s>>>(tri, flags);
    SRG_LAUNCHED();
    cudaFreeAsync(tri, s);
  }
  return SRG_OK;
}

extern "C" int srg_sym_norm_csr(const int32_t *indptr, const int32_t *indices, const void *data,
                                int val_dtype, int64_t n, int64_t nnz, int32_t *out_indptr,
                                double r, double ppr_alpha, int32_t *out_indices,
                                double *out_degree, double *out_val_f64, float *out_val_f32,
                                int32_t *out_flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "sym_norm: negative size");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && out_indptr && out_indices && out_flags, "sym_norm: NULL pointer");
  const int dt = val_dtype & 0xff;
  SRG_REQUIRE(dt >= 0 && dt <= 2, "sym_norm: bad val_dtype %d", val_dtype);
  SRG_REQUIRE(dt == SRG_VAL_ONES || data != nullptr, "sym_norm: data is NULL but val_dtype says values");
  SRG_REQUIRE(nnz + n <= 2147483647LL, "sym_norm: nnz + n exceeds the int32 CSR range");
  cudaStream_t s = as_stream(stream);
  // scratch: dl, dr (n doubles each) [+ degree] [+ A~ values when the dtype can carry weights]
  const int64_t cap = nnz + n;
  const int64_t n_doubles = 2 * n + (out_degree ? 0 : n) + (dt == SRG_VAL_ONES ? 0 : cap);
  double *scratch = nullptr;
  SRG_CUDA(cudaMallocAsync(&scratch, (size_t)n_doubles * sizeof(double), s));
  double *dl = scratch, *dr = scratch + n;
  double *deg = out_degree ? out_degree : scratch + 2 * n;
  double *at_val = (dt == SRG_VAL_ONES) ? nullptr : scratch + 2 * n + (out_degree ? 0 : n);
  rc = rows_fill_launch(indptr, indices, data, val_dtype, n, nnz, 0, out_indptr, out_indices, at_val, deg, out_flags, 0, s);
  if (!rc) rc = srg_pow_tables_f64(deg, n, r, dl, dr, s);
  if (!rc)
    rc = srg_norm_values_rows_csr(out_indptr, out_indices, at_val, deg, n, cap, 0, n, dl, dr, ppr_alpha, 1,
                                  out_val_f64, out_val_f32, out_flags, s);
  cudaFreeAsync(scratch, s);
  return rc;
}
