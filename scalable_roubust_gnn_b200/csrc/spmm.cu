// spmm.cu — fp32 CSR x dense propagation hop for sm_100a (B200).
//
// Replaces SSRG/operators/csrc/matmul.c:23-40 (FloatCSRMulDenseOMP): for every row i and feature k
//   answer[i,k] = fma(data[j], mat[indices[j],k], answer[i,k])   for j = indptr[i] .. indptr[i+1]-1
// in CSR order, starting from 0.  Every output element here is the same sequential fp32 FMA chain,
// so a hop is bit-identical to the reference (rows longer than the long-row threshold are the one
// exception: they are split into fixed segments whose partial sums are added in order —
// deterministic, within fp32 rounding of the chain).
//
// The path is HBM-bound (0.5 flop/B, no reuse of X on a random graph), so there are no tensor
// cores here; the kernels are built around bytes in flight and DRAM sectors:
//
//   stream kernel (F >= 68)  one warp owns R consecutive rows and walks their neighbours as ONE
//       stream.  Per neighbour each lane issues a 16-byte cp.async (LDGSTS, L2 -> shared, no
//       register staging, 64-byte DRAM fetch granule) into a per-warp ring of 2B row slots; the
//       consumer runs one batch of B neighbours behind, applies the FMAs in CSR order and stores
//       a row of Y whenever it crosses a row end.  The stream is processed in chunks of 32
//       neighbours whose (index, value) pairs sit one per lane (coalesced 128-byte loads,
//       prefetched one chunk ahead); the chunk body is fully unrolled, so shuffle lanes, ring
//       slots and row-end tests are compile-time constants: ~13 issued instructions per gathered
//       row.  Measured 4.35 ms per hop on the products shape = 0.955 of the measured HBM peak
//       in gather-model bytes (profiles/).
//   group kernel (F < 68, unaligned layouts, the accumulate-into-answer shim)  G lanes per
//       (row, feature chunk), U register-staged gathers in flight.
//   long rows (power-law graphs)  rows longer than the threshold are cut into segments that run
//       as independent single-row tasks of the same stream kernel, then a combine kernel adds the
//       segment sums in order.  The plan (list of long rows / segments) is built on the device
//       from indptr; no host synchronisation.
//   push (multi-GPU)  the row store of the stream kernel writes into the next-hop buffer of
//       every rank over NVLink peer mappings: the per-hop all-gather is the kernel's epilogue.
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"

namespace srg {
void set_exact_sym_check(int v);   // norm.cu

// ---- tuning knobs (srg_set_tuning) -----------------------------------------------------------------
static int g_spmm_variant = 1;   // 1 = stream kernel where it applies, 0 = group kernel everywhere
static int g_stream_rows = 4;    // rows per warp task (R)
static int g_stream_batch = 4;   // B: 4 or 8
static int g_gather_l2_64 = 1;   // gathers fetch 64-byte DRAM granules instead of whole 128-byte lines
static int g_group_unroll = 4;   // U of the 32-lane group kernel (4 or 8)
static int g_push_tma = 0;       // push hop: stage the block's rows in shared memory, one TMA bulk store per peer
static int g_long_row = 1024;    // rows with more entries are split (0 = never split)
static int g_bulk_gather = -1;   // -1: bulk (TMA) row gathers for rows of <= g_bulk_auto floats; 0 never; 1 whenever a row is <= 512 B
static int g_bulk_auto = 64;     // widest row (floats, padded) the automatic choice hands to the bulk kernel
static int g_bulk_min = 32;      // ... and the narrowest (a bulk copy of < 128 bytes is not worth a TMA request)
static int g_bulk_stages = 2;    // chunks of 32 neighbour rows in flight per warp (2..4)
static int g_bulk_rows = 8;      // rows per warp task of the bulk kernel (1..32); 8 measured best (profiles/r02_hop_shard_sweep.log)
static int g_push_rows_blocks = 148;   // grid of the input-exchange kernel (one block per SM by default)
static int g_push_rows_tma = 1;        // input exchange through the TMA unit (bulk load + bulk stores) instead of st.global
static int g_push_rows_tma_blocks = 48;   // its grid (one warp per block)
static int g_bulk_tile = 0;      // 1: finished rows are staged in shared memory and leave as ONE bulk store per destination

// ---- vector abstraction for the group kernel: float4 fast path, float scalar path ------------------
template <typename VT> struct VecOps;
template <> struct VecOps<float4> {
  __device__ static __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static __forceinline__ float4 gather(const float4 *p) { return ld_gather_f4(p); }
  __device__ static __forceinline__ void fma(float a, const float4 &x, float4 &acc) {
    acc.x = fmaf(a, x.x, acc.x);
    acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z);
    acc.w = fmaf(a, x.w, acc.w);
  }
};
template <> struct VecOps<float> {
  __device__ static __forceinline__ float zero() { return 0.f; }
  __device__ static __forceinline__ float gather(const float *p) { return __ldg(p); }
  __device__ static __forceinline__ void fma(float a, const float &x, float &acc) { acc = fmaf(a, x, acc); }
};

constexpr int kSpmmThreads = 256;

// rows of <= 512 bytes can be fetched whole by the TMA unit (bulk-gather kernel below)
static inline bool use_bulk_gather(int nvec, int64_t ld_floats) {
  if (nvec > 32 || nvec < 1) return false;
  if (g_bulk_gather > 0) return true;
  return g_bulk_gather < 0 && ld_floats <= g_bulk_auto && ld_floats >= g_bulk_min;
}

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
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
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
  if (active) Y[row * ldy + col] = acc;
}

template <typename VT, int G, bool ACCUM, int U = ((G >= 4) ? 4 : G)>
static int launch_group(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                        const VT *X, int64_t ldx, VT *Y, int64_t ldy, int nvec, cudaStream_t s) {
  constexpr int GROUPS = kSpmmThreads / G;
  const int chunks = (nvec + G - 1) / G;
  const int64_t blocks = ceil_div64(n_rows * (int64_t)chunks, GROUPS);
  if (blocks > 2147483647LL) {
    set_err("spmm: grid too large (%lld blocks)", (long long)blocks);
    return SRG_ERR_RANGE;
  }
  if (blocks == 0) return SRG_OK;
  spmm_group_kernel<VT, G, U, ACCUM><<<(unsigned)blocks, kSpmmThreads, 0, s>>>(indptr, indices, vals, n_rows, X, ldx, Y,
                                                                               ldy, nvec, chunks);
  SRG_LAUNCHED();
  return SRG_OK;
}

template <typename VT, bool ACCUM>
static int dispatch_group(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                          const VT *X, int64_t ldx, VT *Y, int64_t ldy, int nvec, cudaStream_t s) {
#define SRG_CASE(GG) return launch_group<VT, GG, ACCUM>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s)
  if (nvec <= 1) SRG_CASE(1);
  if (nvec <= 2) SRG_CASE(2);
  if (nvec <= 4) SRG_CASE(4);
  if (nvec <= 8) SRG_CASE(8);
  if (nvec <= 16) SRG_CASE(16);
  if (g_group_unroll == 8) return launch_group<VT, 32, ACCUM, 8>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
  SRG_CASE(32);
#undef SRG_CASE
}

// ---- stream kernel ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16s(unsigned smem_addr, const void *g, bool l2_64) {
  if (l2_64)
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
  return v;
}

constexpr int kStreamWarps = 8;
constexpr int kMaxPeers = 9;      // 8 row blocks + the rank's own copy of the hop (the K+1 matrices of the API)

struct PeerDests {
  float4 *p[kMaxPeers];
  int count;
};

// long rows (power-law hubs): the kernel that meets a row longer than long_len registers it here (device-side plan,
// no separate pass over indptr) and leaves it to the segment tasks that run next in stream order
struct LongPlan {
  int *counts;     // [0] number of long rows, [1] number of segments
  int *row;        // [cap_rows]   row id
  int *first;      // [cap_rows]   first segment of the row
  int *nseg;       // [cap_rows]   its segment count
  int *seg_lo;     // [cap_segs]
  int *seg_hi;     // [cap_segs]
  float *partial;  // [cap_segs x ldp] segment sums
  int cap_rows, cap_segs;
  void *base;      // the single allocation behind all of the above
};

__device__ __forceinline__ void plan_add_long_row(const LongPlan &p, int row, int st, int ed, int L) {
  const int ns = (ed - st + L - 1) / L;
  const int i = atomicAdd(p.counts, 1);
  const int s0 = atomicAdd(p.counts + 1, ns);
  if (i >= p.cap_rows || s0 + ns > p.cap_segs) return;  // cannot happen: caps come from nnz / L
  p.row[i] = row;
  p.first[i] = s0;
  p.nseg[i] = ns;
  for (int k = 0; k < ns; ++k) {
    p.seg_lo[s0 + k] = st + k * L;
    p.seg_hi[s0 + k] = min(ed, st + (k + 1) * L);
  }
}

struct StreamArgs {
  const int *row_lo;         // first entry of row r   (CSR: indptr)
  const int *row_hi;         // one past the last entry (CSR: indptr + 1); rows of one task are contiguous
  const int *indices;
  const float *vals;
  long long n_rows;
  const int *n_rows_dev;     // when non-NULL the row count is min(*n_rows_dev, n_rows) (segment tasks)
  const float4 *X;
  unsigned ldx;              // in float4
  float4 *Y;
  long long ldy;             // in float4
  int nvec, chunks, R;
  int long_len;              // rows longer than this are left to the segment kernels (0: none are)
  LongPlan plan;             // where such rows are registered (long_len > 0)
  PeerDests peers;
};

template <bool PUSH>
__device__ __forceinline__ void store_row(const StreamArgs &a, float4 *yrow, const float4 &acc) {
  if (PUSH) {
    const long long off = yrow - a.Y;
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
      if (d < a.peers.count) a.peers.p[d][off] = acc;
  } else {
    *yrow = acc;
  }
}

// TMA = the push hop's bulk-store form: the warp writes its finished rows into the block's shared-memory tile
// (`tile`, kStreamWarps * R rows of ldy float4) instead of global memory; the kernel's epilogue sends the tile.
template <int B, bool L2_64, bool PUSH, bool TMA>
__device__ __forceinline__ void stream_warp_task(const StreamArgs &a, float4 *smem4, float4 *tile, const int w,
                                                 const int lane, const long long task, const long long n_rows) {
  constexpr int S = 2 * B;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned ring = (unsigned)__cvta_generic_to_shared(smem4 + (size_t)w * S * 32 + lane);

  const long long rblock = task / a.chunks;
  const int chunk = (int)(task - rblock * a.chunks);
  const long long r0 = rblock * a.R;
  if (r0 >= n_rows) return;
  const int nr = (int)min((long long)a.R, n_rows - r0);
  const int col = chunk * 32 + lane;
  const bool active = col < a.nvec;
  const float4 *Xc = a.X + col;
  float4 *yrow = TMA ? (tile + (long long)(w * a.R) * a.ldy + col) : (a.Y + r0 * a.ldy + col);
  const unsigned ldx = a.ldx;
  const long long ldy = a.ldy;

  const int my_end = (lane < nr) ? __ldg(a.row_hi + r0 + lane) : 0;
  const int e0 = __ldg(a.row_lo + r0);
  const int e1 = __shfl_sync(FULL, my_end, nr - 1);
  int prev_end = __shfl_up_sync(FULL, my_end, 1);
  if (lane == 0) prev_end = e0;
  const int my_len = my_end - prev_end;
  const bool odd_row = lane < nr && (my_len == 0 || (a.long_len > 0 && my_len > a.long_len));

  if (__any_sync(FULL, odd_row)) {
    // an empty or an over-long row in this task: rows one by one, long rows left to the segment
    // kernels (A^ has a full diagonal, so empty rows only occur for caller-supplied matrices)
    for (int r = 0; r < nr; ++r) {
      const int st = __shfl_sync(FULL, prev_end, r), ed = __shfl_sync(FULL, my_end, r);
      if (a.long_len > 0 && ed - st > a.long_len) {
        if (lane == 0 && chunk == 0) plan_add_long_row(a.plan, (int)(r0 + r), st, ed, a.long_len);
        continue;
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = st; j < ed; ++j) {
        const int c = __ldg(a.indices + j);
        const float v = __ldg(a.vals + j);
        if (active) {
          const float4 x = ld_gather_f4(Xc + (unsigned long long)(unsigned)c * ldx);
          acc.x = fmaf(v, x.x, acc.x);
          acc.y = fmaf(v, x.y, acc.y);
          acc.z = fmaf(v, x.z, acc.z);
          acc.w = fmaf(v, x.w, acc.w);
        }
      }
      if (active) store_row<PUSH && !TMA>(a, yrow + (long long)r * ldy, acc);
    }
    return;
  }

  int cbase = e0;
  int my_c = 0, nx_c = 0;
  float my_v = 0.f, nx_v = 0.f, pv_v = 0.f;
  if (cbase + lane < e1) {
    my_c = ld_stream_i32(a.indices + cbase + lane);
    my_v = ld_stream_f32(a.vals + cbase + lane);
  }
  if (cbase + 32 + lane < e1) {
    nx_c = ld_stream_i32(a.indices + cbase + 32 + lane);
    nx_v = ld_stream_f32(a.vals + cbase + 32 + lane);
  }
  int prev_cnt = 0;
  unsigned prev_mask = 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  for (;;) {
    const int cnt = max(0, min(32, e1 - cbase));
    unsigned bit = 0;
    {
      const int d = my_end - cbase - 1;
      if (lane < nr && d >= 0 && d < 32) bit = 1u << d;
    }
    const unsigned endmask = __reduce_or_sync(FULL, bit);
#pragma unroll
    for (int b = 0; b < 32 / B; ++b) {
      if (b * B < cnt) {
#pragma unroll
        for (int u = 0; u < B; ++u) {
          const int i = b * B + u;
          const int c = __shfl_sync(FULL, my_c, i);
          if (i < cnt && active)
            cp_async_16s(ring + ((i & (S - 1)) << 9), Xc + (unsigned long long)(unsigned)c * ldx, L2_64);
        }
      }
      cp_async_commit();
      cp_async_wait<1>();
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int j = (b - 1) * B + u;  // position consumed now (one batch behind the producer)
        const int jj = j & 31;          // lane / bit inside its own chunk
        const bool valid = (j < 0) ? (jj < prev_cnt) : (j < cnt);
        if (valid) {
          const float v = __shfl_sync(FULL, (j < 0) ? pv_v : my_v, jj);
          if (active) {
            const float4 x = lds_f4(ring + ((jj & (S - 1)) << 9));
            acc.x = fmaf(v, x.x, acc.x);
            acc.y = fmaf(v, x.y, acc.y);
            acc.z = fmaf(v, x.z, acc.z);
            acc.w = fmaf(v, x.w, acc.w);
          }
          if ((((j < 0) ? prev_mask : endmask) >> jj) & 1u) {
            if (active) store_row<PUSH && !TMA>(a, yrow, acc);
            yrow += ldy;
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      if (b * B >= cnt) break;  // nothing was issued in this round: the chunk is drained
    }
    if (cnt <= 32 - B) break;   // everything issued in this chunk has been consumed
    prev_cnt = cnt;
    prev_mask = endmask;
    pv_v = my_v;
    my_c = nx_c;
    my_v = nx_v;
    cbase += 32;
    if (cbase + 32 + lane < e1) {
      nx_c = ld_stream_i32(a.indices + cbase + 32 + lane);
      nx_v = ld_stream_f32(a.vals + cbase + 32 + lane);
    }
  }
  cp_async_wait<0>();
}

template <int B, bool L2_64, bool PUSH, bool TMA>
__global__ void __launch_bounds__(kStreamWarps * 32) spmm_stream_kernel(const StreamArgs a) {
  extern __shared__ float4 smem4[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *tile = smem4 + (size_t)kStreamWarps * (2 * B) * 32;   // behind the gather rings
  if (TMA) {
    // rows this kernel leaves alone (hub rows go to the segment kernels) and the pad columns travel as zeros
    const int tile_vec = kStreamWarps * a.R * (int)a.ldy;
    for (int i = threadIdx.x; i < tile_vec; i += kStreamWarps * 32) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
  }
  {
    // the main launch has one task per warp; the segment launch (row count only known on the device) runs a
    // bounded grid whose warps stride over the tasks
    const long long n_rows = a.n_rows_dev ? min((long long)*a.n_rows_dev, a.n_rows) : a.n_rows;
    const long long n_tasks = ((n_rows + a.R - 1) / a.R) * a.chunks;
    for (long long task = (long long)blockIdx.x * kStreamWarps + w; task < n_tasks; task += (long long)gridDim.x * kStreamWarps) {
      stream_warp_task<B, L2_64, PUSH, TMA>(a, smem4, tile, w, lane, task, n_rows);
      if (TMA) break;
    }
  }
  if (TMA) {
    // generic-proxy writes of the tile -> visible to the async proxy, then ONE bulk store per destination:
    // the exchange leaves the SM through the TMA unit, not through the load/store path the gathers use
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long row0 = (long long)blockIdx.x * kStreamWarps * a.R;
      const long long rows = min((long long)kStreamWarps * a.R, a.n_rows - row0);
      if (rows > 0) {
        const unsigned bytes = (unsigned)(rows * a.ldy * 16);
        const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
        for (int d = 0; d < a.peers.count; ++d) {
          float4 *dst = a.peers.p[d] + row0 * a.ldy;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }
}

// ---- bulk-gather stream kernel: neighbour rows of <= 512 bytes fetched by the TMA unit -----------------------------
// Same task shape and the same arithmetic as the stream kernel (one warp walks the neighbours of R consecutive rows
// as ONE stream and applies the FMAs in CSR order, so the result is bit-identical), but a neighbour row is moved by
// ONE 1-D bulk copy (cp.async.bulk.shared.global, SASS UBLKCP) issued by the lane that holds its index instead of
// one 16-byte cp.async per lane: the warp spends its issue slots on the FMA chain, not on addressing.  The narrow
// rows of the multi-GPU grid (F/2 = 50 floats -> 224 bytes) were bound by issued instructions per gathered row
// (~56) in the LDGSTS form; here a gathered row costs ~6 instructions to fetch and ~8 to consume.
//   per warp: S stages of 32 row slots (+ the 32 edge weights of the chunk), one mbarrier per stage; the lane that
//   owns position i of a chunk issues the copy of X[indices[i]] into slot i with complete_tx on the stage barrier;
//   the warp waits for the stage, consumes its 32 positions in order and refills the stage.
//   TILE: finished rows are staged in a shared-memory tile (R x ldy) and leave as one bulk store per destination
//   (own buffer, peers over NVLink, the caller's copy of the hop) - the exchange never touches the load/store path.
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(mb), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mb) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mb)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float lds_f1(unsigned smem_addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_addr));
  return v;
}
__device__ __forceinline__ void sts_f1(unsigned smem_addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_f4(unsigned smem_addr, const float4 &v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

static inline size_t bulk_warp_smem(int S, int nvec, int R, long long ldy, bool tile) {
  size_t b = (size_t)S * 32 * nvec * 16 + (size_t)S * 128 + (tile ? (size_t)R * ldy * 16 : 0) + (size_t)S * 8;
  return (b + 127) / 128 * 128;
}

template <int S, bool PUSH, bool TILE>
__global__ void __launch_bounds__(128) spmm_bulk_kernel(const StreamArgs a, const unsigned warp_smem) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const unsigned rowbytes = (unsigned)a.nvec * 16u;
  const unsigned slots = (unsigned)__cvta_generic_to_shared(bulk_smem) + (unsigned)w * warp_smem;
  const unsigned sv = slots + (unsigned)S * 32u * rowbytes;
  const unsigned tile = sv + (unsigned)S * 128u;
  const unsigned mb = tile + (TILE ? (unsigned)a.R * (unsigned)a.ldy * 16u : 0u);

  const long long n_rows = a.n_rows_dev ? min((long long)*a.n_rows_dev, a.n_rows) : a.n_rows;
  const long long n_tasks = (n_rows + a.R - 1) / a.R;
  // the main launch has one task per warp; the segment launch (count only known on the device) strides
  for (long long task = (long long)blockIdx.x * wpb + w; task < n_tasks; task += (long long)gridDim.x * wpb) {
  const long long r0 = task * a.R;
  const int nr = (int)min((long long)a.R, n_rows - r0);
  const bool active = lane < a.nvec;
  const float4 *Xc = a.X + lane;
  const unsigned ldx = a.ldx;
  const long long ldy = a.ldy;
  float4 *yrow = a.Y + r0 * ldy + lane;   // !TILE
  unsigned trow = tile + (unsigned)lane * 16u;   // TILE: the lane's float4 of the current tile row

  const int my_end = (lane < nr) ? __ldg(a.row_hi + r0 + lane) : 0;
  const int e0 = __ldg(a.row_lo + r0);
  const int e1 = __shfl_sync(FULL, my_end, nr - 1);
  int prev_end = __shfl_up_sync(FULL, my_end, 1);
  if (lane == 0) prev_end = e0;
  const int my_len = my_end - prev_end;
  const bool odd_row = lane < nr && (my_len == 0 || (a.long_len > 0 && my_len > a.long_len));
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = zero4;

#define SRG_BULK_ROW_DONE()                                                    \
  do {                                                                         \
    if (TILE) {                                                                \
      if (lane < (int)ldy) sts_f4(trow, active ? acc : zero4);                 \
      trow += (unsigned)ldy * 16u;                                             \
    } else {                                                                   \
      if (active) store_row<PUSH>(a, yrow, acc);                               \
      yrow += ldy;                                                             \
    }                                                                          \
    acc = zero4;                                                               \
  } while (0)

  if (__any_sync(FULL, odd_row)) {
    // an empty or an over-long row in this task: rows one by one with direct loads (A^ has a full diagonal, so
    // empty rows only occur for caller-supplied matrices; over-long rows are left to the segment kernels and
    // travel as zeros here - the combine kernel, later in stream order, writes them)
    for (int r = 0; r < nr; ++r) {
      const int st = __shfl_sync(FULL, prev_end, r), ed = __shfl_sync(FULL, my_end, r);
      const bool skip = a.long_len > 0 && ed - st > a.long_len;
      if (skip && lane == 0) plan_add_long_row(a.plan, (int)(r0 + r), st, ed, a.long_len);
      if (!skip) {
        for (int j = st; j < ed; ++j) {
          const int c = __ldg(a.indices + j);
          const float v = __ldg(a.vals + j);
          if (active) {
            const float4 x = ld_gather_f4(Xc + (unsigned long long)(unsigned)c * ldx);
            acc.x = fmaf(v, x.x, acc.x);
            acc.y = fmaf(v, x.y, acc.y);
            acc.z = fmaf(v, x.z, acc.z);
            acc.w = fmaf(v, x.w, acc.w);
          }
        }
      }
      if (TILE || !skip) {
        SRG_BULK_ROW_DONE();
      } else {
        yrow += ldy;
      }
    }
  } else {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(mb + 8u * s, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int n_chunks = (e1 - e0 + 31) >> 5;
    int nx_c = 0;
    float nx_v = 0.f;
    if (e0 + lane < e1) {
      nx_c = ld_stream_i32(a.indices + e0 + lane);
      nx_v = ld_stream_f32(a.vals + e0 + lane);
    }
    int issued = 0, ist = 0;   // chunks issued so far, stage of the next issue

#define SRG_BULK_ISSUE()                                                                                    \
  do {                                                                                                      \
    const int cb_ = e0 + (issued << 5);                                                                     \
    const int cnt_ = min(32, e1 - cb_);                                                                     \
    const int c_ = nx_c;                                                                                    \
    sts_f1(sv + (unsigned)ist * 128u + (unsigned)lane * 4u, nx_v);                                          \
    if (cb_ + 32 + lane < e1) {                                                                             \
      nx_c = ld_stream_i32(a.indices + cb_ + 32 + lane);                                                    \
      nx_v = ld_stream_f32(a.vals + cb_ + 32 + lane);                                                       \
    }                                                                                                       \
    if (lane == 0) mbar_expect_tx(mb + 8u * ist, (unsigned)cnt_ * rowbytes);                                \
    __syncwarp();                                                                                           \
    if (lane < cnt_)                                                                                        \
      bulk_g2s(slots + ((unsigned)ist * 32u + (unsigned)lane) * rowbytes, a.X + (unsigned long long)(unsigned)c_ * ldx, \
               rowbytes, mb + 8u * ist);                                                                    \
    ++issued;                                                                                               \
    if (++ist == S) ist = 0;                                                                                \
  } while (0)

    for (int k = 0; k < S && k < n_chunks; ++k) SRG_BULK_ISSUE();
    int cst = 0;
    unsigned par = 0;
    for (int k = 0; k < n_chunks; ++k) {
      const int cb = e0 + (k << 5);
      const int cnt = min(32, e1 - cb);
      unsigned bit = 0;
      {
        const int d = my_end - cb - 1;
        if (lane < nr && d >= 0 && d < 32) bit = 1u << d;
      }
      const unsigned endmask = __reduce_or_sync(FULL, bit);
      const unsigned sl = slots + (unsigned)cst * 32u * rowbytes + (unsigned)lane * 16u;
      const unsigned svp = sv + (unsigned)cst * 128u;
      mbar_wait(mb + 8u * cst, par);
      if (cnt == 32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v4 = lds_f4(svp + q * 16);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = q * 4 + u;
            const float v = (u == 0) ? v4.x : (u == 1) ? v4.y : (u == 2) ? v4.z : v4.w;
            if (active) {
              const float4 x = lds_f4(sl + (unsigned)j * rowbytes);
              acc.x = fmaf(v, x.x, acc.x);
              acc.y = fmaf(v, x.y, acc.y);
              acc.z = fmaf(v, x.z, acc.z);
              acc.w = fmaf(v, x.w, acc.w);
            }
            if ((endmask >> j) & 1u) SRG_BULK_ROW_DONE();
          }
        }
      } else {
        for (int j = 0; j < cnt; ++j) {
          const float v = lds_f1(svp + j * 4);
          if (active) {
            const float4 x = lds_f4(sl + (unsigned)j * rowbytes);
            acc.x = fmaf(v, x.x, acc.x);
            acc.y = fmaf(v, x.y, acc.y);
            acc.z = fmaf(v, x.z, acc.z);
            acc.w = fmaf(v, x.w, acc.w);
          }
          if ((endmask >> j) & 1u) SRG_BULK_ROW_DONE();
        }
      }
      if (++cst == S) {
        cst = 0;
        par ^= 1u;
      }
      __syncwarp();   // every lane has read the stage before it is refilled
      if (issued < n_chunks) SRG_BULK_ISSUE();
    }
#undef SRG_BULK_ISSUE
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb + 8u * s) : "memory");
    }
    __syncwarp();
  }
#undef SRG_BULK_ROW_DONE

  if (TILE) {
    // generic-proxy writes of the tile -> visible to the async proxy, then ONE bulk store per destination
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned bytes = (unsigned)nr * (unsigned)ldy * 16u;
    const long long off = r0 * ldy;
    bool sent = false;
    if (PUSH) {
#pragma unroll
      for (int d = 0; d < kMaxPeers; ++d)
        if (d < a.peers.count && lane == d) {
          bulk_s2g(a.peers.p[d] + off, tile, bytes);
          sent = true;
        }
    } else if (lane == 0) {
      bulk_s2g(a.Y + off, tile, bytes);
      sent = true;
    }
    if (sent) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  }  // task loop
}

template <int S, bool PUSH, bool TILE>
static int launch_bulk_t(const StreamArgs &a, int64_t max_rows, cudaStream_t s, int64_t grid_cap) {
  const size_t per_warp = bulk_warp_smem(S, a.nvec, a.R, a.ldy, TILE);
  int wpb = 4;
  while (wpb > 1 && per_warp * wpb > 113 * 1024) wpb >>= 1;   // at least two blocks per SM when a warp's share allows it
  const size_t smem = per_warp * wpb;
  if (smem > 227 * 1024) {
    set_err("spmm: bulk-gather kernel needs %zu bytes of shared memory", smem);
    return SRG_ERR_RANGE;
  }
  const int64_t tasks = ceil_div64(max_rows, a.R);
  int64_t blocks = ceil_div64(tasks, wpb);
  if (grid_cap > 0) blocks = std::min(blocks, grid_cap);
  if (blocks > 2147483647LL) {
    set_err("spmm: grid too large (%lld blocks)", (long long)blocks);
    return SRG_ERR_RANGE;
  }
  if (blocks == 0) return SRG_OK;
  SRG_CUDA(cudaFuncSetAttribute(spmm_bulk_kernel<S, PUSH, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spmm_bulk_kernel<S, PUSH, TILE><<<(unsigned)blocks, wpb * 32, smem, s>>>(a, (unsigned)per_warp);
  SRG_LAUNCHED();
  return SRG_OK;
}

template <bool PUSH>
static int launch_bulk(const StreamArgs &a, int64_t max_rows, bool tile, cudaStream_t s, int64_t grid_cap = 0) {
  const int S = g_bulk_stages < 2 ? 2 : (g_bulk_stages > 4 ? 4 : g_bulk_stages);
#define SRG_CASE(SS)                                                     \
  if (S == SS)                                                           \
    return tile ? launch_bulk_t<SS, PUSH, true>(a, max_rows, s, grid_cap) : launch_bulk_t<SS, PUSH, false>(a, max_rows, s, grid_cap)
  SRG_CASE(2);
  SRG_CASE(3);
  SRG_CASE(4);
#undef SRG_CASE
  return SRG_ERR_INVALID;
}

// ---- long rows: device-side plan, segment tasks, ordered combine ------------------------------------------
template <bool PUSH>
__global__ void __launch_bounds__(256)
combine_long_rows_kernel(LongPlan p, long long ldp4, float4 *Y, long long ldy, int nvec, PeerDests peers) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int n_long = min(p.counts[0], p.cap_rows);
  const float4 *P = reinterpret_cast<const float4 *>(p.partial);
  for (long long i = warp0; i < n_long; i += nwarps) {
    const int row = p.row[i], s0 = p.first[i], ns = p.nseg[i];
    for (int col = lane; col < nvec; col += 32) {
      float4 acc = P[(long long)s0 * ldp4 + col];
      for (int k = 1; k < ns; ++k) {
        const float4 t = P[(long long)(s0 + k) * ldp4 + col];
        acc.x = __fadd_rn(acc.x, t.x);
        acc.y = __fadd_rn(acc.y, t.y);
        acc.z = __fadd_rn(acc.z, t.z);
        acc.w = __fadd_rn(acc.w, t.w);
      }
      const long long off = (long long)row * ldy + col;
      if (PUSH) {
        for (int d = 0; d < peers.count; ++d) peers.p[d][off] = acc;
      } else {
        Y[off] = acc;
      }
    }
  }
}

static int alloc_long_plan(int64_t nnz, int L, int64_t ldp, LongPlan *p, cudaStream_t s) {
  const int64_t cap_rows = nnz / L + 1, cap_segs = 2 * (nnz / L) + 2;
  const size_t ints = 2 + 3 * (size_t)cap_rows + 2 * (size_t)cap_segs;
  const size_t int_bytes = (ints * sizeof(int) + 255) / 256 * 256;
  const size_t bytes = int_bytes + (size_t)cap_segs * ldp * sizeof(float);
  char *base = nullptr;
  SRG_CUDA(cudaMallocAsync(&base, bytes, s));
  int *ip = reinterpret_cast<int *>(base);
  p->base = base;
  p->counts = ip;
  p->row = ip + 2;
  p->first = p->row + cap_rows;
  p->nseg = p->first + cap_rows;
  p->seg_lo = p->nseg + cap_rows;
  p->seg_hi = p->seg_lo + cap_segs;
  p->partial = reinterpret_cast<float *>(base + int_bytes);
  p->cap_rows = (int)cap_rows;
  p->cap_segs = (int)cap_segs;
  SRG_CUDA(cudaMemsetAsync(p->counts, 0, 2 * sizeof(int), s));
  return SRG_OK;
}

template <int B, bool L2_64, bool PUSH, bool TMA>
static int launch_stream_t(const StreamArgs &a, int64_t blocks, cudaStream_t s) {
  size_t smem = (size_t)kStreamWarps * (2 * B) * 32 * sizeof(float4);
  if (TMA) smem += (size_t)kStreamWarps * a.R * a.ldy * sizeof(float4);
  SRG_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<B, L2_64, PUSH, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spmm_stream_kernel<B, L2_64, PUSH, TMA><<<(unsigned)blocks, kStreamWarps * 32, smem, s>>>(a);
  SRG_LAUNCHED();
  return SRG_OK;
}

constexpr int64_t kSegGridCap = 148 * 8;   // segment launches: bounded grid, warps stride over the device-side task count

template <bool PUSH>
static int launch_stream(const StreamArgs &a, int64_t max_rows, cudaStream_t s, bool tma = false, int64_t grid_cap = 0) {
  const int64_t tasks = ceil_div64(max_rows, a.R) * a.chunks;
  int64_t blocks = ceil_div64(tasks, kStreamWarps);
  if (grid_cap > 0) blocks = std::min(blocks, grid_cap);
  if (blocks > 2147483647LL) {
    set_err("spmm: grid too large (%lld blocks)", (long long)blocks);
    return SRG_ERR_RANGE;
  }
  if (blocks == 0) return SRG_OK;
  if (PUSH && tma) {
    if (g_stream_batch == 8)
      return g_gather_l2_64 ? launch_stream_t<8, true, PUSH, PUSH>(a, blocks, s) : launch_stream_t<8, false, PUSH, PUSH>(a, blocks, s);
    return g_gather_l2_64 ? launch_stream_t<4, true, PUSH, PUSH>(a, blocks, s) : launch_stream_t<4, false, PUSH, PUSH>(a, blocks, s);
  }
  if (g_stream_batch == 8)
    return g_gather_l2_64 ? launch_stream_t<8, true, PUSH, false>(a, blocks, s) : launch_stream_t<8, false, PUSH, false>(a, blocks, s);
  return g_gather_l2_64 ? launch_stream_t<4, true, PUSH, false>(a, blocks, s) : launch_stream_t<4, false, PUSH, false>(a, blocks, s);
}

// one hop through the stream kernel (+ the long-row pipeline when nnz is known)
template <bool PUSH>
static int stream_hop(const int *indptr, const int *indices, const float *vals, int64_t n_rows, int64_t nnz,
                      const float4 *X, int64_t ldx4, float4 *Y, int64_t ldy4, int nvec, const PeerDests *peers,
                      cudaStream_t s) {
  if (ldx4 > 0xffffffffLL) {
    set_err("spmm: leading dimension too large");
    return SRG_ERR_RANGE;
  }
  StreamArgs a;
  a.row_lo = indptr;
  a.row_hi = indptr + 1;
  a.indices = indices;
  a.vals = vals;
  a.n_rows = n_rows;
  a.n_rows_dev = nullptr;
  a.X = X;
  a.ldx = (unsigned)ldx4;
  a.Y = Y;
  a.ldy = ldy4;
  a.nvec = nvec;
  a.chunks = (nvec + 31) / 32;
  a.R = g_stream_rows < 1 ? 1 : (g_stream_rows > 32 ? 32 : g_stream_rows);
  a.peers.count = 0;
  for (int d = 0; d < kMaxPeers; ++d) a.peers.p[d] = nullptr;
  if (peers) a.peers = *peers;
  const int L = g_long_row;
  const bool split = L > 0 && nnz > L && n_rows > 0;  // a row longer than L needs nnz > L
  a.long_len = split ? L : 0;
  LongPlan plan;
  int rc;
  memset(&a.plan, 0, sizeof(a.plan));
  if (split) {
    const int64_t ldp = (int64_t)a.chunks * 32 * 4;  // floats per partial row (whole chunks)
    if ((rc = alloc_long_plan(nnz, L, ldp, &plan, s))) return rc;
    a.plan = plan;   // the hop kernel registers the long rows it meets
  }
  // bulk-store form of the push hop: one row chunk per task (nvec <= 32), whole rows of <= 32 float4, tile <= 32 KB
  const bool tma = PUSH && g_push_tma && a.chunks == 1 && a.ldy <= 32 && a.R <= 8 && a.peers.count > 0;
  // rows of <= 512 bytes: the TMA unit fetches whole neighbour rows (bulk-gather kernel)
  const bool bulk = use_bulk_gather(nvec, ldx4 * 4);
  if (bulk) {
    StreamArgs b = a;
    b.R = g_bulk_rows < 1 ? 1 : (g_bulk_rows > 32 ? 32 : g_bulk_rows);
    const bool tile = g_bulk_tile && b.ldy <= 32;
    rc = launch_bulk<PUSH>(b, n_rows, tile, s);
  } else {
    rc = launch_stream<PUSH>(a, n_rows, s, tma);
  }
  if (split && !rc) {
    // segments as single-row tasks into the partial buffer, then the ordered combine
    StreamArgs g = a;
    g.row_lo = plan.seg_lo;
    g.row_hi = plan.seg_hi;
    g.n_rows = plan.cap_segs;        // the device count is clamped to the capacity
    g.n_rows_dev = plan.counts + 1;
    g.Y = reinterpret_cast<float4 *>(plan.partial);
    g.ldy = (long long)a.chunks * 32;
    g.R = 1;
    g.long_len = 0;
    g.peers.count = 0;
    rc = bulk ? launch_bulk<false>(g, plan.cap_segs, false, s, kSegGridCap) : launch_stream<false>(g, plan.cap_segs, s, false, kSegGridCap);
    if (!rc) {
      const int blocks = (int)std::min<int64_t>(ceil_div64((int64_t)plan.cap_rows * 32, 256), 148 * 8);
      combine_long_rows_kernel<PUSH><<<blocks, 256, 0, s>>>(plan, (long long)a.chunks * 32, Y, ldy4, nvec, a.peers);
      SRG_LAUNCHED();
    }
  }
  if (split) cudaFreeAsync(plan.base, s);
  return rc;
}

int spmm_csr_f32_impl(const int32_t *indptr, const int32_t *indices, const float *vals,
                      int64_t n_rows, int64_t nnz, const float *X, int64_t ldx, float *Y,
                      int64_t ldy, int32_t F, bool accumulate, cudaStream_t s) {
  SRG_REQUIRE(n_rows >= 0 && F >= 0, "spmm: negative size (n_rows=%lld, F=%d)", (long long)n_rows, F);
  if (n_rows == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(indptr && X && Y, "spmm: NULL pointer argument");
  SRG_REQUIRE((indices && vals) || nnz == 0, "spmm: indices / vals are NULL but nnz says the matrix has entries");
  SRG_REQUIRE(ldx >= F && ldy >= F, "spmm: leading dimension smaller than F (ldx=%lld ldy=%lld F=%d)",
              (long long)ldx, (long long)ldy, F);
  SRG_REQUIRE(X != Y, "spmm: in-place hop (X == Y) is not supported");
  const bool vec_ok = (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0);
  if (vec_ok) {
    const int nvec = (F + 3) / 4;
    const float4 *X4 = reinterpret_cast<const float4 *>(X);
    float4 *Y4 = reinterpret_cast<float4 *>(Y);
    if (!accumulate && g_spmm_variant == 1 && (nvec > 16 || use_bulk_gather(nvec, ldx)))
      return stream_hop<false>(indptr, indices, vals, n_rows, nnz, X4, ldx / 4, Y4, ldy / 4, nvec, nullptr, s);
    return accumulate ? dispatch_group<float4, true>(indptr, indices, vals, n_rows, X4, ldx / 4, Y4, ldy / 4, nvec, s)
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

// ---- message-operator aggregation over the hop list (SSRG/operators/message_operator/*.py) ---------
// acc (n x ld_acc) is updated from one hop matrix x (n x ld_x); arithmetic mirrors the reference's
// torch expressions: `sum(list)` = sequential fp32 adds in hop order, weighted = products first
// (`feat * w`), then the same sequential sum, mean = sum / count (true division).
__global__ void __launch_bounds__(256)
agg_update_kernel(float *__restrict__ acc, long long ld_acc, int col0, const float *__restrict__ x,
                  long long ld_x, long long n, int F, int mode, float w, int first) {
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n) return;
  float *a = acc + row * ld_acc + col0;
  const float *xr = x ? x + row * ld_x : nullptr;
  for (int c = threadIdx.x; c < F; c += 32) {
    float v;
    switch (mode) {
      case SRG_AGG_SUM:
      case SRG_AGG_MEAN:
        v = first ? xr[c] : __fadd_rn(a[c], xr[c]);
        break;
      case SRG_AGG_WEIGHTED:
        v = first ? __fmul_rn(xr[c], w) : __fadd_rn(a[c], __fmul_rn(xr[c], w));
        break;
      case SRG_AGG_MAX:
        v = first ? xr[c] : fmaxf(a[c], xr[c]);
        break;
      case SRG_AGG_MIN:
        v = first ? xr[c] : fminf(a[c], xr[c]);
        break;
      case -1:  // finalise the mean: acc / count
        v = __fdiv_rn(a[c], w);
        break;
      default:  // SRG_AGG_LAST / SRG_AGG_CONCAT: plain copy into the column block
        v = xr[c];
        break;
    }
    a[c] = v;
  }
}

}  // namespace srg

using namespace srg;

extern "C" int srg_aggregate_update_f32(float *acc, int64_t ld_acc, int32_t col0, const float *x,
                                        int64_t ld_x, int64_t n, int32_t F, int32_t mode, float weight,
                                        int32_t first, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && F >= 0 && col0 >= 0, "aggregate_update: bad sizes");
  SRG_REQUIRE((mode >= SRG_AGG_LAST && mode <= SRG_AGG_WEIGHTED) || mode == -1, "aggregate_update: bad mode %d", mode);
  if (n == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(acc && (x || mode == -1), "aggregate_update: NULL pointer");
  SRG_REQUIRE(ld_acc >= col0 + F && (mode == -1 || ld_x >= F), "aggregate_update: leading dimension too small");
  const int64_t blocks = ceil_div64(n, 8);
  SRG_REQUIRE(blocks <= 2147483647LL, "aggregate_update: too many rows");
  agg_update_kernel<<<(unsigned)blocks, dim3(32, 8), 0, as_stream(stream)>>>(acc, ld_acc, col0, x, ld_x, n, F, mode,
                                                                            weight, first);
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_set_tuning(const char *key, int64_t value) {
  SRG_REQUIRE(key != nullptr, "set_tuning: NULL key");
  const std::string k(key);
  if (k == "spmm_variant") g_spmm_variant = (int)value;
  else if (k == "stream_rows") g_stream_rows = (int)value;
  else if (k == "stream_batch") g_stream_batch = (int)value;
  else if (k == "group_unroll") g_group_unroll = (int)value;
  else if (k == "gather_l2_64") g_gather_l2_64 = (int)value;
  else if (k == "long_row") g_long_row = (int)value;
  else if (k == "push_tma") g_push_tma = (int)value;
  else if (k == "exact_sym_check") set_exact_sym_check((int)value);
  else if (k == "push_rows_blocks") g_push_rows_blocks = (int)std::max<int64_t>(1, value);
  else if (k == "push_rows_tma") g_push_rows_tma = (int)value;
  else if (k == "push_rows_tma_blocks") g_push_rows_tma_blocks = (int)std::max<int64_t>(1, value);
  else if (k == "bulk_gather") g_bulk_gather = (int)value;
  else if (k == "bulk_auto") g_bulk_auto = (int)value;
  else if (k == "bulk_min") g_bulk_min = (int)value;
  else if (k == "bulk_stages") g_bulk_stages = (int)value;
  else if (k == "bulk_rows") g_bulk_rows = (int)value;
  else if (k == "bulk_tile") g_bulk_tile = (int)value;
  else {
    set_err("set_tuning: unknown key '%s'", key);
    return SRG_ERR_INVALID;
  }
  return SRG_OK;
}

extern "C" int srg_spmm_csr_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                                int64_t n_rows, int64_t nnz, const float *X, int64_t ldx, float *Y,
                                int64_t ldy, int32_t F, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  return spmm_csr_f32_impl(indptr, indices, vals, n_rows, nnz, X, ldx, Y, ldy, F, false, as_stream(stream));
}

extern "C" int srg_propagate_khop_f32(const int32_t *indptr, const int32_t *indices,
                                      const float *vals, int64_t n, int64_t nnz, float *const *hops,
                                      int64_t ld, int32_t F, int32_t K, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(K >= 0, "propagate: K must be >= 0 (got %d)", K);
  SRG_REQUIRE(hops != nullptr, "propagate: hops is NULL");
  for (int k = 1; k <= K; ++k) {
    rc = spmm_csr_f32_impl(indptr, indices, vals, n, nnz, hops[k - 1], ld, hops[k], ld, F, false, as_stream(stream));
    if (rc) return rc;
  }
  return SRG_OK;
}

extern "C" int srg_spmm_csr_f32_push(const int32_t *indptr, const int32_t *indices, const float *vals,
                                     int64_t n_rows, int64_t nnz, const float *X, int64_t ldx,
                                     float *const *dests, int32_t n_dests, int64_t dest_row0, int64_t ldy,
                                     int32_t F, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && F >= 0 && dest_row0 >= 0, "spmm_push: negative size");
  SRG_REQUIRE(n_dests >= 1 && n_dests <= kMaxPeers, "spmm_push: n_dests must be 1..%d", kMaxPeers);
  if (n_rows == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(indptr && X && dests, "spmm_push: NULL pointer");
  SRG_REQUIRE((indices && vals) || nnz == 0, "spmm_push: indices / vals are NULL but nnz says the matrix has entries");
  SRG_REQUIRE(ldx >= F && ldy >= F && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)X % 16 == 0,
              "spmm_push: needs ld %% 4 == 0 and 16-byte aligned matrices");
  const int nvec = (F + 3) / 4;
  PeerDests pd;
  pd.count = n_dests;
  for (int d = 0; d < kMaxPeers; ++d) pd.p[d] = nullptr;
  for (int d = 0; d < n_dests; ++d) {
    SRG_REQUIRE(dests[d] && (uintptr_t)dests[d] % 16 == 0, "spmm_push: dests[%d] NULL or unaligned", d);
    SRG_REQUIRE((const float *)dests[d] != X, "spmm_push: destination aliases the input");
    pd.p[d] = reinterpret_cast<float4 *>(dests[d]) + dest_row0 * (ldy / 4);
  }
  // Y is only the origin the kernel measures row offsets from
  return stream_hop<true>(indptr, indices, vals, n_rows, nnz, reinterpret_cast<const float4 *>(X), ldx / 4, pd.p[0],
                          ldy / 4, nvec, &pd, as_stream(stream));
}

// the same with one row offset per destination: the peers' full buffers take the rows at dest_row0s[d] = the
// rank's first global row, the caller's own n_rows x ldy copy of the hop (element k of the API's K+1 list) is
// simply one more destination with offset 0 - no clone of the slice after the hop
extern "C" int srg_spmm_csr_f32_push2(const int32_t *indptr, const int32_t *indices, const float *vals,
                                      int64_t n_rows, int64_t nnz, const float *X, int64_t ldx,
                                      float *const *dests, const int64_t *dest_row0s, int32_t n_dests, int64_t ldy,
                                      int32_t F, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && F >= 0, "spmm_push2: negative size");
  SRG_REQUIRE(n_dests >= 1 && n_dests <= kMaxPeers, "spmm_push2: n_dests must be 1..%d", kMaxPeers);
  if (n_rows == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(indptr && X && dests && dest_row0s, "spmm_push2: NULL pointer");
  SRG_REQUIRE((indices && vals) || nnz == 0, "spmm_push2: indices / vals are NULL but nnz says the matrix has entries");
  SRG_REQUIRE(ldx >= F && ldy >= F && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)X % 16 == 0,
              "spmm_push2: needs ld %% 4 == 0 and 16-byte aligned matrices");
  const int nvec = (F + 3) / 4;
  PeerDests pd;
  pd.count = n_dests;
  for (int d = 0; d < kMaxPeers; ++d) pd.p[d] = nullptr;
  for (int d = 0; d < n_dests; ++d) {
    SRG_REQUIRE(dests[d] && (uintptr_t)dests[d] % 16 == 0 && dest_row0s[d] >= 0, "spmm_push2: dests[%d] NULL, unaligned or negative offset", d);
    SRG_REQUIRE((const float *)dests[d] != X, "spmm_push2: destination aliases the input");
    pd.p[d] = reinterpret_cast<float4 *>(dests[d]) + dest_row0s[d] * (ldy / 4);
  }
  return stream_hop<true>(indptr, indices, vals, n_rows, nnz, reinterpret_cast<const float4 *>(X), ldx / 4, pd.p[0],
                          ldy / 4, nvec, &pd, as_stream(stream));
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

extern "C" int srg_apply_feature_mask_f32(const float *x, int64_t ld_x, const int32_t *mask, float *out,
                                          int64_t ld_out, int64_t n, int32_t F, void *stream) {
  SRG_REQUIRE(mask != nullptr, "apply_feature_mask: mask is NULL");
  return srg_pack_features_f32(x, ld_x, out, ld_out, n, F, mask, stream);
}

// copy this rank's rows into the same rows of every destination buffer (peer stores over NVLink):
// the input exchange of the multi-GPU path without a collective
// One block per SM, 8 independent 16-byte loads in flight per thread: enough bytes in flight to fill the NVLink
// egress (~2 us x 700 GB/s) while leaving most of every SM to the normalisation kernels this exchange overlaps.
__global__ void __launch_bounds__(256)
push_rows_kernel(const float4 *__restrict__ src, long long n_vec, PeerDests peers) {
  constexpr int U = 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n_vec) v[u] = ld_gather_f4(src + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n_vec) {
#pragma unroll
        for (int d = 0; d < kMaxPeers; ++d)
          if (d < peers.count) peers.p[d][i] = v[u];
      }
    }
  }
}

// The same exchange driven by the TMA unit: one warp per SM streams 32 KB chunks global -> shared (bulk load,
// mbarrier) -> every destination (one bulk store per peer, issued by different lanes).  No load/store-unit traffic at
// all: per-lane remote stores back-pressure the SM's LSU while they wait for NVLink credits and slow down every other
// kernel resident on the SM (measured: the normalisation's count pass took 0.58 ms instead of ~0.1 ms next to the
// st.global form of this kernel on 8 GPUs).
constexpr int kPushChunk = 32 * 1024;
__global__ void __launch_bounds__(32)
push_rows_tma_kernel(const float4 *__restrict__ src, long long n_bytes, PeerDests peers) {
  extern __shared__ __align__(128) unsigned char push_smem[];
  const int lane = threadIdx.x;
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(push_smem);
  const unsigned mb = sm0 + 2u * kPushChunk;
  if (lane == 0) {
    mbar_init(mb, 1u);
    mbar_init(mb + 8u, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long n_chunks = (n_bytes + kPushChunk - 1) / kPushChunk;
  const char *base = reinterpret_cast<const char *>(src);
  unsigned par[2] = {0u, 0u};
  long long c = blockIdx.x;
  if (c < n_chunks && lane == 0) {
    const unsigned bytes = (unsigned)min((long long)kPushChunk, n_bytes - c * kPushChunk);
    mbar_expect_tx(mb, bytes);
    bulk_g2s(sm0, base + c * kPushChunk, bytes, mb);
  }
  int st = 0;
  for (; c < n_chunks; c += gridDim.x) {
    const unsigned bytes = (unsigned)min((long long)kPushChunk, n_bytes - c * kPushChunk);
    mbar_wait(mb + 8u * st, par[st]);
    par[st] ^= 1u;
    // (the bulk load wrote through the async proxy and the bulk stores read through it: no proxy fence needed)
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
      if (d < peers.count && lane == d)
        bulk_s2g(reinterpret_cast<char *>(peers.p[d]) + c * kPushChunk, sm0 + (unsigned)st * kPushChunk, bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the other stage is free once the stores issued from it one iteration ago have read their source
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const long long nc = c + gridDim.x;
    if (nc < n_chunks && lane == 0) {
      const unsigned nb = (unsigned)min((long long)kPushChunk, n_bytes - nc * kPushChunk);
      mbar_expect_tx(mb + 8u * (st ^ 1), nb);
      bulk_g2s(sm0 + (unsigned)(st ^ 1) * kPushChunk, base + nc * kPushChunk, nb, mb + 8u * (st ^ 1));
    }
    st ^= 1;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

extern "C" int srg_push_rows_f32(const float *src, int64_t n_rows, int64_t ld, float *const *dests,
                                 int32_t n_dests, int64_t dest_row0, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && ld >= 0 && dest_row0 >= 0, "push_rows: negative size");
  SRG_REQUIRE(n_dests >= 1 && n_dests <= kMaxPeers, "push_rows: n_dests must be 1..%d", kMaxPeers);
  if (n_rows == 0 || ld == 0) return SRG_OK;
  SRG_REQUIRE(src && dests && ld % 4 == 0 && (uintptr_t)src % 16 == 0, "push_rows: needs ld %% 4 == 0 and 16-byte alignment");
  PeerDests pd;
  pd.count = n_dests;
  for (int d = 0; d < kMaxPeers; ++d) pd.p[d] = nullptr;
  for (int d = 0; d < n_dests; ++d) {
    SRG_REQUIRE(dests[d] && (uintptr_t)dests[d] % 16 == 0, "push_rows: dests[%d] NULL or unaligned", d);
    pd.p[d] = reinterpret_cast<float4 *>(dests[d]) + dest_row0 * (ld / 4);
  }
  const long long n_vec = n_rows * (ld / 4);
  if (g_push_rows_tma) {
    const long long n_bytes = n_vec * 16;
    // 48 warps (2 x 32 KB in flight each) already fill the NVLink egress and leave two thirds of the SMs undisturbed
    // for the normalisation this exchange overlaps: 8 GPUs, ms per step with 24 / 48 / 96 / 148 blocks = 3.98 / 3.44 /
    // 3.51 / 3.54 (profiles/r02_bench_n8_input_exchange_ab.txt)
    const int blocks = (int)std::min<int64_t>(ceil_div64(n_bytes, kPushChunk), std::min(g_push_rows_blocks, g_push_rows_tma_blocks));
    const size_t smem = 2 * kPushChunk + 16;
    SRG_CUDA(cudaFuncSetAttribute(push_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    push_rows_tma_kernel<<<blocks, 32, smem, as_stream(stream)>>>(reinterpret_cast<const float4 *>(src), n_bytes, pd);
    SRG_LAUNCHED();
    return SRG_OK;
  }
  const int blocks = (int)std::min<int64_t>(ceil_div64(n_vec, 256 * 8), g_push_rows_blocks);
  push_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(src), n_vec, pd);
  SRG_LAUNCHED();
  return SRG_OK;
}

// ---- cross-rank ordering of the push hops without a collective ------------------------------------------------------
// Every rank owns a small flag array in peer-mapped memory, one 32-bit slot per peer.  After a hop (stream order)
// thread q of this one-block kernel publishes the hop's epoch into slot `my_slot` of peer q's array with a
// system-scope release (the hop kernel's remote stores were performed before this kernel started) and then waits,
// with system-scope acquires, until peer q's epoch shows up in the local array: when the kernel ends every peer has
// finished writing this rank's next-hop buffer AND has finished reading the buffer the next hop overwrites.
// A bounded wait (2 s of %globaltimer) turns a lost peer into an error flag instead of a hung GPU.
__global__ void __launch_bounds__(32)
peer_barrier_kernel(unsigned *local_flags, PeerDests peer_flags, int my_slot, unsigned epoch, int *timeout_flag) {
  const int q = threadIdx.x;
  if (q >= peer_flags.count) return;
  unsigned *remote = reinterpret_cast<unsigned *>(peer_flags.p[q]) + my_slot;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_flags + q) : "memory");
    if ((int)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) {
      if (timeout_flag) atomicExch(timeout_flag, 1);
      break;
    }
    __nanosleep(64);
  }
}

extern "C" int srg_peer_barrier(void *local_flags, void *const *peer_flags, int32_t n_peers, int32_t my_slot,
                                uint32_t epoch, int32_t *timeout_flag, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(local_flags && peer_flags && n_peers >= 1 && n_peers <= kMaxPeers && my_slot >= 0 && my_slot < 32,
              "peer_barrier: bad arguments");
  PeerDests pd;
  pd.count = n_peers;
  for (int d = 0; d < kMaxPeers; ++d) pd.p[d] = nullptr;
  for (int d = 0; d < n_peers; ++d) {
    SRG_REQUIRE(peer_flags[d] != nullptr, "peer_barrier: peer_flags[%d] is NULL", d);
    pd.p[d] = reinterpret_cast<float4 *>(peer_flags[d]);
  }
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<unsigned *>(local_flags), pd, my_slot, epoch, timeout_flag);
  SRG_LAUNCHED();
  return SRG_OK;
}

// DMA copy between (peer-mapped) device buffers on `stream`: the exchange of the "copy" multi-GPU mode runs
// on the copy engines over NVLink, so it takes no SM, LSU or L1 bandwidth from the hop kernel it overlaps.
extern "C" int srg_copy_async(void *dst, const void *src, int64_t bytes, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(bytes >= 0, "copy_async: negative size");
  if (bytes == 0) return SRG_OK;
  SRG_REQUIRE(dst && src, "copy_async: NULL pointer");
  SRG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, as_stream(stream)));
  return SRG_OK;
}

// ---- peer-mapped buffers (CUDA IPC) for the push hop ---------------------------------------------------
extern "C" int srg_ipc_alloc(void **ptr, int64_t bytes) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(ptr && bytes > 0, "ipc_alloc: bad arguments");
  SRG_CUDA(cudaMalloc(ptr, (size_t)bytes));
  return SRG_OK;
}
extern "C" int srg_ipc_free(void *ptr) {
  if (ptr) SRG_CUDA(cudaFree(ptr));
  return SRG_OK;
}
extern "C" int srg_ipc_get_handle(void *ptr, void *handle64) {
  SRG_REQUIRE(ptr && handle64, "ipc_get_handle: NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  SRG_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return SRG_OK;
}
extern "C" int srg_ipc_open(const void *handle64, void **ptr) {
  SRG_REQUIRE(ptr && handle64, "ipc_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SRG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SRG_OK;
}
extern "C" int srg_ipc_close(void *ptr) {
  if (ptr) SRG_CUDA(cudaIpcCloseMemHandle(ptr));
  return SRG_OK;
}
