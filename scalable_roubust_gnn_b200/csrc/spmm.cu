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
#include <cstring>
#include <string>

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

static int g_group_unroll = 4;  // tuning: gathers in flight per lane for the 32-lane group kernel

template <typename VT, int G, bool ACCUM, int U = ((G >= 4) ? 4 : G)>
static int launch_group(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                        const VT *X, int64_t ldx, VT *Y, int64_t ldy, int nvec, cudaStream_t s) {
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
  if (g_group_unroll == 8)
    return launch_group<VT, 32, ACCUM, 8>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
  SRG_CASE(32);
#undef SRG_CASE
}


// ---- stream kernel: deep asynchronous gather pipeline through shared memory ---------------------
// One warp owns R consecutive rows and treats their neighbours as ONE stream of positions
// [indptr[r0], indptr[r0+R)).  A producer cursor issues, for every position, one 16-byte
// cp.async (LDGSTS, L2 -> shared, no register staging) per lane into a per-warp ring of S row
// slots; a consumer cursor S positions behind applies the FMAs in CSR order and flushes a row of
// Y whenever it crosses a row end.  The in-flight depth is S rows (S*16 B per lane) independent
// of the register budget, the index/value chunks are prefetched one chunk ahead, and nothing
// stalls at row boundaries, so the warp keeps the memory system busy across short rows.
// Each lane reads back only the 16 bytes it copied itself: no barrier is needed for the ring.
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// same, fetching 64-byte granules from DRAM on an L2 miss instead of the default full 128-byte
// line (measured: tools/micro/fetch_gran.cu, profiles/fetch_granularity.md)
__device__ __forceinline__ void cp_async_16_l2_64(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kStreamWarps = 8;

template <int S, int B, bool L2_64>
__global__ void __launch_bounds__(kStreamWarps * 32)
spmm_stream_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                   const float *__restrict__ vals, long long n_rows, const float4 *__restrict__ X,
                   long long ldx, float4 *__restrict__ Y, long long ldy, int nvec, int chunks,
                   int R, int stride) {
  static_assert(S % B == 0 && (S & (S - 1)) == 0, "ring must be a power of two multiple of B");
  constexpr int NB = S / B;
  extern __shared__ float4 smem4[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *ring = smem4 + (size_t)w * S * stride;  // slot = `stride` float4 (active lanes only)
  float *vring = reinterpret_cast<float *>(smem4 + (size_t)kStreamWarps * S * stride) + w * S;

  const long long task = (long long)blockIdx.x * kStreamWarps + w;
  const long long rblock = task / chunks;
  const int chunk = (int)(task - rblock * chunks);
  const long long r0 = rblock * R;
  if (r0 >= n_rows) return;
  const int nr = (int)min((long long)R, n_rows - r0);
  const int col = chunk * 32 + lane;
  const bool active = col < nvec;

  const int my_end = (lane < nr) ? __ldg(indptr + r0 + lane + 1) : 0;
  const int e0 = __ldg(indptr + r0);
  const int e1 = __shfl_sync(0xffffffffu, my_end, nr - 1);

  int p = e0, q = e0;
  // index / value chunks: lane l holds position cbase + l; the next chunk is always in flight
  int cbase = e0;
  int my_c = 0, nx_c = 0;
  float my_v = 0.f, nx_v = 0.f;
  if (cbase + lane < e1) {
    my_c = ld_stream_i32(indices + cbase + lane);
    my_v = ld_stream_f32(vals + cbase + lane);
  }
  if (cbase + 32 + lane < e1) {
    nx_c = ld_stream_i32(indices + cbase + 32 + lane);
    nx_v = ld_stream_f32(vals + cbase + 32 + lane);
  }

  const float4 *Xc = X + col;
  float4 *Yc = Y + r0 * ldy + col;
  int cr = 0;
  int cur_end = __shfl_sync(0xffffffffu, my_end, 0);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush_rows = [&]() {
    while (cr < nr && q == cur_end) {  // also walks over empty rows
      if (active) Yc[(long long)cr * ldy] = acc;
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      ++cr;
      cur_end = __shfl_sync(0xffffffffu, my_end, min(cr, 31));
    }
  };
  auto issue_batch = [&]() {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      if (p < e1) {
        if (p == cbase + 32) {
          cbase += 32;
          my_c = nx_c;
          my_v = nx_v;
          if (cbase + 32 + lane < e1) {
            nx_c = ld_stream_i32(indices + cbase + 32 + lane);
            nx_v = ld_stream_f32(vals + cbase + 32 + lane);
          }
        }
        const int c = __shfl_sync(0xffffffffu, my_c, p - cbase);
        const float v = __shfl_sync(0xffffffffu, my_v, p - cbase);
        const int slot = p & (S - 1);
        if (lane == 0) vring[slot] = v;
        if (active) {
          if (L2_64)
            cp_async_16_l2_64(ring + slot * stride + lane, Xc + (long long)c * ldx);
          else
            cp_async_16(ring + slot * stride + lane, Xc + (long long)c * ldx);
        }
        ++p;
      }
    }
    cp_async_commit();
  };
  auto consume_batch = [&]() {
    __syncwarp();
#pragma unroll
    for (int b = 0; b < B; ++b) {
      if (q < e1) {
        const int slot = q & (S - 1);
        const float v = vring[slot];
        if (active) {
          const float4 x = ring[slot * stride + lane];
          acc.x = fmaf(v, x.x, acc.x);
          acc.y = fmaf(v, x.y, acc.y);
          acc.z = fmaf(v, x.z, acc.z);
          acc.w = fmaf(v, x.w, acc.w);
        }
        ++q;
        flush_rows();
      }
    }
    __syncwarp();
  };

  flush_rows();  // leading empty rows
  const int nbatches = (e1 - e0 + B - 1) / B;
#pragma unroll
  for (int i = 0; i < NB - 1; ++i) issue_batch();
  for (int it = 0; it < nbatches; ++it) {
    issue_batch();
    cp_async_wait<NB - 1>();
    consume_batch();
  }
  cp_async_wait<0>();
}

// ---- stream kernel, unrolled form ------------------------------------------------------------------
// Same pipeline as above with the bookkeeping taken out of the per-position path: the stream is
// walked in chunks of 32 positions whose (index, value) live one per lane, the chunk body is fully
// unrolled so every shuffle lane, ring slot and row-end test is a compile-time constant, row ends
// are a 32-bit mask built once per chunk (REDUX), and the consumer lags the producer by exactly
// one batch of B positions (ring of 2B slots), carrying the tail of a chunk into the next one.
// ~13 issued instructions per gathered row instead of ~85: the kernel is DRAM-bound, not
// issue-bound (profiles/spmm_stream.md).
__device__ __forceinline__ void cp_async_16s(unsigned smem_addr, const void *g, bool l2_64) {
  if (l2_64)
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
  return v;
}

// PUSH: the row-partitioned multi-GPU hop.  Instead of one Y, every finished row is stored into the
// next-hop feature buffer of EVERY rank (peer pointers mapped over NVLink, own buffer included) at
// its global row index, so the all-gather of the next hop happens inside this kernel's epilogue and
// overlaps the gathers of the rows still in flight.
constexpr int kMaxPeers = 8;
struct PeerDests {
  float4 *p[kMaxPeers];
  int count;
};

template <int B, bool L2_64, bool PUSH>
__global__ void __launch_bounds__(kStreamWarps * 32)
spmm_stream2_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                    const float *__restrict__ vals, long long n_rows, const float4 *__restrict__ X,
                    unsigned ldx, float4 *__restrict__ Y, long long ldy, int nvec, int chunks, int R,
                    PeerDests peers) {
  constexpr int S = 2 * B;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ float4 smem4[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned ring = (unsigned)__cvta_generic_to_shared(smem4 + (size_t)w * S * 32 + lane);

  const long long task = (long long)blockIdx.x * kStreamWarps + w;
  const long long rblock = task / chunks;
  const int chunk = (int)(task - rblock * chunks);
  const long long r0 = rblock * R;
  if (r0 >= n_rows) return;
  const int nr = (int)min((long long)R, n_rows - r0);
  const int col = chunk * 32 + lane;
  const bool active = col < nvec;
  const float4 *Xc = X + col;
  float4 *yrow = Y + r0 * ldy + col;

  const int my_end = (lane < nr) ? __ldg(indptr + r0 + lane + 1) : 0;
  const int e0 = __ldg(indptr + r0);
  const int e1 = __shfl_sync(FULL, my_end, nr - 1);
  int prev_end = __shfl_up_sync(FULL, my_end, 1);
  if (lane == 0) prev_end = e0;

  if (__any_sync(FULL, lane < nr && my_end == prev_end)) {
    // an empty row in this task (never happens for A^, which has a full diagonal): plain loop
    for (int r = 0; r < nr; ++r) {
      const int st = __shfl_sync(FULL, prev_end, r), ed = __shfl_sync(FULL, my_end, r);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = st; j < ed; ++j) {
        const int c = __ldg(indices + j);
        const float v = __ldg(vals + j);
        if (active) {
          const float4 x = ld_gather_f4(Xc + (unsigned long long)(unsigned)c * ldx);
          acc.x = fmaf(v, x.x, acc.x);
          acc.y = fmaf(v, x.y, acc.y);
          acc.z = fmaf(v, x.z, acc.z);
          acc.w = fmaf(v, x.w, acc.w);
        }
      }
      if (active) {
        if (PUSH) {
          for (int d = 0; d < peers.count; ++d) (peers.p[d] + (yrow - Y))[(long long)r * ldy] = acc;
        } else {
          yrow[(long long)r * ldy] = acc;
        }
      }
    }
    return;
  }

  int cbase = e0;
  int my_c = 0, nx_c = 0;
  float my_v = 0.f, nx_v = 0.f, pv_v = 0.f;
  if (cbase + lane < e1) {
    my_c = ld_stream_i32(indices + cbase + lane);
    my_v = ld_stream_f32(vals + cbase + lane);
  }
  if (cbase + 32 + lane < e1) {
    nx_c = ld_stream_i32(indices + cbase + 32 + lane);
    nx_v = ld_stream_f32(vals + cbase + 32 + lane);
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
        const int j = (b - 1) * B + u;  // position consumed now (one batch behind)
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
            if (active) {
              if (PUSH) {
                const long long off = yrow - Y;
#pragma unroll
                for (int d = 0; d < kMaxPeers; ++d)
                  if (d < peers.count) peers.p[d][off] = acc;
              } else {
                *yrow = acc;
              }
            }
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
      nx_c = ld_stream_i32(indices + cbase + 32 + lane);
      nx_v = ld_stream_f32(vals + cbase + 32 + lane);
    }
  }
  cp_async_wait<0>();
}

// tuning knobs (srg_set_tuning): which kernel serves 16 < nvec and its shape
static int g_spmm_variant = 1;     // 0 = group kernel, 1 = stream kernel
static int g_stream_rows = 4;      // rows per warp task
static int g_stream_cfg = 10;      // 10/11/12: unrolled stream kernel B4/B8/B2; 0: S8/B4, 1: S8/B2, 2: S4/B2, 3: S4/B4, 4: S16/B4, 5: S16/B8, 6: S8/B1
static int g_stream_compact = 0;   // ring slot = nvec float4 instead of 32
static int g_gather_l2_64 = 1;     // gathers fetch 64-byte DRAM granules instead of 128-byte lines

template <int S, int B, bool L2_64>
static int launch_stream(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                         const float4 *X, int64_t ldx, float4 *Y, int64_t ldy, int nvec,
                         cudaStream_t s) {
  const int R = g_stream_rows < 1 ? 1 : (g_stream_rows > 32 ? 32 : g_stream_rows);
  const int chunks = (nvec + 31) / 32;
  const int stride = (g_stream_compact && nvec < 32) ? nvec : 32;
  const int64_t tasks = ceil_div64(n_rows, R) * chunks;
  const int64_t blocks = ceil_div64(tasks, kStreamWarps);
  if (blocks > 2147483647LL) {
    set_err("spmm: grid too large (%lld blocks)", (long long)blocks);
    return SRG_ERR_RANGE;
  }
  const size_t smem = (size_t)kStreamWarps * S * stride * sizeof(float4) + (size_t)kStreamWarps * S * sizeof(float);
  SRG_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<S, B, L2_64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spmm_stream_kernel<S, B, L2_64><<<(unsigned)blocks, kStreamWarps * 32, smem, s>>>(
      indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, chunks, R, stride);
  SRG_LAUNCHED();
  return SRG_OK;
}

template <int B, bool L2_64, bool PUSH = false>
static int launch_stream2(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                          const float4 *X, int64_t ldx, float4 *Y, int64_t ldy, int nvec,
                          cudaStream_t s, const PeerDests *peers = nullptr) {
  const int R = g_stream_rows < 1 ? 1 : (g_stream_rows > 32 ? 32 : g_stream_rows);
  const int chunks = (nvec + 31) / 32;
  const int64_t tasks = ceil_div64(n_rows, R) * chunks;
  const int64_t blocks = ceil_div64(tasks, kStreamWarps);
  if (blocks > 2147483647LL || ldx > 0xffffffffLL) {
    set_err("spmm: problem too large for the stream kernel (%lld blocks, ldx %lld)", (long long)blocks, (long long)ldx);
    return SRG_ERR_RANGE;
  }
  const size_t smem = (size_t)kStreamWarps * (2 * B) * 32 * sizeof(float4);
  PeerDests pd;
  pd.count = 0;
  for (int d = 0; d < kMaxPeers; ++d) pd.p[d] = nullptr;
  if (peers) pd = *peers;
  SRG_CUDA(cudaFuncSetAttribute(spmm_stream2_kernel<B, L2_64, PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spmm_stream2_kernel<B, L2_64, PUSH><<<(unsigned)blocks, kStreamWarps * 32, smem, s>>>(
      indptr, indices, vals, n_rows, X, (unsigned)ldx, Y, ldy, nvec, chunks, R, pd);
  SRG_LAUNCHED();
  return SRG_OK;
}

template <bool L2_64>
static int dispatch_stream(const int *indptr, const int *indices, const float *vals, int64_t n_rows,
                           const float4 *X, int64_t ldx, float4 *Y, int64_t ldy, int nvec,
                           cudaStream_t s) {
  switch (g_stream_cfg) {
    case 10: return launch_stream2<4, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 11: return launch_stream2<8, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 12: return launch_stream2<2, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 1: return launch_stream<8, 2, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 2: return launch_stream<4, 2, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 3: return launch_stream<4, 4, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 4: return launch_stream<16, 4, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 5: return launch_stream<16, 8, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    case 6: return launch_stream<8, 1, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
    default: return launch_stream<8, 4, L2_64>(indptr, indices, vals, n_rows, X, ldx, Y, ldy, nvec, s);
  }
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
    if (!accumulate && g_spmm_variant == 1 && nvec > 16) {
      return g_gather_l2_64 ? dispatch_stream<true>(indptr, indices, vals, n_rows, X4, ldx / 4, Y4, ldy / 4, nvec, s)
                            : dispatch_stream<false>(indptr, indices, vals, n_rows, X4, ldx / 4, Y4, ldy / 4, nvec, s);
    }
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

extern "C" int srg_set_tuning(const char *key, int64_t value) {
  SRG_REQUIRE(key != nullptr, "set_tuning: NULL key");
  const std::string k(key);
  if (k == "spmm_variant") g_spmm_variant = (int)value;
  else if (k == "stream_rows") g_stream_rows = (int)value;
  else if (k == "group_unroll") g_group_unroll = (int)value;
  else if (k == "gather_l2_64") g_gather_l2_64 = (int)value;
  else if (k == "stream_compact") g_stream_compact = (int)value;
  else if (k == "stream_cfg") g_stream_cfg = (int)value;
  else if (k == "l2_fetch_granularity") {
    int rc = require_device();
    if (rc) return rc;
    SRG_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
  } else {
    set_err("set_tuning: unknown key '%s'", key);
    return SRG_ERR_INVALID;
  }
  return SRG_OK;
}

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

extern "C" int srg_apply_feature_mask_f32(const float *x, int64_t ld_x, const int32_t *mask, float *out,
                                          int64_t ld_out, int64_t n, int32_t F, void *stream) {
  SRG_REQUIRE(mask != nullptr, "apply_feature_mask: mask is NULL");
  return srg_pack_features_f32(x, ld_x, out, ld_out, n, F, mask, stream);
}

// ---- multi-GPU push hop -------------------------------------------------------------------------------
extern "C" int srg_spmm_csr_f32_push(const int32_t *indptr, const int32_t *indices, const float *vals,
                                     int64_t n_rows, const float *X, int64_t ldx, float *const *dests,
                                     int32_t n_dests, int64_t dest_row0, int64_t ldy, int32_t F,
                                     void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n_rows >= 0 && F >= 0 && dest_row0 >= 0, "spmm_push: negative size");
  SRG_REQUIRE(n_dests >= 1 && n_dests <= kMaxPeers, "spmm_push: n_dests must be 1..%d", kMaxPeers);
  if (n_rows == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && vals && X && dests, "spmm_push: NULL pointer");
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
  float4 *origin = pd.p[0];
  if (g_gather_l2_64)
    return launch_stream2<4, true, true>(indptr, indices, vals, n_rows, reinterpret_cast<const float4 *>(X), ldx / 4,
                                         origin, ldy / 4, nvec, as_stream(stream), &pd);
  return launch_stream2<4, false, true>(indptr, indices, vals, n_rows, reinterpret_cast<const float4 *>(X), ldx / 4,
                                        origin, ldy / 4, nvec, as_stream(stream), &pd);
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
