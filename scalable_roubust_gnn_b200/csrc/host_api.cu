// host_api.cu — library core (errors, counters) and the HOST-buffer entry points.
//
//   srg_propagate_host      = GraphOp.propagate           (SSRG/operators/base_operator.py:19-36)
//   srg_construct_adj_host  = GraphOp.construct_adj        (SSRG/operators/graph_operator/*.py)
//   FloatCSRMulDenseOMP     = literal ABI of SSRG/operators/csrc/matmul.h:5
//   FloatCSRMulDense        = literal ABI of SSRG/operators/csrc/cudamatmul.c:28
//
// Pipeline of srg_propagate_host (three library-owned streams, no host threads):
//   copy-in stream : H2D raw CSR, then H2D features (+ pack into the padded device layout)
//   compute stream : normalisation kernels (overlap the feature upload), then K SpMM hops
//   copy-out stream: after hop k: unpack to the host layout and D2H, overlapping hop k+1
// Device memory comes from the stream-ordered pool (cudaMallocAsync) whose release threshold is
// raised so repeated calls reuse the same blocks.
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace srg {

std::atomic<long long> g_launches{0};

char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
void set_err(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
}

int require_device() {
  static std::atomic<int> cached{-1};
  int c = cached.load(std::memory_order_relaxed);
  if (c < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
      n = 0;
      (void)cudaGetLastError();
    }
    cached.store(n, std::memory_order_relaxed);
    c = n;
  }
  if (c <= 0) {
    set_err("no CUDA device visible: libsrgnn_b200 has no CPU fallback");
    return SRG_ERR_NODEV;
  }
  // the kernels take their scratch from the stream-ordered pool: keep freed blocks cached across
  // synchronisation points instead of returning them to the driver (once per device)
  static std::atomic<unsigned long long> pool_done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !(pool_done.load() >> dev & 1ULL)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thresh = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    pool_done.fetch_or(1ULL << dev);
  }
  return SRG_OK;
}

int spmm_csr_f32_impl(const int32_t *indptr, const int32_t *indices, const float *vals,
                      int64_t n_rows, int64_t nnz, const float *X, int64_t ldx, float *Y, int64_t ldy,
                      int32_t F, bool accumulate, cudaStream_t s);

// ---- per-device state of the host entry points -------------------------------------------------
constexpr int kStageSlots = 4;
constexpr size_t kStageChunk = 16u << 20;   // bytes per slot of the pinned staging ring

struct DeviceState {
  bool init = false;
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
  cudaEvent_t ev_csr = nullptr, ev_x = nullptr, ev_norm = nullptr;
  std::vector<cudaEvent_t> ev_hop;
  // the streams, events and the staging ring are shared by every call on this device: one call at a time
  std::mutex mu;
  // pinned staging ring for pageable inputs (allocated on first use)
  void *stage_buf[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  unsigned stage_next = 0;
};
static std::mutex g_state_mu;
static DeviceState g_state[64];

static int get_state(int device, DeviceState **out) {
  SRG_REQUIRE(device >= 0 && device < 64, "bad device ordinal %d", device);
  std::lock_guard<std::mutex> lk(g_state_mu);
  DeviceState &st = g_state[device];
  if (!st.init) {
    SRG_CUDA(cudaStreamCreateWithFlags(&st.s_in, cudaStreamNonBlocking));
    SRG_CUDA(cudaStreamCreateWithFlags(&st.s_compute, cudaStreamNonBlocking));
    SRG_CUDA(cudaStreamCreateWithFlags(&st.s_out, cudaStreamNonBlocking));
    SRG_CUDA(cudaEventCreateWithFlags(&st.ev_csr, cudaEventDisableTiming));
    SRG_CUDA(cudaEventCreateWithFlags(&st.ev_x, cudaEventDisableTiming));
    SRG_CUDA(cudaEventCreateWithFlags(&st.ev_norm, cudaEventDisableTiming));
    cudaMemPool_t pool;
    SRG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    SRG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    st.init = true;
  }
  *out = &st;
  return SRG_OK;
}

// ---- pageable host -> device copies -----------------------------------------------------------------------------
// The reference's callers hand over ordinary numpy arrays (SSRG/models/base_scalable/base_model.py:36).  A
// cudaMemcpyAsync from pageable memory is staged by the driver on ONE thread and blocks the caller; here the copy is
// cut into 16 MB chunks that a small pool of host threads moves into a ring of pinned buffers while the DMA engine
// drains the previous chunks, so the upload runs close to the pinned PCIe rate.
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this, i] { run(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  // dst <- src, split over the workers and the calling thread
  void copy(void *dst, const void *src, size_t n) {
    const int parts = (int)th_.size() + 1;
    const size_t per = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = static_cast<char *>(dst);
      src_ = static_cast<const char *>(src);
      n_ = n;
      per_ = per;
      pending_ = (int)th_.size();
      ++gen_;
    }
    cv_.notify_all();
    const size_t lo = std::min(n, per * (size_t)th_.size());
    if (lo < n) memcpy(dst_ + lo, src_ + lo, n - lo);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(int i) {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m_);
      cv_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      char *d = dst_;
      const char *s = src_;
      const size_t lo = std::min(n_, per_ * (size_t)i), hi = std::min(n_, lo + per_);
      lk.unlock();
      if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t n_ = 0, per_ = 0;
  int pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

static bool host_ptr_is_pageable(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

static int staging_threads() {
  static const int v = [] {
    const char *e = getenv("SRG_STAGE_THREADS");
    // measured at the products shape (1.24 GB of pageable input, 16 host CPUs): driver path 172 ms end to end,
    // 2 threads 146, 4 threads 106, 8 threads 99, pinned inputs 85; on a second, slower box 8 / 12 / 16 threads =
    // 107.5 / 99.8 / 97.3 ms (pinned 89): the staging copy is bound by host memory bandwidth, so take every CPU
    const int hw = (int)std::thread::hardware_concurrency();
    int t = e ? atoi(e) : std::max(2, std::min(16, hw));
    return std::max(0, std::min(t, 16));
  }();
  return v;   // 0: leave pageable copies to the driver
}

// H2D copy of `bytes` from a host buffer that may be pageable; stream-ordered on `s` (the caller holds st->mu)
static int h2d_copy(DeviceState *st, void *dst, const void *src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return SRG_OK;
  const int threads = staging_threads();
  if (threads == 0 || bytes < (4u << 20) || !host_ptr_is_pageable(src)) {
    SRG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return SRG_OK;
  }
  static CopyPool pool(threads - 1);
  static std::mutex pool_mu;   // one staged copy at a time across devices (the pool is shared)
  std::lock_guard<std::mutex> lk(pool_mu);
  for (int i = 0; i < kStageSlots; ++i)
    if (!st->stage_buf[i]) {
      SRG_CUDA(cudaHostAlloc(&st->stage_buf[i], kStageChunk, cudaHostAllocDefault));
      SRG_CUDA(cudaEventCreateWithFlags(&st->stage_ev[i], cudaEventDisableTiming));
    }
  for (size_t off = 0; off < bytes; off += kStageChunk) {
    const size_t len = std::min(kStageChunk, bytes - off);
    const unsigned slot = st->stage_next++ % kStageSlots;
    SRG_CUDA(cudaEventSynchronize(st->stage_ev[slot]));   // the DMA that last read this slot has finished
    pool.copy(st->stage_buf[slot], static_cast<const char *>(src) + off, len);
    SRG_CUDA(cudaMemcpyAsync(static_cast<char *>(dst) + off, st->stage_buf[slot], len, cudaMemcpyHostToDevice, s));
    SRG_CUDA(cudaEventRecord(st->stage_ev[slot], s));
  }
  return SRG_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// stream-ordered allocations released at scope exit
struct PoolAllocs {
  cudaStream_t s;
  std::vector<void *> ptrs;
  bool clean = false;  // set once every stream that touched the blocks has been synchronised
  explicit PoolAllocs(cudaStream_t s_) : s(s_) {}
  template <typename T> int alloc(T **p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    void *q = nullptr;
    SRG_CUDA(cudaMallocAsync(&q, (size_t)count * sizeof(T), s));
    ptrs.push_back(q);
    *p = static_cast<T *>(q);
    return SRG_OK;
  }
  ~PoolAllocs() {
    if (!clean) cudaDeviceSynchronize();  // error path: other streams may still use the blocks
    for (void *q : ptrs) cudaFreeAsync(q, s);
  }
};

static inline size_t val_bytes(int dt) { return dt == SRG_VAL_F32 ? 4 : dt == SRG_VAL_F64 ? 8 : 0; }
static inline int64_t pad8(int64_t f) { return (f + 7) / 8 * 8; }

static int check_flags(int flags) {
  if (flags & SRG_FLAG_BAD_INDEX) {
    set_err("adjacency has a column index outside [0, n)");
    return SRG_ERR_INVALID;
  }
  if (flags & (SRG_FLAG_UNSORTED | SRG_FLAG_ASYMMETRIC | SRG_FLAG_EXPLICIT_ZERO)) {
    set_err("internal: normalisation retry did not resolve flags 0x%x", flags);
    return SRG_ERR_CUDA;
  }
  if (flags & SRG_FLAG_ZERO_PRODUCT) {
    set_err("a normalised value is exactly 0 (scipy would drop the entry); compaction path required");
    return SRG_ERR_UNSUPPORTED;
  }
  return SRG_OK;
}

// normalisation on the compute stream from device copies of the raw CSR
struct NormOut {
  int32_t *indptr = nullptr, *indices = nullptr;
  double *val64 = nullptr;
  float *val32 = nullptr;
  int32_t *flags = nullptr;
};

static int run_norm(PoolAllocs &pa, const int32_t *d_indptr, const int32_t *d_indices,
                    const void *d_data, int val_dtype, int64_t n, int64_t nnz, double r,
                    double ppr_alpha, bool want64, bool want32, cudaStream_t s, int mode,
                    NormOut *o) {
  const bool canon = mode & 1, general = mode & 2, zeros = mode & 4;
  int rc;
  if ((rc = pa.alloc(&o->flags, 1))) return rc;
  SRG_CUDA(cudaMemsetAsync(o->flags, 0, sizeof(int32_t), s));
  if (canon) {
    // unsorted rows / duplicate entries: sort + sum first (scipy's tocoo() + eye does the same)
    int32_t *c_indptr, *c_indices, *c_nnz;
    double *c_vals;
    if ((rc = pa.alloc(&c_indptr, n + 1))) return rc;
    if ((rc = pa.alloc(&c_indices, nnz))) return rc;
    if ((rc = pa.alloc(&c_vals, nnz))) return rc;
    if ((rc = pa.alloc(&c_nnz, 1))) return rc;
    if ((rc = srg_csr_canonicalize(d_indptr, d_indices, d_data, val_dtype, n, nnz, c_indptr, c_indices, c_vals,
                                   c_nnz, o->flags, s)))
      return rc;
    d_indptr = c_indptr;
    d_indices = c_indices;
    d_data = c_vals;
    val_dtype = SRG_VAL_F64;
  }
  if (zeros) val_dtype |= SRG_VAL_HAS_ZEROS;  // explicit zeros: the compacting kernels
  int32_t *at_indptr;
  if ((rc = pa.alloc(&at_indptr, n + 1))) return rc;
  if ((rc = pa.alloc(&o->indices, nnz + n))) return rc;
  if (want64 && (rc = pa.alloc(&o->val64, nnz + n))) return rc;
  if (want32 && (rc = pa.alloc(&o->val32, nnz + n))) return rc;
  rc = srg_degree_selfloop_csr(d_indptr, d_indices, d_data, val_dtype, n, nnz, at_indptr, nullptr, o->flags, s);
  if (rc) return rc;
  if (!general) {
    o->indptr = at_indptr;
    return srg_sym_norm_csr(d_indptr, d_indices, d_data, val_dtype, n, nnz, at_indptr, r, ppr_alpha,
                            o->indices, nullptr, o->val64, o->val32, o->flags, s);
  }
  if ((rc = pa.alloc(&o->indptr, n + 1))) return rc;
  return srg_sym_norm_csr_general(d_indptr, d_indices, d_data, val_dtype, n, nnz, at_indptr, r, ppr_alpha,
                                  o->indptr, o->indices, nullptr, o->val64, o->val32, o->flags, s);
}

// aggregation of the hop list requested by srg_propagate_aggregate_host (mode SRG_AGG_NONE: none)
struct AggSpec {
  int mode = SRG_AGG_NONE;
  int start = 0, end = 0;
  const float *weights = nullptr;  // host, end - start entries (WEIGHTED)
  float *out = nullptr;            // host, n x f_out
};

// which retry the flags ask for: returns true when another attempt with (canon, general) makes sense
// mode bits: 1 canonicalise first, 2 general (transpose) path, 4 input holds explicit zeros
static bool next_attempt(int flags, int *mode) {
  if (flags & SRG_FLAG_BAD_INDEX) return false;
  int want = *mode;
  if (flags & SRG_FLAG_UNSORTED) want |= 1;
  if (flags & SRG_FLAG_EXPLICIT_ZERO) want |= 4;
  if (flags & SRG_FLAG_ASYMMETRIC) want |= 2;
  if (want == *mode) return false;
  *mode = want;
  return true;
}

// ---- "is the adjacency unweighted?" on the host --------------------------------------------------------------
// scipy hands A.data as float64 even when every stored value is 1.0 (the usual unweighted graph): 8 bytes per
// entry of PCIe traffic that carry no information (products shape: 495 MB of the 1.73 GB upload).  With
// The host pipeline therefore starts WITHOUT uploading the values (val_dtype ONES) while worker threads verify,
// in the shadow of the transfers, that the array really is all ones; if not, the result is discarded and the
// regular path runs (SRG_ONES_SHORTCUT=0 turns the shortcut off; measured at the products shape: 92.9 -> 84.0 ms
// end to end, identical outputs).  The device arithmetic is identical either way (the kernels detect weightedness at run
// time and use exact integer degrees for all-ones input).
template <typename T>
static bool range_all_ones(const T *p, int64_t lo, int64_t hi, const std::atomic<int> &stop) {
  const T one = (T)1;
  for (int64_t i = lo; i < hi; i += 8192) {
    if (stop.load(std::memory_order_relaxed)) return false;
    const int64_t e = std::min<int64_t>(hi, i + 8192);
    int bad = 0;
    for (int64_t j = i; j < e; ++j) bad |= (p[j] != one);
    if (bad) return false;
  }
  return true;
}

struct OnesProbe {
  std::vector<std::thread> workers;
  std::atomic<int> bad{0};
  void start(const void *data, int val_dtype, int64_t nnz, int threads) {
    threads = std::max(1, std::min(threads, 64));
    const int64_t per = (nnz + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
      const int64_t lo = std::min<int64_t>(nnz, t * per), hi = std::min<int64_t>(nnz, lo + per);
      if (lo >= hi) break;
      workers.emplace_back([this, data, val_dtype, lo, hi] {
        const bool ok = (val_dtype == SRG_VAL_F64) ? range_all_ones(static_cast<const double *>(data), lo, hi, bad)
                                                   : range_all_ones(static_cast<const float *>(data), lo, hi, bad);
        if (!ok) bad.store(1, std::memory_order_relaxed);
      });
    }
  }
  bool finish() {
    for (auto &w : workers) w.join();
    workers.clear();
    return bad.load() == 0;
  }
  ~OnesProbe() { finish(); }
};

// cheap pre-filter: the first entries and a strided sample
static bool sampled_all_ones(const void *data, int val_dtype, int64_t nnz) {
  std::atomic<int> never{0};
  const int64_t head = std::min<int64_t>(nnz, 4096);
  const bool f64 = val_dtype == SRG_VAL_F64;
  if (f64 ? !range_all_ones(static_cast<const double *>(data), 0, head, never)
          : !range_all_ones(static_cast<const float *>(data), 0, head, never))
    return false;
  const int64_t stride = std::max<int64_t>(1, nnz / 4096);
  for (int64_t i = 0; i < nnz; i += stride)
    if (f64 ? static_cast<const double *>(data)[i] != 1.0 : static_cast<const float *>(data)[i] != 1.0f) return false;
  return true;
}

static bool ones_shortcut_enabled() {
  static const int v = [] {
    const char *e = getenv("SRG_ONES_SHORTCUT");
    return e ? atoi(e) : 1;
  }();
  return v != 0;
}

static bool ones_shortcut_applies(const void *data, int val_dtype, int64_t nnz) {
  return ones_shortcut_enabled() && data && (val_dtype == SRG_VAL_F64 || val_dtype == SRG_VAL_F32) && nnz >= (1 << 20) &&
         sampled_all_ones(data, val_dtype, nnz);
}

}  // namespace srg

using namespace srg;

extern "C" int srg_abi_version(void) { return SRG_ABI_VERSION; }
extern "C" const char *srg_last_error(void) { return err_buf(); }
extern "C" int srg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" int64_t srg_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int srg_host_all_ones(const void *data, int val_dtype, int64_t nnz, int32_t threads) {
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2 && nnz >= 0, "host_all_ones: bad arguments");
  if (val_dtype == SRG_VAL_ONES || nnz == 0) return 1;
  SRG_REQUIRE(data != nullptr, "host_all_ones: data is NULL");
  OnesProbe probe;
  probe.start(data, val_dtype, nnz, threads);
  return probe.finish() ? 1 : 0;
}

extern "C" int srg_release_workspace(void) {
  std::lock_guard<std::mutex> lk(g_state_mu);
  for (int d = 0; d < 64; ++d) {
    DeviceState &st = g_state[d];
    if (!st.init) continue;
    DeviceGuard g(d);
    cudaDeviceSynchronize();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  return SRG_OK;
}

static int construct_attempt(const int32_t *indptr, const int32_t *indices, const void *data,
                             int val_dtype, int64_t n, int64_t nnz, double r, double ppr_alpha,
                             int32_t *out_indptr, int32_t *out_indices, double *out_data,
                             int64_t *out_nnz, int device, int mode, int *flags_out) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "construct_adj: negative size");
  SRG_REQUIRE(indptr && out_indptr && out_nnz, "construct_adj: NULL pointer");
  SRG_REQUIRE(nnz == 0 || indices, "construct_adj: indices is NULL");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "construct_adj: bad val_dtype");
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || nnz == 0 || data, "construct_adj: data is NULL");
  SRG_REQUIRE(nnz + n <= 2147483647LL, "construct_adj: nnz + n exceeds the int32 CSR range");
  DeviceGuard guard(device);
  SRG_REQUIRE(guard.ok, "construct_adj: cannot select device %d", device);
  DeviceState *st;
  if ((rc = get_state(device, &st))) return rc;
  cudaStream_t s = st->s_compute;
  if (n == 0) {
    out_indptr[0] = 0;
    *out_nnz = 0;
    return SRG_OK;
  }
  std::lock_guard<std::mutex> dev_lock(st->mu);
  int flags = 0;
  int32_t nnz_out = 0;
  {
    PoolAllocs pa(s);
    int32_t *d_indptr, *d_indices;
    char *d_data = nullptr;
    if ((rc = pa.alloc(&d_indptr, n + 1))) return rc;
    if ((rc = pa.alloc(&d_indices, nnz))) return rc;
    if ((rc = h2d_copy(st, d_indptr, indptr, (size_t)(n + 1) * 4, s))) return rc;
    if (nnz && (rc = h2d_copy(st, d_indices, indices, (size_t)nnz * 4, s))) return rc;
    if (val_dtype != SRG_VAL_ONES) {
      if ((rc = pa.alloc(&d_data, nnz * (int64_t)val_bytes(val_dtype)))) return rc;
      if (nnz && (rc = h2d_copy(st, d_data, data, (size_t)nnz * val_bytes(val_dtype), s))) return rc;
    }
    NormOut no;
    if ((rc = run_norm(pa, d_indptr, d_indices, d_data, val_dtype, n, nnz, r, ppr_alpha,
                       out_data != nullptr, false, s, mode, &no)))
      return rc;
    SRG_CUDA(cudaMemcpyAsync(&flags, no.flags, 4, cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaMemcpyAsync(out_indptr, no.indptr, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));
    flags &= ~SRG_FLAG_WEIGHTED;  // informational
    *flags_out = flags;
    if (flags) {
      pa.clean = true;
      return SRG_OK;  // the caller decides between a retry and an error
    }
    nnz_out = out_indptr[n];
    if (out_indices) SRG_CUDA(cudaMemcpyAsync(out_indices, no.indices, (size_t)nnz_out * 4, cudaMemcpyDeviceToHost, s));
    if (out_data) SRG_CUDA(cudaMemcpyAsync(out_data, no.val64, (size_t)nnz_out * 8, cudaMemcpyDeviceToHost, s));
    SRG_CUDA(cudaStreamSynchronize(s));
    pa.clean = true;
  }
  *out_nnz = nnz_out;
  return SRG_OK;
}

extern "C" int srg_construct_adj_host(const int32_t *indptr, const int32_t *indices,
                                      const void *data, int val_dtype, int64_t n, int64_t nnz,
                                      double r, double ppr_alpha, int32_t *out_indptr,
                                      int32_t *out_indices, double *out_data, int64_t *out_nnz,
                                      int device) {
  int mode = 0;
  for (;;) {
    int flags = 0;
    int rc = construct_attempt(indptr, indices, data, val_dtype, n, nnz, r, ppr_alpha, out_indptr, out_indices,
                               out_data, out_nnz, device, mode, &flags);
    if (rc) return rc;
    if (!flags) return SRG_OK;
    if (!next_attempt(flags, &mode)) return check_flags(flags);
  }
}

static int propagate_attempt(const int32_t *indptr, const int32_t *indices, const void *data,
                             int val_dtype, int64_t n, int64_t nnz, const float *features,
                             int32_t F, const int32_t *feature_mask, int32_t K, double r,
                             double ppr_alpha, float *const *out_hops, int32_t *out_norm_indptr,
                             int32_t *out_norm_indices, double *out_norm_data, int64_t *out_nnz,
                             int device, int mode, int *flags_out, const AggSpec *agg = nullptr) {
  int rc = require_device();
  if (rc) return rc;
  const bool do_agg = agg && agg->mode != SRG_AGG_NONE;
  SRG_REQUIRE(n >= 0 && nnz >= 0 && F >= 0 && K >= 0, "propagate_host: negative size");
  SRG_REQUIRE(indptr, "propagate_host: indptr is NULL");
  SRG_REQUIRE(nnz == 0 || indices, "propagate_host: indices is NULL");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "propagate_host: bad val_dtype");
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || nnz == 0 || data, "propagate_host: data is NULL");
  SRG_REQUIRE(n * (int64_t)F == 0 || features, "propagate_host: features is NULL");
  SRG_REQUIRE(K == 0 || out_hops || do_agg, "propagate_host: out_hops is NULL");
  SRG_REQUIRE(nnz + n <= 2147483647LL, "propagate_host: nnz + n exceeds the int32 CSR range");
  for (int k = 0; k < K && !do_agg; ++k)
    SRG_REQUIRE(n * (int64_t)F == 0 || out_hops[k], "propagate_host: out_hops[%d] is NULL", k);
  if (out_nnz) *out_nnz = 0;
  if (n == 0) {
    if (out_norm_indptr) out_norm_indptr[0] = 0;
    return SRG_OK;
  }
  DeviceGuard guard(device);
  SRG_REQUIRE(guard.ok, "propagate_host: cannot select device %d", device);
  DeviceState *st;
  if ((rc = get_state(device, &st))) return rc;
  std::lock_guard<std::mutex> dev_lock(st->mu);   // shared streams / events / staging ring: one call per device at a time
  while ((int)st->ev_hop.size() < K) {
    cudaEvent_t e;
    SRG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    st->ev_hop.push_back(e);
  }
  cudaStream_t s_in = st->s_in, s_c = st->s_compute, s_out = st->s_out;
  const int64_t ld = pad8(F);
  const bool padded = (ld != F);
  int flags = 0;
  int32_t nnz_out = 0;
  {
    // all allocations are made on the compute stream first; the copy streams are ordered
    // behind an event recorded after the allocations.
    PoolAllocs pa(s_c);
    int32_t *d_indptr, *d_indices;
    char *d_data = nullptr;
    if ((rc = pa.alloc(&d_indptr, n + 1))) return rc;
    if ((rc = pa.alloc(&d_indices, nnz))) return rc;
    if (val_dtype != SRG_VAL_ONES && (rc = pa.alloc(&d_data, nnz * (int64_t)val_bytes(val_dtype)))) return rc;
    std::vector<float *> hops(K + 1, nullptr);
    const bool have_feat = (int64_t)F * n > 0;
    // with an aggregation only two hop buffers ping-pong and an accumulator collects the result
    const int n_bufs = (do_agg && agg->mode != SRG_AGG_NAFS) ? std::min(K + 1, 2) : K + 1;
    if (have_feat)
      for (int k = 0; k <= K; ++k) {
        if (k < n_bufs) {
          if ((rc = pa.alloc(&hops[k], n * ld))) return rc;
        } else {
          hops[k] = hops[k - 2];
        }
      }
    const int agg_cnt = do_agg ? agg->end - agg->start : 0;
    const int64_t f_out = do_agg ? (agg->mode == SRG_AGG_CONCAT ? (int64_t)agg_cnt * F : F) : 0;
    const int64_t ld_acc = pad8(f_out);
    float *d_acc = nullptr, *d_acc_flat = nullptr;
    if (do_agg && have_feat && agg->mode != SRG_AGG_LAST) {
      if ((rc = pa.alloc(&d_acc, n * ld_acc))) return rc;
    }
    if (do_agg && have_feat && (ld_acc != f_out)) {
      if ((rc = pa.alloc(&d_acc_flat, n * f_out))) return rc;
    }
    float *d_flat_in = nullptr, *d_flat_out[2] = {nullptr, nullptr};
    int32_t *d_mask = nullptr;
    const bool need_pack = have_feat && (padded || feature_mask);
    if (need_pack && (rc = pa.alloc(&d_flat_in, n * (int64_t)F))) return rc;
    if (have_feat && feature_mask && (rc = pa.alloc(&d_mask, n * (int64_t)F))) return rc;
    if (have_feat && padded && K > 0 && !do_agg) {
      if ((rc = pa.alloc(&d_flat_out[0], n * (int64_t)F))) return rc;
      if (K > 1 && (rc = pa.alloc(&d_flat_out[1], n * (int64_t)F))) return rc;
    }
    SRG_CUDA(cudaEventRecord(st->ev_csr, s_c));
    SRG_CUDA(cudaStreamWaitEvent(s_in, st->ev_csr, 0));
    SRG_CUDA(cudaStreamWaitEvent(s_out, st->ev_csr, 0));

    // copy-in: raw CSR first (normalisation can start), then features
    if ((rc = h2d_copy(st, d_indptr, indptr, (size_t)(n + 1) * 4, s_in))) return rc;
    if (nnz && (rc = h2d_copy(st, d_indices, indices, (size_t)nnz * 4, s_in))) return rc;
    if (d_data && nnz && (rc = h2d_copy(st, d_data, data, (size_t)nnz * val_bytes(val_dtype), s_in))) return rc;
    SRG_CUDA(cudaEventRecord(st->ev_csr, s_in));
    // compute: normalisation (queued before the feature upload so it overlaps it even when the
    // host buffers are pageable and cudaMemcpyAsync blocks the host while staging)
    SRG_CUDA(cudaStreamWaitEvent(s_c, st->ev_csr, 0));
    NormOut no;
    if ((rc = run_norm(pa, d_indptr, d_indices, d_data, val_dtype, n, nnz, r, ppr_alpha,
                       out_norm_data != nullptr, true, s_c, mode, &no)))
      return rc;
    SRG_CUDA(cudaEventRecord(st->ev_norm, s_c));

    if (have_feat) {
      float *dst = need_pack ? d_flat_in : hops[0];
      if ((rc = h2d_copy(st, dst, features, (size_t)n * F * 4, s_in))) return rc;
      if (feature_mask && (rc = h2d_copy(st, d_mask, feature_mask, (size_t)n * F * 4, s_in))) return rc;
      if (need_pack && (rc = srg_pack_features_f32(d_flat_in, F, hops[0], ld, n, F, d_mask, s_in))) return rc;
      SRG_CUDA(cudaEventRecord(st->ev_x, s_in));
    }

    // hops + overlapped copy-out
    if (have_feat && !do_agg) {
      SRG_CUDA(cudaStreamWaitEvent(s_c, st->ev_x, 0));
      for (int k = 1; k <= K; ++k) {
        if ((rc = spmm_csr_f32_impl(no.indptr, no.indices, no.val32, n, nnz + n, hops[k - 1], ld, hops[k], ld, F, false, s_c)))
          return rc;
        SRG_CUDA(cudaEventRecord(st->ev_hop[k - 1], s_c));
        SRG_CUDA(cudaStreamWaitEvent(s_out, st->ev_hop[k - 1], 0));
        const float *src = hops[k];
        if (padded) {
          float *flat = d_flat_out[(k - 1) & 1];
          if ((rc = srg_unpack_features_f32(hops[k], ld, flat, F, n, F, s_out))) return rc;
          src = flat;
        }
        SRG_CUDA(cudaMemcpyAsync(out_hops[k - 1], src, (size_t)n * F * 4, cudaMemcpyDeviceToHost, s_out));
      }
    }
    if (have_feat && do_agg) {
      // hop k is folded into the accumulator as soon as it exists; only the aggregate goes back
      SRG_CUDA(cudaStreamWaitEvent(s_c, st->ev_x, 0));
      bool first = true;
      for (int k = 0; k <= K; ++k) {
        if (k > 0 &&
            (rc = spmm_csr_f32_impl(no.indptr, no.indices, no.val32, n, nnz + n, hops[k - 1], ld, hops[k], ld, F, false, s_c)))
          return rc;
        if (agg->mode == SRG_AGG_LAST || agg->mode == SRG_AGG_NAFS || k < agg->start || k >= agg->end) continue;
        const int slot = k - agg->start;
        const float wgt = (agg->mode == SRG_AGG_WEIGHTED) ? agg->weights[slot] : 1.0f;
        const int col0 = (agg->mode == SRG_AGG_CONCAT) ? slot * F : 0;
        if ((rc = srg_aggregate_update_f32(d_acc, ld_acc, col0, hops[k], ld, n, F, agg->mode, wgt,
                                           (first || agg->mode == SRG_AGG_CONCAT) ? 1 : 0, s_c)))
          return rc;
        first = false;
      }
      if (agg->mode == SRG_AGG_MEAN &&
          (rc = srg_aggregate_update_f32(d_acc, ld_acc, 0, nullptr, 0, n, F, -1, (float)agg_cnt, 0, s_c)))
        return rc;
      if (agg->mode == SRG_AGG_NAFS &&
          (rc = srg_nafs_combine_f32(hops.data(), K + 1, ld, n, F, d_acc, ld_acc, nullptr, s_c)))
        return rc;
      const float *res = (agg->mode == SRG_AGG_LAST) ? hops[K] : d_acc;
      const int64_t ld_res = (agg->mode == SRG_AGG_LAST) ? ld : ld_acc;
      if (ld_res != f_out) {
        if ((rc = srg_unpack_features_f32(res, ld_res, d_acc_flat, f_out, n, (int32_t)f_out, s_c))) return rc;
        res = d_acc_flat;
      }
      SRG_CUDA(cudaMemcpyAsync(agg->out, res, (size_t)n * f_out * 4, cudaMemcpyDeviceToHost, s_c));
    }
    // normalised CSR back to the host if requested (after the hop copies are queued)
    SRG_CUDA(cudaStreamWaitEvent(s_in, st->ev_norm, 0));
    SRG_CUDA(cudaMemcpyAsync(&flags, no.flags, 4, cudaMemcpyDeviceToHost, s_in));
    int32_t h_nnz_out = 0;
    SRG_CUDA(cudaMemcpyAsync(&h_nnz_out, no.indptr + n, 4, cudaMemcpyDeviceToHost, s_in));
    if (out_norm_indptr)
      SRG_CUDA(cudaMemcpyAsync(out_norm_indptr, no.indptr, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, s_in));
    SRG_CUDA(cudaStreamSynchronize(s_in));
    nnz_out = h_nnz_out;
    flags &= ~SRG_FLAG_WEIGHTED;  // informational
    *flags_out = flags;
    const int frc = flags ? 1 : 0;  // flagged: results are discarded, the caller retries or fails
    if (!frc) {
      if (out_norm_indices)
        SRG_CUDA(cudaMemcpyAsync(out_norm_indices, no.indices, (size_t)nnz_out * 4, cudaMemcpyDeviceToHost, s_in));
      if (out_norm_data)
        SRG_CUDA(cudaMemcpyAsync(out_norm_data, no.val64, (size_t)nnz_out * 8, cudaMemcpyDeviceToHost, s_in));
    }
    SRG_CUDA(cudaStreamSynchronize(s_in));
    SRG_CUDA(cudaStreamSynchronize(s_out));
    SRG_CUDA(cudaStreamSynchronize(s_c));
    pa.clean = true;
    if (frc) return SRG_OK;
  }
  if (out_nnz) *out_nnz = nnz_out;
  return SRG_OK;
}

extern "C" int srg_propagate_host(const int32_t *indptr, const int32_t *indices, const void *data,
                                  int val_dtype, int64_t n, int64_t nnz, const float *features,
                                  int32_t F, const int32_t *feature_mask, int32_t K, double r,
                                  double ppr_alpha, float *const *out_hops,
                                  int32_t *out_norm_indptr, int32_t *out_norm_indices,
                                  double *out_norm_data, int64_t *out_nnz, int device) {
  auto run = [&](const void *vals, int vt) -> int {
    int mode = 0;
    for (;;) {
      int flags = 0;
      int rc = propagate_attempt(indptr, indices, vals, vt, n, nnz, features, F, feature_mask, K, r, ppr_alpha,
                                 out_hops, out_norm_indptr, out_norm_indices, out_norm_data, out_nnz, device, mode,
                                 &flags);
      if (rc) return rc;
      if (!flags) return SRG_OK;
      if (!next_attempt(flags, &mode)) return check_flags(flags);
    }
  };
  if (ones_shortcut_applies(data, val_dtype, nnz)) {
    OnesProbe probe;
    probe.start(data, val_dtype, nnz, 8);
    const int rc = run(nullptr, SRG_VAL_ONES);
    if (probe.finish()) return rc;   // really unweighted: the result (or the error) stands
  }
  return run(data, val_dtype);
}

extern "C" int srg_propagate_aggregate_host(const int32_t *indptr, const int32_t *indices, const void *data,
                                            int val_dtype, int64_t n, int64_t nnz, const float *features,
                                            int32_t F, const int32_t *feature_mask, int32_t K, double r,
                                            double ppr_alpha, int32_t agg_mode, int32_t agg_start,
                                            int32_t agg_end, const float *agg_weights, float *out_agg,
                                            int device) {
  SRG_REQUIRE(agg_mode >= SRG_AGG_LAST && agg_mode <= SRG_AGG_NAFS, "propagate_aggregate: bad agg_mode %d", agg_mode);
  SRG_REQUIRE(K >= 0, "propagate_aggregate: K must be >= 0");
  if (agg_mode == SRG_AGG_LAST) {
    agg_start = K;
    agg_end = K + 1;
  }
  if (agg_mode == SRG_AGG_NAFS) {
    agg_start = 0;
    agg_end = K + 1;
    if (K + 1 > 64) {
      set_err("propagate_aggregate: NAFS aggregation supports at most 64 hop matrices (K = %d)", K);
      return SRG_ERR_UNSUPPORTED;
    }
  }
  SRG_REQUIRE(agg_start >= 0 && agg_start < agg_end && agg_end <= K + 1,
              "propagate_aggregate: hop slice [%d, %d) outside [0, %d]", agg_start, agg_end, K + 1);
  SRG_REQUIRE(agg_mode != SRG_AGG_WEIGHTED || agg_weights, "propagate_aggregate: weights are NULL");
  SRG_REQUIRE(n * (int64_t)F == 0 || out_agg, "propagate_aggregate: out_agg is NULL");
  AggSpec agg;
  agg.mode = agg_mode;
  agg.start = agg_start;
  agg.end = agg_end;
  agg.weights = agg_weights;
  agg.out = out_agg;
  auto run = [&](const void *vals, int vt) -> int {
    int mode = 0;
    for (;;) {
      int flags = 0;
      int rc = propagate_attempt(indptr, indices, vals, vt, n, nnz, features, F, feature_mask, K, r, ppr_alpha,
                                 nullptr, nullptr, nullptr, nullptr, nullptr, device, mode, &flags, &agg);
      if (rc) return rc;
      if (!flags) return SRG_OK;
      if (!next_attempt(flags, &mode)) return check_flags(flags);
    }
  };
  if (ones_shortcut_applies(data, val_dtype, nnz)) {
    OnesProbe probe;
    probe.start(data, val_dtype, nnz, 8);
    const int rc = run(nullptr, SRG_VAL_ONES);
    if (probe.finish()) return rc;
  }
  return run(data, val_dtype);
}

// ---- literal reference ABI ------------------------------------------------------------------------
static int shim_spmm(float *answer, const float *data, const int *indices, const int *indptr,
                     const float *mat, int mat_row, int mat_col) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(mat_row >= 0 && mat_col >= 0, "FloatCSRMulDense*: negative size");
  if (mat_row == 0 || mat_col == 0) return SRG_OK;
  SRG_REQUIRE(answer && indptr && mat, "FloatCSRMulDense*: NULL pointer");
  int device = 0;
  SRG_CUDA(cudaGetDevice(&device));
  DeviceState *st;
  if ((rc = get_state(device, &st))) return rc;
  cudaStream_t s = st->s_compute;
  std::lock_guard<std::mutex> dev_lock(st->mu);
  const int64_t n = mat_row, F = mat_col, nnz = indptr[mat_row];
  SRG_REQUIRE(nnz >= 0 && (nnz == 0 || (data && indices)), "FloatCSRMulDense*: bad CSR");
  PoolAllocs pa(s);
  int32_t *d_indptr, *d_indices;
  float *d_vals, *d_x, *d_y;
  if ((rc = pa.alloc(&d_indptr, n + 1))) return rc;
  if ((rc = pa.alloc(&d_indices, nnz))) return rc;
  if ((rc = pa.alloc(&d_vals, nnz))) return rc;
  if ((rc = pa.alloc(&d_x, n * F))) return rc;
  if ((rc = pa.alloc(&d_y, n * F))) return rc;
  SRG_CUDA(cudaMemcpyAsync(d_indptr, indptr, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s));
  if (nnz) {
    SRG_CUDA(cudaMemcpyAsync(d_indices, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
    SRG_CUDA(cudaMemcpyAsync(d_vals, data, (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
  }
  SRG_CUDA(cudaMemcpyAsync(d_x, mat, (size_t)n * F * 4, cudaMemcpyHostToDevice, s));
  // the reference accumulates into `answer` (matmul.c:37): start every chain from its content
  SRG_CUDA(cudaMemcpyAsync(d_y, answer, (size_t)n * F * 4, cudaMemcpyHostToDevice, s));
  if ((rc = spmm_csr_f32_impl(d_indptr, d_indices, d_vals, n, nnz, d_x, F, d_y, F, (int32_t)F, true, s))) return rc;
  SRG_CUDA(cudaMemcpyAsync(answer, d_y, (size_t)n * F * 4, cudaMemcpyDeviceToHost, s));
  SRG_CUDA(cudaStreamSynchronize(s));
  pa.clean = true;
  return SRG_OK;
}

extern "C" void FloatCSRMulDenseOMP(float answer[], float data[], int indices[], int indptr[],
                                    float mat[], int mat_row, int mat_col) {
  // the reference symbol has no error channel (matmul.h:5): failures are loud, never silent.
  int rc = shim_spmm(answer, data, indices, indptr, mat, mat_row, mat_col);
  if (rc) {
    fprintf(stderr, "libsrgnn_b200: FloatCSRMulDenseOMP failed (%d): %s\n", rc, err_buf());
    abort();
  }
}

extern "C" int FloatCSRMulDense(float answer[], int data_nnz, float data[], int indices[],
                                int indptr[], float mat[], int mat_row, int mat_col) {
  (void)data_nnz;
  int rc = shim_spmm(answer, data, indices, indptr, mat, mat_row, mat_col);
  if (rc) {
    printf("libsrgnn_b200: FloatCSRMulDense failed (%d): %s\n", rc, err_buf());  // cudamatmul.c:7-25 prints
    return 1;  // EXIT_FAILURE
  }
  return 0;  // EXIT_SUCCESS
}
