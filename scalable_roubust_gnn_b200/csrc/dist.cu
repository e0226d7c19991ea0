// dist.cu — native multi-GPU handle: row-partitioned propagation with one NCCL all-gather per hop.
//
// `srg_dist_init(...)` + `srg_dist_propagate(...)`.  The reference has no multi-GPU code;
// the scheme is the one BASELINE.json's north star names: contiguous equal row blocks
// (rows_per = ceil(N / P), rank p owns [p * rows_per, min(N, (p + 1) * rows_per))), the degree vector
// all-gathered once for the normalisation (SSRG/operators/utils.py:81-93 on the local rows), and before every hop
// an NCCL all-gather of the previous hop's feature slices, X_k = A^ X_{k-1} on the local rows
// (SSRG/operators/base_operator.py:31-36).  Every output row is owned by one rank and reduced in CSR order, so
// the P-way result is bitwise the 1-GPU result.
// This is the C twin of the "allgather" mode of scalable_roubust_gnn_b200/dist.py for callers that do not
// run torch.distributed; the push exchange (peer mappings, fused epilogue) lives in dist.py + spmm.cu.
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy the process already holds when torch is loaded),
// the library itself does not link against it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace srg {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return SRG_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy already in the process (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_err("dist: cannot load libnccl.so.2 (%s)", dlerror());
    return SRG_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.lib = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
    set_err("dist: libnccl.so.2 lacks an expected symbol");
    return SRG_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return SRG_OK;
}

#define SRG_NCCL(call)                                                                      \
  do {                                                                                      \
    ncclResult_t r__ = (call);                                                              \
    if (r__ != ncclSuccess) {                                                               \
      srg::set_err("%s failed: %s (%s:%d)", #call, srg::g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
      return SRG_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

struct DistHandle {
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  int world = 1, rank = 0, device = 0;
  int64_t n = 0, rows_per = 0, row0 = 0, n_local = 0, n_pad = 0, ld = 0;
  int F = 0;
  float *full[2] = {nullptr, nullptr};   // n_pad x ld, ping-pong
  double *deg_all = nullptr;             // n_pad
};

static inline int64_t pad8(int64_t f) { return (f + 7) / 8 * 8; }

static int make_handle(ncclComm_t comm, bool own, int world, int rank, int64_t n, int F, DistHandle **out) {
  DistHandle *h = new DistHandle();
  h->comm = comm;
  h->own_comm = own;
  h->world = world;
  h->rank = rank;
  cudaGetDevice(&h->device);
  h->n = n;
  h->F = F;
  h->rows_per = n > 0 ? (n + world - 1) / world : 0;
  h->row0 = std::min<int64_t>((int64_t)rank * h->rows_per, n);
  h->n_local = std::min<int64_t>((int64_t)(rank + 1) * h->rows_per, n) - h->row0;
  h->n_pad = h->rows_per * world;
  h->ld = pad8(F);
  const size_t bytes = std::max<size_t>((size_t)h->n_pad * h->ld * sizeof(float), 256);
  for (int b = 0; b < 2; ++b) {
    cudaError_t e = cudaMalloc(&h->full[b], bytes);
    if (e == cudaSuccess) e = cudaMemset(h->full[b], 0, bytes);   // rows >= n of the last block travel as zeros
    if (e != cudaSuccess) {
      for (int c = 0; c <= b; ++c) cudaFree(h->full[c]);
      delete h;
      return cuda_fail(e, "cudaMalloc(full feature buffer)", __FILE__, __LINE__);
    }
  }
  cudaError_t e = cudaMalloc(&h->deg_all, std::max<size_t>((size_t)h->n_pad * sizeof(double), 256));
  if (e == cudaSuccess) e = cudaMemset(h->deg_all, 0, std::max<size_t>((size_t)h->n_pad * sizeof(double), 256));
  if (e != cudaSuccess) {
    cudaFree(h->full[0]);
    cudaFree(h->full[1]);
    delete h;
    return cuda_fail(e, "cudaMalloc(degree vector)", __FILE__, __LINE__);
  }
  *out = h;
  return SRG_OK;
}

}  // namespace srg

using namespace srg;

extern "C" int srg_dist_unique_id(void *id128) {
  SRG_REQUIRE(id128 != nullptr, "dist_unique_id: NULL pointer");
  int rc = load_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  SRG_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SRG_OK;
}

extern "C" int srg_dist_init(const void *id128, int32_t world, int32_t rank, int64_t n, int32_t F, void **out_handle) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(id128 && out_handle, "dist_init: NULL pointer");
  SRG_REQUIRE(world >= 1 && rank >= 0 && rank < world && n >= 0 && F >= 0, "dist_init: bad world / rank / sizes");
  if ((rc = load_nccl())) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  SRG_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));   // collective: every rank calls it with the same id
  DistHandle *h = nullptr;
  if ((rc = make_handle(comm, true, world, rank, n, F, &h))) {
    g_nccl.CommDestroy(comm);
    return rc;
  }
  *out_handle = h;
  return SRG_OK;
}

extern "C" int srg_dist_init_comm(void *nccl_comm, int32_t world, int32_t rank, int64_t n, int32_t F,
                                  void **out_handle) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(nccl_comm && out_handle, "dist_init_comm: NULL pointer");
  SRG_REQUIRE(world >= 1 && rank >= 0 && rank < world && n >= 0 && F >= 0, "dist_init_comm: bad world / rank / sizes");
  if ((rc = load_nccl())) return rc;
  DistHandle *h = nullptr;
  if ((rc = make_handle(static_cast<ncclComm_t>(nccl_comm), false, world, rank, n, F, &h))) return rc;
  *out_handle = h;
  return SRG_OK;
}

extern "C" int srg_dist_partition(const void *handle, int64_t *row0, int64_t *n_local, int64_t *rows_per, int64_t *ld) {
  SRG_REQUIRE(handle != nullptr, "dist_partition: NULL handle");
  const DistHandle *h = static_cast<const DistHandle *>(handle);
  if (row0) *row0 = h->row0;
  if (n_local) *n_local = h->n_local;
  if (rows_per) *rows_per = h->rows_per;
  if (ld) *ld = h->ld;
  return SRG_OK;
}

extern "C" int srg_dist_destroy(void *handle) {
  if (!handle) return SRG_OK;
  DistHandle *h = static_cast<DistHandle *>(handle);
  cudaDeviceSynchronize();
  cudaFree(h->full[0]);
  cudaFree(h->full[1]);
  cudaFree(h->deg_all);
  if (h->own_comm && h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  delete h;
  return SRG_OK;
}

extern "C" int srg_dist_propagate(void *handle, const int32_t *indptr, const int32_t *indices, const void *data,
                                  int val_dtype, int64_t nnz, const float *x_local, int64_t ld_x, int32_t K, double r,
                                  double ppr_alpha, float *const *out_hops, int64_t ld_out, int32_t *flags,
                                  void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(handle != nullptr, "dist_propagate: NULL handle");
  DistHandle *h = static_cast<DistHandle *>(handle);
  SRG_REQUIRE(K >= 0 && nnz >= 0, "dist_propagate: negative size");
  SRG_REQUIRE(indptr && flags, "dist_propagate: NULL pointer");
  SRG_REQUIRE(val_dtype >= 0 && val_dtype <= 2, "dist_propagate: bad val_dtype");
  SRG_REQUIRE(nnz == 0 || indices, "dist_propagate: indices is NULL");
  SRG_REQUIRE(val_dtype == SRG_VAL_ONES || nnz == 0 || data, "dist_propagate: data is NULL");
  SRG_REQUIRE(h->n_local * (int64_t)h->F == 0 || (x_local && ld_x >= h->F), "dist_propagate: bad feature slice");
  SRG_REQUIRE(!out_hops || ld_out >= h->F, "dist_propagate: ld_out smaller than F");
  SRG_REQUIRE(nnz + h->n_local <= 2147483647LL, "dist_propagate: nnz + n_local exceeds the int32 CSR range");
  if (h->n == 0 || h->F == 0) return SRG_OK;
  cudaStream_t s = as_stream(stream);
  const int64_t n_loc = h->n_local, cap = std::max<int64_t>(nnz + n_loc, 1), ld = h->ld;
  const int F = h->F;

  // Every rank issues the SAME sequence of collectives whatever happens locally: a rank whose local stage failed
  // keeps taking part (its slices are then meaningless, the error code tells), so no peer is left waiting inside an
  // ncclAllGather.  Scratch is scoped: every return path hands its blocks back.
  StreamScratch scratch(s);
  int32_t *at_indptr = nullptr, *at_indices = nullptr;
  double *at_val = nullptr, *dl = nullptr, *dr = nullptr;
  float *val32 = nullptr;
  rc = scratch.alloc(&at_indptr, (size_t)(n_loc + 1));
  if (!rc) rc = scratch.alloc(&at_indices, (size_t)cap);
  if (!rc && val_dtype != SRG_VAL_ONES) rc = scratch.alloc(&at_val, (size_t)cap);
  if (!rc) rc = scratch.alloc(&dl, (size_t)std::max<int64_t>(h->n_pad, 1));
  if (!rc) rc = scratch.alloc(&dr, (size_t)std::max<int64_t>(h->n_pad, 1));
  if (!rc) rc = scratch.alloc(&val32, (size_t)cap);
  auto keep = [&rc](int r2) {
    if (!rc && r2) rc = r2;
  };
  auto cuda_rc = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess) keep(cuda_fail(e, what, __FILE__, __LINE__));
  };
  auto gather = [&](const void *send, void *recv, size_t count, ncclDataType_t dt) {
    if (h->world <= 1) return;
    const ncclResult_t e = g_nccl.AllGather(send, recv, count, dt, h->comm, s);
    if (e != ncclSuccess) {
      if (!rc) {
        set_err("ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
        rc = SRG_ERR_CUDA;
      }
    }
  };

  // ---- normalisation of the local rows: structure, degrees, ONE all-gather of the degree vector, values ------
  double *deg_loc = h->deg_all + (int64_t)h->rank * h->rows_per;
  if (!rc) keep(srg_selfloop_rows_csr(indptr, indices, data, val_dtype, n_loc, nnz, h->row0, h->n, at_indptr, nullptr, flags, s));
  if (!rc)
    keep(srg_selfloop_fill_rows_csr(indptr, indices, data, val_dtype, n_loc, nnz, h->row0, h->n, at_indptr, at_indices,
                                    at_val, deg_loc, flags, s));
  gather(deg_loc, h->deg_all, (size_t)h->rows_per, ncclDouble);   // in place
  if (!rc) keep(srg_pow_tables_f64(h->deg_all, h->n_pad, r, dl, dr, s));
  if (!rc)
    keep(srg_norm_values_rows_csr(at_indptr, at_indices, at_val, deg_loc, n_loc, nnz + n_loc, h->row0, h->n_pad, dl, dr,
                                  ppr_alpha, 0, nullptr, val32, flags, s));

  // ---- hops: slice into the full buffer, all-gather, local SpMM ------------------------------------------------
  int cur = 0;
  if (!rc && n_loc > 0)
    cuda_rc(cudaMemcpy2DAsync(h->full[0] + h->row0 * ld, (size_t)ld * 4, x_local, (size_t)ld_x * 4, (size_t)F * 4,
                              (size_t)n_loc, cudaMemcpyDeviceToDevice, s), "cudaMemcpy2DAsync");
  if (!rc && out_hops && out_hops[0] && n_loc > 0)
    cuda_rc(cudaMemcpy2DAsync(out_hops[0], (size_t)ld_out * 4, x_local, (size_t)ld_x * 4, (size_t)F * 4, (size_t)n_loc,
                              cudaMemcpyDeviceToDevice, s), "cudaMemcpy2DAsync");
  const size_t slice = (size_t)h->rows_per * ld;
  gather(h->full[0] + (int64_t)h->rank * slice, h->full[0], slice, ncclFloat);
  for (int k = 1; k <= K; ++k) {
    const int nxt = cur ^ 1;
    float *y = h->full[nxt] + h->row0 * ld;
    if (!rc) keep(srg_spmm_csr_f32(at_indptr, at_indices, val32, n_loc, nnz + n_loc, h->full[cur], ld, y, ld, F, s));
    gather(h->full[nxt] + (int64_t)h->rank * slice, h->full[nxt], slice, ncclFloat);
    if (!rc && out_hops && out_hops[k] && n_loc > 0)
      cuda_rc(cudaMemcpy2DAsync(out_hops[k], (size_t)ld_out * 4, y, (size_t)ld * 4, (size_t)F * 4, (size_t)n_loc,
                                cudaMemcpyDeviceToDevice, s), "cudaMemcpy2DAsync");
    cur = nxt;
  }
  return rc;
}
