// augment.cu — device stages of the dataset augmentation that precedes the propagation.
//
// Reference: edge_augument, SSRG/data_augument.py:73-103 — every node whose endpoint count is below `degree_level`
// receives (degree_level - count) new neighbours: the closest, by L2 distance of the soft labels
// (compute_distance, SSRG/utils.py:35-38), among 100 x deficit candidates drawn with Python's `random.sample`
// (generate_numbers, SSRG/utils.py:29-33); the edge list is then symmetrised and de-duplicated (torch.unique).
// The candidate DRAWS are defined by the host RNG stream and stay on the host (augment.py); here:
//   srg_endpoint_counts_i64    Counter(cat(row, col)): count and first position of every node            (:75-81)
//   srg_candidate_topk_f32     distances to the candidates + the `deficit` closest, ascending            (:88-95)
//   srg_csr_to_edge_index_i64  the symmetrised duplicate-free edge list in torch.unique's (row, col) order (:97-102,
//                              on top of srg_edges_to_sym_csr)
#include <float.h>

#include "common.cuh"

namespace srg {

__global__ void __launch_bounds__(256)
endpoint_counts_kernel(const long long *__restrict__ row, const long long *__restrict__ col, long long m, long long n,
                       int *__restrict__ counts, long long *__restrict__ first_pos, int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * m) return;
  const long long v = (i < m) ? row[i] : col[i - m];     // position in cat(edge_row, edge_col)
  if (v < 0 || v >= n) {
    atomicOr(flags, SRG_FLAG_BAD_INDEX);
    return;
  }
  atomicAdd(counts + v, 1);
  atomicMin(reinterpret_cast<unsigned long long *>(first_pos + v), (unsigned long long)i);
}

__global__ void fill_i64_kernel(long long *p, long long n, long long v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// One warp per low-degree node.  distance_c = || soft[node] - soft[cand_c] ||_2: the differences are the reference's
// float32 subtractions; their squares are summed in double and the root is rounded to float32 (torch.norm returns
// float32), so the ORDER of the candidates is the reference's except between candidates whose float32 norms differ
// by a rounding of torch's own vectorised sum.  The k closest are written in ascending order, ties by position in
// the candidate list (a stable sort).
__global__ void __launch_bounds__(128)
candidate_topk_kernel(const float *__restrict__ soft, long long ld, int n_classes, const int *__restrict__ nodes,
                      const int *__restrict__ cand, const int *__restrict__ cand_cnt, const int *__restrict__ k_sel,
                      const int *__restrict__ out_off, int c_max, int n_low, long long *__restrict__ out_src,
                      long long *__restrict__ out_dst) {
  extern __shared__ float dist_sh[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + w;
  if (i >= n_low) return;
  float *dist = dist_sh + (size_t)w * c_max;
  const int node = nodes[i], cnt = cand_cnt[i], k = k_sel[i];
  const float *a = soft + (long long)node * ld;
  const int *cl = cand + (long long)i * c_max;
  for (int c = lane; c < cnt; c += 32) {
    const float *b = soft + (long long)cl[c] * ld;
    double s = 0.0;
    for (int f = 0; f < n_classes; ++f) {
      const float d = __fsub_rn(a[f], b[f]);
      s = __fma_rn((double)d, (double)d, s);
    }
    dist[c] = __double2float_rn(sqrt(s));
  }
  __syncwarp();
  for (int r = 0; r < k && r < cnt; ++r) {
    float best = FLT_MAX;
    int bi = 0x7fffffff;
    for (int c = lane; c < cnt; c += 32) {
      const float d = dist[c];
      if (d < best || (d == best && c < bi)) {
        best = d;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_src[out_off[i] + r] = node;
      out_dst[out_off[i] + r] = cl[bi];
      dist[bi] = FLT_MAX;        // taken (a real distance is never FLT_MAX: soft labels are probabilities)
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
csr_to_edge_index_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, long long n, long long nnz,
                         long long *__restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  for (int p = indptr[i] + lane; p < indptr[i + 1]; p += 32) {
    out[p] = i;
    out[nnz + p] = indices[p];
  }
}

}  // namespace srg

using namespace srg;

extern "C" int srg_endpoint_counts_i64(const int64_t *row, const int64_t *col, int64_t m, int64_t n, int32_t *counts,
                                       int64_t *first_pos, int32_t *flags, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(m >= 0 && n >= 0, "endpoint_counts: negative size");
  if (n == 0) return SRG_OK;
  SRG_REQUIRE(counts && first_pos && flags && (m == 0 || (row && col)), "endpoint_counts: NULL pointer");
  cudaStream_t s = as_stream(stream);
  SRG_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * sizeof(int32_t), s));
  fill_i64_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(reinterpret_cast<long long *>(first_pos), n, 0x7fffffffffffffffLL);
  SRG_LAUNCHED();
  if (m > 0) {
    endpoint_counts_kernel<<<(unsigned)ceil_div64(2 * m, 256), 256, 0, s>>>(
        reinterpret_cast<const long long *>(row), reinterpret_cast<const long long *>(col), m, n, counts,
        reinterpret_cast<long long *>(first_pos), flags);
    SRG_LAUNCHED();
  }
  return SRG_OK;
}

extern "C" int srg_candidate_topk_f32(const float *soft, int64_t ld, int64_t n, int32_t n_classes, const int32_t *nodes,
                                      const int32_t *cand, const int32_t *cand_cnt, const int32_t *k_sel,
                                      const int32_t *out_off, int32_t c_max, int32_t n_low, int64_t *out_src,
                                      int64_t *out_dst, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && n_classes >= 0 && c_max >= 0 && n_low >= 0 && ld >= n_classes, "candidate_topk: bad sizes");
  if (n_low == 0 || c_max == 0) return SRG_OK;
  SRG_REQUIRE(soft && nodes && cand && cand_cnt && k_sel && out_off && out_src && out_dst, "candidate_topk: NULL pointer");
  const int warps = 4;
  const size_t smem = (size_t)warps * c_max * sizeof(float);
  SRG_REQUIRE(smem <= 200 * 1024, "candidate_topk: more than %d candidates per node", (int)(200 * 1024 / warps / 4));
  SRG_CUDA(cudaFuncSetAttribute(candidate_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  candidate_topk_kernel<<<(unsigned)ceil_div64(n_low, warps), warps * 32, smem, as_stream(stream)>>>(
      soft, ld, n_classes, nodes, cand, cand_cnt, k_sel, out_off, c_max, n_low, reinterpret_cast<long long *>(out_src),
      reinterpret_cast<long long *>(out_dst));
  SRG_LAUNCHED();
  return SRG_OK;
}

extern "C" int srg_csr_to_edge_index_i64(const int32_t *indptr, const int32_t *indices, int64_t n, int64_t nnz,
                                         int64_t *out_edge_index, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && nnz >= 0, "csr_to_edge_index: negative size");
  if (n == 0 || nnz == 0) return SRG_OK;
  SRG_REQUIRE(indptr && indices && out_edge_index, "csr_to_edge_index: NULL pointer");
  csr_to_edge_index_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, as_stream(stream)>>>(
      indptr, indices, n, nnz, reinterpret_cast<long long *>(out_edge_index));
  SRG_LAUNCHED();
  return SRG_OK;
}
