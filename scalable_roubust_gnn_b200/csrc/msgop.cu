// msgop.cu — NAFS over-smoothing-distance aggregation of the hop list, on the device.
//
// Reference: OverSmoothDistanceWeightedOp.combine,
// SSRG/operators/message_operator/over_smooth_distance_op.py:11-33 (the aggregator of models/nafs.py:12):
//   score[i][j]  = ((x0[i] . xj[i]) / (||xj[i]||_2 + 1e-10)) / (||x0[i]||_2 + 1e-10)      (:13-19)
//   weight[i][:] = softmax_j(score[i][:])                                                   (:22)
//   out[i]       = 0. + sum_j weight[i][j] * xj[i]          (hop order, product then add)    (:27-31)
// The reference evaluates the last step as an O(N * hops) Python loop over rows; here one warp owns
// a row: pass 1 forms the hop scores (fp32 products and sums, warp tree reduction), the softmax is
// evaluated redundantly by every lane, pass 2 re-reads the row (L1/L2 hits) and writes the weighted sum.
// HBM-bound: hops * n * F * 4 bytes read + n * F * 4 written.
#include "common.cuh"

namespace srg {

constexpr int kNafsMaxHops = 64;
struct NafsHops {
  const float *p[kNafsMaxHops];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
nafs_combine_kernel(NafsHops hops, int n_hops, long long ld, long long n, int F, float *__restrict__ out,
                    long long ld_out, float *__restrict__ weights_out) {
  __shared__ float s_w[8][kNafsMaxHops];
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n) return;  // whole warps leave together (one row per warp)
  const int lane = threadIdx.x;
  float *w = s_w[threadIdx.y];
  const float *x0 = hops.p[0] + row * ld;

  // pass 1: scores (every lane ends up with the same value, lane 0 parks it in shared memory)
  float norm_fea = 0.f, smax = -INFINITY;
  for (int j = 0; j < n_hops; ++j) {
    const float *xj = hops.p[j] + row * ld;
    float dot = 0.f, sq = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float a = x0[c], b = xj[c];
      dot = __fadd_rn(dot, __fmul_rn(a, b));
      sq = __fadd_rn(sq, __fmul_rn(b, b));
    }
    dot = warp_sum(dot);
    sq = warp_sum(sq);
    const float norm_cur = __fadd_rn(__fsqrt_rn(sq), 1e-10f);
    if (j == 0) norm_fea = norm_cur;
    const float score = __fdiv_rn(__fdiv_rn(dot, norm_cur), norm_fea);
    smax = fmaxf(smax, score);
    if (lane == 0) w[j] = score;
  }
  __syncwarp();
  // softmax over the hops: exp(score - max) / sum
  float denom = 0.f;
  for (int j = 0; j < n_hops; ++j) denom = __fadd_rn(denom, expf(__fsub_rn(w[j], smax)));
  __syncwarp();
  if (lane == 0)
    for (int j = 0; j < n_hops; ++j) w[j] = __fdiv_rn(expf(__fsub_rn(w[j], smax)), denom);
  __syncwarp();
  if (weights_out && lane < n_hops)
    for (int j = lane; j < n_hops; j += 32) weights_out[row * n_hops + j] = w[j];

  // pass 2: out = 0. + w0 * x0 + w1 * x1 + ...   (sequential fp32, product rounded before the add)
  float *o = out + row * ld_out;
  for (int c = lane; c < F; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < n_hops; ++j) acc = __fadd_rn(acc, __fmul_rn(w[j], hops.p[j][row * ld + c]));
    o[c] = acc;
  }
}

}  // namespace srg

using namespace srg;

extern "C" int srg_nafs_combine_f32(const float *const *hops, int32_t n_hops, int64_t ld, int64_t n, int32_t F,
                                    float *out, int64_t ld_out, float *weights_out, void *stream) {
  int rc = require_device();
  if (rc) return rc;
  SRG_REQUIRE(n >= 0 && F >= 0, "nafs_combine: negative size");
  SRG_REQUIRE(n_hops >= 1, "nafs_combine: the hop list is empty");
  if (n_hops > kNafsMaxHops) {
    set_err("nafs_combine: %d hop matrices, at most %d are supported", n_hops, kNafsMaxHops);
    return SRG_ERR_UNSUPPORTED;
  }
  if (n == 0 || F == 0) return SRG_OK;
  SRG_REQUIRE(hops && out, "nafs_combine: NULL pointer");
  SRG_REQUIRE(ld >= F && ld_out >= F, "nafs_combine: leading dimension too small");
  NafsHops hp;
  for (int j = 0; j < n_hops; ++j) {
    SRG_REQUIRE(hops[j], "nafs_combine: hops[%d] is NULL", j);
    hp.p[j] = hops[j];
  }
  for (int j = n_hops; j < kNafsMaxHops; ++j) hp.p[j] = nullptr;
  const int64_t blocks = ceil_div64(n, 8);
  SRG_REQUIRE(blocks <= 2147483647LL, "nafs_combine: too many rows");
  nafs_combine_kernel<<<(unsigned)blocks, dim3(32, 8), 0, as_stream(stream)>>>(hp, n_hops, ld, n, F, out, ld_out,
                                                                              weights_out);
  SRG_LAUNCHED();
  return SRG_OK;
}
