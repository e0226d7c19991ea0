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

// Vector form for the device layout (ld % 4 == 0, <= 8 hop matrices, F <= 512): every lane keeps its
// float4 columns of ALL hop rows in registers, so the row is read from HBM exactly once with every load
// in flight at the same time and the weighted sum needs no second pass over memory.  The 2 * hops partial
// sums (dot_j, |x_j|^2) are reduced together by a halving butterfly (V values -> V/2 -> ... -> 1, then
// plain xor steps): 9 shuffles instead of 40 for 4 hops, and each total lands in its own lane group, so
// sqrt / divide / exp run ONCE for all hops (every hop in a different lane) instead of once per hop.
template <int HALF>
__device__ __forceinline__ void butterfly_step(float *a, bool up, int o) {
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const float send = up ? a[i] : a[i + HALF];
    const float keep = up ? a[i + HALF] : a[i];
    a[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, o));
  }
}

template <int VPL, int HMAX>
__global__ void __launch_bounds__(256)
nafs_combine_vec_kernel(NafsHops hops, int n_hops, long long ld4, long long n, int nvec, int F,
                        float4 *__restrict__ out, long long ldo4, float *__restrict__ weights_out) {
  static_assert(HMAX == 4 || HMAX == 8, "hop capacity");
  constexpr int V = 2 * HMAX;              // values reduced together
  constexpr int SH = (V == 8) ? 2 : 1;     // total of value i ends up in lanes [i << SH, (i + 1) << SH)
  constexpr unsigned FULL = 0xffffffffu;
  const long long row = (long long)blockIdx.x * 8 + threadIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x;
  float4 v[HMAX][VPL];
#pragma unroll
  for (int j = 0; j < HMAX; ++j) {
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      const int c = lane + 32 * p;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < n_hops && c < nvec) {
        t = __ldcs(reinterpret_cast<const float4 *>(hops.p[j]) + row * ld4 + c);
        const int e = F - 4 * c;               // columns >= F (tail of the last float4) do not take part
        if (e < 4) t.w = 0.f;
        if (e < 3) t.z = 0.f;
        if (e < 2) t.y = 0.f;
      }
      v[j][p] = t;
    }
  }
  // per-lane partials: a[2j] = x0 . xj, a[2j + 1] = xj . xj
  float a[V];
#pragma unroll
  for (int j = 0; j < HMAX; ++j) {
    float dot = 0.f, sq = 0.f;
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      const float4 x = v[0][p], b = v[j][p];
      dot = __fadd_rn(dot, __fmul_rn(x.x, b.x));
      dot = __fadd_rn(dot, __fmul_rn(x.y, b.y));
      dot = __fadd_rn(dot, __fmul_rn(x.z, b.z));
      dot = __fadd_rn(dot, __fmul_rn(x.w, b.w));
      sq = __fadd_rn(sq, __fmul_rn(b.x, b.x));
      sq = __fadd_rn(sq, __fmul_rn(b.y, b.y));
      sq = __fadd_rn(sq, __fmul_rn(b.z, b.z));
      sq = __fadd_rn(sq, __fmul_rn(b.w, b.w));
    }
    a[2 * j] = dot;
    a[2 * j + 1] = sq;
  }
  if (V == 16) butterfly_step<V / 2>(a, lane & 16, 16);
  butterfly_step<4>(a, lane & (V == 16 ? 8 : 16), V == 16 ? 8 : 16);
  butterfly_step<2>(a, lane & (V == 16 ? 4 : 8), V == 16 ? 4 : 8);
  butterfly_step<1>(a, lane & (V == 16 ? 2 : 4), V == 16 ? 2 : 4);
  float tot = a[0];
  if (V == 8) tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, 2));
  tot = __fadd_rn(tot, __shfl_xor_sync(FULL, tot, 1));
  // lanes of group 2j hold dot_j; |x_j|^2 sits one group up, |x_0|^2 in group 1
  const float sq_j = __shfl_sync(FULL, tot, (lane + (1 << SH)) & 31);
  const float sq_0 = __shfl_sync(FULL, tot, 1 << SH);
  const float norm_cur = __fadd_rn(__fsqrt_rn(sq_j), 1e-10f);
  const float norm_fea = __fadd_rn(__fsqrt_rn(sq_0), 1e-10f);
  const float score = __fdiv_rn(__fdiv_rn(tot, norm_cur), norm_fea);   // meaningful in the even groups
  float smax = -INFINITY;
#pragma unroll
  for (int j = 0; j < HMAX; ++j)
    if (j < n_hops) smax = fmaxf(smax, __shfl_sync(FULL, score, (2 * j) << SH));
  const float ex = expf(__fsub_rn(score, smax));
  float denom = 0.f;
#pragma unroll
  for (int j = 0; j < HMAX; ++j)
    if (j < n_hops) denom = __fadd_rn(denom, __shfl_sync(FULL, ex, (2 * j) << SH));
  const float wgt = __fdiv_rn(ex, denom);
  float w[HMAX];
#pragma unroll
  for (int j = 0; j < HMAX; ++j) {
    w[j] = 0.f;
    if (j < n_hops) {
      w[j] = __shfl_sync(FULL, wgt, (2 * j) << SH);
      if (weights_out && lane == ((2 * j) << SH)) weights_out[row * n_hops + j] = wgt;
    }
  }
#pragma unroll
  for (int p = 0; p < VPL; ++p) {
    const int c = lane + 32 * p;
    if (c >= nvec) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < HMAX; ++j)
      if (j < n_hops) {
        acc.x = __fadd_rn(acc.x, __fmul_rn(w[j], v[j][p].x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(w[j], v[j][p].y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(w[j], v[j][p].z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(w[j], v[j][p].w));
      }
    __stcs(out + row * ldo4 + c, acc);
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
  bool vec = (ld % 4 == 0) && (ld_out % 4 == 0) && ((uintptr_t)out % 16 == 0) && n_hops <= 8 && F <= 512;
  for (int j = 0; j < n_hops && vec; ++j) vec = ((uintptr_t)hops[j] % 16 == 0);
  if (vec) {
    const int nvec = (F + 3) / 4;
    float4 *o4 = reinterpret_cast<float4 *>(out);
    const dim3 blk(32, 8);
    cudaStream_t st = as_stream(stream);
#define SRG_NAFS_LAUNCH(VPL_, H_) \
  nafs_combine_vec_kernel<VPL_, H_><<<(unsigned)blocks, blk, 0, st>>>(hp, n_hops, ld / 4, n, nvec, F, o4, ld_out / 4, weights_out)
    if (n_hops <= 4) {
      if (nvec <= 32) SRG_NAFS_LAUNCH(1, 4);
      else if (nvec <= 64) SRG_NAFS_LAUNCH(2, 4);
      else SRG_NAFS_LAUNCH(4, 4);
    } else {
      if (nvec <= 32) SRG_NAFS_LAUNCH(1, 8);
      else if (nvec <= 64) SRG_NAFS_LAUNCH(2, 8);
      else SRG_NAFS_LAUNCH(4, 8);
    }
#undef SRG_NAFS_LAUNCH
    SRG_LAUNCHED();
    return SRG_OK;
  }
  nafs_combine_kernel<<<(unsigned)blocks, dim3(32, 8), 0, as_stream(stream)>>>(hp, n_hops, ld, n, F, out, ld_out,
                                                                              weights_out);
  SRG_LAUNCHED();
  return SRG_OK;
}
