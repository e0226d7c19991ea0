// Micro-benchmark: is a feature-sliced hop (X slice resident in L2) faster than the streaming hop?
//
// Products shape: N = 2 449 029 rows, 26 stored entries per row (uniform random columns), F = 100.
// The streaming kernel (csrc/spmm.cu) gathers 416-byte rows from DRAM: 27.2 GB per hop, 4.35 ms.
// Alternative measured here: cut X into slices of W floats stored slice-major ([N][W], 2.45M x 32 B
// = 78 MB for W = 8, inside the 126 MB L2), one pass over the CSR per slice, every gather an L2 hit.
// DRAM traffic per hop drops to passes x 0.51 GB (CSR) + 2 GB (X, Y once), but every pass re-walks
// the CSR and all gathers go through the L2 slices at sector granularity.
//
// Prints ms per pass and the projected hop time (ceil(100 / W) passes) for W = 4, 8, 16, 32.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_slice l2_slice.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int HINT>
__device__ __forceinline__ float4 ldx(const float4* p) {
    float4 v;
    if (HINT == 0) {
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    } else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    }
    return v;
}
template <int HINT>
__device__ __forceinline__ int2 ldcsr(const int2* p) {
    int2 v;
    if (HINT == 0) {
        asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    } else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    }
    return v;
}
template <int HINT>
__device__ __forceinline__ void sty(float4* p, float4 v) {
    if (HINT == 0) {
        *p = v;
    } else {
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
    }
}

// One thread per row, W floats of the slice per thread; the (index, value) pairs of the block's rows are
// staged in shared memory with coalesced loads (fixed degree DEG keeps the micro-benchmark simple).
template <int W, int HINT, int DEG, int ROWS>
__global__ void __launch_bounds__(ROWS) slice_pass(const int2* __restrict__ ent, const float* __restrict__ x,
                                                   float* __restrict__ y, int n) {
    __shared__ int2 s_ent[ROWS * DEG];
    const int row0 = blockIdx.x * ROWS;
    const int rows = min(ROWS, n - row0);
    const int2* src = ent + (size_t)row0 * DEG;
    for (int i = threadIdx.x; i < rows * DEG; i += ROWS) s_ent[i] = ldcsr<HINT>(src + i);
    __syncthreads();
    if ((int)threadIdx.x >= rows) return;
    constexpr int V = W / 4;
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int2* e = s_ent + threadIdx.x * DEG;
#pragma unroll 2
    for (int j = 0; j < DEG; ++j) {
        const int2 iv = e[j];
        const float a = __int_as_float(iv.y);
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)iv.x * W);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 t = ldx<HINT>(xr + v);
            acc[v].x = fmaf(a, t.x, acc[v].x);
            acc[v].y = fmaf(a, t.y, acc[v].y);
            acc[v].z = fmaf(a, t.z, acc[v].z);
            acc[v].w = fmaf(a, t.w, acc[v].w);
        }
    }
    float4* yr = reinterpret_cast<float4*>(y + (size_t)(row0 + threadIdx.x) * W);
#pragma unroll
    for (int v = 0; v < V; ++v) sty<HINT>(yr + v, acc[v]);
}

template <int W, int HINT>
static float run(const int2* ent, int n, int slices_alloc, float* xs, float* ys, int reps) {
    constexpr int DEG = 26, ROWS = 128;
    const int grid = (n + ROWS - 1) / ROWS;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    // rotate over distinct slices like a real hop would (each pass meets a cold slice)
    for (int i = 0; i < 3; ++i) {
        size_t off = (size_t)(i % slices_alloc) * n * W;
        slice_pass<W, HINT, DEG, ROWS><<<grid, ROWS>>>(ent, xs + off, ys + off, n);
    }
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) {
        size_t off = (size_t)(i % slices_alloc) * n * W;
        slice_pass<W, HINT, DEG, ROWS><<<grid, ROWS>>>(ent, xs + off, ys + off, n);
    }
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2449029;
    const int deg = 26, f = 100;
    const size_t nnz = (size_t)n * deg;
    std::vector<int2> h(nnz);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < nnz; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        h[i].x = (int)(s % (uint64_t)n);
        float v = 1.0f / 27.0f;
        h[i].y = *reinterpret_cast<int*>(&v);
    }
    int2* ent;
    CK(cudaMalloc(&ent, nnz * sizeof(int2)));
    CK(cudaMemcpy(ent, h.data(), nnz * sizeof(int2), cudaMemcpyHostToDevice));
    // whole X and Y (n x 104 floats each), viewed as a sequence of slice-major slices
    const size_t tot = (size_t)n * 104;
    float *xs, *ys;
    CK(cudaMalloc(&xs, tot * 4));
    CK(cudaMalloc(&ys, tot * 4));
    CK(cudaMemset(xs, 0, tot * 4));
    int l2 = 0;
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
    printf("n=%d nnz=%zu L2=%d MB\n", n, nnz, l2 >> 20);
    const int reps = 26;
#define RUN(W, H)                                                                                          \
    {                                                                                                      \
        float ms = run<W, H>(ent, n, 104 / W, xs, ys, reps);                                               \
        int passes = (f + W - 1) / W;                                                                      \
        printf("W=%2d hint=%d slice=%6.1f MB  pass %.4f ms  x %2d passes = %.3f ms per hop (stream kernel: 4.35)\n", W, H, \
               (double)n * W * 4 / 1e6, ms, passes, ms * passes);                                          \
    }
    RUN(4, 0) RUN(4, 1) RUN(8, 0) RUN(8, 1) RUN(16, 0) RUN(16, 1) RUN(32, 0) RUN(32, 1)
    return 0;
}
