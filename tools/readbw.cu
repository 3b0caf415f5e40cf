// Micro-benchmark behind the colsum_kernel design: how fast can one B200 stream a
// [rows][16384] fp32 matrix from HBM into per-column fp64 sums?
//   v0  flat grid-stride float4 read, fp32 sum                      (read ceiling, no structure)
//   v1  column-tile x row-range CTAs, float4 __ldcs, fp64 sums      (the shipped access pattern)
//   v2  same decomposition, rows staged by cp.async.bulk (1-D TMA) into an mbarrier ring
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o readbw tools/readbw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) v0_flat(const float4 *__restrict__ x, size_t n4, float *out)
{
    float s = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll 8
    for (; i < n4; i += stride) { float4 v = __ldcs(x + i); s += (v.x + v.y) + (v.z + v.w); }
    if (s == 123.456f) out[0] = s;
}

template <int U>
__global__ void __launch_bounds__(256) v1_tile(const float *__restrict__ x, int rows, int N, int parts, double *out)
{
    const int tile = blockIdx.x % (N / 1024), part = blockIdx.x / (N / 1024);
    const int r0 = (int)((long long)rows * part / parts), r1 = (int)((long long)rows * (part + 1) / parts);
    const float *p = x + (size_t)r0 * N + tile * 1024 + threadIdx.x * 4;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int r = r0;
    for (; r + U <= r1; r += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldcs(reinterpret_cast<const float4 *>(p + (size_t)u * N));
#pragma unroll
        for (int u = 0; u < U; u++) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
        p += (size_t)U * N;
    }
    for (; r < r1; r++) { float4 v = __ldcs(reinterpret_cast<const float4 *>(p)); a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; p += N; }
    double *o = out + ((size_t)part * N + tile * 1024 + threadIdx.x * 4);
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
}

// ---- v2: bulk-copy ring ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int R, int S>
__global__ void __launch_bounds__(288) v2_bulk(const float *__restrict__ x, int rows, int N, int parts, double *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *buf = reinterpret_cast<float *>(smem);                    // [S][R][1024]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * R * 4096), *empty = full + S;
    const int tile = blockIdx.x % (N / 1024), part = blockIdx.x / (N / 1024);
    const int r0 = (int)((long long)rows * part / parts), r1 = (int)((long long)rows * (part + 1) / parts);
    const int nst = (r1 - r0 + R - 1) / R;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float *src = x + (size_t)r0 * N + tile * 1024;
    if (threadIdx.x >= 256) {                                        // producer warp
        if (threadIdx.x == 256) {
            for (int i = 0; i < nst; i++) {
                const int s = i % S;
                if (i >= S) mbar_wait(&empty[s], ((i / S) - 1) & 1);
                const int nr = min(R, r1 - r0 - i * R);
                mbar_expect_tx(&full[s], nr * 4096);
                for (int r = 0; r < nr; r++) bulk_g2s(buf + ((size_t)s * R + r) * 1024, src + (size_t)(i * R + r) * N, 4096, &full[s]);
            }
        }
        return;
    }
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int i = 0; i < nst; i++) {
        const int s = i % S;
        mbar_wait(&full[s], (i / S) & 1);
        const int nr = min(R, r1 - r0 - i * R);
        const float4 *b4 = reinterpret_cast<const float4 *>(buf + (size_t)s * R * 1024) + threadIdx.x;
        if (nr == R) {
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = b4[r * 256];
#pragma unroll
            for (int r = 0; r < R; r++) { a0 += v[r].x; a1 += v[r].y; a2 += v[r].z; a3 += v[r].w; }
        } else {
            for (int r = 0; r < nr; r++) { float4 v = b4[r * 256]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
    }
    double *o = out + ((size_t)part * N + tile * 1024 + threadIdx.x * 4);
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
}

__global__ void fill_kernel(float *x, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned int h = (unsigned int)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        x[i] = (float)(int)(h & 0xffffff) * (1.0f / 8388608.0f) - 1.0f;
    }
}

template <typename F>
static float timeit(F f, int reps = 20)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) f();
    std::vector<float> t;
    for (int i = 0; i < reps; i++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main()
{
    const int rows = 8192, N = 16384;
    const size_t n = (size_t)rows * N;
    float *x; double *out; float *o32;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&out, (size_t)256 * N * 8)); CK(cudaMalloc(&o32, 64));
    if (getenv("ZERO")) CK(cudaMemset(x, 0, n * 4)); else fill_kernel<<<148 * 8, 256>>>(x, n);
    CK(cudaDeviceSynchronize());
    auto rep = [&](const char *name, float ms) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, n * 4 / (ms * 1e-3) / 1e9); };
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32})
        { char nm[64]; snprintf(nm, 64, "v0 flat float4, grid %d", g); rep(nm, timeit([&] { v0_flat<<<g, 256>>>((const float4 *)x, n / 4, o32); })); }
    for (int parts : {37, 55, 74, 111, 148}) {
        char nm[64];
        snprintf(nm, 64, "v1 tile U=4 parts=%d (grid %d)", parts, parts * 16); rep(nm, timeit([&] { v1_tile<4><<<parts * 16, 256>>>(x, rows, N, parts, out); }));
        snprintf(nm, 64, "v1 tile U=8 parts=%d (grid %d)", parts, parts * 16); rep(nm, timeit([&] { v1_tile<8><<<parts * 16, 256>>>(x, rows, N, parts, out); }));
    }
#define RUN_V2(R, S, parts) do { \
        const int smem = S * R * 4096 + 2 * S * 8; \
        CK(cudaFuncSetAttribute(v2_bulk<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        char nm[64]; snprintf(nm, 64, "v2 bulk R=%d S=%d parts=%d (grid %d, %d KB)", R, S, parts, parts * 16, smem / 1024); \
        rep(nm, timeit([&] { v2_bulk<R, S><<<parts * 16, 288, smem>>>(x, rows, N, parts, out); })); } while (0)
    RUN_V2(2, 4, 37); RUN_V2(2, 8, 37); RUN_V2(4, 4, 37); RUN_V2(4, 3, 37);
    RUN_V2(4, 8, 18); RUN_V2(4, 12, 9); RUN_V2(8, 6, 9); RUN_V2(2, 4, 55); RUN_V2(2, 3, 74); RUN_V2(4, 4, 18); RUN_V2(1, 8, 37); RUN_V2(1, 6, 74);
    return 0;
}
