/*
 * tspws_mma.cu -- the resampling contraction on the 5th-generation tensor cores (sm_100a only).
 *
 * Single-stage bootstrap / sub-sampling / jackknife (reference src/ts_pws1f_lib.c:528-565): for every
 * realisation m the phase stack is  PS_m[i] = sum_t c[m][t] * phasor_t[i]  with c the selection (0/1) or
 * count (with-replacement bootstrap) matrix -- a dense [M x T] . [T x 2 ncoef] contraction (1.05 TFLOP at
 * BASELINE configs[3]).  ncu of the SIMT kernel it replaces (profiles/r02_select_sum_ncu.csv): 4.6 TFLOP/s,
 * 10 % of DRAM, tensor pipe 0 % -- contraction bound, so it belongs on tcgen05:
 *
 *   A = counts        [M  x T] bf16 (exact: integers <= 256)
 *   B = phasor matrix [2 ncoef x T] as NTERM bf16 terms (p = hi + lo [+ lo2]; 2 terms: 2^-18 per component)
 *   D = PS            [M x 2 ncoef] fp32, accumulated in TMEM
 *
 * Operands live in global memory as TILE IMAGES: 128 rows x 64 bf16 (one 128-byte swizzle atom along K) stored
 * exactly as the K-major SWIZZLE_128B shared-memory layout tcgen05.mma reads, so a tile is ONE 16 KiB bulk copy
 * (cp.async.bulk -> UBLKCP, the TMA engine's linear mode) completing on an mbarrier -- no tensor maps needed.
 *
 * Kernel = the canonical warp-specialised shape: warp 0 / lane 0 bulk-copy producer, warp 1 / lane 0 issues
 * tcgen05.mma (accumulators in TMEM, tcgen05.commit frees the stage), warps 2..5 drain TMEM with tcgen05.ld.
 */
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tspws_internal.h"

#define MMA_TILE_MN   128                    /* rows of a tile image (UMMA M and N)           */
#define MMA_TILE_K    64                     /* bf16 per row = one 128-byte swizzle atom      */
#define MMA_TILE_BYTES (MMA_TILE_MN * MMA_TILE_K * 2)   /* 16 KiB                            */
#define MMA_UMMA_K    16                     /* bf16 per tcgen05.mma                          */
#define MMA_MT        2                      /* M tiles (accumulators) per CTA                */
#define MMA_STAGES(NTERM) ((NTERM) == 2 ? 3 : 2)   /* 3 x 64 KiB or 2 x 80 KiB of the 227 KiB */
#define MMA_THREADS   192

/* ---- element (row r, column k) of a tile image: 16-byte chunk k / 8 goes to chunk position (k / 8) ^ (r % 8) ---- */
__host__ __device__ __forceinline__ unsigned mma_img_offset(const unsigned r, const unsigned k)
{
    return r * MMA_TILE_K + ((((k >> 3) ^ (r & 7u)) << 3) | (k & 7u));      /* in bf16 elements */
}

/* ---- PTX wrappers ---- */
__device__ __forceinline__ unsigned m_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void m_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void m_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(m_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void m_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(m_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(m_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void m_tmem_alloc(uint32_t *smem_dst, unsigned ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(m_smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void m_tmem_dealloc(uint32_t taddr, unsigned ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void m_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void m_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
/* D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32, one elected thread */
__device__ __forceinline__ void m_umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
/* all tcgen05.mma issued so far by this thread complete -> one arrival on the mbarrier (implies fence::before_thread_sync) */
__device__ __forceinline__ void m_umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(m_smem_u32(bar)) : "memory");
}
/* 32 lanes x 32 columns of fp32: thread = TMEM lane, v[j] = column j */
__device__ __forceinline__ void m_tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* K-major SWIZZLE_128B shared-memory matrix descriptor of a tile image at `p` (1024-byte aligned), advanced by
 * `kbytes` along K inside the swizzle atom: start address >> 4 | LBO (unused for swizzled K-major) |
 * SBO = 8 rows x 128 B = 1024 B >> 4 at bit 32 | version 1 at bit 46 | layout type 2 (SWIZZLE_128B) at bit 61 */
__device__ __forceinline__ uint64_t m_smem_desc(const void *p, unsigned kbytes)
{
    const uint64_t addr = (uint64_t)((m_smem_u32(p) + kbytes) >> 4) & 0x3fffu;
    return addr | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
/* instruction descriptor: D fp32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24 */
#define MMA_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MMA_TILE_MN >> 3) << 17) | ((uint32_t)(MMA_TILE_MN >> 4) << 24))

/* ===================================================================================== */
/* D[m][n] = sum_k A[m][k] * sum_term B_term[n][k]                                        */
/*   a_img: [mtiles][kblocks] tile images; b_img: [ntiles][kblocks][NTERM] tile images   */
/*   out:   [M][ld_out] fp32, columns [0, N)                                             */
/*   grid = (ceil(mtiles / MMA_MT), ntiles): M tile pairs fastest, so the CTAs that share */
/*   a B column block run together and its second..nth read hits L2                       */
/* ===================================================================================== */
template <int NTERM>
__global__ void __launch_bounds__(MMA_THREADS, 1)
resample_mma_kernel(const __nv_bfloat16 *__restrict__ a_img, const __nv_bfloat16 *__restrict__ b_img,
                    int mtiles, int kblocks, int M, int N, float *__restrict__ out, size_t ld_out)
{
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    constexpr int STAGE_TILES = MMA_MT + NTERM;
    constexpr int STAGES = MMA_STAGES(NTERM);
    constexpr unsigned STAGE_BYTES = STAGE_TILES * MMA_TILE_BYTES;
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(mma_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t *full  = reinterpret_cast<uint64_t *>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *done  = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt0 = blockIdx.x * MMA_MT, nt = blockIdx.y;
    const int nmt = min(MMA_MT, mtiles - mt0);
    constexpr unsigned TMEM_COLS = MMA_MT * MMA_TILE_MN;        /* 256: a power of two >= 32 */

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { m_mbar_init(full + s, 1); m_mbar_init(empty + s, 1); }
        m_mbar_init(done, 1);
        m_fence_barrier_init();
    } else if (warp == 2) {
        m_tmem_alloc(tmem_slot, TMEM_COLS);
    }
    m_tc_fence_before();
    __syncthreads();
    m_tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        /* ---- producer: one bulk copy per tile image ---- */
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES;
            if (kb >= STAGES) m_mbar_wait(empty + s, ((kb / STAGES) & 1) ^ 1);
            unsigned char *st = tiles + (size_t)s * STAGE_BYTES;
            m_mbar_expect_tx(full + s, (unsigned)(nmt + NTERM) * MMA_TILE_BYTES);
            for (int i = 0; i < nmt; i++)
                m_bulk_g2s(st + (size_t)i * MMA_TILE_BYTES, a_img + ((size_t)(mt0 + i) * kblocks + kb) * (MMA_TILE_MN * MMA_TILE_K),
                           MMA_TILE_BYTES, full + s);
            for (int t = 0; t < NTERM; t++)
                m_bulk_g2s(st + (size_t)(MMA_MT + t) * MMA_TILE_BYTES,
                           b_img + (((size_t)nt * kblocks + kb) * NTERM + t) * (MMA_TILE_MN * MMA_TILE_K), MMA_TILE_BYTES, full + s);
        }
    } else if (warp == 1 && lane == 0) {
        /* ---- MMA issuer ---- */
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % STAGES;
            m_mbar_wait(full + s, (kb / STAGES) & 1);
            m_tc_fence_after();
            const unsigned char *st = tiles + (size_t)s * STAGE_BYTES;
            for (int i = 0; i < nmt; i++)
                for (int t = 0; t < NTERM; t++)
#pragma unroll
                    for (int k = 0; k < MMA_TILE_K / MMA_UMMA_K; k++)
                        m_umma_bf16(tmem + (uint32_t)i * MMA_TILE_MN,
                                    m_smem_desc(st + (size_t)i * MMA_TILE_BYTES, k * MMA_UMMA_K * 2),
                                    m_smem_desc(st + (size_t)(MMA_MT + t) * MMA_TILE_BYTES, k * MMA_UMMA_K * 2),
                                    MMA_IDESC, (kb | t | k) != 0);
            m_umma_commit(empty + s);                    /* the stage is free once these MMAs have read it */
        }
        m_umma_commit(done);                             /* accumulators complete */
    } else if (warp >= 2) {
        /* ---- epilogue: TMEM lane quarter (warp % 4) -> rows of the M tile; 32 columns per tcgen05.ld ---- */
        m_mbar_wait(done, 0);
        m_tc_fence_after();
        const int q = warp & 3;
        for (int i = 0; i < nmt; i++) {
            const int m = (mt0 + i) * MMA_TILE_MN + q * 32 + lane;
            for (int c0 = 0; c0 < MMA_TILE_MN; c0 += 32) {
                uint32_t v[32];
                m_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * MMA_TILE_MN + c0), v);
                const int n0 = nt * MMA_TILE_MN + c0;
                if (m < M && n0 < N) {
                    float *o = out + (size_t)m * ld_out + n0;
                    if (n0 + 32 <= N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (n0 + j < N) o[j] = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }
    m_tc_fence_before();
    __syncthreads();
    if (warp == 2) m_tmem_dealloc(tmem, TMEM_COLS);
}

static size_t mma_smem_bytes(int nterm)
{
    return (size_t)MMA_STAGES(nterm) * (MMA_MT + nterm) * MMA_TILE_BYTES + 1024 /* alignment */ + (2 * MMA_STAGES(nterm) + 1) * 8 + 16;
}

/* out[m][n] (m < M, n < N) = sum over the tile images; mtiles = ceil(M / 128), ntiles = ceil(N / 128) */
extern "C" int tspws_cu_resample_mma(const void *d_a_img, const void *d_b_img, int nterm, int M, int N, int kblocks,
                                     float *d_out, size_t ld_out, tspws_stream_t st)
{
    if (M <= 0 || N <= 0 || kblocks <= 0) return 0;
    if (nterm != 2 && nterm != 3) return (int)cudaErrorInvalidValue;
    const int mtiles = (M + MMA_TILE_MN - 1) / MMA_TILE_MN, ntiles = (N + MMA_TILE_MN - 1) / MMA_TILE_MN;
    dim3 grid((mtiles + MMA_MT - 1) / MMA_MT, ntiles);
    const size_t smem = mma_smem_bytes(nterm);
    cudaError_t e;
    if (nterm == 2) {
        if ((e = cudaFuncSetAttribute(resample_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        resample_mma_kernel<2><<<grid, MMA_THREADS, smem, (cudaStream_t)st>>>((const __nv_bfloat16 *)d_a_img, (const __nv_bfloat16 *)d_b_img,
                                                                               mtiles, kblocks, M, N, d_out, ld_out);
    } else {
        if ((e = cudaFuncSetAttribute(resample_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        resample_mma_kernel<3><<<grid, MMA_THREADS, smem, (cudaStream_t)st>>>((const __nv_bfloat16 *)d_a_img, (const __nv_bfloat16 *)d_b_img,
                                                                               mtiles, kblocks, M, N, d_out, ld_out);
    }
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* ===================================================================================== */
/* tile images                                                                           */
/* ===================================================================================== */
/* counts [M][ld] (signed char, >= 0; traces [t0, t0 + T)) -> A image [mtiles][kblocks]; zero padded */
__global__ void __launch_bounds__(256)
count_image_kernel(const signed char *__restrict__ cnt, int ld, int M, int T, int kblocks, __nv_bfloat16 *__restrict__ img)
{
    const int mt = blockIdx.y, kb = blockIdx.x;
    __nv_bfloat16 *tile = img + ((size_t)mt * kblocks + kb) * (MMA_TILE_MN * MMA_TILE_K);
    for (int e = threadIdx.x; e < MMA_TILE_MN * MMA_TILE_K; e += blockDim.x) {
        const int r = e / MMA_TILE_K, k = e % MMA_TILE_K;
        const int m = mt * MMA_TILE_MN + r, t = kb * MMA_TILE_K + k;
        float v = 0.f;
        if (m < M && t < T) { const int c = cnt[(size_t)m * ld + t]; v = c > 0 ? (float)c : 0.f; }
        tile[mma_img_offset(r, k)] = __float2bfloat16_rn(v);
    }
}

extern "C" int tspws_cu_count_image(const signed char *d_cnt, int ld, int M, int T, void *d_img, tspws_stream_t st)
{
    if (M <= 0 || T <= 0) return 0;
    const int mtiles = (M + MMA_TILE_MN - 1) / MMA_TILE_MN, kblocks = (T + MMA_TILE_K - 1) / MMA_TILE_K;
    count_image_kernel<<<dim3(kblocks, mtiles), 256, 0, (cudaStream_t)st>>>(d_cnt, ld, M, T, kblocks, (__nv_bfloat16 *)d_img);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* unit phasors P [nt][ncoef] float2 of traces [t0, t0 + nt) (t0 % 64 == 0) -> B image rows n = 2 i + (re | im), K = trace,
 * NTERM bf16 terms p = hi + lo (+ lo2).  CTA = one (n tile, k block): the 64 x 64 block of P goes through shared memory
 * so that both the global reads (along i) and the image writes (16 KiB per term, linear) are coalesced. */
template <int NTERM>
__global__ void __launch_bounds__(256)
phasor_image_kernel(const float2 *__restrict__ P, int ncoef, int nt, int t0, int kblocks_total, __nv_bfloat16 *__restrict__ img)
{
    __shared__ __align__(16) __nv_bfloat16 s_img[NTERM][MMA_TILE_MN * MMA_TILE_K];
    const int ntile = blockIdx.y, kbl = blockIdx.x;                /* k block within this batch */
    const int i0 = ntile * (MMA_TILE_MN / 2);
    for (int e = threadIdx.x; e < (MMA_TILE_MN / 2) * MMA_TILE_K; e += blockDim.x) {
        const int tl = e / (MMA_TILE_MN / 2), il = e % (MMA_TILE_MN / 2);   /* consecutive threads: consecutive coefficients */
        const int t = kbl * MMA_TILE_K + tl, i = i0 + il;
        float2 p = make_float2(0.f, 0.f);
        if (t < nt && i < ncoef) p = P[(size_t)t * ncoef + i];
        float re = p.x, im = p.y;
#pragma unroll
        for (int term = 0; term < NTERM; term++) {
            const __nv_bfloat16 hr = __float2bfloat16_rn(re), hi = __float2bfloat16_rn(im);
            s_img[term][mma_img_offset(2 * il, tl)]     = hr;
            s_img[term][mma_img_offset(2 * il + 1, tl)] = hi;
            re -= __bfloat162float(hr); im -= __bfloat162float(hi);
        }
    }
    __syncthreads();
    const int kb = t0 / MMA_TILE_K + kbl;
    uint4 *dst = reinterpret_cast<uint4 *>(img + (((size_t)ntile * kblocks_total + kb) * NTERM) * (MMA_TILE_MN * MMA_TILE_K));
    const uint4 *src = reinterpret_cast<const uint4 *>(&s_img[0][0]);
    for (int e = threadIdx.x; e < NTERM * MMA_TILE_BYTES / 16; e += blockDim.x) dst[e] = src[e];
}

extern "C" int tspws_cu_phasor_image(const void *d_P, int ncoef, int nt, int t0, int kblocks_total, int nterm, void *d_img,
                                     tspws_stream_t st)
{
    if (nt <= 0 || ncoef <= 0) return 0;
    if ((t0 % MMA_TILE_K) != 0 || (nterm != 2 && nterm != 3)) return (int)cudaErrorInvalidValue;
    const int ntiles = (2 * ncoef + MMA_TILE_MN - 1) / MMA_TILE_MN, kbl = (nt + MMA_TILE_K - 1) / MMA_TILE_K;
    dim3 grid(kbl, ntiles);
    if (nterm == 2) phasor_image_kernel<2><<<grid, 256, 0, (cudaStream_t)st>>>((const float2 *)d_P, ncoef, nt, t0, kblocks_total, (__nv_bfloat16 *)d_img);
    else            phasor_image_kernel<3><<<grid, 256, 0, (cudaStream_t)st>>>((const float2 *)d_P, ncoef, nt, t0, kblocks_total, (__nv_bfloat16 *)d_img);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

extern "C" size_t tspws_cu_mma_a_image_bytes(int M, int T)
{
    return (size_t)((M + MMA_TILE_MN - 1) / MMA_TILE_MN) * ((T + MMA_TILE_K - 1) / MMA_TILE_K) * MMA_TILE_BYTES;
}
extern "C" size_t tspws_cu_mma_b_image_bytes(int N, int T, int nterm)
{
    return (size_t)((N + MMA_TILE_MN - 1) / MMA_TILE_MN) * ((T + MMA_TILE_K - 1) / MMA_TILE_K) * nterm * MMA_TILE_BYTES;
}

/* Self-test hook (tests/test_gpu_mma.py): counts [M][T] int8 and P [T][ncoef] float2 on the host -> PS [M][2 ncoef] fp32
 * through count_image / phasor_image / resample_mma.  Returns 0 or a CUDA error. */
extern "C" int tspws_cu_mma_selftest(const signed char *h_cnt, const float *h_P, int M, int T, int ncoef, int nterm, float *h_out)
{
    signed char *d_cnt = nullptr; float *d_P = nullptr, *d_out = nullptr; void *d_a = nullptr, *d_b = nullptr;
    const int N = 2 * ncoef, kblocks = (T + MMA_TILE_K - 1) / MMA_TILE_K;
    int rc = 0;
#define ST(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto out; } } while (0)
    ST(cudaMalloc((void **)&d_cnt, (size_t)M * T));
    ST(cudaMalloc((void **)&d_P, (size_t)T * ncoef * 8));
    ST(cudaMalloc((void **)&d_out, (size_t)M * N * 4));
    ST(cudaMalloc(&d_a, tspws_cu_mma_a_image_bytes(M, T)));
    ST(cudaMalloc(&d_b, tspws_cu_mma_b_image_bytes(N, T, nterm)));
    ST(cudaMemcpy(d_cnt, h_cnt, (size_t)M * T, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(d_P, h_P, (size_t)T * ncoef * 8, cudaMemcpyHostToDevice));
    ST(cudaMemset(d_b, 0, tspws_cu_mma_b_image_bytes(N, T, nterm)));
    ST(cudaMemset(d_out, 0xff, (size_t)M * N * 4));
    if ((rc = tspws_cu_count_image(d_cnt, T, M, T, d_a, nullptr))) goto out;
    for (int t0 = 0; t0 < T; t0 += 128) {               /* batches like the resampling driver's */
        const int nt = T - t0 < 128 ? T - t0 : 128;
        if ((rc = tspws_cu_phasor_image(d_P + (size_t)t0 * ncoef * 2, ncoef, nt, t0, kblocks, nterm, d_b, nullptr))) goto out;
    }
    if ((rc = tspws_cu_resample_mma(d_a, d_b, nterm, M, N, kblocks, d_out, (size_t)N, nullptr))) goto out;
    ST(cudaDeviceSynchronize());
    ST(cudaMemcpy(h_out, d_out, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
#undef ST
out:
    cudaFree(d_cnt); cudaFree(d_P); cudaFree(d_out); cudaFree(d_a); cudaFree(d_b);
    return rc;
}
