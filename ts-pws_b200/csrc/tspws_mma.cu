/*
 * tspws_mma.cu -- the tcgen05 / TMEM kernels (sm_100a only):
 *   1. the resampling contraction (resample_mma_kernel), described right below;
 *   2. the octaves with decimation D >= 16 of the single-stage forward transform as dense (D >= 256) and banded
 *      (16 <= D < 256) GEMMs with a phasor epilogue -- second half of the file (fwd_dense_mma_kernel<BAND>,
 *      fwd_mma_persistent_kernel, trace_image_kernel, w_image_kernel, band_w_image_kernel).
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
#include <stdlib.h>

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
    tspws_cu_profile_launched("resample_mma_kernel", st);
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
    tspws_cu_profile_launched("count_image_kernel", st);
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
    tspws_cu_profile_launched("phasor_image_kernel", st);
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

/* ===================================================================================== */
/* Coarse scales of the single-stage forward transform as ONE dense GEMM on tcgen05       */
/*                                                                                       */
/* For the scales with few, long filters (D >= 256: L = 2560 .. N taps, Ns <= 64 coefficients)  */
/* the transform of a batch of traces (src/FWTa/wavelet_v7.c:43-64 -> cdotx_dc, cdotx.c:35-72)  */
/* is a plain matrix product  Y[t][(s,m,c)] = sum_n x_t[n] W[(s,m,c)][n],  W = the conjugated  */
/* filter of scale s placed (circularly) at coefficient m -- [traces x N] . [N x 2 ncoarse].    */
/* On the SIMT polyphase kernel these octaves were 4/7 of the D >= 32 time.  Here:              */
/*   A = traces      [mtr x N]  as THREE bf16 terms x = x0 + x1 + x2 (24 mantissa bits: fp32-exact)  */
/*   B = W           [2 ncoarse x N] as three bf16 terms (built once per plan on the device)     */
/*   D = x0 W0 + x0 W1 + x1 W0 + x1 W1 + x0 W2 + x2 W0   (everything above 2^-25), fp32 in TMEM  */
/* Two terms (2^-17) are NOT enough: a trace whose coefficient happens to be ~1e-4 of the typical     */
/* magnitude has an ill-conditioned phase, and one such (trace, coefficient) pair in 150 000 moved   */
/* the coherence by 1.2e-4 (tools/dense_check2.py); fp32-faithful products keep it at the SIMT      */
/* path's 2e-6.                                                                                    */
/* and the epilogue is the phasor accumulate of src/ts_pws1f_lib.c:486-494: unit phasor of     */
/* every (trace, coefficient), summed over the 128 traces of the M tile (warp transpose-reduce   */
/* + shared memory) into one partial row per M tile; the rows are added in tile order later      */
/* (deterministic).  The per-trace coefficients never leave the SM.                              */
/* ===================================================================================== */
#define FD_STAGES 2
#define FD_NT 3                                          /* bf16 terms per operand */
#define FD_STAGE_BYTES (2 * FD_NT * MMA_TILE_BYTES)    /* A terms 0..2, B terms 0..2 */

__device__ __forceinline__ void fd_phasor(float re, float im, float &pr, float &pi)
{
    const float s = re * re + im * im;
    if (s > 1e-30f && s < 1e30f) { const float inv = rsqrtf(s); pr = re * inv; pi = im * inv; }
    else {
        const float a = hypotf(re, im);
        if (a > 0.f && a <= 3.0e38f) { pr = re / a; pi = im / a; } else { pr = 0.f; pi = 0.f; }   /* skipped like :492 */
    }
}

/* one stage of the transposing butterfly: lanes l and l ^ HALF exchange halves of the 2 HALF live values */
template <int HALF>
__device__ __forceinline__ void fd_transpose_stage(float (&f)[32], const int lane)
{
    const bool up = (lane & HALF) != 0;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
        const float send = up ? f[i] : f[i + HALF];
        const float keep = up ? f[i + HALF] : f[i];
        f[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
    }
}

/* The tensor core adds into its fp32 accumulator with truncation, not round-to-nearest: measured on B200
 * (tools/mma_acc_check.py) the result of a K = 16384 accumulation is biased by -1.7e-5 relative, linearly in K
 * -- 8e-5 on the coherence of the coarse scales.  So the K loop is cut into chunks of FD_CHUNK k blocks (512 samples:
 * bias < 1e-6): every chunk accumulates into one of TWO TMEM accumulators, and the epilogue warps drain the finished
 * one into fp32 registers (round-to-nearest adds) while the next chunk's MMAs run -- the promotion trick of FP8 GEMMs. */
#define FD_CHUNK 8
/* BAND = false: dense GEMM over all kblocks row blocks, column tile blockIdx.x of w_img, output columns col0 + ...
 * BAND = true : banded GEMM of one (octave, group) = blockIdx.x (see tspws_band_t): nj row blocks starting at
 *               group * rows_per_group + qmin (circular), B tiles tile_base .. tile_base + nj of w_img, UMMA N = ncols */
template <bool BAND>
__global__ void __launch_bounds__(MMA_THREADS, 1)
fwd_dense_mma_kernel(const __nv_bfloat16 *__restrict__ x_img, const __nv_bfloat16 *__restrict__ w_img,
                     int kblocks_total, int ncol, float *__restrict__ part /* [mtiles][ld_part] */, int ld_part, int col0,
                     int term_mask, const tspws_band_t *__restrict__ bands)
{
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(mma_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t *full   = reinterpret_cast<uint64_t *>(tiles + (size_t)FD_STAGES * FD_STAGE_BYTES);
    uint64_t *empty  = full + FD_STAGES;
    uint64_t *tfull  = empty + FD_STAGES;       /* [2] accumulator complete */
    uint64_t *tempty = tfull + 2;               /* [2] accumulator drained (4 epilogue warps) */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    float (*colsum)[MMA_TILE_MN] = reinterpret_cast<float (*)[MMA_TILE_MN]>(tiles);   /* [4 warps][128 columns]: reuses stage 0 after the MMAs */

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y;                           /* column tiles fastest: the CTAs that share an A tile run together */
    int nt = blockIdx.x, kblocks = kblocks_total, row0 = 0, btile0 = 0, ncols = MMA_TILE_MN, grp = 0, ob = 0;
    if (BAND) {
        grp = blockIdx.x;
        while (grp >= bands[ob].ngroups) { grp -= bands[ob].ngroups; ob++; }
        kblocks = bands[ob].nj; ncols = bands[ob].ncols; btile0 = bands[ob].tile_base;
        row0 = (grp * bands[ob].rows_per_group + bands[ob].qmin) % kblocks_total;
        if (row0 < 0) row0 += kblocks_total;
        nt = 0;
    }
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(MMA_TILE_MN >> 4) << 24);
    constexpr unsigned TMEM_COLS = 2 * MMA_TILE_MN;
    const int nchunk = (kblocks + FD_CHUNK - 1) / FD_CHUNK;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < FD_STAGES; s++) { m_mbar_init(full + s, 1); m_mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; s++) { m_mbar_init(tfull + s, 1); m_mbar_init(tempty + s, 4); }
        m_fence_barrier_init();
    } else if (warp == 2) {
        m_tmem_alloc(tmem_slot, TMEM_COLS);
    }
    m_tc_fence_before();
    __syncthreads();
    m_tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < kblocks; kb++) {
            const int s = kb % FD_STAGES;
            if (kb >= FD_STAGES) m_mbar_wait(empty + s, ((kb / FD_STAGES) & 1) ^ 1);
            unsigned char *st = tiles + (size_t)s * FD_STAGE_BYTES;
            m_mbar_expect_tx(full + s, BAND ? FD_NT * MMA_TILE_BYTES + FD_NT * ncols * 128 : FD_STAGE_BYTES);
            /* images: [tile][kb][term] -> the terms of a tile are one contiguous 48 KiB copy */
            int ka = row0 + kb;                          /* BAND: circular row block of the trace; nj <= kblocks_total */
            if (ka >= kblocks_total) ka -= kblocks_total;
            const size_t kw = BAND ? (size_t)(btile0 + kb) : (size_t)nt * kblocks + kb;
            m_bulk_g2s(st, x_img + ((size_t)mt * kblocks_total + ka) * FD_NT * (MMA_TILE_MN * MMA_TILE_K), FD_NT * MMA_TILE_BYTES, full + s);
            if (BAND) {                                  /* only the ncols rows (128 B each) the MMA reads, per term */
                for (int t = 0; t < FD_NT; t++)
                    m_bulk_g2s(st + (FD_NT + t) * MMA_TILE_BYTES, w_img + (kw * FD_NT + t) * (MMA_TILE_MN * MMA_TILE_K), ncols * 128, full + s);
            } else
                m_bulk_g2s(st + FD_NT * MMA_TILE_BYTES, w_img + kw * FD_NT * (MMA_TILE_MN * MMA_TILE_K), FD_NT * MMA_TILE_BYTES, full + s);
        }
    } else if (warp == 1 && lane == 0) {
        for (int c = 0; c < nchunk; c++) {
            const int buf = c & 1;
            if (c >= 2) { m_mbar_wait(tempty + buf, ((c >> 1) & 1) ^ 1); m_tc_fence_after(); }
            const uint32_t acc = tmem + (uint32_t)buf * MMA_TILE_MN;
            const int kb1 = min(kblocks, (c + 1) * FD_CHUNK);
            for (int kb = c * FD_CHUNK; kb < kb1; kb++) {
                const int s = kb % FD_STAGES;
                m_mbar_wait(full + s, (kb / FD_STAGES) & 1);
                m_tc_fence_after();
                const unsigned char *st = tiles + (size_t)s * FD_STAGE_BYTES;
                const unsigned char *a = st, *b = st + FD_NT * MMA_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < MMA_TILE_K / MMA_UMMA_K; k++) {
                    const unsigned kby = k * MMA_UMMA_K * 2;
                    /* smallest products first: they are added while the accumulator is still small */
                    bool first = (kb == c * FD_CHUNK) && k == 0;
#define FD_MMA(I, J) do { if (term_mask & (1 << (I + J))) { m_umma_bf16(acc, m_smem_desc(a + (I) * MMA_TILE_BYTES, kby), m_smem_desc(b + (J) * MMA_TILE_BYTES, kby), idesc, !first); first = false; } } while (0)
                    FD_MMA(2, 0); FD_MMA(0, 2); FD_MMA(1, 1); FD_MMA(1, 0); FD_MMA(0, 1); FD_MMA(0, 0);
#undef FD_MMA
                }
                m_umma_commit(empty + s);
            }
            m_umma_commit(tfull + buf);
        }
    } else if (warp >= 2) {
        /* ---- epilogue warps: thread = trace (TMEM lane); drain every finished chunk into registers ---- */
        const int q = warp & 3;
        float y[MMA_TILE_MN];
#pragma unroll
        for (int j = 0; j < MMA_TILE_MN; j++) y[j] = 0.f;
        for (int c = 0; c < nchunk; c++) {
            const int buf = c & 1;
            m_mbar_wait(tfull + buf, (c >> 1) & 1);
            m_tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < MMA_TILE_MN; c0 += 32) {
                if (BAND && c0 >= ncols) break;          /* columns the MMA never wrote */
                uint32_t v[32];
                m_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MMA_TILE_MN + c0), v);
#pragma unroll
                for (int j = 0; j < 32; j++) y[c0 + j] += __uint_as_float(v[j]);
            }
            m_tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m_smem_u32(tempty + buf)) : "memory");
        }
        /* phasor of every (trace, coefficient), then the sum over this warp's 32 traces */
#pragma unroll
        for (int c0 = 0; c0 < MMA_TILE_MN; c0 += 32) {
            if (BAND && c0 >= ncols) break;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) fd_phasor(y[c0 + j], y[c0 + j + 1], f[j], f[j + 1]);
            /* transposing butterfly over the 32 lanes (= traces): lane j ends with the sum of column j */
            fd_transpose_stage<16>(f, lane); fd_transpose_stage<8>(f, lane); fd_transpose_stage<4>(f, lane);
            fd_transpose_stage<2>(f, lane);  fd_transpose_stage<1>(f, lane);
            colsum[q][c0 + lane] = f[0];
        }
    }
    m_tc_fence_before();
    __syncthreads();
    if (threadIdx.x < MMA_TILE_MN) {
        const float sum = ((colsum[2][threadIdx.x] + colsum[3][threadIdx.x]) + colsum[0][threadIdx.x]) + colsum[1][threadIdx.x];
        if (BAND) {
            /* column = ((v Pn + p') U + u) 2 + (re | im)  ->  coefficient m = (grp Pn + p') U + u of voice v */
            const tspws_band_t *bd = bands + ob;
            const int col = threadIdx.x, t2 = col >> 1, U = bd->U, Pn = bd->Pn;
            const int u = t2 % U, t3 = t2 / U, pp = t3 % Pn, v = t3 / Pn;
            const int m = (grp * Pn + pp) * U + u;
            if (col < ncols && v < bd->nv && m < bd->Ns)
                part[(size_t)mt * ld_part + 2 * (bd->coef_rel[v] + m) + (col & 1)] = sum;
        } else {
            const int col = nt * MMA_TILE_MN + threadIdx.x;
            if (col < ncol) part[(size_t)mt * ld_part + col0 + col] = sum;
        }
    }
    if (warp == 2) m_tmem_dealloc(tmem, TMEM_COLS);
}

/* ---- persistent variant: ONE launch for all dense and banded jobs of a stack ------------------------------------
 * One CTA per SM walks the job list j = blockIdx.x, + gridDim.x, ...; jobs are ordered M tile by M tile (the CTAs that
 * run at the same time share their A tiles through L2), within an M tile the dense column tiles (256 row blocks each)
 * come first, then the (octave, group) items of the banded GEMMs.  The three roles keep running counters of row blocks
 * and chunks across jobs, so the smem ring and the two TMEM accumulators never drain between jobs: while the epilogue
 * warps turn the last chunk of job i into phasors and column sums, the producer and the MMA thread are already two
 * chunks into job i + 1.  (The one-job-per-CTA launches spent half of the band's time outside the tensor pipe.) */
struct fd_job {
    int mt, kblocks, row0, ncols, ob, grp, nt;
    size_t wtile0;                      /* first B tile (3-term triple) of the job in its image */
    bool band;
};
__device__ __forceinline__ fd_job fd_get_job(int j, int jobs_per_mt, int ndense, int kblocks_total, const tspws_band_t *__restrict__ bands)
{
    fd_job q;
    q.mt = j / jobs_per_mt;
    int r = j - q.mt * jobs_per_mt;
    q.band = r >= ndense;
    q.ob = 0; q.grp = 0; q.nt = 0;
    if (!q.band) {
        q.nt = r; q.kblocks = kblocks_total; q.row0 = 0; q.ncols = MMA_TILE_MN;
        q.wtile0 = (size_t)r * kblocks_total;
    } else {
        r -= ndense;
        while (r >= bands[q.ob].ngroups) { r -= bands[q.ob].ngroups; q.ob++; }
        q.grp = r;
        q.kblocks = bands[q.ob].nj; q.ncols = bands[q.ob].ncols; q.wtile0 = (size_t)bands[q.ob].tile_base;
        q.row0 = (r * bands[q.ob].rows_per_group + bands[q.ob].qmin) % kblocks_total;
        if (q.row0 < 0) q.row0 += kblocks_total;
    }
    return q;
}

__global__ void __launch_bounds__(MMA_THREADS, 1)
fwd_mma_persistent_kernel(const __nv_bfloat16 *__restrict__ x_img, const __nv_bfloat16 *__restrict__ w_img /* dense */,
                          const __nv_bfloat16 *__restrict__ b_img /* band */, int kblocks_total, int ncol_dense, int ndense,
                          int nband_groups, int mtiles, float *__restrict__ part /* [mtiles][ld_part] */, int ld_part, int col0_dense,
                          int term_mask, const tspws_band_t *__restrict__ bands)
{
    extern __shared__ __align__(1024) unsigned char mma_smem[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(mma_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t *full   = reinterpret_cast<uint64_t *>(tiles + (size_t)FD_STAGES * FD_STAGE_BYTES);
    uint64_t *empty  = full + FD_STAGES;
    uint64_t *tfull  = empty + FD_STAGES;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    float (*colsum)[MMA_TILE_MN] = reinterpret_cast<float (*)[MMA_TILE_MN]>(tmem_slot + 4);     /* [4 warps][128 columns], its own 2 KiB */

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr unsigned TMEM_COLS = 2 * MMA_TILE_MN;
    const int jobs_per_mt = ndense + nband_groups, njobs = jobs_per_mt * mtiles;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < FD_STAGES; s++) { m_mbar_init(full + s, 1); m_mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; s++) { m_mbar_init(tfull + s, 1); m_mbar_init(tempty + s, 4); }
        m_fence_barrier_init();
    } else if (warp == 2) {
        m_tmem_alloc(tmem_slot, TMEM_COLS);
    }
    m_tc_fence_before();
    __syncthreads();
    m_tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        /* ---- producer ---- */
        unsigned kc = 0;
        for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
            const fd_job q = fd_get_job(j, jobs_per_mt, ndense, kblocks_total, bands);
            const __nv_bfloat16 *wsrc = q.band ? b_img : w_img;
            const unsigned bbytes = q.band ? (unsigned)q.ncols * 128u : (unsigned)MMA_TILE_BYTES;
            for (int kb = 0; kb < q.kblocks; kb++, kc++) {
                const unsigned s = kc % FD_STAGES;
                if (kc >= FD_STAGES) m_mbar_wait(empty + s, ((kc / FD_STAGES) & 1) ^ 1);
                unsigned char *st = tiles + (size_t)s * FD_STAGE_BYTES;
                m_mbar_expect_tx(full + s, FD_NT * MMA_TILE_BYTES + FD_NT * bbytes);
                int ka = q.row0 + kb;
                if (ka >= kblocks_total) ka -= kblocks_total;
                m_bulk_g2s(st, x_img + ((size_t)q.mt * kblocks_total + ka) * FD_NT * (MMA_TILE_MN * MMA_TILE_K), FD_NT * MMA_TILE_BYTES, full + s);
                for (int t = 0; t < FD_NT; t++)
                    m_bulk_g2s(st + (FD_NT + t) * MMA_TILE_BYTES, wsrc + ((q.wtile0 + kb) * FD_NT + t) * (MMA_TILE_MN * MMA_TILE_K), bbytes, full + s);
            }
        }
    } else if (warp == 1 && lane == 0) {
        /* ---- MMA issuer ---- */
        unsigned kc = 0, cc = 0;
        for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
            const fd_job q = fd_get_job(j, jobs_per_mt, ndense, kblocks_total, bands);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(q.ncols >> 3) << 17) | ((uint32_t)(MMA_TILE_MN >> 4) << 24);
            const int nchunk = (q.kblocks + FD_CHUNK - 1) / FD_CHUNK;
            for (int c = 0; c < nchunk; c++, cc++) {
                const unsigned buf = cc & 1;
                if (cc >= 2) { m_mbar_wait(tempty + buf, ((cc >> 1) & 1) ^ 1); m_tc_fence_after(); }
                const uint32_t acc = tmem + buf * MMA_TILE_MN;
                const int kb1 = min(q.kblocks, (c + 1) * FD_CHUNK);
                for (int kb = c * FD_CHUNK; kb < kb1; kb++, kc++) {
                    const unsigned s = kc % FD_STAGES;
                    m_mbar_wait(full + s, (kc / FD_STAGES) & 1);
                    m_tc_fence_after();
                    const unsigned char *st = tiles + (size_t)s * FD_STAGE_BYTES;
                    const unsigned char *a = st, *b = st + FD_NT * MMA_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < MMA_TILE_K / MMA_UMMA_K; k++) {
                        const unsigned kby = k * MMA_UMMA_K * 2;
                        bool first = (kb == c * FD_CHUNK) && k == 0;
#define FD_MMA(I, J) do { if (term_mask & (1 << (I + J))) { m_umma_bf16(acc, m_smem_desc(a + (I) * MMA_TILE_BYTES, kby), m_smem_desc(b + (J) * MMA_TILE_BYTES, kby), idesc, !first); first = false; } } while (0)
                        FD_MMA(2, 0); FD_MMA(0, 2); FD_MMA(1, 1); FD_MMA(1, 0); FD_MMA(0, 1); FD_MMA(0, 0);
#undef FD_MMA
                    }
                    m_umma_commit(empty + s);
                }
                m_umma_commit(tfull + buf);
            }
        }
    } else if (warp >= 2) {
        /* ---- epilogue warps: thread = trace (TMEM lane) ---- */
        const int q4 = warp & 3, e = threadIdx.x - 64;          /* e: 0..127 over the four epilogue warps */
        unsigned cc = 0;
        for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
            const fd_job q = fd_get_job(j, jobs_per_mt, ndense, kblocks_total, bands);
            const int nchunk = (q.kblocks + FD_CHUNK - 1) / FD_CHUNK;
            float y[MMA_TILE_MN];
#pragma unroll
            for (int i = 0; i < MMA_TILE_MN; i++) y[i] = 0.f;
            for (int c = 0; c < nchunk; c++, cc++) {
                const unsigned buf = cc & 1;
                m_mbar_wait(tfull + buf, (cc >> 1) & 1);
                m_tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < MMA_TILE_MN; c0 += 32) {
                    if (c0 >= q.ncols) break;
                    uint32_t v[32];
                    m_tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * MMA_TILE_MN + c0), v);
#pragma unroll
                    for (int i = 0; i < 32; i++) y[c0 + i] += __uint_as_float(v[i]);
                }
                m_tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m_smem_u32(tempty + buf)) : "memory");
            }
#pragma unroll
            for (int c0 = 0; c0 < MMA_TILE_MN; c0 += 32) {
                if (c0 >= q.ncols) break;
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) fd_phasor(y[c0 + i], y[c0 + i + 1], f[i], f[i + 1]);
                fd_transpose_stage<16>(f, lane); fd_transpose_stage<8>(f, lane); fd_transpose_stage<4>(f, lane);
                fd_transpose_stage<2>(f, lane);  fd_transpose_stage<1>(f, lane);
                colsum[q4][c0 + lane] = f[0];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e < q.ncols) {
                const float sum = ((colsum[2][e] + colsum[3][e]) + colsum[0][e]) + colsum[1][e];
                if (q.band) {
                    const tspws_band_t *bd = bands + q.ob;
                    const int t2 = e >> 1, U = bd->U, Pn = bd->Pn;
                    const int u = t2 % U, t3 = t2 / U, pp = t3 % Pn, v = t3 / Pn;
                    const int m = (q.grp * Pn + pp) * U + u;
                    if (v < bd->nv && m < bd->Ns)
                        part[(size_t)q.mt * ld_part + 2 * (bd->coef_rel[v] + m) + (e & 1)] = sum;
                } else {
                    const int col = q.nt * MMA_TILE_MN + e;
                    if (col < ncol_dense) part[(size_t)q.mt * ld_part + col0_dense + col] = sum;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");       /* colsum is free for the next job */
        }
    }
    m_tc_fence_before();
    __syncthreads();
    if (warp == 2) m_tmem_dealloc(tmem, TMEM_COLS);
}

/* traces x [ntr][N] fp32 -> A image [mtiles][kblocks][hi | lo], rows beyond ntr zero; and per M tile the column sums
 * of its traces (the linear stack of the coarse scales is the transform of the summed trace, by linearity) */
__global__ void __launch_bounds__(256)
trace_image_kernel(const float *__restrict__ x, int ntr, int N, int kblocks, __nv_bfloat16 *__restrict__ img, float *__restrict__ tsum /* [mtiles][N] */)
{
    extern __shared__ __align__(16) unsigned char ti_smem[];
    __nv_bfloat16 (*s_img)[MMA_TILE_MN * MMA_TILE_K] = reinterpret_cast<__nv_bfloat16 (*)[MMA_TILE_MN * MMA_TILE_K]>(ti_smem);
    __shared__ float s_sum[4][MMA_TILE_K];
    const int kb = blockIdx.x, mt = blockIdx.y;
    const int k = threadIdx.x & 63, r0 = threadIdx.x >> 6;             /* 4 rows per pass, 64 consecutive samples per row */
    float acc = 0.f;
    for (int r = r0; r < MMA_TILE_MN; r += 4) {
        const int t = mt * MMA_TILE_MN + r;
        const float v = t < ntr ? __ldg(x + (size_t)t * N + (size_t)kb * MMA_TILE_K + k) : 0.f;
        acc += v;
        const unsigned o = mma_img_offset(r, k);
        float rem = v;
#pragma unroll
        for (int term = 0; term < FD_NT; term++) {
            const __nv_bfloat16 h = __float2bfloat16_rn(rem);
            s_img[term][o] = h;
            rem -= __bfloat162float(h);
        }
    }
    s_sum[r0][k] = acc;
    __syncthreads();
    if (threadIdx.x < MMA_TILE_K)
        tsum[(size_t)mt * N + (size_t)kb * MMA_TILE_K + threadIdx.x] = (s_sum[0][threadIdx.x] + s_sum[1][threadIdx.x]) + (s_sum[2][threadIdx.x] + s_sum[3][threadIdx.x]);
    uint4 *dst = reinterpret_cast<uint4 *>(img + ((size_t)mt * kblocks + kb) * FD_NT * (MMA_TILE_MN * MMA_TILE_K));
    const uint4 *src = reinterpret_cast<const uint4 *>(&s_img[0][0]);
    for (int e = threadIdx.x; e < FD_NT * MMA_TILE_BYTES / 16; e += blockDim.x) dst[e] = src[e];
}

/* sum trace = sum over the M tiles (fixed order), fp64 accumulate -> fp32 */
__global__ void __launch_bounds__(256)
tile_sum_kernel(const float *__restrict__ tsum, int mtiles, int N, float *__restrict__ out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double a = 0.0;
    for (int m = 0; m < mtiles; m++) a += (double)tsum[(size_t)m * N + n];
    out[n] = (float)a;
}

/* PS of the dense coefficients = sum over the M tiles of the partial rows (fixed order); ST = the transform of the
 * summed trace.  Both go to chunk 0 of the per-chunk partial arrays (the other chunks hold zeros for these coefficients). */
__global__ void __launch_bounds__(256)
dense_finish_kernel(const float *__restrict__ part, int mtiles, int ncd, const float2 *__restrict__ ysum /* [ncoef] */, int ci0,
                    float2 *__restrict__ pps0, float2 *__restrict__ pst0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncd) return;
    float pr = 0.f, pi = 0.f;
    for (int m = 0; m < mtiles; m++) { pr += part[(size_t)m * 2 * ncd + 2 * i]; pi += part[(size_t)m * 2 * ncd + 2 * i + 1]; }
    pps0[ci0 + i] = make_float2(pr, pi);
    pst0[ci0 + i] = ysum[ci0 + i];
}

extern "C" size_t tspws_cu_dense_ximg_bytes(int ntr, int N)
{
    return (size_t)((ntr + MMA_TILE_MN - 1) / MMA_TILE_MN) * (N / MMA_TILE_K) * FD_NT * MMA_TILE_BYTES;
}
extern "C" size_t tspws_cu_dense_wimg_bytes(int ncol, int N)
{
    return (size_t)((ncol + MMA_TILE_MN - 1) / MMA_TILE_MN) * (N / MMA_TILE_K) * FD_NT * MMA_TILE_BYTES;
}

/* x [ntr][N] (N % 64 == 0) -> PS partial rows of the dense coefficients + the summed trace */
extern "C" int tspws_cu_forward_dense(const float *d_x, int ntr, int N, const void *d_wimg, int ncd, void *d_ximg,
                                      float *d_tsum /* [mtiles][N] */, float *d_sumtrace /* [N] */, float *d_part /* [mtiles][ld_part] */,
                                      int ld_part, int col0, tspws_stream_t st)
{
    if (ntr <= 0 || ncd <= 0) return 0;
    if (N % MMA_TILE_K) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)st;
    const int mtiles = (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN, kblocks = N / MMA_TILE_K, ncol = 2 * ncd;
    const int ntiles = (ncol + MMA_TILE_MN - 1) / MMA_TILE_MN;
    cudaFuncSetAttribute(trace_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FD_NT * MMA_TILE_BYTES);
    trace_image_kernel<<<dim3(kblocks, mtiles), 256, FD_NT * MMA_TILE_BYTES, s>>>(d_x, ntr, N, kblocks, (__nv_bfloat16 *)d_ximg, d_tsum);
    tspws_cu_profile_launched("trace_image_kernel", st);
    tile_sum_kernel<<<(N + 255) / 256, 256, 0, s>>>(d_tsum, mtiles, N, d_sumtrace);
    tspws_cu_profile_launched("tile_sum_kernel", st);
    const size_t smem = (size_t)FD_STAGES * FD_STAGE_BYTES + 1024 + (2 * FD_STAGES + 4) * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(fwd_dense_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fwd_dense_mma_kernel<false><<<dim3(ntiles, mtiles), MMA_THREADS, smem, s>>>((const __nv_bfloat16 *)d_ximg, (const __nv_bfloat16 *)d_wimg,
                                                                         kblocks, ncol, d_part, ld_part, col0,
                                                                         getenv("TSPWS_DENSE_TERMS") ? atoi(getenv("TSPWS_DENSE_TERMS")) : 7,   /* bit (i + j): products of terms i, j */
                                                                         nullptr);
    tspws_cu_profile_launched("fwd_dense_mma_kernel<dense>", st);
    tspws_cu_launch_count_add(3);
    return (int)cudaGetLastError();
}

/* dense + banded jobs in one persistent launch (d_ximg built by tspws_cu_forward_image); nbands may be 0 */
extern "C" int tspws_cu_forward_mma_all(const void *d_ximg, int ntr, int N, const void *d_wimg, int ncd_dense, const void *d_bimg,
                                        const tspws_band_t *d_bands, int nbands, int ngroups_total, float *d_part, int ld_part,
                                        int col0_dense, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    if (N % MMA_TILE_K) return (int)cudaErrorInvalidValue;
    const int mtiles = (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN, ncol = 2 * ncd_dense;
    const int ndense = (ncol + MMA_TILE_MN - 1) / MMA_TILE_MN;
    if (nbands <= 0) ngroups_total = 0;
    if (ndense + ngroups_total <= 0) return 0;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; } }
    const size_t smem = (size_t)FD_STAGES * FD_STAGE_BYTES + 1024 + (2 * FD_STAGES + 4) * 8 + 16 + 4 * MMA_TILE_MN * sizeof(float) + 16;
    cudaError_t e = cudaFuncSetAttribute(fwd_mma_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long njobs = (long long)(ndense + ngroups_total) * mtiles;
    const int grid = (int)(njobs < sms ? njobs : sms);
    fwd_mma_persistent_kernel<<<grid, MMA_THREADS, smem, (cudaStream_t)st>>>((const __nv_bfloat16 *)d_ximg, (const __nv_bfloat16 *)d_wimg,
        (const __nv_bfloat16 *)d_bimg, N / MMA_TILE_K, ncol, ndense, ngroups_total, mtiles, d_part, ld_part, col0_dense,
        getenv("TSPWS_DENSE_TERMS") ? atoi(getenv("TSPWS_DENSE_TERMS")) : 7, d_bands);
    tspws_cu_profile_launched("fwd_mma_persistent_kernel", st);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* the trace image + summed trace alone (first half of tspws_cu_forward_dense) */
extern "C" int tspws_cu_forward_image(const float *d_x, int ntr, int N, void *d_ximg, float *d_tsum, float *d_sumtrace, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    if (N % MMA_TILE_K) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)st;
    const int mtiles = (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN, kblocks = N / MMA_TILE_K;
    cudaFuncSetAttribute(trace_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FD_NT * MMA_TILE_BYTES);
    trace_image_kernel<<<dim3(kblocks, mtiles), 256, FD_NT * MMA_TILE_BYTES, s>>>(d_x, ntr, N, kblocks, (__nv_bfloat16 *)d_ximg, d_tsum);
    tspws_cu_profile_launched("trace_image_kernel", st);
    tspws_cu_launch_count_add(1);
    if (d_sumtrace) {                                   /* NULL: the caller sums the tiles later (tspws_cu_tile_sum), e.g. after several parts */
        tile_sum_kernel<<<(N + 255) / 256, 256, 0, s>>>(d_tsum, mtiles, N, d_sumtrace);
        tspws_cu_profile_launched("tile_sum_kernel", st);
        tspws_cu_launch_count_add(1);
    }
    return (int)cudaGetLastError();
}

extern "C" int tspws_cu_tile_sum(const float *d_tsum, int ntr, int N, float *d_sumtrace, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    tile_sum_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)st>>>(d_tsum, (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN, N, d_sumtrace);
    tspws_cu_profile_launched("tile_sum_kernel", st);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* middle octaves: banded GEMMs over the trace image tspws_cu_forward_dense built (call it first) */
extern "C" int tspws_cu_forward_band(const void *d_ximg, int ntr, int N, const void *d_bimg, const tspws_band_t *d_bands, int nbands,
                                     int ngroups_total, float *d_part, int ld_part, tspws_stream_t st)
{
    if (ntr <= 0 || nbands <= 0 || ngroups_total <= 0) return 0;
    if (N % MMA_TILE_K) return (int)cudaErrorInvalidValue;
    const int mtiles = (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN;
    const size_t smem = (size_t)FD_STAGES * FD_STAGE_BYTES + 1024 + (2 * FD_STAGES + 4) * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(fwd_dense_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    fwd_dense_mma_kernel<true><<<dim3(ngroups_total, mtiles), MMA_THREADS, smem, (cudaStream_t)st>>>(
        (const __nv_bfloat16 *)d_ximg, (const __nv_bfloat16 *)d_bimg, N / MMA_TILE_K, 0, d_part, ld_part, 0,
        getenv("TSPWS_DENSE_TERMS") ? atoi(getenv("TSPWS_DENSE_TERMS")) : 7, d_bands);
    tspws_cu_profile_launched("fwd_dense_mma_kernel<band>", st);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* B tiles of the banded GEMMs: tile (octave, jj), row = column of the product, element r of the 64-sample row block:
 *   q = jj + qmin - E p',  l = 64 q + r - u D + c_v,  value = conj(w_v[l]) for 0 <= l < L_v, else 0 */
__global__ void __launch_bounds__(256)
band_w_image_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ w32, const tspws_band_t *__restrict__ bands,
                    int nbands, __nv_bfloat16 *__restrict__ img)
{
    const int tile = blockIdx.x;
    int ob = 0;
    while (ob + 1 < nbands && bands[ob + 1].tile_base <= tile) ob++;
    const tspws_band_t bd = bands[ob];
    const int jj = tile - bd.tile_base;
    __nv_bfloat16 *t0 = img + (size_t)tile * FD_NT * (MMA_TILE_MN * MMA_TILE_K);
    for (int e = threadIdx.x; e < MMA_TILE_MN * MMA_TILE_K; e += blockDim.x) {
        const int col = e / MMA_TILE_K, r = e % MMA_TILE_K;
        float v = 0.f;
        const int t2 = col >> 1, u = t2 % bd.U, t3 = t2 / bd.U, pp = t3 % bd.Pn, vi = t3 / bd.Pn;
        if (col < bd.ncols && vi < bd.nv) {
            const tspws_scale_t sc = scales[bands[ob].s_of[vi]];
            const int q = jj + bd.qmin - bd.E * pp;
            const int l = 64 * q + r - u * bd.D + sc.c;
            if (l >= 0 && l < sc.L) {
                const float2 w = w32[sc.tap_off + l];
                v = (col & 1) ? -w.y : w.x;
            }
        }
        const unsigned o = mma_img_offset(col, r);
#pragma unroll
        for (int term = 0; term < FD_NT; term++) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            t0[(size_t)term * (MMA_TILE_MN * MMA_TILE_K) + o] = h;
            v -= __bfloat162float(h);
        }
    }
}

extern "C" int tspws_cu_band_wimage(const tspws_devframe_t *f, const tspws_band_t *d_bands, int nbands, int ntiles, void *d_img, tspws_stream_t st)
{
    if (nbands <= 0 || ntiles <= 0) return 0;
    band_w_image_kernel<<<ntiles, 256, 0, (cudaStream_t)st>>>(f->d_scale, (const float2 *)f->d_w32, d_bands, nbands, (__nv_bfloat16 *)d_img);
    return (int)cudaGetLastError();
}

extern "C" int tspws_cu_dense_finish(const float *d_part, int ntr, int ncd, const void *d_ysum, int ci0, void *d_pps0, void *d_pst0,
                                     tspws_stream_t st)
{
    if (ncd <= 0) return 0;
    const int mtiles = (ntr + MMA_TILE_MN - 1) / MMA_TILE_MN;
    dense_finish_kernel<<<(ncd + 255) / 256, 256, 0, (cudaStream_t)st>>>(d_part, mtiles, ncd, (const float2 *)d_ysum, ci0,
                                                                          (float2 *)d_pps0, (float2 *)d_pst0);
    tspws_cu_profile_launched("dense_finish_kernel", st);
    tspws_cu_launch_count_add(1);
    return (int)cudaGetLastError();
}

/* B image of the dense forward GEMM: row = column of the product = 2 (i - ci0) + (re | im) for the coefficients
 * i >= ci0 (the coarse scales, contiguous at the end of the ragged coefficient array); element n = the conjugated
 * filter of the coefficient's scale placed circularly at its position: W[n] = conj(w_s[(n - m D_s + c_s) mod N]). */
__global__ void __launch_bounds__(256)
w_image_kernel(const tspws_scale_t *__restrict__ scales, int S, int s0, const float2 *__restrict__ w32, int ci0, int ncd, int N, int kblocks,
               __nv_bfloat16 *__restrict__ img)
{
    const int kb = blockIdx.x, nt = blockIdx.y;
    __nv_bfloat16 *t0 = img + ((size_t)nt * kblocks + kb) * FD_NT * (MMA_TILE_MN * MMA_TILE_K);
    for (int e = threadIdx.x; e < MMA_TILE_MN * MMA_TILE_K; e += blockDim.x) {
        const int r = e / MMA_TILE_K, k = e % MMA_TILE_K;
        const int col = nt * MMA_TILE_MN + r, n = kb * MMA_TILE_K + k;
        float v = 0.f;
        if (col < 2 * ncd) {
            const int i = ci0 + (col >> 1);
            int s = s0;
            while (s + 1 < S && scales[s + 1].coef_off <= i) s++;
            const tspws_scale_t sc = scales[s];
            const int m = i - sc.coef_off;
            long long l = ((long long)n - (long long)m * sc.D + sc.c) % N;
            if (l < 0) l += N;
            if (l < sc.L) {
                const float2 w = w32[sc.tap_off + (int)l];
                v = (col & 1) ? -w.y : w.x;
            }
        }
        const unsigned o = mma_img_offset(r, k);
#pragma unroll
        for (int term = 0; term < FD_NT; term++) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            t0[(size_t)term * (MMA_TILE_MN * MMA_TILE_K) + o] = h;
            v -= __bfloat162float(h);
        }
    }
}

extern "C" int tspws_cu_dense_wimage(const tspws_devframe_t *f, int s0, int ci0, int ncd, void *d_img, tspws_stream_t st)
{
    if (ncd <= 0) return 0;
    if (f->N % MMA_TILE_K) return (int)cudaErrorInvalidValue;
    const int kblocks = f->N / MMA_TILE_K, ntiles = (2 * ncd + MMA_TILE_MN - 1) / MMA_TILE_MN;
    w_image_kernel<<<dim3(kblocks, ntiles), 256, 0, (cudaStream_t)st>>>(f->d_scale, f->S, s0, (const float2 *)f->d_w32, ci0, ncd, f->N, kblocks,
                                                                         (__nv_bfloat16 *)d_img);
    return (int)cudaGetLastError();
}
