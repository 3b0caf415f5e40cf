/*
 * tspws_kernels.cu -- sm_100a kernels of the ts-PWS stacking path and their C launchers.
 *
 * Kernel                    replaces (reference file:line)                         bound
 * ------------------------  -----------------------------------------------------  -------
 * prep_kernel               fold src/ts_pws1f_lib.c:76-86, mean removal :160-168   HBM
 * colsum_kernel             partial_linear_stacks src/ts_pws1f_lib.c:866-881       HBM  (headline)
 * rowmean/finalize/post     mean + fold applied to the Kmax group sums (linear)    tiny
 * fwd_kernel                complex_1D_wavelet_dec src/FWTa/wavelet_v7.c:43-64 ->
 *                           cdotx_dc src/FWTa/cdotx.c:35-72, fused with the phasor
 *                           accumulate of src/ts_pws1f_lib.c:486-494 / :897-904    FP32/LSU
 * reduce_weight_kernel      tspws_biased / tspws_unbiased src/ts_pws1f_lib.c:909-984 tiny
 * inv_kernel / inv_finish   Re_complex_1D_wavelet_rec src/FWTa/wavelet_v7.c:124-150
 *                           -> re_cdotx_upsampling_cc / re_cdotx_cc
 *                           src/FWTa/cdotx.c:305-340, :176-211; cast :233-241      L2/LSU
 *
 * Everything is written from the mathematics in SURVEY.md Appendix A.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "tspws_internal.h"

static unsigned long g_launches = 0;
extern "C" unsigned long tspws_cu_launch_count(void) { return g_launches; }

#define TSPWS_LAUNCHED()                              \
    do {                                              \
        g_launches++;                                 \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) return (int)e__;      \
    } while (0)

template <typename T> struct cplx;
template <> struct cplx<float>  { typedef float2  type; };
template <> struct cplx<double> { typedef double2 type; };

/* streaming (evict-first) loads: every trace sample is read exactly once */
__device__ __forceinline__ float4 ld_stream4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float  ld_stream1(const float *p) { return __ldcs(p); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* Block-wide sum of doubles (blockDim.x multiple of 32, <= 1024); result broadcast to all. */
__device__ __forceinline__ double block_sum(double v, double *sh /*[33]*/)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    double r = sh[32];
    __syncthreads();
    return r;
}

/* ===================================================================================== */
/* prep: in-place fold + mean removal of every trace, fp32-faithful to the reference     */
/* ===================================================================================== */
__global__ void __launch_bounds__(256)
prep_kernel(float *__restrict__ x, int N, int do_fold, int do_rm)
{
    __shared__ double sh[33];
    float *row = x + (size_t)blockIdx.x * N;
    double sum = 0.0;
    if (do_fold) {
        const int h = N / 2;
        for (int n = threadIdx.x; n < h; n += blockDim.x) {
            float a = __fadd_rn(row[n], row[N - 1 - n]);     /* float add, then exact halving */
            a *= 0.5f;
            row[n] = a;
            row[N - 1 - n] = a;
            sum += 2.0 * (double)a;
        }
        if ((N & 1) && threadIdx.x == 0) sum += (double)row[h];
    } else if (do_rm) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) sum += (double)row[n];
    }
    if (do_rm) {
        const double tot = block_sum(sum, sh);               /* also orders the fold stores */
        const float  m   = (float)(tot / (double)(unsigned int)N);
        for (int n = threadIdx.x; n < N; n += blockDim.x) row[n] = __fsub_rn(row[n], m);
    }
}

extern "C" int tspws_cu_prep(float *d_x, int ntr, int N, int do_fold, int do_rm, tspws_stream_t st)
{
    if (ntr <= 0 || (!do_fold && !do_rm)) return 0;
    prep_kernel<<<ntr, 256, 0, (cudaStream_t)st>>>(d_x, N, do_fold, do_rm);
    TSPWS_LAUNCHED();
    return 0;
}

/* ===================================================================================== */
/* colsum: fp32 traces -> fp64 partial column sums per work item (HBM-bound stream)      */
/*   grid.x = column tiles of 1024 samples, grid.y = items (trace ranges inside a group) */
/* ===================================================================================== */
#define CS_THREADS 256
#define CS_TILE    1024
#define CS_UNROLL  8

template <bool VEC, bool ROWSUM>
__global__ void __launch_bounds__(CS_THREADS)
colsum_kernel(const float *__restrict__ x, int N, const tspws_csum_item_t *__restrict__ items,
              double *__restrict__ part, float *__restrict__ rowpart, int rowpart_ld)
{
    const tspws_csum_item_t it = items[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;

    if (VEC) {
        const int  col    = blockIdx.x * CS_TILE + threadIdx.x * 4;
        const bool active = col < N;                       /* N % 4 == 0 on this path */
        const float *p = x + (size_t)it.t0 * N + (active ? col : 0);
        int t = it.t0;
        for (; t + CS_UNROLL <= it.t1; t += CS_UNROLL) {
            float4 v[CS_UNROLL];
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) v[j] = ld_stream4(p + (size_t)j * N);
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) {
                if (!active) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                a0 += (double)v[j].x; a1 += (double)v[j].y; a2 += (double)v[j].z; a3 += (double)v[j].w;
                if (ROWSUM) {
                    float r = warp_sum((v[j].x + v[j].y) + (v[j].z + v[j].w));
                    if (lane == 0) rowpart[(size_t)(t + j) * rowpart_ld + blockIdx.x * (CS_THREADS / 32) + warp] = r;
                }
            }
            p += (size_t)CS_UNROLL * N;
        }
        for (; t < it.t1; t++) {
            float4 v = ld_stream4(p);
            if (!active) v = make_float4(0.f, 0.f, 0.f, 0.f);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            if (ROWSUM) {
                float r = warp_sum((v.x + v.y) + (v.z + v.w));
                if (lane == 0) rowpart[(size_t)t * rowpart_ld + blockIdx.x * (CS_THREADS / 32) + warp] = r;
            }
            p += N;
        }
        if (active) {
            double *o = part + (size_t)it.slot * N + col;
            reinterpret_cast<double2 *>(o)[0] = make_double2(a0, a1);
            reinterpret_cast<double2 *>(o)[1] = make_double2(a2, a3);
        }
    } else {
        /* generic path: any N, any alignment; thread owns columns base + j*256 */
        const int base = blockIdx.x * CS_TILE + threadIdx.x;
        for (int t = it.t0; t < it.t1; t++) {
            const float *p = x + (size_t)t * N;
            float v0 = base < N ? ld_stream1(p + base) : 0.f;
            float v1 = base + 256 < N ? ld_stream1(p + base + 256) : 0.f;
            float v2 = base + 512 < N ? ld_stream1(p + base + 512) : 0.f;
            float v3 = base + 768 < N ? ld_stream1(p + base + 768) : 0.f;
            a0 += (double)v0; a1 += (double)v1; a2 += (double)v2; a3 += (double)v3;
            if (ROWSUM) {
                float r = warp_sum((v0 + v1) + (v2 + v3));
                if (lane == 0) rowpart[(size_t)t * rowpart_ld + blockIdx.x * (CS_THREADS / 32) + warp] = r;
            }
        }
        double *o = part + (size_t)it.slot * N;
        if (base < N)       o[base]       = a0;
        if (base + 256 < N) o[base + 256] = a1;
        if (base + 512 < N) o[base + 512] = a2;
        if (base + 768 < N) o[base + 768] = a3;
    }
}

extern "C" int tspws_cu_colsum(const float *d_x, int N, const tspws_csum_item_t *d_items, int nitems,
                               double *d_part, float *d_rowpart, int rowpart_ld, tspws_stream_t st)
{
    if (nitems <= 0) return 0;
    dim3 grid((N + CS_TILE - 1) / CS_TILE, nitems);
    const bool vec = (N % 4 == 0) && (((uintptr_t)d_x & 15) == 0);
    cudaStream_t s = (cudaStream_t)st;
    if (vec) {
        if (d_rowpart) colsum_kernel<true, true><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_rowpart, rowpart_ld);
        else           colsum_kernel<true, false><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_rowpart, rowpart_ld);
    } else {
        if (d_rowpart) colsum_kernel<false, true><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_rowpart, rowpart_ld);
        else           colsum_kernel<false, false><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_rowpart, rowpart_ld);
    }
    TSPWS_LAUNCHED();
    return 0;
}

/* Mean correction of each group: sum over its traces of (double)(float)(rowsum/N)
 * (the reference subtracts the float-cast mean from every sample, src/ts_pws1f_lib.c:166-167;
 * the sum of those subtractions over a group is this constant).  grid = groups. */
__global__ void __launch_bounds__(256)
rowmean_kernel(const float *__restrict__ rowpart, int rowpart_ld, int rowpart_n,
               const int *__restrict__ grp_t0 /*[G+1] local trace ranges*/, int N,
               double *__restrict__ meancorr)
{
    __shared__ double sh[33];
    const int g = blockIdx.x;
    double acc = 0.0;
    for (int t = grp_t0[g] + threadIdx.x; t < grp_t0[g + 1]; t += blockDim.x) {
        const float *r = rowpart + (size_t)t * rowpart_ld;
        double s = 0.0;
        for (int j = 0; j < rowpart_n; j++) s += (double)r[j];
        acc += (double)(float)(s / (double)(unsigned int)N);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) meancorr[g] = acc;
}

/* gsum[g][n] = sum over the group's partial slots - meancorr[g]. grid = (n tiles, G). */
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const double *__restrict__ part, const int *__restrict__ slot_off, int N,
                       const double *__restrict__ meancorr, double *__restrict__ gsum)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (n >= N) return;
    double a = 0.0;
    for (int s = slot_off[g]; s < slot_off[g + 1]; s++) a += part[(size_t)s * N + n];
    if (meancorr) a -= meancorr[g];
    gsum[(size_t)g * N + n] = a;
}

extern "C" int tspws_cu_colsum_finalize(const double *d_part, const int *d_slot_off, int G, int N,
                                        const float *d_rowpart, int rowpart_ld, int rowpart_n,
                                        const int *d_grp_t0, double *d_meancorr, double *d_gsum,
                                        tspws_stream_t st)
{
    cudaStream_t s = (cudaStream_t)st;
    if (d_rowpart) {
        rowmean_kernel<<<G, 256, 0, s>>>(d_rowpart, rowpart_ld, rowpart_n, d_grp_t0, N, d_meancorr);
        TSPWS_LAUNCHED();
    }
    dim3 grid((N + 255) / 256, G);
    colsum_finalize_kernel<<<grid, 256, 0, s>>>(d_part, d_slot_off, N, d_rowpart ? d_meancorr : nullptr, d_gsum);
    TSPWS_LAUNCHED();
    return 0;
}

/* Fold the (already reduced) group sums -- fold is linear, so folding Kmax sums equals summing
 * folded traces -- and emit the fp32 copy the fp32 transform reads. grid = (tiles over ceil(N/2), G). */
__global__ void __launch_bounds__(256)
group_post_kernel(double *__restrict__ gsum, float *__restrict__ gsum32, int N, int do_fold)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = (N + 1) / 2;
    if (n >= half) return;
    double *g = gsum + (size_t)blockIdx.y * N;
    float  *f = gsum32 + (size_t)blockIdx.y * N;
    const int r = N - 1 - n;
    double a = g[n], b = g[r];
    if (do_fold && r != n) { a = 0.5 * (a + b); b = a; g[n] = a; g[r] = a; }
    f[n] = (float)a;
    f[r] = (float)b;
}

extern "C" int tspws_cu_group_post(double *d_gsum, float *d_gsum32, int G, int N, int do_fold, tspws_stream_t st)
{
    dim3 grid(((N + 1) / 2 + 255) / 256, G);
    group_post_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(d_gsum, d_gsum32, N, do_fold);
    TSPWS_LAUNCHED();
    return 0;
}

/* ===================================================================================== */
/* forward frame transform fused with the phasor / linear accumulation                   */
/*   Y_s[m] = sum_l x[(m D_s - c_s + l) mod N] conj(w_s[l])                              */
/*   one warp per work item (a run of coefficients of one scale); G lanes split the taps */
/*   of one coefficient, 32/G coefficients are produced per pass; KB traces share every  */
/*   filter load.  grid.y = chunks of traces; per-chunk PS / ST partials go to scratch.  */
/* ===================================================================================== */
#define FWD_WARPS 8

template <typename T>
__device__ __forceinline__ void phasor_accumulate(T yr, T yi, T &psr, T &psi, T &str_, T &sti)
{
    str_ += yr; sti += yi;
    /* reference: c /= cabs(c); skipped unless |c/|c||^2 <= 1.001, i.e. for c == 0 or non-finite
     * (src/ts_pws1f_lib.c:491-492).  hypot() does not under/overflow for tiny/huge c. */
    T a = (sizeof(T) == 4) ? (T)hypotf((float)yr, (float)yi) : (T)hypot((double)yr, (double)yi);
    if (a > (T)0 && a <= (T)((sizeof(T) == 4) ? 3.0e38f : 1.0e308)) {
        T inv = (T)1 / a;
        psr += yr * inv; psi += yi * inv;
    }
}

template <typename TIN, typename T, int KB>
__global__ void __launch_bounds__(FWD_WARPS * 32)
fwd_kernel(const tspws_scale_t *__restrict__ scales, const typename cplx<T>::type *__restrict__ wall,
           const tspws_fwd_item_t *__restrict__ items, int nitems,
           const TIN *__restrict__ x, int N, int ntr, int chunk, int ncoef,
           typename cplx<T>::type *__restrict__ part_ps, typename cplx<T>::type *__restrict__ part_st)
{
    typedef typename cplx<T>::type T2;
    const int witem = blockIdx.x * FWD_WARPS + (threadIdx.x >> 5);
    if (witem >= nitems) return;                       /* warp-uniform */
    const tspws_fwd_item_t it = items[witem];
    const tspws_scale_t sc = scales[it.s];
    const int lane = threadIdx.x & 31;
    const int G = sc.G, g = lane & (G - 1), og = lane / G, opp = 32 / G;
    const int L = sc.L, D = sc.D;
    const T2 *__restrict__ w = wall + sc.tap_off;
    const int k_begin = blockIdx.y * chunk;
    const int k_end   = min(ntr, k_begin + chunk);
    const size_t obase = (size_t)blockIdx.y * ncoef + sc.coef_off;

    for (int mb = 0; mb < it.nout; mb += opp) {
        const bool valid = (mb + og) < it.nout;
        const int  m     = it.m0 + (valid ? mb + og : it.nout - 1);
        long long pll = ((long long)m * D - sc.c + g) % N;
        if (pll < 0) pll += N;
        const int p0 = (int)pll;
        T psr = 0, psi = 0, str_ = 0, sti = 0;
        for (int k0 = k_begin; k0 < k_end; k0 += KB) {
            const TIN *xk[KB];
#pragma unroll
            for (int kk = 0; kk < KB; kk++) xk[kk] = x + (size_t)min(k0 + kk, k_end - 1) * N;
            T ar[KB], ai[KB];
#pragma unroll
            for (int kk = 0; kk < KB; kk++) { ar[kk] = 0; ai[kk] = 0; }
            int p = p0;
            for (int l = g; l < L; l += G) {
                const T2 wv = w[l];
#pragma unroll
                for (int kk = 0; kk < KB; kk++) {
                    const T xv = (T)xk[kk][p];
                    ar[kk] = fma(xv, wv.x, ar[kk]);
                    ai[kk] = fma(-xv, wv.y, ai[kk]);       /* conj(w) */
                }
                p += G;
                if (p >= N) p -= N;
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int kk = 0; kk < KB; kk++) {
                    ar[kk] += __shfl_xor_sync(0xffffffffu, ar[kk], o);
                    ai[kk] += __shfl_xor_sync(0xffffffffu, ai[kk], o);
                }
            }
            if (g == 0) {
#pragma unroll
                for (int kk = 0; kk < KB; kk++)
                    if (k0 + kk < k_end) phasor_accumulate<T>(ar[kk], ai[kk], psr, psi, str_, sti);
            }
        }
        if (g == 0 && valid) {
            T2 a; a.x = psr; a.y = psi;
            T2 b; b.x = str_; b.y = sti;
            part_ps[obase + m] = a;
            part_st[obase + m] = b;
        }
    }
}

template <typename TIN, typename T>
static int fwd_launch(const tspws_devframe_t *f, const void *d_x, int ntr, int chunk,
                      void *d_pps, void *d_pst, cudaStream_t s)
{
    typedef typename cplx<T>::type T2;
    const int nchunk = (ntr + chunk - 1) / chunk;
    dim3 grid((f->nitems + FWD_WARPS - 1) / FWD_WARPS, nchunk);
    const T2 *w = (const T2 *)(sizeof(T) == 4 ? f->d_w32 : f->d_w64);
    if (chunk >= 4)
        fwd_kernel<TIN, T, 4><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, w, f->d_items, f->nitems,
            (const TIN *)d_x, f->N, ntr, chunk, f->ncoef, (T2 *)d_pps, (T2 *)d_pst);
    else if (chunk >= 2)
        fwd_kernel<TIN, T, 2><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, w, f->d_items, f->nitems,
            (const TIN *)d_x, f->N, ntr, chunk, f->ncoef, (T2 *)d_pps, (T2 *)d_pst);
    else
        fwd_kernel<TIN, T, 1><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, w, f->d_items, f->nitems,
            (const TIN *)d_x, f->N, ntr, chunk, f->ncoef, (T2 *)d_pps, (T2 *)d_pst);
    TSPWS_LAUNCHED();
    return 0;
}

extern "C" int tspws_cu_forward(const tspws_devframe_t *f, const void *d_x, int in64, int acc64,
                                int ntr, int chunk, void *d_part_ps, void *d_part_st, tspws_stream_t st)
{
    if (ntr <= 0 || f->nitems <= 0) return 0;
    if (chunk < 1) chunk = 1;
    cudaStream_t s = (cudaStream_t)st;
    if (acc64) {
        if (!f->d_w64) return (int)cudaErrorInvalidValue;
        return in64 ? fwd_launch<double, double>(f, d_x, ntr, chunk, d_part_ps, d_part_st, s)
                    : fwd_launch<float, double>(f, d_x, ntr, chunk, d_part_ps, d_part_st, s);
    }
    return in64 ? fwd_launch<double, float>(f, d_x, ntr, chunk, d_part_ps, d_part_st, s)
                : fwd_launch<float, float>(f, d_x, ntr, chunk, d_part_ps, d_part_st, s);
}

/* ===================================================================================== */
/* chunk reduction + coherence weighting                                                 */
/*   segment sg sums chunks [sg*seg, (sg+1)*seg) -> PS, ST; Zout = ST * weight(|PS|)     */
/* ===================================================================================== */
__device__ __forceinline__ double weight_factor(const tspws_weight_t r, double pr, double pi)
{
    const double a2 = pr * pr + pi * pi;
    switch (r.mode) {
    case 0:  return a2 * (1.0 / (r.K * r.K * r.T));                       /* nu = 2          */
    case 1:  return sqrt(a2) * (1.0 / (r.K * r.T));                       /* nu = 1          */
    case 3: {                                                             /* unbiased, nu = 2 */
        const double ik = 1.0 / r.K;
        const double c  = (pr * ik) * (pr * ik) + (pi * ik) * (pi * ik);
        return (r.K * c - 1.0) * (1.0 / (r.K - 1.0)) * (1.0 / r.T);
    }
    default: return pow(sqrt(a2) / r.K, r.wu) / r.T;                      /* general nu      */
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
reduce_weight_kernel(const typename cplx<T>::type *part_ps, const typename cplx<T>::type *part_st,
                     int ncoef, int seg, const tspws_weight_t *__restrict__ rules,
                     typename cplx<T>::type *ps, typename cplx<T>::type *st,
                     typename cplx<T>::type *__restrict__ zout)
{
    typedef typename cplx<T>::type T2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoef) return;
    const int sg = blockIdx.y;
    double pr = 0, pi = 0, sr = 0, si = 0;
    for (int c = 0; c < seg; c++) {
        const size_t idx = ((size_t)sg * seg + c) * ncoef + i;
        const T2 a = part_ps[idx], b = part_st[idx];
        pr += (double)a.x; pi += (double)a.y; sr += (double)b.x; si += (double)b.y;
    }
    const double f = weight_factor(rules[sg], pr, pi);
    const size_t o = (size_t)sg * ncoef + i;
    T2 v;
    v.x = (T)pr; v.y = (T)pi; ps[o] = v;
    v.x = (T)sr; v.y = (T)si; st[o] = v;
    v.x = (T)(sr * f); v.y = (T)(si * f); zout[o] = v;
}

extern "C" int tspws_cu_reduce_weight(const void *d_part_ps, const void *d_part_st, int acc64, int ncoef,
                                      int nseg, int seg, const tspws_weight_t *d_rules,
                                      void *d_ps, void *d_st, void *d_zout, tspws_stream_t st)
{
    if (ncoef <= 0 || nseg <= 0) return 0;
    dim3 grid((ncoef + 255) / 256, nseg);
    cudaStream_t s = (cudaStream_t)st;
    if (acc64)
        reduce_weight_kernel<double><<<grid, 256, 0, s>>>((const double2 *)d_part_ps, (const double2 *)d_part_st,
            ncoef, seg, d_rules, (double2 *)d_ps, (double2 *)d_st, (double2 *)d_zout);
    else
        reduce_weight_kernel<float><<<grid, 256, 0, s>>>((const float2 *)d_part_ps, (const float2 *)d_part_st,
            ncoef, seg, d_rules, (float2 *)d_ps, (float2 *)d_st, (float2 *)d_zout);
    TSPWS_LAUNCHED();
    return 0;
}

extern "C" int tspws_cu_weight(const void *d_ps, const void *d_st, int acc64, int ncoef, int nseg,
                               const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st)
{
    /* in-place re-reduction with one chunk per segment: PS/ST are rewritten with themselves */
    return tspws_cu_reduce_weight(d_ps, d_st, acc64, ncoef, nseg, 1, d_rules,
                                  (void *)d_ps, (void *)d_st, d_zout, st);
}

/* ===================================================================================== */
/* inverse frame transform (two coefficient sets at once: Zout -> tsPWS, ST -> ls)       */
/*   x^[n] = sum_s gain_s sum_m Re( Z_s[m] w_s[l] ),  (m D_s - c_s + l) mod N == n       */
/*   grid = (n tiles, scale groups, segments); partials per scale group go to scratch    */
/* ===================================================================================== */
template <typename T, bool WITH_ST>
__global__ void __launch_bounds__(256)
inv_kernel(const tspws_scale_t *__restrict__ scales, const typename cplx<T>::type *__restrict__ wall,
           const tspws_inv_group_t *__restrict__ groups, int ninv, int N, int ncoef,
           const typename cplx<T>::type *__restrict__ zout, const typename cplx<T>::type *__restrict__ st,
           T *__restrict__ scratch, int nseg)
{
    typedef typename cplx<T>::type T2;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const tspws_inv_group_t grp = groups[blockIdx.y];
    const int sg = blockIdx.z;
    const T2 *__restrict__ Z = zout + (size_t)sg * ncoef;
    const T2 *__restrict__ S = WITH_ST ? st + (size_t)sg * ncoef : nullptr;
    T accz = 0, accs = 0;
    for (int s = grp.s0; s < grp.s1; s++) {
        const tspws_scale_t sc = scales[s];
        const T2 *__restrict__ w  = wall + sc.tap_off;
        const T2 *__restrict__ Zs = Z + sc.coef_off;
        const T2 *__restrict__ Ss = WITH_ST ? S + sc.coef_off : nullptr;
        T az = 0, as = 0;
#pragma unroll 1
        for (int j = -1; j <= 1; j++) {
            const long long base = (long long)n + sc.c + (long long)j * N;   /* l = base - m D */
            if (base < 0) continue;
            long long hi = base / sc.D;
            if (hi > sc.Ns - 1) hi = sc.Ns - 1;
            const long long lnum = base - (sc.L - 1);
            long long lo = lnum <= 0 ? 0 : (lnum + sc.D - 1) / sc.D;
            int l = (int)(base - lo * sc.D);
            for (int m = (int)lo; m <= (int)hi; m++, l -= sc.D) {
                const T2 wv = w[l];
                const T2 z  = Zs[m];
                az = fma(z.x, wv.x, az);
                az = fma(-z.y, wv.y, az);
                if (WITH_ST) {
                    const T2 q = Ss[m];
                    as = fma(q.x, wv.x, as);
                    as = fma(-q.y, wv.y, as);
                }
            }
        }
        accz = fma((T)sc.gain, az, accz);
        accs = fma((T)sc.gain, as, accs);
    }
    const size_t o = ((size_t)blockIdx.y * nseg + sg) * N + n;
    scratch[o] = accz;
    if (WITH_ST) scratch[(size_t)ninv * nseg * N + o] = accs;
}

/* Sum the scale-group partials, cast like src/ts_pws1f_lib.c:233-241: tsPWS = (float)x,
 * ls = (float)x / mtr (float division). */
template <typename T, bool WITH_ST>
__global__ void __launch_bounds__(256)
inv_finish_kernel(const T *__restrict__ scratch, int ninv, int nseg, int N,
                  const tspws_weight_t *__restrict__ rules, float *__restrict__ out_ts, float *__restrict__ out_ls)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int sg = blockIdx.y;
    double az = 0, as = 0;
    for (int g = 0; g < ninv; g++) {
        const size_t o = ((size_t)g * nseg + sg) * N + n;
        az += (double)scratch[o];
        if (WITH_ST) as += (double)scratch[(size_t)ninv * nseg * N + o];
    }
    out_ts[(size_t)sg * N + n] = (float)az;
    if (WITH_ST) out_ls[(size_t)sg * N + n] = __fdiv_rn((float)as, (float)(unsigned int)rules[sg].T);
}

template <typename T>
static int inv_launch(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int nseg,
                      const tspws_weight_t *d_rules, void *d_scratch, float *d_ts, float *d_ls, cudaStream_t s)
{
    typedef typename cplx<T>::type T2;
    const T2 *w = (const T2 *)(sizeof(T) == 4 ? f->d_w32 : f->d_w64);
    dim3 grid((f->N + 255) / 256, f->ninv, nseg);
    dim3 gridf((f->N + 255) / 256, nseg);
    if (d_st) {
        inv_kernel<T, true><<<grid, 256, 0, s>>>(f->d_scale, w, f->d_invgrp, f->ninv, f->N, f->ncoef,
            (const T2 *)d_zout, (const T2 *)d_st, (T *)d_scratch, nseg);
        TSPWS_LAUNCHED();
        inv_finish_kernel<T, true><<<gridf, 256, 0, s>>>((const T *)d_scratch, f->ninv, nseg, f->N, d_rules, d_ts, d_ls);
        TSPWS_LAUNCHED();
    } else {
        inv_kernel<T, false><<<grid, 256, 0, s>>>(f->d_scale, w, f->d_invgrp, f->ninv, f->N, f->ncoef,
            (const T2 *)d_zout, nullptr, (T *)d_scratch, nseg);
        TSPWS_LAUNCHED();
        inv_finish_kernel<T, false><<<gridf, 256, 0, s>>>((const T *)d_scratch, f->ninv, nseg, f->N, d_rules, d_ts, d_ls);
        TSPWS_LAUNCHED();
    }
    return 0;
}

extern "C" int tspws_cu_inverse(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int acc64,
                                int nseg, const tspws_weight_t *d_rules, void *d_scratch,
                                float *d_out_ts, float *d_out_ls, tspws_stream_t st)
{
    if (nseg <= 0) return 0;
    if (acc64 && !f->d_w64) return (int)cudaErrorInvalidValue;
    return acc64 ? inv_launch<double>(f, d_zout, d_st, nseg, d_rules, d_scratch, d_out_ts, d_out_ls, (cudaStream_t)st)
                 : inv_launch<float>(f, d_zout, d_st, nseg, d_rules, d_scratch, d_out_ts, d_out_ls, (cudaStream_t)st);
}
