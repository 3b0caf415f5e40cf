/*
 * tspws_kernels.cu -- sm_100a kernels of the ts-PWS stacking path and their C launchers.
 *
 * Kernel                       replaces (reference file:line)                            bound
 * ---------------------------  --------------------------------------------------------  ---------
 * prep_kernel, prep_rm_kernel  fold src/ts_pws1f_lib.c:76-86, mean removal :160-168      HBM
 * colsum2_kernel (colsum_kernel when N % 4 != 0)
 *                              partial_linear_stacks src/ts_pws1f_lib.c:866-881          HBM (headline)
 * group_post2_kernel<PEER>     slot sums + mean removal / fold (:76-86, :160-168 applied
 *                              to the group sums) + fp64 mean split; PEER: the sum over
 *                              the ranks of a node, read from blocks pushed over NVLink  latency
 * fwd2_acc_kernel<LG>          complex_1D_wavelet_dec src/FWTa/wavelet_v7.c:43-64 ->
 *   fwd2_pertrace_kernel         cdotx_dc src/FWTa/cdotx.c:35-72 (polyphase register
 *   fwd_tail_kernel              tile, FFMA2), fused with the phasor accumulate of
 *                                src/ts_pws1f_lib.c:486-494 / :897-904 (acc), or one
 *                                coefficient set per trace (pertrace / tail)             FP32 issue
 * fwd_f64_kernel / f64_cta_body  the same transform in fp64 for the group mean           issue
 * fwd_kernel, fwd_pertrace_generic_kernel
 *                              lane-split form for any geometry (D = 7, 15, 31 ..., fp64) LSU
 * phasor_combine_kernel        tspws_stacks_double src/ts_pws1f_lib.c:885-906 + weights  latency
 * reduce_weight_kernel, weight_* kernels
 *                              tspws_biased / tspws_unbiased src/ts_pws1f_lib.c:909-984  tiny
 * inv2_kernel, inv4_kernel, inv_finish_kernel (fp32); inv_kernel (fp64 / generic)
 *                              Re_complex_1D_wavelet_rec src/FWTa/wavelet_v7.c:124-150
 *                              -> re_cdotx_upsampling_cc / re_cdotx_cc
 *                              src/FWTa/cdotx.c:305-340, :176-211; cast :233-241         L2 / issue
 * regroup, group_total, coef_matrix, select_sum kernels
 *                              tspws_subsmpl_float / TwoStage_subsmpl_float / jackknife
 *                              src/ts_pws1f_lib.c:501-831                                 HBM / L2
 * prefix_weight, simil, lsconv kernels
 *                              convergence curves src/ts_pws1f_lib.c:247-314, :433-462    L2
 * peer_publish_f32 / peer_collect_f32 kernels
 *                              single-stage PS | ST sums across the ranks of a node        NVLink
 *
 * The tcgen05 / TMEM kernels are in tspws_mma.cu: the octaves with D >= 16 of the single-stage forward transform
 * (fwd_mma_persistent_kernel: dense and banded GEMMs) and the resampling contraction (resample_mma_kernel).
 * The links of the two-stage tail are launched with programmatic dependent launch (LAUNCH_DEP / pdl_wait).
 *
 * Everything is written from the mathematics in SURVEY.md Appendix A.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "tspws_internal.h"

static unsigned long g_launches = 0;
extern "C" unsigned long tspws_cu_launch_count(void) { return g_launches; }
extern "C" void tspws_cu_launch_count_add(unsigned long n) { g_launches += n; }

/* ---- optional per-kernel timing with CUDA events on the launching stream: an event is recorded
 * after every launch (and at tspws_cu_profile_mark), and a kernel is charged the time since the
 * previous event -- warm, in-stream durations including launch gaps (ncu's are cold-cache and
 * serialised).  Off by default: zero cost. ---- */
#define PROF_MAX_KERNELS 32
#define PROF_MAX_EVENTS  2048
static struct {
    int         on, nk, nev, created;
    const char *name[PROF_MAX_KERNELS];
    double      ms[PROF_MAX_KERNELS];
    long        count[PROF_MAX_KERNELS];
    cudaEvent_t ev[PROF_MAX_EVENTS];
    int         evk[PROF_MAX_EVENTS];       /* kernel index, -1 for a bare mark */
} g_prof;

static void prof_flush(void)
{
    for (int i = 1; i < g_prof.nev; i++) {
        if (g_prof.evk[i] < 0) continue;
        float ms = 0.f;
        if (cudaEventSynchronize(g_prof.ev[i]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, g_prof.ev[i - 1], g_prof.ev[i]) == cudaSuccess) {
            g_prof.ms[g_prof.evk[i]] += ms;
            g_prof.count[g_prof.evk[i]]++;
        }
    }
    g_prof.nev = 0;
}

static void prof_record(const char *name, cudaStream_t s)
{
    if (g_prof.nev >= PROF_MAX_EVENTS) {
        prof_flush();
        if (name) return;                    /* lost its start marker: skip this one */
    }
    int k = -1;
    if (name) {
        for (k = 0; k < g_prof.nk; k++) if (!strcmp(g_prof.name[k], name)) break;
        if (k == g_prof.nk) { if (k >= PROF_MAX_KERNELS) return; g_prof.name[k] = name; g_prof.nk++; }
        if (g_prof.nev == 0) return;         /* no preceding event to measure from */
    }
    const int i = g_prof.nev;
    if (i >= g_prof.created) { cudaEventCreate(&g_prof.ev[i]); g_prof.created = i + 1; }
    g_prof.evk[i] = k;
    cudaEventRecord(g_prof.ev[i], s);
    g_prof.nev = i + 1;
}

extern "C" int tspws_cu_profile_enabled(void) { return g_prof.on; }
extern "C" void tspws_cu_profile_mark(tspws_stream_t st) { if (g_prof.on) prof_record(nullptr, (cudaStream_t)st); }
/* kernels launched from other translation units (tspws_mma.cu): name must be a string literal */
extern "C" void tspws_cu_profile_launched(const char *name, tspws_stream_t st) { if (g_prof.on) prof_record(name, (cudaStream_t)st); }
extern "C" int tspws_cu_profile_enable(int on)
{
    if (!on) prof_flush();
    g_prof.on = on;
    return 0;
}
extern "C" int tspws_cu_profile_reset(void)
{
    prof_flush();
    for (int k = 0; k < g_prof.nk; k++) { g_prof.ms[k] = 0; g_prof.count[k] = 0; }
    return 0;
}
/* Writes "name count total_ms\n" lines; returns the number of bytes written. */
extern "C" int tspws_cu_profile_read(char *buf, int cap)
{
    prof_flush();
    int off = 0;
    if (cap > 0) buf[0] = 0;
    for (int k = 0; k < g_prof.nk && off < cap - 1; k++)
        off += snprintf(buf + off, cap - off, "%s %ld %.6f\n", g_prof.name[k], g_prof.count[k], g_prof.ms[k]);
    return off;
}

#define TSPWS_LAUNCHED_N(name, stream)                \
    do {                                              \
        if (g_prof.on) prof_record(name, stream);     \
        g_launches++;                                 \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) return (int)e__;      \
    } while (0)

/* ---- per-CTA timeline of the tail kernels (diagnostics, TSPWS_TIMELINE=file): every CTA of an instrumented
 * kernel records (start ns, end ns, SM) of %globaltimer; dumped after the inverse launch.  Off: one
 * uniform load and a predicated-off branch per CTA. ---- */
#ifdef TSPWS_TIMELINE_BUILD        /* make EXTRA=-DTSPWS_TIMELINE_BUILD: never in the product build */
/* Append-only log: every CTA of an instrumented kernel takes the next record (atomic counter) and writes
 * (start ns, end ns, SM | kernel id << 32, CTA); nothing is reset per launch, so it also works under graph replay and
 * across streams.  tspws_cu_timeline_flush(file) (or the end of an inverse launch when TSPWS_TIMELINE_EACH=1) dumps it. */
#define TL_LOG (1u << 21)
__device__ unsigned long long *g_tl_dev = nullptr;
__device__ unsigned int g_tl_count = 0;
static unsigned long long *g_tl_host = nullptr;       /* device buffer [TL_LOG][4] */
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct tl_scope {
    unsigned long long *p;
    __device__ __forceinline__ tl_scope(const int kid)
    {
        __shared__ unsigned int s_rec;
        unsigned long long *base = g_tl_dev;
        p = nullptr;
        if (base) {
            if (threadIdx.x == 0) s_rec = atomicAdd(&g_tl_count, 1u);
            __syncthreads();
            const unsigned int rec = s_rec;
            if (rec < TL_LOG) {
                p = base + (size_t)rec * 4;
                if (threadIdx.x == 0) {
                    unsigned int smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    p[0] = gtimer(); p[2] = (unsigned long long)smid | ((unsigned long long)kid << 32);
                    p[3] = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
                }
            }
        }
    }
    __device__ __forceinline__ ~tl_scope() { if (p && (threadIdx.x & 31) == 0) atomicMax(p + 1, gtimer()); }
};
static void tl_host_init(cudaStream_t s)
{
    static int tried = 0;
    if (!tried) {
        tried = 1;
        if (getenv("TSPWS_TIMELINE") && cudaMalloc((void **)&g_tl_host, (size_t)TL_LOG * 4 * sizeof(unsigned long long)) == cudaSuccess) {
            cudaMemset(g_tl_host, 0, (size_t)TL_LOG * 4 * sizeof(unsigned long long));
            cudaMemcpyToSymbol(g_tl_dev, &g_tl_host, sizeof g_tl_host);
        }
    }
    (void)s;
}
extern "C" int tspws_cu_timeline_flush(const char *file)
{
    if (!g_tl_host) return -1;
    unsigned int cnt = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(&cnt, g_tl_count, sizeof cnt) != cudaSuccess) return -2;
    if (cnt > TL_LOG) cnt = TL_LOG;
    unsigned long long *h = (unsigned long long *)malloc((size_t)(cnt ? cnt : 1) * 4 * sizeof *h);
    if (h && cudaMemcpy(h, g_tl_host, (size_t)cnt * 4 * sizeof *h, cudaMemcpyDeviceToHost) == cudaSuccess) {
        FILE *fp = fopen(file, "w");
        if (fp) {
            for (size_t i = 0; i < (size_t)cnt * 4; i += 4)
                fprintf(fp, "%d %llu %llu %llu %d\n", (int)(h[i + 2] >> 32), h[i + 3], h[i], h[i + 1], (int)(h[i + 2] & 0xffffffffu));
            fclose(fp);
        }
    }
    free(h);
    cnt = 0;
    cudaMemcpyToSymbol(g_tl_count, &cnt, sizeof cnt);
    return 0;
}
static void tl_host_dump(cudaStream_t s)
{
    (void)s;
    if (g_tl_host && getenv("TSPWS_TIMELINE_EACH")) tspws_cu_timeline_flush(getenv("TSPWS_TIMELINE"));
}
#else
struct tl_scope { __device__ __forceinline__ tl_scope(const int) {} };
static void tl_host_init(cudaStream_t) {}
static void tl_host_dump(cudaStream_t) {}
#endif

template <typename T> struct cplx;
template <> struct cplx<float>  { typedef float2  type; };
template <> struct cplx<double> { typedef double2 type; };

/* streaming (evict-first) loads: every trace sample is read exactly once */
__device__ __forceinline__ float4 ld_stream4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float  ld_stream1(const float *p) { return __ldcs(p); }

/* ---- programmatic dependent launch between the short kernels of the two-stage tail: a dependent kernel launched with
 * LAUNCH_DEP may become resident as soon as every CTA of its predecessor in the stream has started (pdl_trigger) and
 * waits in pdl_wait until the predecessor has completed and its writes are visible -- the launch gap and the CTA ramp
 * of every link of the chain disappear under the predecessor's last wave.  Both instructions are no-ops in a kernel
 * that was launched the ordinary way.  TSPWS_PDL=0 switches the launch attribute off. ---- */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static int pdl_on()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("TSPWS_PDL"); on = e ? atoi(e) != 0 : 1; }
    return on;
}
#define LAUNCH_DEP(kernel, grid_, block_, smem_, stream_, ...)                                            \
    do {                                                                                             \
        cudaLaunchConfig_t c_ = {};                                                                  \
        c_.gridDim = (grid_); c_.blockDim = (block_); c_.dynamicSmemBytes = (smem_); c_.stream = (stream_); \
        cudaLaunchAttribute a_[1];                                                                   \
        a_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                               \
        a_[0].val.programmaticStreamSerializationAllowed = 1;                                        \
        c_.attrs = a_; c_.numAttrs = pdl_on() ? 1 : 0;                                                \
        cudaLaunchKernelEx(&c_, kernel, __VA_ARGS__);                                                \
    } while (0)

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

/* Mean removal alone (the usual call: `rm` without `fold`), rows of N % 4 == 0 <= 4 * 256 * PREP_V floats: one pass --
 * the row stays in registers between the sum and the subtraction (PREP_V float4 per thread, all loads in flight at
 * once), so every sample crosses HBM once in each direction.  The two-pass kernel above took 0.40 ms for 8192 x 16384
 * (1 GiB of traffic in 0.16 ms at the roofline): scalar loads, the second pass out of L2. */
#define PREP_V 16
__global__ void __launch_bounds__(256)
prep_rm_kernel(float *__restrict__ x, int N)
{
    __shared__ double sh[33];
    float4 *row = reinterpret_cast<float4 *>(x + (size_t)blockIdx.x * N);
    const int n4 = N >> 2;
    float4 v[PREP_V];
#pragma unroll
    for (int i = 0; i < PREP_V; i++) {
        const int j = threadIdx.x + i * 256;
        v[i] = j < n4 ? row[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < PREP_V; i++) sum += ((double)v[i].x + (double)v[i].y) + ((double)v[i].z + (double)v[i].w);
    const double tot = block_sum(sum, sh);
    const float  m   = (float)(tot / (double)(unsigned int)N);
#pragma unroll
    for (int i = 0; i < PREP_V; i++) {
        const int j = threadIdx.x + i * 256;
        if (j < n4) row[j] = make_float4(__fsub_rn(v[i].x, m), __fsub_rn(v[i].y, m), __fsub_rn(v[i].z, m), __fsub_rn(v[i].w, m));
    }
}

extern "C" int tspws_cu_prep(float *d_x, int ntr, int N, int do_fold, int do_rm, tspws_stream_t st)
{
    if (ntr <= 0 || (!do_fold && !do_rm)) return 0;
    if (!do_fold && do_rm && (N & 3) == 0 && N <= 4 * 256 * PREP_V && ((uintptr_t)d_x & 15) == 0) {
        prep_rm_kernel<<<ntr, 256, 0, (cudaStream_t)st>>>(d_x, N);
        TSPWS_LAUNCHED_N("prep_kernel", (cudaStream_t)st);
        return 0;
    }
    prep_kernel<<<ntr, 256, 0, (cudaStream_t)st>>>(d_x, N, do_fold, do_rm);
    TSPWS_LAUNCHED_N("prep_kernel", (cudaStream_t)st);
    return 0;
}

/* ===================================================================================== */
/* colsum: fp32 traces -> fp64 group sums (HBM-bound stream; the headline kernel)        */
/*   grid.x = column tiles of 1024 samples, grid.y = items (trace ranges inside a group) */
/*   Every CTA leaves the partial column sums of its item in a private slot; the LAST    */
/*   CTA of each (group, tile) to finish -- found with a fence + atomic ticket -- adds   */
/*   the group's slots in slot order (deterministic, no float atomics) into gsum and     */
/*   leaves the tile's total in gtile[g][tile] (used for the mean removal).              */
/* ===================================================================================== */
#define CS_THREADS 256
#define CS_TILE    1024
#define CS_UNROLL  8

template <bool VEC>
__global__ void __launch_bounds__(CS_THREADS)
colsum_kernel(const float *__restrict__ x, int N, const tspws_csum_item_t *__restrict__ items,
              double *part, const int *__restrict__ slot_off, unsigned int *tickets,
              double *__restrict__ gsum, double *__restrict__ gtile)
{
    __shared__ double sh[33];
    __shared__ int is_last;
    const tspws_csum_item_t it = items[blockIdx.y];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int c0, c1, c2, c3;                                   /* the four columns this thread owns */

    if (VEC) {
        const int  col    = blockIdx.x * CS_TILE + threadIdx.x * 4;
        const bool active = col < N;                       /* N % 4 == 0 on this path */
        c0 = col; c1 = col + 1; c2 = col + 2; c3 = col + 3;
        const float *p = x + (size_t)it.t0 * N + (active ? col : 0);
        int t = it.t0;
        for (; t + CS_UNROLL <= it.t1; t += CS_UNROLL) {
            float4 v[CS_UNROLL];
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) v[j] = ld_stream4(p + (size_t)j * N);
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) {
                a0 += (double)v[j].x; a1 += (double)v[j].y; a2 += (double)v[j].z; a3 += (double)v[j].w;
            }
            p += (size_t)CS_UNROLL * N;
        }
        for (; t < it.t1; t++) {
            const float4 v = ld_stream4(p);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            p += N;
        }
        if (!active) { a0 = a1 = a2 = a3 = 0.0; }
    } else {
        /* generic path: any N, any alignment; thread owns columns base + j*256 */
        const int base = blockIdx.x * CS_TILE + threadIdx.x;
        c0 = base; c1 = base + 256; c2 = base + 512; c3 = base + 768;
        for (int t = it.t0; t < it.t1; t++) {
            const float *p = x + (size_t)t * N;
            a0 += c0 < N ? (double)ld_stream1(p + c0) : 0.0;
            a1 += c1 < N ? (double)ld_stream1(p + c1) : 0.0;
            a2 += c2 < N ? (double)ld_stream1(p + c2) : 0.0;
            a3 += c3 < N ? (double)ld_stream1(p + c3) : 0.0;
        }
    }

    const int s_begin = slot_off[it.group], s_end = slot_off[it.group + 1];
    if (s_end - s_begin > 1) {
        /* publish this item's partial, then take a ticket for (group, tile) */
        double *o = part + (size_t)it.slot * N;
        if (c0 < N) o[c0] = a0;
        if (c1 < N) o[c1] = a1;
        if (c2 < N) o[c2] = a2;
        if (c3 < N) o[c3] = a3;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int n = atomicAdd(&tickets[it.group * gridDim.x + blockIdx.x], 1u);
            is_last = (n == (unsigned int)(s_end - s_begin - 1));
            if (is_last) tickets[it.group * gridDim.x + blockIdx.x] = 0;      /* re-arm for the next launch */
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        a0 = a1 = a2 = a3 = 0.0;
        for (int sidx = s_begin; sidx < s_end; sidx++) {   /* fixed order: deterministic */
            const double *q = part + (size_t)sidx * N;
            if (c0 < N) a0 += __ldcg(q + c0);
            if (c1 < N) a1 += __ldcg(q + c1);
            if (c2 < N) a2 += __ldcg(q + c2);
            if (c3 < N) a3 += __ldcg(q + c3);
        }
    }
    double *go = gsum + (size_t)it.group * N;
    if (c0 < N) go[c0] = a0;
    if (c1 < N) go[c1] = a1;
    if (c2 < N) go[c2] = a2;
    if (c3 < N) go[c3] = a3;
    const double tot = block_sum((a0 + a1) + (a2 + a3), sh);
    if (threadIdx.x == 0) gtile[it.group * gridDim.x + blockIdx.x] = tot;
}

extern "C" int tspws_cu_colsum(const float *d_x, int N, const tspws_csum_item_t *d_items, int nitems,
                               double *d_part, const int *d_slot_off, unsigned int *d_tickets,
                               double *d_gsum, double *d_gtile, tspws_stream_t st)
{
    if (nitems <= 0) return 0;
    dim3 grid((N + CS_TILE - 1) / CS_TILE, nitems);
    const bool vec = (N % 4 == 0) && (((uintptr_t)d_x & 15) == 0);
    cudaStream_t s = (cudaStream_t)st;
    if (vec) colsum_kernel<true><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_slot_off, d_tickets, d_gsum, d_gtile);
    else     colsum_kernel<false><<<grid, CS_THREADS, 0, s>>>(d_x, N, d_items, d_part, d_slot_off, d_tickets, d_gsum, d_gtile);
    TSPWS_LAUNCHED_N("colsum_kernel", s);
    return 0;
}

/* Group post-processing on the group sums (group_post2_kernel below):
 *   mean removal : the reference subtracts every trace's mean from its samples
 *                  (src/ts_pws1f_lib.c:160-168); summed over a group that is the constant
 *                  sum_t mean_t = (1/N) sum_n G[n], taken here from the tile totals;
 *   fold         : 0.5 (G[n] + G[N-1-n]) (src/ts_pws1f_lib.c:76-86), linear as well;
 *   mean split   : the K group sums share a large coherent part (the stacked signal, K-fold
 *                  amplified against the noise).  Their mean M[n] is split off in fp64 and only the
 *                  residuals G_k - M go through the fp32 transform: Y_k = T64(M) + T32(G_k - M).
 *                  The phase of a coefficient that is small in one group is then no longer limited
 *                  by the fp32 rounding of the large common part (headline config: coherence error
 *                  2e-5 -> 4.5e-6 against the reference's double).
 * Emits the fp32 traces the fp32 transform reads. */

/* ------------------------------------------------------------------------------------- */
/* Single-wave group sums (N % 4 == 0): grid = column tiles x row parts, sized on the host to  */
/* 148 SMs x 4 resident CTAs, every CTA streaming the same number of rows (+-1).  A part may   */
/* cross group boundaries: the running sums are flushed to the next slot at each boundary, so  */
/* slots are numbered in row order and those of one group are consecutive.  No tickets, fences */
/* or second pass here -- group_post2_kernel adds a group's slots in slot order (deterministic) */
/* while it applies mean removal / fold / the mean split.  tools/readbw.cu: this access pattern */
/* streams at 6.3 TB/s (97 % of the measured copy peak); the ticketed kernel above at 5.6 TB/s. */
/* part record: (first row, end row, group of the first row, slot of the first flush)          */
/* ------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(CS_THREADS, 4)
colsum2_kernel(const float *__restrict__ x, int N, const int4 *__restrict__ parts, const int *__restrict__ grp_t0,
               double *__restrict__ slots, double *__restrict__ slot_tot, int ntile)
{
    tl_scope tls(0);
    __shared__ double sh[33];
    const int4 pi = parts[blockIdx.y];
    const int  col = blockIdx.x * CS_TILE + threadIdx.x * 4;
    const bool active = col < N;
    const float *p = x + (size_t)pi.x * N + (active ? col : 0);
    int r = pi.x, g = pi.z, slot = pi.w;
    while (r < pi.y) {
        const int rend = min(pi.y, grp_t0[g + 1]);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (; r + CS_UNROLL <= rend; r += CS_UNROLL) {
            float4 v[CS_UNROLL];
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) v[j] = ld_stream4(p + (size_t)j * N);
#pragma unroll
            for (int j = 0; j < CS_UNROLL; j++) {
                a0 += (double)v[j].x; a1 += (double)v[j].y; a2 += (double)v[j].z; a3 += (double)v[j].w;
            }
            p += (size_t)CS_UNROLL * N;
        }
        for (; r < rend; r++) {
            const float4 v = ld_stream4(p);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            p += N;
        }
        if (active) {
            double2 *o = reinterpret_cast<double2 *>(slots + (size_t)slot * N + col);
            o[0] = make_double2(a0, a1); o[1] = make_double2(a2, a3);
        } else {
            a0 = a1 = a2 = a3 = 0.0;
        }
        const double tot = block_sum((a0 + a1) + (a2 + a3), sh);
        if (threadIdx.x == 0) slot_tot[(size_t)slot * ntile + blockIdx.x] = tot;
        g++; slot++;
    }
}

extern "C" int tspws_cu_colsum2(const float *d_x, int N, const void *d_parts, int nparts, const int *d_grp_t0,
                                double *d_slots, double *d_slot_tot, int ntile, tspws_stream_t st)
{
    if (nparts <= 0) return 0;
    if ((N & 3) || ((uintptr_t)d_x & 15)) return (int)cudaErrorInvalidValue;
    tl_host_init((cudaStream_t)st);
    dim3 grid(ntile, nparts);
    colsum2_kernel<<<grid, CS_THREADS, 0, (cudaStream_t)st>>>(d_x, N, (const int4 *)d_parts, d_grp_t0, d_slots, d_slot_tot, ntile);
    TSPWS_LAUNCHED_N("colsum2_kernel", (cudaStream_t)st);
    return 0;
}

/* Slot sums -> group sums, with the linear post-processing described above fused in.
 * Block = 32 consecutive samples x 8 group lanes (warp w handles groups w, w+8, ...): every global
 * access is a coalesced 256 B row, and all groups of a sample are in flight together.
 *   mode 0  combine only: gsum = sum of slots, gtile = sum of slot totals (the exchange block of a
 *           sharded plan; mean removal etc. follow the all-reduce, in a mode-1 launch over the block)
 *   mode 1  full: mean removal (from the slot totals), fold, mean split, fp32 copy / residuals    */
#define GP2_LANES 8
/* Peer exchange (one process per GPU, ranks of one node, regions mapped with CUDA IPC): the cross-GPU sum
 * of the group sums is fused into the two group_post2 launches instead of an all-reduce between them.
 *   publish (mode 0 with peer args): the rank's block (group sums + tile totals of its shard) is PUSHED
 *       into slot [parity][rank] of every rank's region with plain stores over NVLink / NVSwitch (remote
 *       stores pipeline; remote loads measured 10x slower at 8 GPUs: 124 us for 9 MB).  A shard is a
 *       contiguous trace range, so it touches only a few of the K groups: only those are pushed, and a
 *       presence mark per group tells the readers which ranks to add (at 8 GPUs 2-3 of 10 groups);
 *   the last CTA of the publishing launch: system fence, then the step number into every rank's flag word;
 *   PEER = true launch: waits until the flags of all ranks have reached this step, then treats the world
 *       blocks of its OWN region as the slots of every group (rank order: deterministic).
 * Blocks are double-buffered by step parity, so a rank that runs ahead never overwrites a block that a
 * slower rank still reads (it cannot pass its next wait before the slower rank has signalled). */
struct tspws_peer_args {
    double *const *regions;          /* [world] base of every rank's region: [2][world][block_len] doubles */
    double              *own;        /* base of the own region */
    const unsigned int  *flags;      /* [world] own flag words, written by the peers */
    unsigned int        *step;       /* own step counter (advanced by the last CTA of the publishing launch) */
    unsigned int        *err;        /* set when a wait timed out */
    unsigned int        *ticket;     /* publishing launch: CTAs done (the last one signals the peers) */
    unsigned int *const *peer_flags; /* [world] every rank's flag words */
    int world, rank;
    long long block_len;             /* doubles per block = K*N group sums + K*ntile tile totals + K presence marks */
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

/* peer memory is written by other GPUs between steps: system-scope loads, never a stale cached line */
__device__ __forceinline__ double ld_sys_f64(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <bool PEER>
__global__ void __launch_bounds__(32 * GP2_LANES)
group_post2_kernel(const double *__restrict__ slots, const double *__restrict__ slot_tot, const int *__restrict__ slot_off,
                   int N, int K, int ntile, int mode, int do_rm, int do_fold, int split,
                   double *__restrict__ gsum, double *__restrict__ gtile, float *__restrict__ out32, double *__restrict__ gmean,
                   const tspws_peer_args pa, const double inv_n, const double inv_k)
{
    tl_scope tls(1);
    pdl_trigger();
    pdl_wait();
    extern __shared__ double s_c[];                      /* [K] group-mean corrections, then [K][world] presence bytes */
    __shared__ double s_m[GP2_LANES][32];
    unsigned char *s_pres = reinterpret_cast<unsigned char *>(s_c + K);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    size_t par_off = 0;                                  /* first block of this step's parity */
    bool dead = false;                                   /* a peer never arrived: poison the sums (NaN) instead of stacking a partial exchange */
    if (PEER) {
        const unsigned int step = *pa.step;              /* already advanced for this step */
        par_off = (size_t)((step - 1) & 1) * (size_t)pa.world * (size_t)pa.block_len;
        if (threadIdx.x < pa.world) {
            long long spins = 0;
            while (ld_acquire_sys(pa.flags + threadIdx.x) < step) {
                __nanosleep(200);
                if (++spins > 20000000LL) { *pa.err = 1; break; }      /* ~4 s: a rank is gone; never hang the GPU */
            }
        }
        __syncthreads();
        dead = *reinterpret_cast<volatile unsigned int *>(pa.err) != 0;
        /* which rank contributes to which group (marks pushed with the blocks) */
        for (int i = threadIdx.x; i < K * pa.world; i += blockDim.x) {
            const int k = i / pa.world, j = i % pa.world;
            s_pres[i] = ld_sys_f64(pa.own + par_off + (size_t)j * pa.block_len + (size_t)K * N + (size_t)K * ntile + k) != 0.0;
        }
        __syncthreads();
    } else if (mode == 0 && pa.step) {
        /* publishing launch of a peer plan: slot [parity][rank] of every region */
        par_off = ((size_t)(*pa.step & 1) * (size_t)pa.world + (size_t)pa.rank) * (size_t)pa.block_len;
    }
    const bool push = !PEER && mode == 0 && pa.step;
    /* slot j of group k: value at sample n, total of tile t */
    auto cnt = [&](int k) { return PEER ? pa.world : slot_off[k + 1] - slot_off[k]; };
    auto sval = [&](int k, int j, int n) {
        if (PEER) return s_pres[k * pa.world + j] ? ld_sys_f64(pa.own + par_off + (size_t)j * pa.block_len + (size_t)k * N + n) : 0.0;
        return __ldcg(slots + (size_t)(slot_off[k] + j) * N + n);
    };
    auto tval = [&](int k, int j, int t) {
        if (PEER) return s_pres[k * pa.world + j] ? ld_sys_f64(pa.own + par_off + (size_t)j * pa.block_len + (size_t)K * N + (size_t)k * ntile + t) : 0.0;
        return slot_tot[(size_t)(slot_off[k] + j) * ntile + t];
    };
    if (mode == 0 ? blockIdx.x == 0 : do_rm) {
        /* totals of every group: lanes stride over (slot, tile) in a fixed pattern, fixed shuffle tree */
        for (int k = w; k < K; k += GP2_LANES) {
            const int c_k = cnt(k);
            if (mode == 0) {
                if (push && lane == 0)                     /* presence mark of group k, rewritten every step */
                    for (int q = 0; q < pa.world; q++) pa.regions[q][par_off + (size_t)K * N + (size_t)K * ntile + k] = c_k ? 1.0 : 0.0;
                for (int t = lane; t < ntile; t += 32) {
                    double c = 0.0;
                    for (int j = 0; j < c_k; j++) c += tval(k, j, t);
                    if (push) { if (c_k) for (int q = 0; q < pa.world; q++) pa.regions[q][par_off + (size_t)K * N + (size_t)k * ntile + t] = c; }
                    else gtile[(size_t)k * ntile + t] = c;
                }
            } else {
                double c = 0.0;
                if (PEER) {
                    for (int i = lane; i < c_k * ntile; i += 32) c += tval(k, i / ntile, i % ntile);
                } else {                                   /* the totals of a group's slots are contiguous */
                    const double *t0 = slot_tot + (size_t)slot_off[k] * ntile;
                    for (int i = lane; i < c_k * ntile; i += 32) c += t0[i];
                }
                c = warp_sum(c);
                if (lane == 0) s_c[k] = c * inv_n;         /* host-side reciprocal: no fp64 division chain here */
            }
        }
    }
    if (mode == 1 && do_rm) __syncthreads();
    const int n = blockIdx.x * 32 + lane;
    const bool ok = n < N;
    const int r = N - 1 - n;
    /* the first GP2_KR rounds of this warp's groups stay in registers between the two passes */
    constexpr int GP2_KR = 4;
    double vreg[GP2_KR];
    double acc = 0.0;
#pragma unroll
    for (int kk = 0; kk < GP2_KR; kk++) vreg[kk] = 0.0;
    auto group_value = [&](int k) {
        const int c_k = cnt(k);
        double v = 0.0, vr = 0.0;
#pragma unroll 4
        for (int j = 0; j < c_k; j++) v += sval(k, j, n);
        if (mode == 1 && do_fold) {
#pragma unroll 4
            for (int j = 0; j < c_k; j++) vr += sval(k, j, r);
            v = 0.5 * (v + vr);                            /* the constant correction is fold-invariant */
        }
        if (mode == 1 && do_rm) v -= s_c[k];
        if (PEER && dead) v = __longlong_as_double(0x7ff8000000000000LL);
        if (push) { if (c_k) for (int q = 0; q < pa.world; q++) pa.regions[q][par_off + (size_t)k * N + n] = v; }
        else gsum[(size_t)k * N + n] = v;
        return v;
    };
    if (ok) {
#pragma unroll
        for (int kk = 0; kk < GP2_KR; kk++) {
            const int k = w + kk * GP2_LANES;
            if (k < K) { vreg[kk] = group_value(k); acc += vreg[kk]; }
        }
        for (int k = w + GP2_KR * GP2_LANES; k < K; k += GP2_LANES) acc += group_value(k);
    }
    if (mode == 0) {
        if (push) {
            /* every CTA has pushed its columns; the LAST one to get here makes the whole block visible and stores
             * the step number into every rank's flag word (what a separate signalling launch used to do) */
            __shared__ unsigned int s_last;
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int t = atomicAdd(pa.ticket, 1u);
                s_last = (t == gridDim.x - 1);
                if (s_last) *pa.ticket = 0;
            }
            __syncthreads();
            if (s_last) {
                const unsigned int st = *pa.step + 1;
                __threadfence_system();
                if ((int)threadIdx.x < pa.world)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pa.peer_flags[threadIdx.x] + pa.rank), "r"(st) : "memory");
                __syncthreads();
                if (threadIdx.x == 0) *pa.step = st;
            }
        }
        return;
    }
    double m = 0.0;
    if (split) {
        s_m[w][lane] = acc;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < GP2_LANES; i++) m += s_m[i][lane];
        m *= inv_k;
        if (w == 0 && ok && gmean) { gmean[n] = m; gmean[n + N] = m; gmean[n + 2 * N] = m; }   /* three periods for fwd_f64_kernel */
    }
    if (!ok) return;
#pragma unroll
    for (int kk = 0; kk < GP2_KR; kk++) {
        const int k = w + kk * GP2_LANES;
        if (k < K) {
            out32[(size_t)k * N + n] = (float)(vreg[kk] - m);
            if (split > 1) out32[(size_t)(K + k) * N + n] = (float)vreg[kk];   /* rows K..2K-1: unsplit (long filters) */
        }
    }
    for (int k = w + GP2_KR * GP2_LANES; k < K; k += GP2_LANES) {
        const double v = gsum[(size_t)k * N + n];
        out32[(size_t)k * N + n] = (float)(v - m);
        if (split > 1) out32[(size_t)(K + k) * N + n] = (float)v;
    }
}

/* ---- peer exchange of a flat fp32 block (single-stage stacks: the rank's phasor and linear sums, 4 ncoef floats):
 * publish = every thread stores its elements into slot [parity][rank] of every rank's region, the last CTA signals;
 * collect = wait for all ranks, then the sum of the world blocks of the own region in rank order (deterministic,
 * identical on all ranks).  Same regions, flags, step counter and parity double buffering as the two-stage exchange. ---- */
__global__ void __launch_bounds__(256)
peer_publish_f32_kernel(const float *__restrict__ src, long long n, const tspws_peer_args pa)
{
    __shared__ unsigned int s_last;
    const size_t blk = (size_t)pa.block_len * 2;                         /* floats per block (block_len counts doubles) */
    const size_t off = ((size_t)(*pa.step & 1) * (size_t)pa.world + (size_t)pa.rank) * blk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = src[i];
        for (int q = 0; q < pa.world; q++) reinterpret_cast<float *>(pa.regions[q])[off + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(pa.ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *pa.ticket = 0;
    }
    __syncthreads();
    if (s_last) {
        const unsigned int st = *pa.step + 1;
        __threadfence_system();
        if ((int)threadIdx.x < pa.world)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pa.peer_flags[threadIdx.x] + pa.rank), "r"(st) : "memory");
        __syncthreads();
        if (threadIdx.x == 0) *pa.step = st;
    }
}

__global__ void __launch_bounds__(256)
peer_collect_f32_kernel(float *__restrict__ dst, long long n, const tspws_peer_args pa)
{
    __shared__ int s_dead;
    const unsigned int step = *pa.step;                                    /* already advanced by the publishing launch */
    if (threadIdx.x == 0) s_dead = 0;
    __syncthreads();
    if (threadIdx.x < pa.world) {
        long long spins = 0;
        while (ld_acquire_sys(pa.flags + threadIdx.x) < step) {
            __nanosleep(200);
            if (++spins > 20000000LL) { *pa.err = 1; s_dead = 1; break; }
        }
    }
    __syncthreads();
    const bool dead = s_dead || *reinterpret_cast<volatile unsigned int *>(pa.err) != 0;
    const size_t blk = (size_t)pa.block_len * 2;
    const float *own = reinterpret_cast<const float *>(pa.own) + (size_t)((step - 1) & 1) * (size_t)pa.world * blk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int r = 0; r < pa.world; r++) {
            float v;
            asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(own + (size_t)r * blk + i) : "memory");
            a += v;
        }
        dst[i] = dead ? __int_as_float(0x7fc00000) : a;
    }
}

extern "C" int tspws_cu_peer_f32(int publish, float *d_buf, long long n, long long block_len_doubles, void *d_regions, double *d_own,
                                 const unsigned int *d_flags, unsigned int *d_step, unsigned int *d_err, void *d_peer_flags,
                                 int rank, int world, tspws_stream_t st)
{
    if (world > 256) return (int)cudaErrorInvalidValue;
    tspws_peer_args pa;
    pa.regions = (double *const *)d_regions; pa.own = d_own; pa.flags = d_flags; pa.step = d_step; pa.err = d_err;
    pa.ticket = d_step + 1; pa.peer_flags = (unsigned int *const *)d_peer_flags;
    pa.world = world; pa.rank = rank; pa.block_len = block_len_doubles;
    const int grid = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    if (publish) peer_publish_f32_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(d_buf, n, pa);
    else         peer_collect_f32_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(d_buf, n, pa);
    TSPWS_LAUNCHED_N(publish ? "peer_publish_f32_kernel" : "peer_collect_f32_kernel", (cudaStream_t)st);
    return 0;
}

/* peer variants of the two group_post2 launches: publish (push into every region) and the sum over ranks */
extern "C" int tspws_cu_group_post2_peer(int publish, const double *d_slots, const double *d_slot_tot, const int *d_slot_off, int N, int K,
                                         int ntile, int do_rm, int do_fold, int split, double *d_gsum,
                                         float *d_out32, double *d_gmean, void *d_regions, double *d_own, const unsigned int *d_flags,
                                         unsigned int *d_step, unsigned int *d_err, void *d_peer_flags, int rank, int world, tspws_stream_t st)
{
    const size_t smem = (size_t)K * sizeof(double) + (size_t)K * world;
    if (smem > 40 * 1024 || world > 32 * GP2_LANES) return (int)cudaErrorInvalidValue;
    tspws_peer_args pa;
    pa.regions = (double *const *)d_regions; pa.own = d_own; pa.flags = d_flags; pa.step = d_step; pa.err = d_err;
    pa.ticket = d_step + 1; pa.peer_flags = (unsigned int *const *)d_peer_flags;
    pa.world = world; pa.rank = rank; pa.block_len = (long long)K * N + (long long)K * ntile + K;
    if (publish)
        group_post2_kernel<false><<<(N + 31) / 32, 32 * GP2_LANES, smem, (cudaStream_t)st>>>(d_slots, d_slot_tot, d_slot_off, N, K, ntile,
            0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, pa, 1.0 / (double)N, 1.0 / (double)K);
    else
        group_post2_kernel<true><<<(N + 31) / 32, 32 * GP2_LANES, smem, (cudaStream_t)st>>>(nullptr, nullptr, nullptr, N, K, ntile,
            1, do_rm, do_fold, split, d_gsum, nullptr, d_out32, d_gmean, pa, 1.0 / (double)N, 1.0 / (double)K);
    TSPWS_LAUNCHED_N(publish ? "group_post2_kernel" : "group_post2_peer_kernel", (cudaStream_t)st);
    return 0;
}

extern "C" int tspws_cu_group_post2(const double *d_slots, const double *d_slot_tot, const int *d_slot_off, int N, int K,
                                    int ntile, int mode, int do_rm, int do_fold, int split, double *d_gsum, double *d_gtile,
                                    float *d_out32, double *d_gmean, tspws_stream_t st)
{
    const size_t smem = (size_t)K * sizeof(double);
    if (smem > 40 * 1024) return (int)cudaErrorInvalidValue;
    tspws_peer_args pa;
    memset(&pa, 0, sizeof pa);
    LAUNCH_DEP(group_post2_kernel<false>, dim3((N + 31) / 32), dim3(32 * GP2_LANES), smem, (cudaStream_t)st, d_slots, d_slot_tot, d_slot_off, N, K, ntile,
               mode, do_rm, do_fold, split, d_gsum, d_gtile, d_out32, d_gmean, pa, 1.0 / (double)N, 1.0 / (double)K);
    TSPWS_LAUNCHED_N("group_post2_kernel", (cudaStream_t)st);
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
    TSPWS_LAUNCHED_N("fwd_kernel", s);
    return 0;
}

extern "C" int tspws_cu_forward_one(const tspws_devframe_t *f, const float *d_x, int item0, void *d_ps, void *d_st, tspws_stream_t st)
{
    const int n = f->nitems - item0;
    if (n <= 0) return 0;
    dim3 grid((n + FWD_WARPS - 1) / FWD_WARPS, 1);
    fwd_kernel<float, float, 1><<<grid, FWD_WARPS * 32, 0, (cudaStream_t)st>>>(f->d_scale, (const float2 *)f->d_w32, f->d_items + item0, n,
        d_x, f->N, 1, 1, f->ncoef, (float2 *)d_ps, (float2 *)d_st);
    TSPWS_LAUNCHED_N("fwd_kernel", (cudaStream_t)st);
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
/* polyphase register-tiled forward transform (fp32 fast path)                           */
/*                                                                                       */
/*   With c = A D - e (0 <= e < D) and the conjugated filter padded by e leading zeros,  */
/*   g[t] = conj(w[t - e]):                                                              */
/*       Y[m] = sum_{rho<D} sum_{q<Qs} x[((m - A + q) D + rho) mod N] * g[q D + rho]     */
/*   i.e. D short correlations (Qs ~ 11..19 taps) of the polyphase components of x.      */
/*   A lane owns one phase rho and FW_R = 16 consecutive coefficients: it slides a       */
/*   register window of R+Qs-1 samples (stride D) past Qs filter taps -> 2 R Qs FMAs for */
/*   R+Qs-1 + Qs loads.  Lanes of a warp cover                                           */
/*     D = 2^lg < 32 : D phases x 32/D coefficient tiles; x is staged in (skewed,        */
/*                     bank-conflict-free) shared memory, warp-private;                  */
/*     D >= 32       : 32 consecutive phases of one tile; x is read from global/L1       */
/*                     (consecutive lanes = consecutive samples), looping over blocks of */
/*                     32 phases.                                                        */
/*   The per-lane partial sums (over phases) are combined with a transposing butterfly   */
/*   (log2 steps, halving the live values each step) that leaves every lane holding      */
/*   complete coefficients, which go straight into the phasor epilogue.                  */
/* ===================================================================================== */
#define FW_R 16
#ifndef FW_WARPS
#define FW_WARPS 4
#endif
#ifndef FW_MINB
#define FW_MINB 4
#endif

__device__ __forceinline__ double weight_factor(const tspws_weight_t r, double pr, double pi);

__device__ __forceinline__ void phasor_fast(float re, float im, float &pr, float &pi)
{
    const float s = re * re + im * im;
    if (s > 1e-30f && s < 1e30f) {            /* no under/overflow in the square: one MUFU */
        const float inv = rsqrtf(s);
        pr = re * inv; pi = im * inv;
    } else {                                  /* tiny, huge, zero or non-finite: careful path */
        const float a = hypotf(re, im);
        if (a > 0.f && a <= 3.0e38f) { pr = re / a; pi = im / a; }
        else { pr = 0.f; pi = 0.f; }          /* skipped like src/ts_pws1f_lib.c:492 */
    }
}

template <int LGR>
__device__ __forceinline__ void lane_transpose_reduce(float (&v)[2 * FW_R], const int lane)
{
#pragma unroll
    for (int b = LGR - 1; b >= 0; b--) {
        const int  half = 16 >> (LGR - 1 - b);
        const bool up   = (lane >> b) & 1;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << b);
        }
    }
}

/* ---- cp.async helpers (Ampere+ LDGSTS; fine on sm_100a) ---- */
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

/* One pass: 16 coefficients x Qs taps of one phase.
 *   loadx(j): window sample j of the current 32-tap block (j is a compile-time constant after
 *   unrolling, so addresses are immediates), loadg(q): filter tap q of the block, advance(): step
 *   both to the next 32-tap block.  Qs is a multiple of 4 (the host pads the filters with zero
 *   taps), the window source holds Qs + 16 samples (exactly what the taps reach) and the filter array
 *   Qs + 8 taps; window loads and refills are guarded per group of 4 with uniform predicates.
 * The window lives in a 32-slot register ring (slot = j mod 32): every index is static after
 * unrolling 32 taps; a slot is refilled right after its last use, 16 taps before it is needed again,
 * filter taps 4 taps ahead -- loads never stall the FMAs.
 * Arithmetic is Blackwell's packed FFMA2 (fma.rn.f32x2): one instruction updates (Re, Im) of a
 * coefficient, acc[r] += x * (g.re, g.im). */
#define FW_WS 32
#ifndef FW_GQ5
#define FW_GQ5 4          /* filter taps in flight per lane when D >= 32; 8 measured no faster (450 vs 462 us, 64 traces) */
#endif
template <int GQ, typename LoadX, typename LoadG, typename Advance>
__device__ __forceinline__ void fw_taps(float2 (&acc)[FW_R], const int Qs, const int Qw, LoadX loadx, LoadG loadg, Advance advance)
{
    float2 win[FW_WS];
    float2 gq[GQ];
    /* only the Qs + FW_R - 1 window samples the taps reach are loaded (groups of 4; Qs % 4 == 0); Qw = Qs,
     * or Qs + 16 when the full ring is staged */
#pragma unroll
    for (int j = 0; j < FW_WS; j += 4) {
        if (j < Qw + FW_R) {
#pragma unroll
            for (int u = 0; u < 4; u++) { const float t = loadx(j + u); win[j + u] = make_float2(t, t); }
        }
    }
#pragma unroll
    for (int q = 0; q < GQ; q++) gq[q] = loadg(q);
    for (int t0 = 0; t0 < Qs; t0 += FW_WS) {
#pragma unroll
        for (int tq = 0; tq < FW_WS; tq += 4) {
            if (t0 + tq < Qs) {
                const bool refill = t0 + tq + FW_R < Qw;        /* sample t0 + tq + u + 32 is still needed */
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int tt = tq + u;
                    const float2 gv = gq[tt & (GQ - 1)];
#pragma unroll
                    for (int r = 0; r < FW_R; r++) acc[r] = __ffma2_rn(win[(tt + r) % FW_WS], gv, acc[r]);
                    if (refill) { const float t = loadx(tt + FW_WS); win[tt] = make_float2(t, t); }
                    gq[tt & (GQ - 1)] = loadg(tt + GQ);
                }
            }
        }
        advance();
    }
}

/* Stage the window [base, base+W) (circular) of one trace into a warp-private, skewed smem buffer:
 * sample i goes to i + (i / (16 D)) D, i.e. tiles of 16 D samples at a pitch of 17 D floats, which
 * makes the compute-side accesses of lanes (tile a, phase rho) bank-conflict free. */
template <int LG>
__device__ __forceinline__ void fw_stage(float *buf, const float *__restrict__ xk, const int base,
                                         const int W, const int N, const int lane, const bool vec_ok)
{
    constexpr int LP  = 1 << LG;
    constexpr int CPF = LP >= 4 ? 4 : LP;             /* floats per copy: 16 B when D >= 4 */
    if (vec_ok) {
        for (int i = lane * CPF; i < W; i += 32 * CPF) {
            int pos = base + i;
            if (pos >= N) pos -= N;
            cp_async<CPF * 4>(buf + i + (i >> (4 + LG)) * LP, xk + pos);
        }
    } else {
        for (int i = lane; i < W; i += 32) {
            int pos = base + i;
            if (pos >= N) pos -= N;
            cp_async<4>(buf + i + (i >> (4 + LG)) * LP, xk + pos);
        }
    }
}

template <int LG, bool PER_TRACE>
__device__ __forceinline__ void
fwd2_body(const tspws_scale_t &sc, const tspws_fwd2_item_t &it, const float2 *__restrict__ g_all,
          const float *__restrict__ x, const int N, const int k_begin, const int k_end,
          const int ncoef, const int chunk_idx, float2 *__restrict__ part_ps, float2 *__restrict__ part_st,
          float2 *__restrict__ ypart, const int ypart_len, float *xs, const int xs_half, const int lane)
{
    constexpr int LGR  = LG;                       /* lane bits that index phases            */
    constexpr int NVAL = 32 >> LGR;                /* floats of complete coefficients / lane */
    constexpr int LGS  = LG < 5 ? LG : 0;
    constexpr int LP   = 1 << LGS, TILES = 32 >> LGS;
    const int D = sc.D, Qs = sc.Qs;
    const float2 *__restrict__ g = g_all + sc.g_off;
    float ps[NVAL], st[NVAL];
#pragma unroll
    for (int i = 0; i < NVAL; i++) { ps[i] = 0.f; st[i] = 0.f; }

    const int base = it.base;                      /* sample of (tile 0, tap 0, phase 0), computed on the host */
    const int Qw = Qs;

    /* LG < 5: D = 2^LG phases x (32 >> LG) tiles; window staged by cp.async, double buffered:
     * exactly the (TILES - 1) 16 + Qs + 16 samples per phase the taps reach. */
    const int  W      = (TILES * FW_R + Qw) * LP;
    const bool vec_ok = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    /* LG == 5: per (trace, block of 32 phases) the Qs + 16 window rows of 32 consecutive samples are
     * staged by cp.async into a warp-private double buffer one pass ahead (16-byte copies when the
     * geometry allows), so the FMA loop only ever waits on shared memory.  (One 128-byte bulk copy per
     * row -- cp.async.bulk on an mbarrier -- was measured slower: the copy engine is rate-bound on small copies.) */
    const int npb = it.pb1 - it.pb0;
    const int nrows = Qw + FW_R;
    const bool vec5 = vec_ok && (D % 4 == 0);
    const int GS = (D + 31) & ~31;
    auto stage5 = [&](float *buf, const int k, const int pb) {
        const float *__restrict__ xk = x + (size_t)k * N;
        if (vec5) {
            int row = lane >> 3;
            const int q4 = (lane & 7) * 4;
            int pos = (int)(((long long)base + pb * 32 + q4 + (long long)row * D) % N);
            const int step = (int)((4LL * D) % N);
            for (; row < nrows; row += 4) {
                cp_async<16>(buf + row * 32 + q4, xk + pos);
                pos += step;
                if (pos >= N) pos -= N;
            }
        } else {
            int pos = (int)(((long long)base + pb * 32 + lane) % N);
            const int step = D % N;
            for (int row = 0; row < nrows; row++) {
                cp_async<4>(buf + row * 32 + lane, xk + pos);
                pos += step;
                if (pos >= N) pos -= N;
            }
        }
    };
    if (LG < 5) {
        __syncwarp();
        fw_stage<LGS>(xs, x + (size_t)k_begin * N, base, W, N, lane, vec_ok);
        cp_async_commit();
    }
    if (LG == 5) {
        __syncwarp();
        stage5(xs, k_begin, it.pb0);
        cp_async_commit();
    }
    int pass = 0;

    for (int k = k_begin; k < k_end; k++) {
        const float *__restrict__ xk = x + (size_t)k * N;
        float2 acc[FW_R];
#pragma unroll
        for (int i = 0; i < FW_R; i++) acc[i] = make_float2(0.f, 0.f);

        if (LG < 5) {
            const float *cur = xs + ((k - k_begin) & 1) * xs_half;
            if (k + 1 < k_end) {                   /* next trace's window flies in during the FMAs */
                fw_stage<LGS>(xs + (((k - k_begin) & 1) ^ 1) * xs_half, xk + N, base, W, N, lane, vec_ok);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const int a = lane >> LGS, rho = lane & (LP - 1);
            const float  *xb = cur + a * ((FW_R + 1) * LP) + rho;      /* sample j of the block at (j + (j >> 4)) LP */
            const float2 *gb = g + rho;
            fw_taps<4>(acc, Qs, Qw,
                    [&](const int j) { return xb[(j + (j >> 4)) * LP]; },
                    [&](const int q) { return __ldg(gb + q * LP); },
                    [&]() { xb += 34 * LP; gb += 32 * LP; });
            __syncwarp();
        } else {
            /* D >= 32: 32 consecutive phases per block; the filter rows are zero-padded to a multiple
             * of 32 phases, so idle lanes add zeros */
            for (int pb = it.pb0; pb < it.pb1; pb++, pass++) {
                const float *cur = xs + (pass & 1) * xs_half;
                int kn = k, pbn = pb + 1;
                if (pbn == it.pb1) { pbn = it.pb0; kn = k + 1; }
                if (kn < k_end) {
                    stage5(xs + ((pass & 1) ^ 1) * xs_half, kn, pbn);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                const float  *xb = cur + lane;
                const float2 *gp = g + pb * 32 + lane;                  /* taps are consumed strictly in order */
                fw_taps<FW_GQ5>(acc, Qs, Qw,
                        [&](const int j) { return xb[j * 32]; },
                        [&](const int) { const float2 v = __ldg(gp); gp += GS; return v; },
                        [&]() { xb += 32 * 32; });
                __syncwarp();
            }
            (void)npb;
        }
        float(&v)[2 * FW_R] = reinterpret_cast<float(&)[2 * FW_R]>(acc);

        lane_transpose_reduce<LGR>(v, lane);

        if (NVAL >= 2) {
            const int a = lane >> LG, rho = lane & ((1 << LG) - 1);
            const int m_first = it.m0 + a * FW_R + (rho * NVAL) / 2;
#pragma unroll
            for (int i = 0; i < NVAL / 2; i++) {
                const float re = v[2 * i], im = v[2 * i + 1];
                if (PER_TRACE) {
                    const int m = m_first + i;
                    if (m < sc.Ns)
                        ypart[(size_t)k * ypart_len + sc.ypart_off + (size_t)it.split * sc.Ns + m] = make_float2(re, im);
                } else {
                    float pr, pi;
                    phasor_fast(re, im, pr, pi);
                    ps[2 * i] += pr; ps[2 * i + 1] += pi;
                    st[2 * i] += re; st[2 * i + 1] += im;
                }
            }
        } else {
            /* one float per lane: even lanes hold Re, odd lanes Im of coefficient lane/2 */
            const float mine  = v[0];
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const bool  odd   = lane & 1;
            if (PER_TRACE) {
                const int m = it.m0 + (lane >> 1);
                if (m < sc.Ns) {
                    float *o = reinterpret_cast<float *>(ypart + (size_t)k * ypart_len + sc.ypart_off + (size_t)it.split * sc.Ns + m);
                    o[odd] = mine;
                }
            } else {
                float pr, pi;
                phasor_fast(odd ? other : mine, odd ? mine : other, pr, pi);
                ps[0] += odd ? pi : pr;
                st[0] += mine;
            }
        }
    }

    if (!PER_TRACE) {
        const size_t obase = (size_t)chunk_idx * ncoef + sc.coef_off;
        if (NVAL >= 2) {
            const int a = lane >> LG, rho = lane & ((1 << LG) - 1);
            const int m_first = it.m0 + a * FW_R + (rho * NVAL) / 2;
#pragma unroll
            for (int i = 0; i < NVAL / 2; i++) {
                const int m = m_first + i;
                if (m < sc.Ns) {
                    part_ps[obase + m] = make_float2(ps[2 * i], ps[2 * i + 1]);
                    part_st[obase + m] = make_float2(st[2 * i], st[2 * i + 1]);
                }
            }
        } else {
            const int m = it.m0 + (lane >> 1);
            if (m < sc.Ns) {
                reinterpret_cast<float *>(part_ps + obase + m)[lane & 1] = ps[0];
                reinterpret_cast<float *>(part_st + obase + m)[lane & 1] = st[0];
            }
        }
    }
}

/* chunk-accumulate mode: one instantiation per LG (register budgets differ); items of that LG
 * are [item_begin, item_end).  grid = (ceil(items / 8 warps), chunks). */
template <int LG>
__global__ void __launch_bounds__(FW_WARPS * 32, FW_MINB)
fwd2_acc_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ g_all,
                const tspws_fwd2_item_t *__restrict__ items, int item_begin, int item_end,
                const float *__restrict__ x, int N, int ntr, int chunk, int ncoef,
                float2 *__restrict__ part_ps, float2 *__restrict__ part_st, int smem_per_warp, int cb)
{
    extern __shared__ __align__(128) float fw_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    /* 1-D grid, two-level order: the chunks are taken in super-blocks of `cb` chunks whose traces fit
     * in L2 together; inside a super-block the chunk index runs fastest (CTAs running together share
     * the item's filter rows in L1), then the item group, so every trace comes from HBM once. */
    const int nchunk = (ntr + chunk - 1) / chunk;
    const int nig = (item_end - item_begin + FW_WARPS - 1) / FW_WARPS;
    const int per_super = nig * cb;
    const int super = blockIdx.x / per_super, rem = blockIdx.x - super * per_super;
    const int cb_here = min(cb, nchunk - super * cb);          /* the last super-block may be short */
    const int ig = rem / cb_here, cidx = super * cb + rem - ig * cb_here;
    if (ig >= nig || cidx >= nchunk) return;
    const int idx = item_begin + ig * FW_WARPS + w;
    if (idx >= item_end) return;
    const tspws_fwd2_item_t it = items[idx];
    const tspws_scale_t sc = scales[it.s];
    const int k0 = cidx * chunk, k1 = min(ntr, k0 + chunk);
    fwd2_body<LG, false>(sc, it, g_all, x, N, k0, k1, ncoef, cidx, part_ps, part_st, nullptr, 0,
                         fw_smem + (size_t)w * smem_per_warp, smem_per_warp / 2, lane);
}

/* per-trace mode: one kernel for all LG (no accumulators -> one register budget fits all).
 * grid = (ceil(items / 8 warps), traces). */
#ifndef FW_PT_MINB
#define FW_PT_MINB FW_MINB
#endif
__global__ void __launch_bounds__(FW_WARPS * 32, FW_PT_MINB)
fwd2_pertrace_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ g_all,
                     const tspws_fwd2_item_t *__restrict__ items, int nitems,
                     const float *__restrict__ x, int N, float2 *__restrict__ ypart, int ypart_len,
                     int smem_per_warp, int ntr, int tchunk, int long_L, int long_row0)
{
    extern __shared__ __align__(128) float fw_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    /* blockIdx.x = run of tchunk traces (fastest): the traces of one item run together -> one code
     * variant and one set of filter rows live at a time.  Inside a run the item set-up (scale record,
     * index arithmetic) is paid once and the next trace's window is staged under the FMAs. */
    const int idx = blockIdx.y * FW_WARPS + w;
    if (idx >= nitems) return;
    const tspws_fwd2_item_t it = items[idx];
    const tspws_scale_t sc = scales[it.s];
    const int k = blockIdx.x * tchunk, k1 = min(ntr, k + tchunk);
    float *xs = fw_smem + (size_t)w * smem_per_warp;
    /* mean split: scales with long filters read the unsplit traces stored from row long_row0 on */
    if (sc.L >= long_L) x += (size_t)long_row0 * N;
    switch (sc.lg) {
    case 0: fwd2_body<0, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 1: fwd2_body<1, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 2: fwd2_body<2, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 3: fwd2_body<3, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 4: fwd2_body<4, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    default: fwd2_body<5, true>(sc, it, g_all, x, N, k, k1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    }
}

/* generic kernel in per-trace mode: for the scales the polyphase kernel does not cover (fp32), and
 * -- in double -- for the fp64 transform of the group mean.  out: [ntr][out_ld] complex at
 * out_off(scale) + m, where out_off is ypart_off (fp32 partial-Y layout) or coef_off (plain). */
template <typename T, bool PLAIN>
__global__ void __launch_bounds__(FWD_WARPS * 32)
fwd_pertrace_generic_kernel(const tspws_scale_t *__restrict__ scales, const typename cplx<T>::type *__restrict__ wall,
                            const tspws_fwd_item_t *__restrict__ items, int nitems,
                            const T *__restrict__ x, int N, typename cplx<T>::type *__restrict__ out, int out_ld,
                            int long_L, int long_row0)
{
    typedef typename cplx<T>::type T2;
    const int witem = blockIdx.x * FWD_WARPS + (threadIdx.x >> 5);
    if (witem >= nitems) return;
    const tspws_fwd_item_t it = items[witem];
    const tspws_scale_t sc = scales[it.s];
    const int lane = threadIdx.x & 31;
    const int G = sc.G, g = lane & (G - 1), og = lane / G, opp = 32 / G;
    const T2 *__restrict__ w = wall + sc.tap_off;
    const T *__restrict__ xk = x + ((size_t)blockIdx.y + (sc.L >= long_L ? long_row0 : 0)) * N;
    for (int mb = 0; mb < it.nout; mb += opp) {
        const bool valid = (mb + og) < it.nout;
        const int  m     = it.m0 + (valid ? mb + og : it.nout - 1);
        long long pll = ((long long)m * sc.D - sc.c + g) % N;
        if (pll < 0) pll += N;
        int p = (int)pll;
        T ar = 0, ai = 0;
#pragma unroll 4
        for (int l = g; l < sc.L; l += G) {
            const T2 wv = w[l];
            const T  xv = xk[p];
            ar = fma(xv, wv.x, ar);
            ai = fma(-xv, wv.y, ai);
            p += G;
            if (p >= N) p -= N;
        }
        for (int o = G >> 1; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (g == 0 && valid) {
            T2 v; v.x = ar; v.y = ai;
            out[(size_t)blockIdx.y * out_ld + (PLAIN ? sc.coef_off : sc.ypart_off) + m] = v;
        }
    }
}

/* fp64 forward transform of a few traces (the group mean of the mean split, see group_post2_kernel).
 *   x: [ntr][3N], every trace stored three times back to back, so that the circular index
 *      (m D - c + l) mod N is a plain offset base + r D + l < 3N and the tap loop has no wrap test;
 *   every lane keeps F64_R consecutive coefficients of one scale in registers: a filter tap is loaded
 *   once per F64_R coefficients, 2 DFMA per loaded sample, ~1.7 instructions per DFMA.
 *   CTAs [0, nshort_cta): 8 warp items each; G lanes split the taps of a coefficient group and the
 *   32/G groups of a warp advance through the item's run together (filters shorter than F64_LONG);
 *   remaining CTAs: the whole CTA splits the taps of one group of F64_R coefficients (long filters). */
#define F64_R 4
template <int G, int STRIDE>
__device__ __forceinline__ void f64_taps(const double2 *__restrict__ pw, const double *__restrict__ px, const int (&off)[F64_R],
                                         int l, const int L, double (&ar)[F64_R], double (&ai)[F64_R])
{
    pw += l; px += l;
    for (; l + 3 * STRIDE < L; l += 4 * STRIDE) {
        double2 wv[4];
        double  xv[4][F64_R];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            wv[j] = __ldg(pw + j * STRIDE);
#pragma unroll
            for (int r = 0; r < F64_R; r++) xv[j][r] = __ldg(px + off[r] + j * STRIDE);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < F64_R; r++) { ar[r] = fma(xv[j][r], wv[j].x, ar[r]); ai[r] = fma(-xv[j][r], wv[j].y, ai[r]); }
        pw += 4 * STRIDE; px += 4 * STRIDE;
    }
    for (; l < L; l += STRIDE) {
        const double2 wv = __ldg(pw);
#pragma unroll
        for (int r = 0; r < F64_R; r++) { const double xv = __ldg(px + off[r]); ar[r] = fma(xv, wv.x, ar[r]); ai[r] = fma(-xv, wv.y, ai[r]); }
        pw += STRIDE; px += STRIDE;
    }
}

template <int G>
__device__ __forceinline__ void f64_short_item(const tspws_scale_t &sc, const tspws_fwd_item_t &it, const double2 *__restrict__ w,
                                               const double *__restrict__ xk, int N, double2 *__restrict__ out, int lane)
{
    const int g = lane & (G - 1), og = lane / G, opp = 32 / G;
    for (int mb = 0; mb < it.nout; mb += opp * F64_R) {
        const int  mrel  = mb + og * F64_R;
        const bool valid = mrel < it.nout;
        const int  m     = it.m0 + (valid ? mrel : 0);
        const int  nr    = valid ? min(F64_R, it.nout - mrel) : 0;
        long long pll = ((long long)m * sc.D - sc.c) % N;
        if (pll < 0) pll += N;
        int off[F64_R];
#pragma unroll
        for (int r = 0; r < F64_R; r++) off[r] = min(r, max(nr - 1, 0)) * sc.D;
        double ar[F64_R], ai[F64_R];
#pragma unroll
        for (int r = 0; r < F64_R; r++) ar[r] = ai[r] = 0.0;
        f64_taps<G, G>(w, xk + (int)pll, off, g, sc.L, ar, ai);
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < F64_R; r++) {
                ar[r] += __shfl_xor_sync(0xffffffffu, ar[r], o);
                ai[r] += __shfl_xor_sync(0xffffffffu, ai[r], o);
            }
        if (g == 0)
#pragma unroll
            for (int r = 0; r < F64_R; r++)
                if (r < nr) out[sc.coef_off + m + r] = make_double2(ar[r], ai[r]);
    }
}

/* body of one CTA of NT threads; bid = CTA index within the fp64 work list, xk / o = this trace */
template <int NT>
__device__ __forceinline__ void f64_cta_body(const tspws_scale_t *__restrict__ scales, const double2 *__restrict__ wall,
                                             const tspws_fwd_item_t *__restrict__ items, int nitems, int nshort_cta,
                                             const int2 *__restrict__ coefs, const double *__restrict__ xk, int N,
                                             double2 *__restrict__ o, size_t split_stride, int bid, double (*sh)[2 * F64_R])
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (bid < nshort_cta) {
        const int witem = bid * NW + warp;
        if (witem >= nitems) return;
        const tspws_fwd_item_t it = items[witem];
        const tspws_scale_t sc = scales[it.s];
        const double2 *__restrict__ w = wall + sc.tap_off;
        switch (sc.G) {
        case 1:  f64_short_item<1>(sc, it, w, xk, N, o, lane); break;
        case 2:  f64_short_item<2>(sc, it, w, xk, N, o, lane); break;
        case 4:  f64_short_item<4>(sc, it, w, xk, N, o, lane); break;
        case 8:  f64_short_item<8>(sc, it, w, xk, N, o, lane); break;
        case 16: f64_short_item<16>(sc, it, w, xk, N, o, lane); break;
        default: f64_short_item<32>(sc, it, w, xk, N, o, lane); break;
        }
        return;
    }
    const int2 cm = coefs[bid - nshort_cta];               /* (scale | tap split << 16, first coefficient of the group) */
    const tspws_scale_t sc = scales[cm.x & 0xffff];
    const int q = cm.x >> 16, nsp = TSPWS_F64_NSPLIT(sc.L), Lq = (sc.L + nsp - 1) / nsp;
    o += (size_t)q * split_stride;
    const int nr = min(F64_R, sc.Ns - cm.y);
    long long pll = ((long long)cm.y * sc.D - sc.c) % N;
    if (pll < 0) pll += N;
    int off[F64_R];
#pragma unroll
    for (int r = 0; r < F64_R; r++) off[r] = min(r, nr - 1) * sc.D;
    double ar[F64_R], ai[F64_R];
#pragma unroll
    for (int r = 0; r < F64_R; r++) ar[r] = ai[r] = 0.0;
    f64_taps<NT, NT>(wall + sc.tap_off, xk + (int)pll, off, q * Lq + (int)threadIdx.x, min(sc.L, (q + 1) * Lq), ar, ai);
#pragma unroll
    for (int r = 0; r < F64_R; r++) { ar[r] = warp_sum(ar[r]); ai[r] = warp_sum(ai[r]); }
    if (lane == 0)
#pragma unroll
        for (int r = 0; r < F64_R; r++) { sh[warp][2 * r] = ar[r]; sh[warp][2 * r + 1] = ai[r]; }
    __syncthreads();
    if (threadIdx.x < 2 * F64_R) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < NW; i++) t += sh[i][threadIdx.x];
        sh[0][threadIdx.x] = t;
    }
    __syncthreads();
    if ((int)threadIdx.x < nr) o[sc.coef_off + cm.y + threadIdx.x] = make_double2(sh[0][2 * threadIdx.x], sh[0][2 * threadIdx.x + 1]);
}

__global__ void __launch_bounds__(256)
fwd_f64_kernel(const tspws_scale_t *__restrict__ scales, const double2 *__restrict__ wall,
               const tspws_fwd_item_t *__restrict__ items, int nitems, int nshort_cta,
               const int2 *__restrict__ coefs, const double *__restrict__ x, int N,
               double2 *__restrict__ out, int out_ld)
{
    __shared__ double sh[8][2 * F64_R];
    f64_cta_body<256>(scales, wall, items, nitems, nshort_cta, coefs, x + (size_t)blockIdx.y * 3 * N, N,
                      out + (size_t)blockIdx.y * out_ld, (size_t)gridDim.y * out_ld, (int)blockIdx.x, sh);
}

/* ---- fused phasor combine of the tail (src/ts_pws1f_lib.c:897-904 + :909-984 in the same launch) ----
 * Every writer of per-trace coefficients (fp32 item x trace, fp64 mean item) takes a ticket on each
 * 64-coefficient tile it wrote into; the writer that completes a tile combines it right away: sum of the
 * phase / tap splits + mean -> unit phasors -> PS, ST, weights, (Zout, ST) for the inverse.  The partial
 * coefficients make an L2 round trip only (never HBM), and no separate combine launch follows. */
struct tspws_tailfuse_t {
    const int4 *tiles;            /* (ypart offset, coefficient offset, nvalid | nsplit << 8 | nsp64 << 16, Ns) */
    const int2 *expect;           /* writers of a tile: (fp32 items, fp64 items) */
    unsigned int *tickets;        /* [ntile], zero between launches (the last writer re-arms) */
    const tspws_weight_t *rule;
    float2 *ps, *st, *zout;
    int ncoef;
};

/* all lanes have written; returns true in the warp that completes the tile (uniform) */
__device__ __forceinline__ bool ticket_arrive(unsigned int *tk, const unsigned int expect, const int lane)
{
    __threadfence();                                   /* every lane publishes its own stores */
    __syncwarp();
    unsigned int old = 0;
    if (lane == 0) old = atomicAdd(tk, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    const bool last = (old + 1u == expect);
    if (last) {
        if (lane == 0) *tk = 0;                        /* re-arm for the next launch */
        __threadfence();
    }
    return last;
}

__device__ __forceinline__ void tail_combine_tile(const int4 tile, const float2 *__restrict__ ypart, const int ypart_len,
                                                  const int ntr, const double2 *__restrict__ ymean, const tspws_tailfuse_t &fz,
                                                  const int lane)
{
    const int nvalid = tile.z & 255, nsplit = (tile.z >> 8) & 255, nsp64 = tile.z >> 16, Ns = tile.w;
    const tspws_weight_t rule = *fz.rule;
    /* lane owns coefficients lane and lane + 32 of the tile; the loads of both and of several traces are in flight together */
    const int  c0 = lane, c1 = lane + 32;
    const bool v0 = c0 < nvalid, v1 = c1 < nvalid;
    const int  i0 = v0 ? c0 : 0, i1 = v1 ? c1 : 0;
    double m0r = 0.0, m0i = 0.0, m1r = 0.0, m1i = 0.0;
    for (int sp = 0; sp < nsp64; sp++) {                /* tap splits of the fp64 mean transform, in slot order */
        const double2 a = __ldcg(ymean + (size_t)sp * fz.ncoef + tile.y + i0);
        const double2 b = __ldcg(ymean + (size_t)sp * fz.ncoef + tile.y + i1);
        m0r += a.x; m0i += a.y; m1r += b.x; m1i += b.y;
    }
    float  p0r = 0.f, p0i = 0.f, p1r = 0.f, p1i = 0.f;
    double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
    const float2 *yp = ypart + tile.x;
    if (nsplit == 1) {
#pragma unroll 5
        for (int k = 0; k < ntr; k++) {
            const float2 a = __ldcg(yp + (size_t)k * ypart_len + i0);
            const float2 b = __ldcg(yp + (size_t)k * ypart_len + i1);
            const double ar = (double)a.x + m0r, ai = (double)a.y + m0i, br = (double)b.x + m1r, bi = (double)b.y + m1i;
            float u, v;
            phasor_fast((float)ar, (float)ai, u, v); p0r += u; p0i += v; s0r += ar; s0i += ai;
            phasor_fast((float)br, (float)bi, u, v); p1r += u; p1i += v; s1r += br; s1i += bi;
        }
    } else {
        for (int k = 0; k < ntr; k++) {
            const float2 *y = yp + (size_t)k * ypart_len;
            float are = 0.f, aim = 0.f, bre = 0.f, bim = 0.f;
#pragma unroll 8
            for (int sp = 0; sp < nsplit; sp++) {
                const float2 a = __ldcg(y + (size_t)sp * Ns + i0);
                const float2 b = __ldcg(y + (size_t)sp * Ns + i1);
                are += a.x; aim += a.y; bre += b.x; bim += b.y;
            }
            const double ar = (double)are + m0r, ai = (double)aim + m0i, br = (double)bre + m1r, bi = (double)bim + m1i;
            float u, v;
            phasor_fast((float)ar, (float)ai, u, v); p0r += u; p0i += v; s0r += ar; s0i += ai;
            phasor_fast((float)br, (float)bi, u, v); p1r += u; p1i += v; s1r += br; s1i += bi;
        }
    }
    if (v0) {
        const double f = weight_factor(rule, (double)p0r, (double)p0i);
        const size_t o = (size_t)tile.y + c0;
        fz.ps[o] = make_float2(p0r, p0i);
        fz.st[o] = make_float2((float)s0r, (float)s0i);
        reinterpret_cast<float4 *>(fz.zout)[o] = make_float4((float)(s0r * f), (float)(s0i * f), (float)s0r, (float)s0i);
    }
    if (v1) {
        const double f = weight_factor(rule, (double)p1r, (double)p1i);
        const size_t o = (size_t)tile.y + c1;
        fz.ps[o] = make_float2(p1r, p1i);
        fz.st[o] = make_float2((float)s1r, (float)s1i);
        reinterpret_cast<float4 *>(fz.zout)[o] = make_float4((float)(s1r * f), (float)(s1i * f), (float)s1r, (float)s1i);
    }
}

/* Mean-split tail in ONE launch: the fp64 transform of the group mean (CTAs [0, nf64), one trace), the fp32
 * per-trace transforms of the residuals (the rest: trace fastest, then item group) and -- when fz.tickets
 * is set -- the phasor combine of every coefficient tile by the warp that completes it.
 * Launched separately the two transforms exclude each other -- either one fills the register files -- and
 * their times add; in one grid their CTAs share the SMs and the issue slots. */
/* Mean-split tail in ONE launch: the fp64 transform of the group mean (one trace), the fp32 per-trace
 * transforms of the residuals (trace fastest, then item group) and -- when fz.tickets is set -- the phasor
 * combine of every coefficient tile by the warp that completes it.
 * Launched separately the two transforms exclude each other -- either one fills the register files -- and
 * their times add; in one grid their CTAs share the SMs and the issue slots.
 * CTA order = longest first: the item groups with D >= 32 (coarsest scale first: up to 8 blocks of 32 phases
 * per item), then the fp64 CTAs, then the single-pass D < 32 groups.  Measured with the per-CTA timeline
 * (profiles/): the heavy CTAs (12..20 us each) otherwise start at 40 % of the grid and finish alone. */
__device__ __forceinline__ void
fwd_tail_cta(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ g_all,
                const tspws_fwd2_item_t *__restrict__ items, int nitems,
                const float *__restrict__ x, int N, float2 *__restrict__ ypart, int ypart_len,
                int smem_per_warp, int ntr, int long_L, int long_row0,
                const double2 *__restrict__ w64, const tspws_fwd_item_t *__restrict__ items64, int nitems64, int nshort_cta,
                const int2 *__restrict__ coefs64, int nf64, const double *__restrict__ x64, double2 *__restrict__ y64, int y64_ld,
                const tspws_tailfuse_t &fz, const int g5, float *fw_smem, double (*sh64)[2 * F64_R])
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngrp = (nitems + FW_WARPS - 1) / FW_WARPS;
    const int nheavy = (ngrp - g5) * ntr;                  /* CTAs of the groups [g5, ngrp) */
    int bid = blockIdx.x, b;
    if (bid < nheavy) b = g5 * ntr + bid;
    else if (bid < nheavy + nf64) b = -1 - (bid - nheavy);
    else b = bid - nheavy - nf64;
    if (b < 0) {
        const int fb = -1 - b;
        f64_cta_body<FW_WARPS * 32>(scales, w64, items64, nitems64, nshort_cta, coefs64, x64, N, y64, (size_t)y64_ld,
                                    fb, sh64);
        if (!fz.tickets) return;
        if (fb < nshort_cta) {
            const int witem = fb * FW_WARPS + w;
            if (witem >= nitems64) return;
            const tspws_fwd_item_t it = items64[witem];
            const int tile0 = scales[it.s].tile0;
            for (int t = tile0 + it.m0 / 64; t <= tile0 + (it.m0 + it.nout - 1) / 64; t++) {
                const int2 e = fz.expect[t];
                if (ticket_arrive(fz.tickets + t, (unsigned int)(e.x * ntr + e.y), lane))
                    tail_combine_tile(fz.tiles[t], ypart, ypart_len, ntr, y64, fz, lane);
            }
        } else if (w == 0) {                               /* the coefficients were stored by threads of warp 0 */
            const int2 cm = coefs64[fb - nshort_cta];
            const int t = scales[cm.x & 0xffff].tile0 + cm.y / 64;
            const int2 e = fz.expect[t];
            if (ticket_arrive(fz.tickets + t, (unsigned int)(e.x * ntr + e.y), lane))
                tail_combine_tile(fz.tiles[t], ypart, ypart_len, ntr, y64, fz, lane);
        }
        return;
    }
    const int idx = (b / ntr) * FW_WARPS + w;
    if (idx >= nitems) return;
    const tspws_fwd2_item_t it = items[idx];
    const tspws_scale_t sc = scales[it.s];
    const int k = b % ntr;
    float *xs = fw_smem + (size_t)w * smem_per_warp;
    if (sc.L >= long_L) x += (size_t)long_row0 * N;
    switch (sc.lg) {
    case 0: fwd2_body<0, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 1: fwd2_body<1, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 2: fwd2_body<2, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 3: fwd2_body<3, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    case 4: fwd2_body<4, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    default: fwd2_body<5, true>(sc, it, g_all, x, N, k, k + 1, 0, 0, nullptr, nullptr, ypart, ypart_len, xs, smem_per_warp / 2, lane); break;
    }
    if (!fz.tickets) return;
    {
        const int per_item = sc.lg < 5 ? (512 >> sc.lg) : FW_R;
        const int mend = min(it.m0 + per_item, sc.Ns) - 1;
        for (int t = sc.tile0 + it.m0 / 64; t <= sc.tile0 + mend / 64; t++) {
            const int2 e = fz.expect[t];
            if (ticket_arrive(fz.tickets + t, (unsigned int)(e.x * ntr + e.y), lane))
                tail_combine_tile(fz.tiles[t], ypart, ypart_len, ntr, y64, fz, lane);
        }
    }
}

__global__ void __launch_bounds__(FW_WARPS * 32, FW_PT_MINB)
fwd_tail_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ g_all,
                const tspws_fwd2_item_t *__restrict__ items, int nitems,
                const float *__restrict__ x, int N, float2 *__restrict__ ypart, int ypart_len,
                int smem_per_warp, int ntr, int long_L, int long_row0,
                const double2 *__restrict__ w64, const tspws_fwd_item_t *__restrict__ items64, int nitems64, int nshort_cta,
                const int2 *__restrict__ coefs64, int nf64, const double *__restrict__ x64, double2 *__restrict__ y64, int y64_ld,
                const tspws_tailfuse_t fz, const int g5)
{
    tl_scope tls(2);
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) float fw_smem[];
    __shared__ double sh64[FW_WARPS][2 * F64_R];
    fwd_tail_cta(scales, g_all, items, nitems, x, N, ypart, ypart_len, smem_per_warp, ntr, long_L, long_row0, w64, items64,
                 nitems64, nshort_cta, coefs64, nf64, x64, y64, y64_ld, fz, g5, fw_smem, sh64);
}

/* d_x: [ntr][3N] period-padded traces (see above) */
extern "C" int tspws_cu_forward_f64(const tspws_devframe_t *f, const double *d_x, int ntr, void *d_y, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    if (!f->d_w64) return (int)cudaErrorInvalidValue;
    const int nshort_cta = (f->nitems_short + 7) / 8;
    if (nshort_cta + f->nlong <= 0) return 0;
    dim3 grid(nshort_cta + f->nlong, ntr);
    fwd_f64_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(f->d_scale, (const double2 *)f->d_w64, f->d_items_short, f->nitems_short,
        nshort_cta, (const int2 *)f->d_long, d_x, f->N, (double2 *)d_y, f->ncoef);
    TSPWS_LAUNCHED_N("fwd_f64_kernel", (cudaStream_t)st);
    return 0;
}

extern "C" int tspws_cu_forward_fast(const tspws_devframe_t *f, const float *d_x, int ntr, int chunk,
                                     void *d_part_ps, void *d_part_st, int skip5, int lg_end, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    if (f->per_trace) return (int)cudaErrorInvalidValue;
    if (chunk < 1) chunk = 1;
    cudaStream_t s = (cudaStream_t)st;
    const int nchunk = (ntr + chunk - 1) / chunk;
    const size_t smem = (size_t)f->fw_smem_per_warp * FW_WARPS * sizeof(float);
    /* chunks per super-block.  Measured at 8192 x 16384 (w0 = 6): all chunks in one super-block (pure
     * chunk-fastest order, the traces are re-streamed from HBM once per item group) 16.3 ms for the
     * D >= 32 kernel; super-blocks of 12 chunks (traces L2-resident, but many different items, hence
     * filter rows, live per SM) 19.4 ms.  Filter-row locality wins; TSPWS_FW_CB overrides. */
    int cb = nchunk;
    { const char *e = getenv("TSPWS_FW_CB"); if (e && atoi(e) > 0) cb = atoi(e); }
    if (cb > nchunk) cb = nchunk;
#define FW_ACC_LAUNCH(LGV)                                                                              \
    if ((LGV == 5 || LGV < lg_end) && f->fw_lg_begin[LGV + 1] > f->fw_lg_begin[LGV] + (LGV == 5 ? skip5 : 0)) { \
        const int ib = f->fw_lg_begin[LGV] + (LGV == 5 ? skip5 : 0);                                    \
        const int nb = f->fw_lg_begin[LGV + 1] - ib;                                                    \
        const int nig = (nb + FW_WARPS - 1) / FW_WARPS;                                                 \
        /* full super-blocks have nig*cb CTAs; a short last one is padded to the same count (extra CTAs exit) */ \
        dim3 grid((unsigned)(((nchunk + cb - 1) / cb) * nig * cb));                                      \
        if (smem > 40 * 1024)                                                                           \
            cudaFuncSetAttribute(fwd2_acc_kernel<LGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        fwd2_acc_kernel<LGV><<<grid, FW_WARPS * 32, smem, s>>>(f->d_scale, (const float2 *)f->d_g32, \
            f->d_items2, ib, f->fw_lg_begin[LGV + 1], d_x, f->N, ntr, chunk, f->ncoef,                  \
            (float2 *)d_part_ps, (float2 *)d_part_st, f->fw_smem_per_warp, cb);                          \
        TSPWS_LAUNCHED_N("fwd2_acc_kernel<" #LGV ">", s);                                                                               \
    }
    FW_ACC_LAUNCH(0) FW_ACC_LAUNCH(1) FW_ACC_LAUNCH(2) FW_ACC_LAUNCH(3) FW_ACC_LAUNCH(4) FW_ACC_LAUNCH(5)
#undef FW_ACC_LAUNCH
    if (f->nitems_gen > 0) {
        dim3 grid((f->nitems_gen + FWD_WARPS - 1) / FWD_WARPS, nchunk);
        if (chunk >= 4)
            fwd_kernel<float, float, 4><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, (const float2 *)f->d_w32, f->d_items_gen,
                f->nitems_gen, d_x, f->N, ntr, chunk, f->ncoef, (float2 *)d_part_ps, (float2 *)d_part_st);
        else
            fwd_kernel<float, float, 1><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, (const float2 *)f->d_w32, f->d_items_gen,
                f->nitems_gen, d_x, f->N, ntr, chunk, f->ncoef, (float2 *)d_part_ps, (float2 *)d_part_st);
        TSPWS_LAUNCHED_N("fwd_kernel", s);
    }
    return 0;
}

/* fp64 transform of one period-padded trace d_x64 -> d_y64 [split][ncoef] together with the per-trace fp32
 * transforms of ntr traces (see fwd_tail_kernel); only for frames fully covered by the polyphase kernel */
extern "C" int tspws_cu_forward_tail(const tspws_devframe_t *f, const float *d_x, int ntr, int long_row0, void *d_ypart,
                                     const double *d_x64, void *d_y64, unsigned int *d_tickets, const tspws_weight_t *d_rule,
                                     void *d_ps, void *d_st, void *d_zout, tspws_stream_t st)
{
    if (ntr <= 0 || !f->per_trace || !f->d_w64 || f->nitems2 <= 0 || f->nitems_gen > 0) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)st;
    const size_t smem = (size_t)f->fw_smem_per_warp * FW_WARPS * sizeof(float);
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(fwd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int cpt = (f->nitems2 + FW_WARPS - 1) / FW_WARPS;
    const int nshort_cta = (f->nitems_short + FW_WARPS - 1) / FW_WARPS;
    const int nf64 = nshort_cta + f->nlong;
    /* d_tickets != NULL: the phasor combine runs inside this launch (PS, ST, Zout are complete when it ends) */
    tspws_tailfuse_t fz;
    fz.tiles = (const int4 *)f->d_tile16; fz.expect = (const int2 *)f->d_tile_expect; fz.tickets = d_tickets; fz.rule = d_rule;
    fz.ps = (float2 *)d_ps; fz.st = (float2 *)d_st; fz.zout = (float2 *)d_zout; fz.ncoef = f->ncoef;
    int g5 = f->fw_lg_begin[5] / FW_WARPS;               /* first item group with D >= 32 items */
    { const char *e = getenv("TSPWS_TAIL_ORDER"); if (e && atoi(e) == 0) g5 = cpt; }   /* 0: fp64 CTAs first, then all groups */
    LAUNCH_DEP(fwd_tail_kernel, dim3(nf64 + ntr * cpt), dim3(FW_WARPS * 32), smem, s, f->d_scale, (const float2 *)f->d_g32, f->d_items2, f->nitems2,
        d_x, f->N, (float2 *)d_ypart, f->ypart_len, f->fw_smem_per_warp, ntr, long_row0 > 0 ? f->f64_maxl : 0x7fffffff, long_row0,
        (const double2 *)f->d_w64, f->d_items_short, f->nitems_short, nshort_cta, (const int2 *)f->d_long, nf64, d_x64,
        (double2 *)d_y64, f->ncoef, fz, g5);
    TSPWS_LAUNCHED_N("fwd_tail_kernel", s);
    return 0;
}

extern "C" int tspws_cu_forward_pertrace(const tspws_devframe_t *f, const float *d_x, int ntr, int long_row0, void *d_ypart,
                                         tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    if (!f->per_trace) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)st;
    if (f->nitems2 > 0) {
        const size_t smem = (size_t)f->fw_smem_per_warp * FW_WARPS * sizeof(float);
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(fwd2_pertrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        /* traces per CTA.  Measured at K = 10 traces x 16384 (profiles/): 1 trace per CTA 27 us, 2 -> 43 us,
         * 10 -> 90 us: the kernel is issue bound and wants the widest grid, so the default stays 1. */
        const int cpt = (f->nitems2 + FW_WARPS - 1) / FW_WARPS;
        static long env_chunk = -1;
        if (env_chunk < 0) { const char *e = getenv("TSPWS_PT_CHUNK"); env_chunk = e ? atol(e) : 1; }
        int tchunk = env_chunk > 0 ? (int)env_chunk : 1;
        if (tchunk > ntr) tchunk = ntr;
        dim3 grid((ntr + tchunk - 1) / tchunk, cpt);
        fwd2_pertrace_kernel<<<grid, FW_WARPS * 32, smem, s>>>(f->d_scale, (const float2 *)f->d_g32, f->d_items2,
            f->nitems2, d_x, f->N, (float2 *)d_ypart, f->ypart_len, f->fw_smem_per_warp, ntr, tchunk,
            long_row0 > 0 ? f->f64_maxl : 0x7fffffff, long_row0);
        TSPWS_LAUNCHED_N("fwd2_pertrace_kernel", s);
    }
    if (f->nitems_gen > 0) {
        dim3 grid((f->nitems_gen + FWD_WARPS - 1) / FWD_WARPS, ntr);
        fwd_pertrace_generic_kernel<float, false><<<grid, FWD_WARPS * 32, 0, s>>>(f->d_scale, (const float2 *)f->d_w32, f->d_items_gen,
            f->nitems_gen, d_x, f->N, (float2 *)d_ypart, f->ypart_len, long_row0 > 0 ? f->f64_maxl : 0x7fffffff, long_row0);
        TSPWS_LAUNCHED_N("fwd_pertrace_generic_kernel", s);
    }
    return 0;
}

/* per-trace coefficients -> PS, ST, Zout.
 * A CTA covers 64 consecutive coefficients of one scale; its 256 threads are 64 coefficients x 4
 * trace slots (slot j takes traces j, j+4, ...), so that the loads of the `seg` traces (and of the
 * phase splits of coarse scales) are in flight together instead of forming one latency chain per
 * coefficient; the 4 slots are then added in slot order through shared memory (deterministic).
 * grid = (64-coefficient tiles of all scales, segments): one wave at the headline size. */
#define PC_TILE 64
__global__ void __launch_bounds__(256)
phasor_combine_kernel(const int4 *__restrict__ tiles,
                      const float2 *__restrict__ ypart, int ypart_len,
                      int ncoef, int seg, const tspws_weight_t *__restrict__ rules, const double2 *__restrict__ ymean,
                      float2 *__restrict__ ps, float2 *__restrict__ st, float2 *__restrict__ zout)
{
    tl_scope tls(3);
    pdl_trigger();
    pdl_wait();
    __shared__ float  s_p[4][PC_TILE][2];
    __shared__ double s_s[4][PC_TILE][2];
    /* tile = (offset in a trace's partial-Y block, offset in the coefficient arrays, valid
     * coefficients | nsplit << 8, Ns = stride between phase splits) */
    const int4 tile = tiles[gridDim.x - 1 - blockIdx.x];   /* coarsest scales (most splits per coefficient) first */
    const int ci = threadIdx.x & (PC_TILE - 1), slot = threadIdx.x / PC_TILE;
    const int nvalid = tile.z & 255, nsplit = (tile.z >> 8) & 255, nsp64 = tile.z >> 16, Ns = tile.w;
    const int sg = blockIdx.y;
    float pr = 0.f, pi = 0.f;
    double sr = 0.0, si = 0.0;
    if (ci < nvalid) {
        const float2 *yp0 = ypart + (size_t)sg * seg * ypart_len + tile.x + ci;
        /* fp64 transform of the group mean (mean split, see group_post2_kernel): Y_k = Ymean + dY_k */
        double mr = 0.0, mi = 0.0;
        if (ymean) {
            for (int sp = 0; sp < nsp64; sp++) {            /* tap splits of fwd_f64_kernel, in slot order */
                const double2 ym = ymean[((size_t)sp * gridDim.y + sg) * ncoef + tile.y + ci];
                mr += ym.x; mi += ym.y;
            }
        }                                                   /* nsp64 == 0: unsplit transform, nothing to add */
        if (nsplit == 1) {
#pragma unroll 4
            for (int k = slot; k < seg; k += 4) {
                const float2 y = __ldcg(yp0 + (size_t)k * ypart_len);
                const double yr = (double)y.x + mr, yi = (double)y.y + mi;
                float a, b;
                phasor_fast((float)yr, (float)yi, a, b);
                pr += a; pi += b; sr += yr; si += yi;
            }
        } else {
            for (int k = slot; k < seg; k += 4) {
                const float2 *yp = yp0 + (size_t)k * ypart_len;
                float re = 0.f, im = 0.f;
#pragma unroll 8
                for (int sp = 0; sp < nsplit; sp++) { const float2 y = __ldcg(yp + (size_t)sp * Ns); re += y.x; im += y.y; }
                const double yr = (double)re + mr, yi = (double)im + mi;
                float a, b;
                phasor_fast((float)yr, (float)yi, a, b);
                pr += a; pi += b; sr += yr; si += yi;
            }
        }
    }
    s_p[slot][ci][0] = pr; s_p[slot][ci][1] = pi;
    s_s[slot][ci][0] = sr; s_s[slot][ci][1] = si;
    __syncthreads();
    if (slot != 0 || ci >= nvalid) return;
    pr = ((s_p[0][ci][0] + s_p[1][ci][0]) + s_p[2][ci][0]) + s_p[3][ci][0];
    pi = ((s_p[0][ci][1] + s_p[1][ci][1]) + s_p[2][ci][1]) + s_p[3][ci][1];
    sr = ((s_s[0][ci][0] + s_s[1][ci][0]) + s_s[2][ci][0]) + s_s[3][ci][0];
    si = ((s_s[0][ci][1] + s_s[1][ci][1]) + s_s[2][ci][1]) + s_s[3][ci][1];
    const double f = weight_factor(rules[sg], (double)pr, (double)pi);
    const size_t o = (size_t)sg * ncoef + tile.y + ci;
    ps[o]   = make_float2(pr, pi);
    st[o]   = make_float2((float)sr, (float)si);
    reinterpret_cast<float4 *>(zout)[o] = make_float4((float)(sr * f), (float)(si * f), (float)sr, (float)si);
}

/* ===================================================================================== */
/* chunk reduction + coherence weighting                                                 */
/*   segment sg sums chunks [sg*seg, (sg+1)*seg) -> PS, ST; Zout = ST * weight(|PS|)     */
/* ===================================================================================== */
__device__ __forceinline__ double weight_factor(const tspws_weight_t r, double pr, double pi)
{
    const double a2 = pr * pr + pi * pi;
    switch (r.mode) {
    case 0:  return a2 * r.inv_f;                                         /* nu = 2          */
    case 1:  return sqrt(a2) * r.inv_f;                                   /* nu = 1          */
    case 3: {                                                             /* unbiased, nu = 2 */
        const double c = (pr * r.inv_k) * (pr * r.inv_k) + (pi * r.inv_k) * (pi * r.inv_k);
        return (r.K * c - 1.0) * r.inv_f;
    }
    default: return pow(sqrt(a2) * r.inv_k, r.wu) * r.inv_f;              /* general nu      */
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
    if (sizeof(T) == 4) {
        reinterpret_cast<float4 *>(zout)[o] = make_float4((float)(sr * f), (float)(si * f), (float)sr, (float)si);
    } else {
        v.x = (T)(sr * f); v.y = (T)(si * f); zout[o] = v;
    }
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
    TSPWS_LAUNCHED_N("reduce_weight_kernel", s);
    return 0;
}

extern "C" int tspws_cu_weight(const void *d_ps, const void *d_st, int acc64, int ncoef, int nseg,
                               const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st)
{
    /* in-place re-reduction with one chunk per segment: PS/ST are rewritten with themselves */
    return tspws_cu_reduce_weight(d_ps, d_st, acc64, ncoef, nseg, 1, d_rules,
                                  (void *)d_ps, (void *)d_st, d_zout, st);
}

extern "C" int tspws_cu_phasor_combine(const tspws_devframe_t *f, const void *d_ypart, int nseg, int seg,
                                       const tspws_weight_t *d_rules, const void *d_ymean, void *d_ps, void *d_st, void *d_zout,
                                       tspws_stream_t st)
{
    if (nseg <= 0 || f->S <= 0) return 0;
    if (f->ntile16 <= 0) return 0;
    dim3 grid(f->ntile16, nseg);
    LAUNCH_DEP(phasor_combine_kernel, grid, dim3(256), 0, (cudaStream_t)st, (const int4 *)f->d_tile16, (const float2 *)d_ypart, f->ypart_len,
        f->ncoef, seg, d_rules, (const double2 *)d_ymean, (float2 *)d_ps, (float2 *)d_st, (float2 *)d_zout);
    TSPWS_LAUNCHED_N("phasor_combine_kernel", (cudaStream_t)st);
    return 0;
}

/* ===================================================================================== */
/* resampling support                                                                    */
/* ===================================================================================== */
/* Regrouped column sums for MB resamples at once (src/ts_pws1f_lib.c:643-654, :759-772):
 * gid[m][t] in {-1, 0..G-1} is the group trace t falls in for resample m (-1: not selected);
 * groups are visited in increasing order as t grows, so a thread only ever holds the CURRENT
 * group's sum per resample and stores it when the group changes -- no atomics, every trace is
 * read once per MB resamples.  out: [M][G][N] fp32 (pre-zeroed by the caller). */
#define RG_MB 8
template <bool VEC>
__global__ void __launch_bounds__(256)
regroup_kernel(const float *__restrict__ x, int N, int ntr, const short *__restrict__ gid, const signed char *__restrict__ cnt,
               int M, int G, float *__restrict__ out)
{
    const int m0 = blockIdx.y * RG_MB;
    const int nm = min(RG_MB, M - m0);
    const int col = VEC ? (blockIdx.x * 256 + threadIdx.x) * 4 : blockIdx.x * 256 + threadIdx.x;
    if (col >= N) return;
    double acc[RG_MB][VEC ? 4 : 1];
    int cur[RG_MB];
#pragma unroll
    for (int i = 0; i < RG_MB; i++) {
        cur[i] = -1;
#pragma unroll
        for (int c = 0; c < (VEC ? 4 : 1); c++) acc[i][c] = 0.0;
    }
    for (int t = 0; t < ntr; t++) {
        float v[VEC ? 4 : 1];
        if (VEC) { const float4 q = __ldg(reinterpret_cast<const float4 *>(x + (size_t)t * N + col)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else v[0] = __ldg(x + (size_t)t * N + col);
#pragma unroll
        for (int i = 0; i < RG_MB; i++) {
            if (i < nm) {
                const int g = gid[(size_t)(m0 + i) * ntr + t];
                if (g >= 0) {
                    if (g != cur[i]) {
                        if (cur[i] >= 0) {
                            float *o = out + ((size_t)(m0 + i) * G + cur[i]) * N + col;
#pragma unroll
                            for (int c = 0; c < (VEC ? 4 : 1); c++) { o[c] = (float)acc[i][c]; acc[i][c] = 0.0; }
                        }
                        cur[i] = g;
                    }
                    const double wgt = cnt ? (double)cnt[(size_t)(m0 + i) * ntr + t] : 1.0;
#pragma unroll
                    for (int c = 0; c < (VEC ? 4 : 1); c++) acc[i][c] += wgt * (double)v[c];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < RG_MB; i++) {
        if (i < nm && cur[i] >= 0) {
            float *o = out + ((size_t)(m0 + i) * G + cur[i]) * N + col;
#pragma unroll
            for (int c = 0; c < (VEC ? 4 : 1); c++) o[c] = (float)acc[i][c];
        }
    }
}

extern "C" int tspws_cu_regroup(const float *d_x, int N, int ntr, const short *d_gid, const signed char *d_cnt, int M, int G,
                                float *d_out, tspws_stream_t st)
{
    if (M <= 0 || ntr <= 0) return 0;
    cudaStream_t s = (cudaStream_t)st;
    if (cudaMemsetAsync(d_out, 0, (size_t)M * G * N * sizeof(float), s) != cudaSuccess) return (int)cudaGetLastError();
    const bool vec = (N % 4 == 0) && (((uintptr_t)d_x & 15) == 0) && (((uintptr_t)d_out & 15) == 0);
    if (vec) {
        dim3 grid((N / 4 + 255) / 256, (M + RG_MB - 1) / RG_MB);
        regroup_kernel<true><<<grid, 256, 0, s>>>(d_x, N, ntr, d_gid, d_cnt, M, G, d_out);
    } else {
        dim3 grid((N + 255) / 256, (M + RG_MB - 1) / RG_MB);
        regroup_kernel<false><<<grid, 256, 0, s>>>(d_x, N, ntr, d_gid, d_cnt, M, G, d_out);
    }
    TSPWS_LAUNCHED_N("regroup_kernel", s);
    return 0;
}

/* out[m][n] = (float)( scale[m] * sum_g (double) gs[m][g][n] )  -- the time-domain linear stack of a
 * resample (src/ts_pws1f_lib.c:680-691).  grid = (n tiles, M). */
__global__ void __launch_bounds__(256)
group_total_kernel(const float *__restrict__ gs, int G, int N, const double *__restrict__ scale, float *__restrict__ out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *p = gs + (size_t)blockIdx.y * G * N + n;
    double a = 0.0;
    for (int g = 0; g < G; g++) a += (double)p[(size_t)g * N];
    out[(size_t)blockIdx.y * N + n] = (float)(a * scale[blockIdx.y]);
}

extern "C" int tspws_cu_group_total(const float *d_gs, int M, int G, int N, const double *d_scale, float *d_out,
                                    tspws_stream_t st)
{
    if (M <= 0) return 0;
    dim3 grid((N + 255) / 256, M);
    group_total_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(d_gs, G, N, d_scale, d_out);
    TSPWS_LAUNCHED_N("group_total_kernel", (cudaStream_t)st);
    return 0;
}

/* Per-trace coefficient matrices from the per-trace forward output: Y[t][i] = sum of phase splits,
 * P[t][i] = unit phasor of Y (zero when skipped).  grid = (tiles over Ns, scales, traces). */
__global__ void __launch_bounds__(256)
coef_matrix_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ ypart, int ypart_len,
                   int ncoef, float2 *__restrict__ Y, float2 *__restrict__ P)
{
    const tspws_scale_t sc = scales[blockIdx.y];
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= sc.Ns) return;
    const float2 *yp = ypart + (size_t)blockIdx.z * ypart_len + sc.ypart_off + m;
    float re = 0.f, im = 0.f;
    for (int sp = 0; sp < sc.nsplit; sp++) { const float2 y = yp[(size_t)sp * sc.Ns]; re += y.x; im += y.y; }
    const size_t o = (size_t)blockIdx.z * ncoef + sc.coef_off + m;
    if (Y) Y[o] = make_float2(re, im);
    if (P) { float a, b; phasor_fast(re, im, a, b); P[o] = make_float2(a, b); }
}

extern "C" int tspws_cu_coef_matrix(const tspws_devframe_t *f, const void *d_ypart, int ntr, void *d_Y, void *d_P,
                                    tspws_stream_t st)
{
    if (ntr <= 0 || f->S <= 0) return 0;
    dim3 grid((f->max_Ns + 255) / 256, f->S, ntr);
    coef_matrix_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(f->d_scale, (const float2 *)d_ypart, f->ypart_len, f->ncoef,
                                                           (float2 *)d_Y, (float2 *)d_P);
    TSPWS_LAUNCHED_N("coef_matrix_kernel", (cudaStream_t)st);
    return 0;
}

/* Selection contraction (src/ts_pws1f_lib.c:555-563): acc[m][i] += sum_t sel[m][t0+t] * P[t][i]
 * for SS_MB resamples per thread; P is read once per SS_MB resamples.  acc: [M][ncoef] double2. */
#define SS_MB 16
__global__ void __launch_bounds__(256)
select_sum_kernel(const float2 *__restrict__ P, int T, int ncoef, const signed char *__restrict__ sel, int sel_ld,
                  int t0, int M, double2 *__restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoef) return;
    const int m0 = blockIdx.y * SS_MB;
    const int nm = min(SS_MB, M - m0);
    double ar[SS_MB], ai[SS_MB];
#pragma unroll
    for (int j = 0; j < SS_MB; j++) { ar[j] = 0.0; ai[j] = 0.0; }
    for (int t = 0; t < T; t++) {
        const float2 p = __ldg(P + (size_t)t * ncoef + i);
#pragma unroll
        for (int j = 0; j < SS_MB; j++)
            if (j < nm && sel[(size_t)(m0 + j) * sel_ld + t0 + t] == 1) { ar[j] += (double)p.x; ai[j] += (double)p.y; }
    }
#pragma unroll
    for (int j = 0; j < SS_MB; j++)
        if (j < nm) {
            double2 *o = acc + (size_t)(m0 + j) * ncoef + i;
            double2 v = *o;
            v.x += ar[j]; v.y += ai[j];
            *o = v;
        }
}

extern "C" int tspws_cu_select_sum(const void *d_P, int T, int ncoef, const signed char *d_sel, int sel_ld, int t0,
                                   int M, void *d_acc, tspws_stream_t st)
{
    if (T <= 0 || M <= 0) return 0;
    dim3 grid((ncoef + 255) / 256, (M + SS_MB - 1) / SS_MB);
    select_sum_kernel<<<grid, 256, 0, (cudaStream_t)st>>>((const float2 *)d_P, T, ncoef, d_sel, sel_ld, t0, M, (double2 *)d_acc);
    TSPWS_LAUNCHED_N("select_sum_kernel", (cudaStream_t)st);
    return 0;
}

/* Zout[m] = ST[m] * weight(PS[m]) with PS in double2 and ST in float2 (single-stage resampling). */
__global__ void __launch_bounds__(256)
weight_mixed_kernel(const double2 *__restrict__ ps, const float2 *__restrict__ st, int ncoef,
                    const tspws_weight_t *__restrict__ rules, float2 *__restrict__ zout)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoef) return;
    const size_t o = (size_t)blockIdx.y * ncoef + i;
    const double2 p = ps[o];
    const float2 s = st[o];
    const double f = weight_factor(rules[blockIdx.y], p.x, p.y);
    reinterpret_cast<float4 *>(zout)[o] = make_float4((float)((double)s.x * f), (float)((double)s.y * f), s.x, s.y);
}

__global__ void __launch_bounds__(256)
weight_mixed32_kernel(const float2 *__restrict__ ps, const float2 *__restrict__ st, int ncoef,
                      const tspws_weight_t *__restrict__ rules, float2 *__restrict__ zout)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoef) return;
    const size_t o = (size_t)blockIdx.y * ncoef + i;
    const float2 p = ps[o];
    const float2 s = st[o];
    const double f = weight_factor(rules[blockIdx.y], (double)p.x, (double)p.y);
    reinterpret_cast<float4 *>(zout)[o] = make_float4((float)((double)s.x * f), (float)((double)s.y * f), s.x, s.y);
}

extern "C" int tspws_cu_weight_mixed32(const void *d_ps32, const void *d_st32, int ncoef, int nseg,
                                       const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st)
{
    if (nseg <= 0) return 0;
    dim3 grid((ncoef + 255) / 256, nseg);
    weight_mixed32_kernel<<<grid, 256, 0, (cudaStream_t)st>>>((const float2 *)d_ps32, (const float2 *)d_st32, ncoef, d_rules,
                                                              (float2 *)d_zout);
    TSPWS_LAUNCHED_N("weight_mixed32_kernel", (cudaStream_t)st);
    return 0;
}

extern "C" int tspws_cu_weight_mixed(const void *d_ps64, const void *d_st32, int ncoef, int nseg,
                                     const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st)
{
    if (nseg <= 0) return 0;
    dim3 grid((ncoef + 255) / 256, nseg);
    weight_mixed_kernel<<<grid, 256, 0, (cudaStream_t)st>>>((const double2 *)d_ps64, (const float2 *)d_st32, ncoef, d_rules,
                                                            (float2 *)d_zout);
    TSPWS_LAUNCHED_N("weight_mixed_kernel", (cudaStream_t)st);
    return 0;
}

/* Convergence (src/ts_pws1f_lib.c:259-286, single-stage branch): running sums over traces
 * t0..t0+B-1 of Y and P on top of the state (st, ps), one weighted coefficient set per prefix. */
__global__ void __launch_bounds__(256)
prefix_weight_kernel(const float2 *__restrict__ Y, const float2 *__restrict__ P, int ncoef, int B,
                     double2 *__restrict__ st_state, double2 *__restrict__ ps_state,
                     const tspws_weight_t *__restrict__ rules, float2 *__restrict__ zout)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncoef) return;
    double2 s = st_state[i], p = ps_state[i];
    for (int b = 0; b < B; b++) {
        const float2 y = Y[(size_t)b * ncoef + i], q = P[(size_t)b * ncoef + i];
        s.x += (double)y.x; s.y += (double)y.y;
        p.x += (double)q.x; p.y += (double)q.y;
        const double f = weight_factor(rules[b], p.x, p.y);
        reinterpret_cast<float4 *>(zout)[(size_t)b * ncoef + i] = make_float4((float)(s.x * f), (float)(s.y * f), (float)s.x, (float)s.y);
    }
    st_state[i] = s; ps_state[i] = p;
}

extern "C" int tspws_cu_prefix_weight(const void *d_Y, const void *d_P, int ncoef, int B, void *d_st_state,
                                      void *d_ps_state, const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st)
{
    if (B <= 0) return 0;
    prefix_weight_kernel<<<(ncoef + 255) / 256, 256, 0, (cudaStream_t)st>>>((const float2 *)d_Y, (const float2 *)d_P, ncoef, B,
        (double2 *)d_st_state, (double2 *)d_ps_state, d_rules, (float2 *)d_zout);
    TSPWS_LAUNCHED_N("prefix_weight_kernel", (cudaStream_t)st);
    return 0;
}

/* similarity (GNCC) and misfit of B candidate traces against a reference trace
 * (src/ts_pws1f_lib.c:433-462): one CTA per candidate; sums in double. */
__global__ void __launch_bounds__(256)
simil_kernel(const float *__restrict__ cand, int N, const float *__restrict__ ref, double *__restrict__ sim,
             double *__restrict__ mis)
{
    __shared__ double sh[33];
    const float *c = cand + (size_t)blockIdx.x * N;
    double xy = 0, xx = 0, yy = 0, dd = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double a = (double)c[n], b = (double)ref[n];
        xy += a * b; xx += a * a; yy += b * b; dd += (a - b) * (a - b);
    }
    xy = block_sum(xy, sh); xx = block_sum(xx, sh); yy = block_sum(yy, sh); dd = block_sum(dd, sh);
    if (threadIdx.x == 0) { sim[blockIdx.x] = xy / sqrt(xx) / sqrt(yy); mis[blockIdx.x] = dd; }
}

extern "C" int tspws_cu_simil(const float *d_cand, int B, int N, const float *d_ref, double *d_sim, double *d_mis,
                              tspws_stream_t st)
{
    if (B <= 0) return 0;
    simil_kernel<<<B, 256, 0, (cudaStream_t)st>>>(d_cand, N, d_ref, d_sim, d_mis);
    TSPWS_LAUNCHED_N("simil_kernel", (cudaStream_t)st);
    return 0;
}

/* Linear-stack convergence (src/ts_pws1f_lib.c:288-308): running time-domain mean after every trace,
 * float-cast, its similarity / misfit against `ref`, and optionally the traces themselves.
 * Each thread owns samples and walks the traces; per-trace dot products are reduced per CTA and
 * finished by simil_finish_kernel.  part: [ntr][gridDim.x][4] doubles. */
__global__ void __launch_bounds__(256)
lsconv_kernel(const float *__restrict__ x, int ntr, int N, const float *__restrict__ ref,
              float *__restrict__ steps, double *__restrict__ part)
{
    __shared__ double sh[33];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = n < N;
    double run = 0.0;
    const double r = ok ? (double)ref[n] : 0.0;
    for (int t = 0; t < ntr; t++) {
        if (ok) run += (double)x[(size_t)t * N + n];
        const float fa1 = (float)(1.0 / (double)(t + 1));            /* fa1 = 1.0/(itr+1) is a float in the reference */
        const double cur = ok ? run * (double)fa1 : 0.0;             /* dsig[n] *= fa1 (double) */
        if (steps && ok) steps[(size_t)t * N + n] = (float)cur;
        /* the reference keeps dsig scaled by fa1*(itr+1) afterwards (:296,:307): reproduce that drift */
        if (ok) run = cur * (double)(t + 1);
        const double xy = block_sum(cur * r, sh), xx = block_sum(cur * cur, sh);
        const double yy = block_sum(r * r, sh), dd = block_sum((cur - r) * (cur - r), sh);
        if (threadIdx.x == 0) {
            double *o = part + ((size_t)t * gridDim.x + blockIdx.x) * 4;
            o[0] = xy; o[1] = xx; o[2] = yy; o[3] = dd;
        }
    }
}
__global__ void simil_finish_kernel(const double *__restrict__ part, int nblk, int ntr, double *__restrict__ sim,
                                    double *__restrict__ mis)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntr) return;
    double xy = 0, xx = 0, yy = 0, dd = 0;
    for (int b = 0; b < nblk; b++) {
        const double *o = part + ((size_t)t * nblk + b) * 4;
        xy += o[0]; xx += o[1]; yy += o[2]; dd += o[3];
    }
    sim[t] = xy / sqrt(xx) / sqrt(yy);
    mis[t] = dd;
}

extern "C" int tspws_cu_lsconv(const float *d_x, int ntr, int N, const float *d_ref, float *d_steps, double *d_part,
                               double *d_sim, double *d_mis, tspws_stream_t st)
{
    if (ntr <= 0) return 0;
    const int nblk = (N + 255) / 256;
    lsconv_kernel<<<nblk, 256, 0, (cudaStream_t)st>>>(d_x, ntr, N, d_ref, d_steps, d_part);
    TSPWS_LAUNCHED_N("lsconv_kernel", (cudaStream_t)st);
    simil_finish_kernel<<<(ntr + 255) / 256, 256, 0, (cudaStream_t)st>>>(d_part, nblk, ntr, d_sim, d_mis);
    TSPWS_LAUNCHED_N("simil_finish_kernel", (cudaStream_t)st);
    return 0;
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
           T *scratch, int nseg, unsigned int *tickets, const tspws_weight_t *__restrict__ rules,
           float *__restrict__ out_ts, float *__restrict__ out_ls)
{
    typedef typename cplx<T>::type T2;
    __shared__ int is_last;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const tspws_inv_group_t grp = groups[blockIdx.y];
    const int sg = blockIdx.z;
    const T2 *__restrict__ Z = zout + (size_t)sg * ncoef;
    const T2 *__restrict__ S = WITH_ST ? st + (size_t)sg * ncoef : nullptr;
    T accz = 0, accs = 0;
    if (n < N) {
        for (int s = grp.s0; s < grp.s1; s++) {
            const tspws_scale_t sc = scales[s];
            const T2 *__restrict__ w  = wall + sc.tap_off;
            const T2 *__restrict__ Zs = Z + sc.coef_off;
            const T2 *__restrict__ Ss = WITH_ST ? S + sc.coef_off : nullptr;
            T az = 0, as = 0;
            const int top = (sc.Ns - 1) * sc.D;                     /* position of the last coefficient */
#pragma unroll 1
            for (int j = -1; j <= 1; j++) {
                const int base = n + sc.c + j * N;                  /* l = base - m D; |base| < 3N fits an int */
                const int lnum = base - (sc.L - 1);
                if (base < 0 || lnum > top) continue;               /* no coefficient reaches n in this image */
                int hi = (int)((unsigned int)base / (unsigned int)sc.D);
                if (hi > sc.Ns - 1) hi = sc.Ns - 1;
                const int lo = lnum <= 0 ? 0 : (int)(((unsigned int)lnum + sc.D - 1) / (unsigned int)sc.D);
                const T2 *wp = w + (base - lo * sc.D);              /* walks down by D taps per coefficient */
                const T2 *zp = Zs + lo;
                const ptrdiff_t sd = WITH_ST ? (Ss - Zs) : 0;
#pragma unroll 4
                for (int cnt = hi - lo + 1; cnt > 0; cnt--, wp -= sc.D, zp++) {
                    const T2 wv = *wp;
                    const T2 z  = *zp;
                    az = fma(z.x, wv.x, az);
                    az = fma(-z.y, wv.y, az);
                    if (WITH_ST) {
                        const T2 q = zp[sd];
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
    /* the last scale group to finish this tile of samples adds the partials in group order and
     * casts like src/ts_pws1f_lib.c:233-241: tsPWS = (float)x, ls = (float)x / mtr (float division) */
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *tk = &tickets[sg * gridDim.x + blockIdx.x];
        const unsigned int c = atomicAdd(tk, 1u);
        is_last = (c == gridDim.y - 1);
        if (is_last) *tk = 0;
    }
    __syncthreads();
    if (!is_last || n >= N) return;
    __threadfence();
    double az = 0, as = 0;
    for (int g = 0; g < ninv; g++) {
        const size_t o = ((size_t)g * nseg + sg) * N + n;
        az += (double)__ldcg(scratch + o);
        if (WITH_ST) as += (double)__ldcg(scratch + (size_t)ninv * nseg * N + o);
    }
    out_ts[(size_t)sg * N + n] = (float)az;
    if (WITH_ST) out_ls[(size_t)sg * N + n] = __fdiv_rn((float)as, (float)(unsigned int)rules[sg].T);
}

/* fp32 inverse, 4 consecutive samples per thread.  For scales with D % 4 == 0 the four samples
 * n0..n0+3 (n0 % 4 == 0) see the same coefficients m and four consecutive taps of the origin-aligned
 * filter: per coefficient one 16-byte load of (Zout, ST) and two of the taps feed 16 FMAs.  Other
 * scales (D = 1, 2, odd) take the one-sample path.  grid = (tiles of 1024 samples, scale groups,
 * segments); the last scale group to finish a tile adds the partials in group order and casts. */
#define INV4_THREADS 64
template <bool WITH_ST, bool FINISH>
__global__ void __launch_bounds__(INV4_THREADS)
inv4_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ wall, const float2 *__restrict__ wi,
            const tspws_inv_group_t *__restrict__ groups, int ninv, int N, int ncoef,
            const float4 *__restrict__ zs, float *scratch, int nseg, unsigned int *tickets,
            const tspws_weight_t *__restrict__ rules, float *__restrict__ out_ts, float *__restrict__ out_ls)
{
    tl_scope tls(5);
    __shared__ int is_last;
    const int n0 = (blockIdx.x * INV4_THREADS + threadIdx.x) * 4;
    const tspws_inv_group_t grp = groups[blockIdx.y];
    const int sg = blockIdx.z;
    const float4 *__restrict__ ZS = zs + (size_t)sg * ncoef;
    float az[4] = {0.f, 0.f, 0.f, 0.f}, as[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 < N) {
        for (int s = grp.s0; s < grp.s1; s++) {
            const tspws_scale_t sc = scales[s];
            const float4 *__restrict__ Zs = ZS + sc.coef_off;
            const int top = (sc.Ns - 1) * sc.D;
            float bz[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
            if (sc.inv4) {
                const float2 *__restrict__ w = wi + sc.wi_off;
#pragma unroll 1
                for (int j = -1; j <= 1; j++) {
                    const int base = n0 + sc.ci + j * N;               /* tap of sample n0 = base - m D, a multiple of 4 */
                    const int lnum = base - (sc.Li4 - 4);
                    if (base < 0 || lnum > top) continue;
                    int hi = (int)((unsigned int)base / (unsigned int)sc.D);
                    if (hi > sc.Ns - 1) hi = sc.Ns - 1;
                    const int lo = lnum <= 0 ? 0 : (int)(((unsigned int)lnum + sc.D - 1) / (unsigned int)sc.D);
                    const float2 *wp = w + (base - lo * sc.D);
                    const float4 *zp = Zs + lo;
#pragma unroll 8
                    for (int cnt = hi - lo + 1; cnt > 0; cnt--, wp -= sc.D, zp++) {
                        const float4 w01 = __ldg(reinterpret_cast<const float4 *>(wp));
                        const float4 w23 = __ldg(reinterpret_cast<const float4 *>(wp + 2));
                        const float4 z   = __ldg(zp);
                        bz[0] = fmaf(z.x, w01.x, bz[0]); bz[0] = fmaf(-z.y, w01.y, bz[0]);
                        bz[1] = fmaf(z.x, w01.z, bz[1]); bz[1] = fmaf(-z.y, w01.w, bz[1]);
                        bz[2] = fmaf(z.x, w23.x, bz[2]); bz[2] = fmaf(-z.y, w23.y, bz[2]);
                        bz[3] = fmaf(z.x, w23.z, bz[3]); bz[3] = fmaf(-z.y, w23.w, bz[3]);
                        if (WITH_ST) {
                            bs[0] = fmaf(z.z, w01.x, bs[0]); bs[0] = fmaf(-z.w, w01.y, bs[0]);
                            bs[1] = fmaf(z.z, w01.z, bs[1]); bs[1] = fmaf(-z.w, w01.w, bs[1]);
                            bs[2] = fmaf(z.z, w23.x, bs[2]); bs[2] = fmaf(-z.w, w23.y, bs[2]);
                            bs[3] = fmaf(z.z, w23.z, bs[3]); bs[3] = fmaf(-z.w, w23.w, bs[3]);
                        }
                    }
                }
            } else {
                const float2 *__restrict__ w = wall + sc.tap_off;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (n0 + i >= N) break;
#pragma unroll 1
                    for (int j = -1; j <= 1; j++) {
                        const int base = n0 + i + sc.c + j * N;
                        const int lnum = base - (sc.L - 1);
                        if (base < 0 || lnum > top) continue;
                        int hi = (int)((unsigned int)base / (unsigned int)sc.D);
                        if (hi > sc.Ns - 1) hi = sc.Ns - 1;
                        const int lo = lnum <= 0 ? 0 : (int)(((unsigned int)lnum + sc.D - 1) / (unsigned int)sc.D);
                        const float2 *wp = w + (base - lo * sc.D);
                        const float4 *zp = Zs + lo;
#pragma unroll 4
                        for (int cnt = hi - lo + 1; cnt > 0; cnt--, wp -= sc.D, zp++) {
                            const float2 wv = __ldg(wp);
                            const float4 z  = __ldg(zp);
                            bz[i] = fmaf(z.x, wv.x, bz[i]); bz[i] = fmaf(-z.y, wv.y, bz[i]);
                            if (WITH_ST) { bs[i] = fmaf(z.z, wv.x, bs[i]); bs[i] = fmaf(-z.w, wv.y, bs[i]); }
                        }
                    }
                }
            }
            const float gn = (float)sc.gain;
#pragma unroll
            for (int i = 0; i < 4; i++) { az[i] = fmaf(gn, bz[i], az[i]); as[i] = fmaf(gn, bs[i], as[i]); }
        }
        /* slot = scale group (FINISH: summed below by the last group) or the scale itself */
        const size_t o = ((size_t)(FINISH ? blockIdx.y : grp.s0) * nseg + sg) * N + n0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (n0 + i < N) {
                scratch[o + i] = az[i];
                if (WITH_ST) scratch[(size_t)ninv * nseg * N + o + i] = as[i];
            }
    }
    if (!FINISH) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *tk = &tickets[sg * gridDim.x + blockIdx.x];
        const unsigned int c = atomicAdd(tk, 1u);
        is_last = (c == gridDim.y - 1);
        if (is_last) *tk = 0;
    }
    __syncthreads();
    if (!is_last || n0 >= N) return;
    __threadfence();
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if ((N & 3) == 0) {
        /* independent 16-byte loads, 8 groups in flight at a time (the adds stay in group order) */
#pragma unroll 8
        for (int g = 0; g < ninv; g++) {
            const size_t o = ((size_t)g * nseg + sg) * N + n0;
            const float4 p = __ldcg(reinterpret_cast<const float4 *>(scratch + o));
            a[0] += (double)p.x; a[1] += (double)p.y; a[2] += (double)p.z; a[3] += (double)p.w;
            if (WITH_ST) {
                const float4 q = __ldcg(reinterpret_cast<const float4 *>(scratch + (size_t)ninv * nseg * N + o));
                b[0] += (double)q.x; b[1] += (double)q.y; b[2] += (double)q.z; b[3] += (double)q.w;
            }
        }
    } else {
        for (int g = 0; g < ninv; g++) {
            const size_t o = ((size_t)g * nseg + sg) * N + n0;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (n0 + i < N) {
                    a[i] += (double)__ldcg(scratch + o + i);
                    if (WITH_ST) b[i] += (double)__ldcg(scratch + (size_t)ninv * nseg * N + o + i);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int n = n0 + i;
        if (n >= N) break;
        out_ts[(size_t)sg * N + n] = (float)a[i];
        if (WITH_ST) out_ls[(size_t)sg * N + n] = __fdiv_rn((float)b[i], (float)(unsigned int)rules[sg].T);
    }
}

/* Polyphase register-tiled inverse for D >= 32 (N % D == 0): the transpose of the forward tile.
 * A lane owns one phase rho and 8 consecutive coefficient positions a; it slides a 16-slot register
 * ring of (Zout, ST) coefficients -- warp-uniform 16-byte loads -- past the Qi taps of its phase
 * (coalesced 8-byte loads, each used for 8 samples x 2 outputs from registers), so the filter is
 * read from L2 once per 8 samples instead of once per sample.  Output sample n = (a D + rho - c') mod N
 * of scale s goes to scratch slot s; inv_finish_kernel adds the slots.
 * One warp per item (scale, a-tile, block of 32 phases). */
#define IV_R 8
#define IV_WS 16
/* The item's Qi tap rows (32 phases each) and its Qi + 7 coefficients are staged into warp-private shared memory
 * with ONE round of cp.async before the tap loop: with loads from global inside the loop, an 8-tap prefetch
 * distance (8 x 16 FFMA2) covers only a third of the L2 latency and every tap stalled (per-CTA timeline: 6 us per
 * item of ~1500 instructions). */
template <bool WITH_ST>
__global__ void __launch_bounds__(128)
inv2_kernel(const tspws_scale_t *__restrict__ scales, const float2 *__restrict__ gi_all, const int4 *__restrict__ items,
            int nitems, int N, int ncoef, int nslots, const float4 *__restrict__ zs, float *__restrict__ scratch, int nseg,
            int smem_per_warp /* bytes */, int qmax)
{
    tl_scope tls(4);
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char iv_smem[];
    const int widx = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (widx >= nitems) return;
    const int lane = threadIdx.x & 31;
    const int4 it = items[widx];
    const tspws_scale_t sc = scales[it.x];
    const int sg = blockIdx.y;
    const int D = sc.D, Qi = sc.Qi, Ns = sc.Ns, a0 = it.y;
    const int rho = it.z * 32 + lane;
    const bool act = rho < D;
    const float4 *__restrict__ Z = zs + (size_t)sg * ncoef + sc.coef_off;
    float2 *gsm = reinterpret_cast<float2 *>(iv_smem + (size_t)(threadIdx.x >> 5) * smem_per_warp);   /* [Qi + 8][32] */
    float4 *zsm = reinterpret_cast<float4 *>(gsm + (size_t)(qmax + 8) * 32);                            /* [Qi + 7]     */
    {
        const float2 *__restrict__ g0 = gi_all + sc.gi_off + it.z * 32;
        const int nrow = Qi + 8;                          /* the tap prefetch runs 8 rows ahead (zero rows in the array) */
        if (((D | sc.gi_off) & 1) == 0 && it.z * 32 + 32 <= D) {
            for (int c = lane; c < nrow * 16; c += 32) {  /* 16-byte chunks: row c / 16, float2 pair c % 16 */
                const int row = c >> 4, col = (c & 15) * 2;
                cp_async<16>(gsm + row * 32 + col, g0 + (size_t)row * D + col);
            }
        } else {
            for (int row = 0; row < nrow; row++) cp_async<8>(gsm + row * 32 + lane, g0 + (size_t)row * D + (act ? lane : 0));
        }
        int zi0 = (a0 - Qi + 1) % Ns;                     /* window slot 0 */
        if (zi0 < 0) zi0 += Ns;
        for (int j = lane; j < Qi + IV_R - 1; j += 32) cp_async<16>(zsm + j, Z + (zi0 + j) % Ns);
        cp_async_commit();
    }
    float2 accr[IV_R];                                    /* (tsPWS, ls) partial of sample r */
#pragma unroll
    for (int r = 0; r < IV_R; r++) accr[r] = make_float2(0.f, 0.f);
    cp_async_wait<0>();
    __syncwarp();
    const float2 *g = gsm + lane;
    int zi = 0;
    float2 wre[IV_WS], wim[IV_WS];                        /* (Zr, Sr) and (Zi, Si) of every ring slot */
    float2 gq[8];
#pragma unroll
    for (int j = 0; j < IV_WS; j++) {
        if (j < Qi + IV_R - 1) {
            const float4 z = zsm[zi++];
            wre[j] = make_float2(z.x, z.z); wim[j] = make_float2(z.y, z.w);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) gq[q] = g[q * 32];
    for (int t0 = 0; t0 < Qi; t0 += IV_WS) {
#pragma unroll
        for (int tt = 0; tt < IV_WS; tt++) {
            if (t0 + tt < Qi) {
                const float2 gv = gq[tt & 7];
                const float2 gx = make_float2(gv.x, gv.x), gy = make_float2(-gv.y, -gv.y);
#pragma unroll
                for (int r = 0; r < IV_R; r++) {
                    accr[r] = __ffma2_rn(wre[(tt + r) % IV_WS], gx, accr[r]);
                    accr[r] = __ffma2_rn(wim[(tt + r) % IV_WS], gy, accr[r]);
                }
                if (t0 + tt + IV_WS - IV_R + 1 < Qi) {
                    const float4 z = zsm[zi++];
                    wre[tt] = make_float2(z.x, z.z); wim[tt] = make_float2(z.y, z.w);
                }
                gq[tt & 7] = g[(t0 + tt + 8) * 32];
            }
        }
    }
    if (!act) return;
    const float gn = (float)sc.gain;
    long long nn = ((long long)a0 * D + rho - sc.ci) % N;
    if (nn < 0) nn += N;
    int n = (int)nn;
    float *o = scratch + ((size_t)it.x * nseg + sg) * N;
#pragma unroll
    for (int r = 0; r < IV_R; r++) {
        if (a0 + r < Ns) {
            o[n] = gn * accr[r].x;
            if (WITH_ST) o[(size_t)nslots * nseg * N + n] = gn * accr[r].y;
        }
        n += D;
        if (n >= N) n -= N;
    }
}

/* out[n] = sum over scratch slots; casts like src/ts_pws1f_lib.c:233-241.
 * A CTA covers 64 groups of 4 samples; its 256 threads are 64 sample groups x 4 slot quarters, each
 * quarter adds its slots (g = quarter, quarter + 4, ...) with independent 16-byte loads, then the four
 * quarter sums are added in quarter order through shared memory (deterministic).
 * grid = (tiles of 256 samples, segments). */
template <bool WITH_ST>
__global__ void __launch_bounds__(256)
inv_finish_kernel(const float *__restrict__ scratch, int nslots, int nseg, int N, const tspws_weight_t *__restrict__ rules,
                  float *__restrict__ out_ts, float *__restrict__ out_ls)
{
    tl_scope tls(6);
    pdl_wait();
    /* 32 float4 columns x 8 slot lanes: thread (col, sl) adds slots sl, sl+8, ... of its 4 samples in
     * fp64, then the 8 lanes of a column are added in lane order -- 12 loads per thread at 44 scales */
    __shared__ double sh[8][8][33];
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * 32 + col) * 4;
    const int sg = blockIdx.y;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (n0 < N) {
        if ((N & 3) == 0) {
#pragma unroll 3
            for (int g = sl; g < nslots; g += 8) {
                const size_t o = ((size_t)g * nseg + sg) * N + n0;
                const float4 p = __ldcg(reinterpret_cast<const float4 *>(scratch + o));
                a[0] += (double)p.x; a[1] += (double)p.y; a[2] += (double)p.z; a[3] += (double)p.w;
                if (WITH_ST) {
                    const float4 q = __ldcg(reinterpret_cast<const float4 *>(scratch + (size_t)nslots * nseg * N + o));
                    b[0] += (double)q.x; b[1] += (double)q.y; b[2] += (double)q.z; b[3] += (double)q.w;
                }
            }
        } else {
            for (int g = sl; g < nslots; g += 8) {
                const size_t o = ((size_t)g * nseg + sg) * N + n0;
                for (int i = 0; i < 4 && n0 + i < N; i++) {
                    a[i] += (double)scratch[o + i];
                    if (WITH_ST) b[i] += (double)scratch[(size_t)nslots * nseg * N + o + i];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { sh[sl][i][col] = a[i]; sh[sl][4 + i][col] = b[i]; }
    __syncthreads();
    /* thread (col, i = sl): value i of the column (i < 4: tsPWS sample n0 + i, else ls sample n0 + i - 4) */
    const int i = sl;
    if (!WITH_ST && i >= 4) return;
    const int n = n0 + (i & 3);
    if (n >= N) return;
    double t = 0.0;
#pragma unroll
    for (int l = 0; l < 8; l++) t += sh[l][i][col];
    if (i < 4) out_ts[(size_t)sg * N + n] = (float)t;
    else       out_ls[(size_t)sg * N + n] = __fdiv_rn((float)t, (float)(unsigned int)rules[sg].T);
}

template <typename T>
static int inv_launch(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int nseg,
                      const tspws_weight_t *d_rules, void *d_scratch, unsigned int *d_tickets,
                      float *d_ts, float *d_ls, cudaStream_t s)
{
    typedef typename cplx<T>::type T2;
    const T2 *w = (const T2 *)(sizeof(T) == 4 ? f->d_w32 : f->d_w64);
    dim3 grid((f->N + 255) / 256, f->ninv, nseg);
    if (d_st)
        inv_kernel<T, true><<<grid, 256, 0, s>>>(f->d_scale, w, f->d_invgrp, f->ninv, f->N, f->ncoef,
            (const T2 *)d_zout, (const T2 *)d_st, (T *)d_scratch, nseg, d_tickets, d_rules, d_ts, d_ls);
    else
        inv_kernel<T, false><<<grid, 256, 0, s>>>(f->d_scale, w, f->d_invgrp, f->ninv, f->N, f->ncoef,
            (const T2 *)d_zout, nullptr, (T *)d_scratch, nseg, d_tickets, d_rules, d_ts, d_ls);
    TSPWS_LAUNCHED_N("inv_kernel", s);
    return 0;
}

extern "C" int tspws_cu_inverse(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int acc64,
                                int nseg, const tspws_weight_t *d_rules, void *d_scratch, unsigned int *d_tickets,
                                float *d_out_ts, float *d_out_ls, tspws_stream_t st, tspws_stream_t side, void *ev_fork, void *ev_join)
{
    if (nseg <= 0) return 0;
    if (acc64 && !f->d_w64) return (int)cudaErrorInvalidValue;
    cudaStream_t g_inv_side = (cudaStream_t)side;
    cudaEvent_t  g_inv_fork = (cudaEvent_t)ev_fork, g_inv_join = (cudaEvent_t)ev_join;
    if (acc64) return inv_launch<double>(f, d_zout, d_st, nseg, d_rules, d_scratch, d_tickets, d_out_ts, d_out_ls, (cudaStream_t)st);
    /* fp32: d_zout is the interleaved (Zout, ST) float4 array; d_st only says whether ls is wanted.
     * Every scale writes its contribution to its own scratch slot (polyphase kernel for D >= 32,
     * 4-sample kernel for the rest); the finish kernel adds the slots in scale order. */
    cudaStream_t s = (cudaStream_t)st;
    const int S = f->S;
    const bool fork = g_inv_side && g_inv_fork && g_inv_join && !g_prof.on && f->nitems_inv2 > 0 && f->ninv4 > 0 && g_inv_side != s;
    cudaStream_t s4 = fork ? g_inv_side : s;
    if (fork) {
        if (cudaEventRecord(g_inv_fork, s) != cudaSuccess || cudaStreamWaitEvent(s4, g_inv_fork, 0) != cudaSuccess) return (int)cudaGetLastError();
    }
    if (f->nitems_inv2 > 0) {
        dim3 grid((f->nitems_inv2 + 3) / 4, nseg);
        /* per warp: (Qi + 8) tap rows of 32 phases + Qi + 7 coefficients (float4) */
        const int spw = ((f->inv2_qmax + 8) * 256 + (f->inv2_qmax + 8) * 16 + 15) & ~15;
        if (4 * spw > 40 * 1024) {
            cudaFuncSetAttribute(inv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * spw);
            cudaFuncSetAttribute(inv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * spw);
        }
        if (d_st) LAUNCH_DEP(inv2_kernel<true>, grid, dim3(128), (size_t)(4 * spw), s, f->d_scale, (const float2 *)f->d_gi32, (const int4 *)f->d_items_inv2, f->nitems_inv2,
                      f->N, f->ncoef, S, (const float4 *)d_zout, (float *)d_scratch, nseg, spw, f->inv2_qmax);
        else      LAUNCH_DEP(inv2_kernel<false>, grid, dim3(128), (size_t)(4 * spw), s, f->d_scale, (const float2 *)f->d_gi32, (const int4 *)f->d_items_inv2, f->nitems_inv2,
                      f->N, f->ncoef, S, (const float4 *)d_zout, (float *)d_scratch, nseg, spw, f->inv2_qmax);
        TSPWS_LAUNCHED_N("inv2_kernel", s);
    }
    if (f->ninv4 > 0) {
        dim3 grid((f->N + 4 * INV4_THREADS - 1) / (4 * INV4_THREADS), f->ninv4, nseg);
        if (d_st)
            inv4_kernel<true, false><<<grid, INV4_THREADS, 0, s4>>>(f->d_scale, (const float2 *)f->d_w32, (const float2 *)f->d_wi32, f->d_invgrp4, S,
                f->N, f->ncoef, (const float4 *)d_zout, (float *)d_scratch, nseg, d_tickets, d_rules, d_out_ts, d_out_ls);
        else
            inv4_kernel<false, false><<<grid, INV4_THREADS, 0, s4>>>(f->d_scale, (const float2 *)f->d_w32, (const float2 *)f->d_wi32, f->d_invgrp4, S,
                f->N, f->ncoef, (const float4 *)d_zout, (float *)d_scratch, nseg, d_tickets, d_rules, d_out_ts, d_out_ls);
        TSPWS_LAUNCHED_N("inv4_kernel", s4);
    }
    if (fork) {
        if (cudaEventRecord(g_inv_join, s4) != cudaSuccess || cudaStreamWaitEvent(s, g_inv_join, 0) != cudaSuccess) return (int)cudaGetLastError();
    }
    {
        dim3 grid((f->N + 127) / 128, nseg);
        if (d_st) LAUNCH_DEP(inv_finish_kernel<true>, grid, dim3(256), 0, s, (const float *)d_scratch, S, nseg, f->N, d_rules, d_out_ts, d_out_ls);
        else      LAUNCH_DEP(inv_finish_kernel<false>, grid, dim3(256), 0, s, (const float *)d_scratch, S, nseg, f->N, d_rules, d_out_ts, d_out_ls);
        TSPWS_LAUNCHED_N("inv_finish_kernel", s);
    }
    tl_host_dump(s);
    return 0;
}
