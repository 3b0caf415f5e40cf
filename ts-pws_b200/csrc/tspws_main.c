/*
 * tspws_main.c -- the drop-in boundary: int tspws_main(t_tsPWS*, t_tsPWS_out*, t_data*)
 * with the reference's semantics (reference src/ts_pws1f_lib.c:48-352), host buffers in and
 * out, all arithmetic on the GPU.  Also the small amount of library state behind the additive
 * exports of include/tspws_b200.h.
 *
 * No CPU fallback: when CUDA is unusable the call fails with TSPWS_ERR_CUDA.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "tspws_internal.h"

/* from tspws_plan.c */
const tspws_frame_t *tspws_plan_frame(tspws_plan_t *pl);
const tspws_devframe_t *tspws_plan_devframe(tspws_plan_t *pl);
/* from tspws_resample.c */
int tspws_resample_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                       unsigned int mtr, int max, const char *sel_override, unsigned int selM,
                       int ss_jackknife, cudaStream_t stream);
int tspws_convergence_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                          unsigned int mtr, int max, cudaStream_t stream);
void tspws_resample_release(void);

#define TSPWS_MAIN_PARTS 8
static struct {
    long          opt_writeback, opt_fp64_tail, opt_device, opt_ss_jackknife;
    int           cuda_state;              /* 0 unknown, 1 ok, -1 unusable */
    cudaStream_t  stream, copy_stream;     /* copy_stream: the traces of a single-stage stack arrive in parts (see tspws_main_locked) */
    cudaEvent_t   part_ev[TSPWS_MAIN_PARTS];
    /* cached plan of the last main stack */
    tspws_plan_t *plan;
    t_tsPWS       plan_key;
    int           plan_max;
    unsigned int  plan_mtr;
    long          plan_fp64;
    /* cached device buffers */
    float        *d_x;   size_t d_x_cap;
    float        *d_out; size_t d_out_cap;
    /* explicit resampling selection */
    char         *sel;   unsigned int selM, selmtr;
} G;

/* The library state above (device buffers, cached plan, stream) is one per process: every entry point that
 * touches it runs under this lock, so concurrent callers are serialised instead of corrupting it (the
 * reference is callable concurrently on distinct data; here the GPU is the shared resource anyway). */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
#define LOCK()   pthread_mutex_lock(&g_lock)
#define UNLOCK() pthread_mutex_unlock(&g_lock)

/* double-buffered slots of tspws_b200_main_batch(): each owns a stream, device buffers and a plan */
#define TSPWS_NSLOT 2
static struct batch_slot {
    cudaStream_t  st;
    tspws_plan_t *plan;
    t_tsPWS       key;
    int           max;
    unsigned int  mtr;
    float        *d_x;   size_t d_x_cap;
    float        *d_out; size_t d_out_cap;
} S[TSPWS_NSLOT];

/* per-kernel timing (CUDA events on the launching stream); see tspws_kernels.cu */
int tspws_b200_profile(int enable) { return tspws_cu_profile_enable(enable); }
int tspws_b200_profile_reset(void) { return tspws_cu_profile_reset(); }
int tspws_b200_profile_read(char *buf, int cap) { return tspws_cu_profile_read(buf, cap); }

const char *tspws_b200_version(void) { return "tspws_b200 0.1 (sm_100a)"; }
unsigned long tspws_b200_kernel_launches(void) { return tspws_cu_launch_count(); }

static void release_device_state(void);

int tspws_b200_set_option(int option, long value)
{
    int rc = 0;
    LOCK();
    switch (option) {
    case TSPWS_OPT_WRITEBACK: G.opt_writeback = value; break;
    case TSPWS_OPT_FP64_TAIL: G.opt_fp64_tail = value; break;
    case TSPWS_OPT_DEVICE:
        /* the stream, the cached plan, the device buffers and the batch slots belong to the previous device:
         * give them back (on that device) before the next call binds to the new one */
        if (value != G.opt_device) {
            if (G.cuda_state == 1) { cudaSetDevice((int)G.opt_device); release_device_state(); }
            G.opt_device = value;
        }
        G.cuda_state = 0;
        break;
    case TSPWS_OPT_SS_JACKKNIFE: G.opt_ss_jackknife = value; break;
    default: rc = TSPWS_ERR_ARG;
    }
    UNLOCK();
    return rc;
}
long tspws_b200_get_option(int option)
{
    switch (option) {
    case TSPWS_OPT_WRITEBACK: return G.opt_writeback;
    case TSPWS_OPT_FP64_TAIL: return G.opt_fp64_tail;
    case TSPWS_OPT_DEVICE:    return G.opt_device;
    case TSPWS_OPT_SS_JACKKNIFE: return G.opt_ss_jackknife;
    default: return -1;
    }
}

int tspws_b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return -(int)e; }
    return n;
}

static int ensure_cuda(void)
{
    if (G.cuda_state == 1) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "tspws_b200: no usable CUDA device (%s); this library has no CPU path\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        G.cuda_state = -1;
        return TSPWS_ERR_CUDA;
    }
    if (G.opt_device < 0 || G.opt_device >= n) {
        fprintf(stderr, "tspws_b200: device %ld out of range (0..%d)\n", G.opt_device, n - 1);
        return TSPWS_ERR_CUDA;
    }
    if (cudaSetDevice((int)G.opt_device) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (!G.stream && cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking) != cudaSuccess) return TSPWS_ERR_CUDA;
    G.cuda_state = 1;
    return 0;
}

static void drop_plan(void)
{
    if (G.plan) { tspws_b200_plan_destroy(G.plan); G.plan = NULL; }
}

/* everything that lives on the current device */
static void release_device_state(void)
{
    drop_plan();
    tspws_resample_release();
    if (G.d_x)   { cudaFree(G.d_x);   G.d_x = NULL;   G.d_x_cap = 0; }
    if (G.d_out) { cudaFree(G.d_out); G.d_out = NULL; G.d_out_cap = 0; }
    if (G.stream) { cudaStreamDestroy(G.stream); G.stream = NULL; }
    if (G.copy_stream) { cudaStreamDestroy(G.copy_stream); G.copy_stream = NULL; }
    for (int i = 0; i < TSPWS_MAIN_PARTS; i++) if (G.part_ev[i]) { cudaEventDestroy(G.part_ev[i]); G.part_ev[i] = NULL; }
    for (int i = 0; i < TSPWS_NSLOT; i++) {
        if (S[i].plan)  tspws_b200_plan_destroy(S[i].plan);
        if (S[i].d_x)   cudaFree(S[i].d_x);
        if (S[i].d_out) cudaFree(S[i].d_out);
        if (S[i].st)    cudaStreamDestroy(S[i].st);
        memset(&S[i], 0, sizeof S[i]);
    }
    G.cuda_state = 0;
}

void tspws_b200_shutdown(void)
{
    LOCK();
    release_device_state();
    free(G.sel); G.sel = NULL; G.selM = G.selmtr = 0;
    UNLOCK();
}

int tspws_b200_set_selection(const char *sel, unsigned int M, unsigned int mtr)
{
    int rc = 0;
    LOCK();
    free(G.sel); G.sel = NULL; G.selM = G.selmtr = 0;
    if (sel) {
        if (!M || !mtr) rc = TSPWS_ERR_ARG;
        else if (!(G.sel = (char *)malloc((size_t)M * mtr))) rc = TSPWS_ERR_ALLOC;
        else { memcpy(G.sel, sel, (size_t)M * mtr); G.selM = M; G.selmtr = mtr; }
    }
    UNLOCK();
    return rc;
}

static int grow(float **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*buf) { cudaFree(*buf); *buf = NULL; *cap = 0; }
    if (cudaMalloc((void **)buf, need * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return TSPWS_ERR_ALLOC; }
    *cap = need;
    return 0;
}

/* the fields of t_tsPWS that shape a plan */
static int same_plan_key(const t_tsPWS *a, const t_tsPWS *b)
{
    return a->type == b->type && a->uni == b->uni && a->J == b->J && a->V == b->V &&
           a->s0 == b->s0 && a->b0 == b->b0 && a->w0 == b->w0 && a->wu == b->wu &&
           a->lrm == b->lrm && a->fold == b->fold && a->unbiased == b->unbiased && a->Kmax == b->Kmax;
}

long tspws_b200_get_stacks(double *st, double *ps, size_t ncoef)
{
    LOCK();
    const long rc = tspws_b200_plan_get_stacks(G.plan, st, ps, ncoef);
    UNLOCK();
    return rc;
}

long tspws_b200_get_coherence(double *coh, size_t ncap)
{
    if (!G.plan) return TSPWS_ERR_NULL;
    const size_t n = tspws_b200_plan_ncoef(G.plan);
    if (!coh) return (long)n;
    if (ncap < n) return TSPWS_ERR_ARG;
    double *ps = (double *)malloc(2 * n * sizeof(double));
    if (!ps) return TSPWS_ERR_ALLOC;
    LOCK();
    long rc = tspws_b200_plan_get_stacks(G.plan, NULL, ps, n);
    const unsigned int K = tspws_b200_plan_K(G.plan);
    UNLOCK();
    if (rc >= 0) {
        for (size_t i = 0; i < n; i++) coh[i] = hypot(ps[2 * i], ps[2 * i + 1]) / (double)K;
    }
    free(ps);
    return rc;
}

static int tspws_check_family(const t_tsPWS *p)
{
    if (p->type <= -1 && p->type >= -3 && p->V) return 0;
    if (p->type == 0)     printf("Error: CreateWaveletFamily: Wavelet type is 0.\n");
    else if (p->type > 0) printf("Error: wavelet type %d: only the complex families (-1 Morlet, -2 exact Morlet, -3 Mexican hat) are supported.\n", p->type);
    else                  printf("Error: CreateWaveletFamily: Complex wavelet not defined.\n");
    return TSPWS_ERR_ALLOC;
}

/* ------------------------------------------------------------------------------------ */
static int tspws_main_locked(t_tsPWS *tspws, t_tsPWS_out *out, t_data *in);

int tspws_main(t_tsPWS *tspws, t_tsPWS_out *out, t_data *in)
{
    if (tspws == NULL || out == NULL || in == NULL) { printf("tspws_main: NULL input\n"); return -1; }
    LOCK();
    const int rc = tspws_main_locked(tspws, out, in);
    UNLOCK();
    return rc;
}

static int tspws_main_locked(t_tsPWS *tspws, t_tsPWS_out *out, t_data *in)
{

    t_hdr *hdr = &in->hdr;
    const int max = hdr->max;
    unsigned int mtr = tspws->Nmax ? tspws->Nmax : hdr->mtr;     /* src/ts_pws1f_lib.c:65 */
    if (mtr > hdr->mtr) mtr = hdr->mtr;                           /* never read past the input */
    const unsigned int nsamp = (unsigned int)max;
    int nerr = 0;

    /* fold only when the lag axis is symmetric (src/ts_pws1f_lib.c:71-75) */
    if (tspws->fold && !tspws_fold_allowed(hdr->beg, hdr->dt, max)) {
        printf("Warning: Folding ignored. B = %f, E = %f, nsamp = %u\n", hdr->beg, hdr->beg + (max - 1) * hdr->dt, nsamp);
        tspws->fold = 0;
    }
    tspws_b200_derive_params(tspws, nsamp, hdr->dt);
    if (tspws->verbose) tspws_print_params(tspws, nsamp, hdr->dt);
    if (max <= 0) return 0;
    /* the frame is built whatever the number of traces; an unusable family makes the reference return 4
     * (pWF == NULL, src/ts_pws1f_lib.c:199-204).  Only the complex families of the ts-PWS are supported here. */
    if ((nerr = tspws_check_family(tspws))) return nerr;
    if (!mtr) return 0;                                           /* nothing to stack (src/ts_pws1f_lib.c:194) */
    if (!in->sigall || !out->ls || !out->tsPWS) { printf("tspws_main: NULL input\n"); return -1; }

    if ((nerr = ensure_cuda())) return nerr;
    cudaStream_t st = G.stream;
    cudaGetLastError();                          /* a stale error of an earlier call must not fail this one */

    /* ---- traces to the device ---- */
    const size_t nelem = (size_t)mtr * (size_t)max;
    if (grow(&G.d_x, &G.d_x_cap, nelem) || grow(&G.d_out, &G.d_out_cap, 2 * (size_t)max)) {
        fprintf(stderr, "tspws_b200: out of device memory\n");
        return TSPWS_ERR_ALLOC;
    }
    const int resampling  = (tspws->subsmpl_N > 0 && tspws->subsmpl_p > 0) || (tspws->jackknife_n > 0 && tspws->jackknife_d > 0);
    const int two_stage   = tspws->Kmax && tspws->Kmax <= mtr;
    /* The host path always cleans the traces themselves, bit-faithfully to the reference's fp32
     * in-place fold / mean removal: it costs one extra pass over HBM, which is hidden behind the
     * PCIe copy, and later passes (resampling, convergence, write-back) need the cleaned traces.
     * (The device-resident plan API instead folds / de-means the Kmax group sums, see tspws_plan.c.) */
    /* ... unless nothing downstream reads the cleaned traces (no write-back, resampling or convergence): then
     * fold / mean removal are applied to the Kmax group sums inside the plan (linear, within 1e-6 of the
     * per-trace fp32 form) and the traces are read from HBM exactly once. */
    const int exact_prep  = two_stage && (tspws->fold || tspws->lrm) &&
                            (G.opt_writeback || resampling || tspws->convergence);
    t_tsPWS pp = *tspws;
    if (exact_prep) { pp.fold = 0; pp.lrm = 0; }

    /* ---- main stack ---- */
    if (!G.plan || !same_plan_key(&G.plan_key, &pp) || G.plan_max != max || G.plan_mtr != mtr ||
        G.plan_fp64 != G.opt_fp64_tail) {
        drop_plan();
        G.plan = tspws_b200_plan_create(&pp, max, mtr, mtr, 0);
        if (!G.plan) { fprintf(stderr, "tspws_b200: cannot create the stacking plan\n"); return TSPWS_ERR_ALLOC; }
        G.plan_key = pp; G.plan_max = max; G.plan_mtr = mtr; G.plan_fp64 = G.opt_fp64_tail;
    }
    /* Single-stage stacks are compute-bound (8 ms of transforms behind a 10 ms copy at 8192 x 16384): the traces go
     * to the device in up to TSPWS_MAIN_PARTS parts on a copy stream and every part is transformed as soon as it has
     * arrived, while the next one is on the bus.  Two-stage stacks read every sample once (0.14 ms behind the same
     * copy): one copy, one graph replay. */
    const unsigned int al = tspws_b200_plan_part_align(G.plan);
    unsigned int part = 0;
    if (al && mtr >= 4 * al && !getenv("TSPWS_MAIN_NO_PARTS")) {
        int np = 6;            /* measured at 8192 x 16384 (w0 = 6): e2e 18.0 ms in one piece, 14.1 / 12.9 / 12.8 / 12.2 / 13.9 ms in 2 / 3 / 4 / 6 / 8 parts */
        { const char *e = getenv("TSPWS_MAIN_PARTS"); if (e && atoi(e) >= 1 && atoi(e) <= TSPWS_MAIN_PARTS) np = atoi(e); }
        part = (mtr + np - 1) / np;
        part = (part + al - 1) / al * al;
        if (!G.copy_stream && cudaStreamCreateWithFlags(&G.copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); part = 0; }
        for (int i = 0; part && i < TSPWS_MAIN_PARTS; i++)
            if (!G.part_ev[i] && cudaEventCreateWithFlags(&G.part_ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); part = 0; }
    }
    if (part) {
        int i = 0;
        for (unsigned int t0 = 0; t0 < mtr; t0 += part, i++) {
            const unsigned int nt = t0 + part < mtr ? part : mtr - t0;
            float *d_p = G.d_x + (size_t)t0 * (size_t)max;
            if (cudaMemcpyAsync(d_p, in->sigall + (size_t)t0 * (size_t)max, (size_t)nt * (size_t)max * sizeof(float),
                                cudaMemcpyHostToDevice, G.copy_stream) != cudaSuccess) goto cuda_fail;
            if (cudaEventRecord(G.part_ev[i], G.copy_stream) != cudaSuccess || cudaStreamWaitEvent(st, G.part_ev[i], 0) != cudaSuccess) goto cuda_fail;
            if ((nerr = tspws_b200_plan_reduce_part(G.plan, d_p, t0, nt, st))) goto cuda_fail;
        }
        if ((nerr = tspws_b200_plan_reduce_done(G.plan, st))) goto cuda_fail;
        if ((nerr = tspws_b200_plan_finish(G.plan, G.d_out, G.d_out + max, st))) goto cuda_fail;
    } else {
        if (cudaMemcpyAsync(G.d_x, in->sigall, nelem * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) goto cuda_fail;
        if (exact_prep && tspws_cu_prep(G.d_x, (int)mtr, max, tspws->fold, tspws->lrm, st)) goto cuda_fail;
        if ((nerr = tspws_b200_plan_stack(G.plan, G.d_x, G.d_out, G.d_out + max, st))) goto cuda_fail;
    }
    if (cudaMemcpyAsync(out->ls, G.d_out, (size_t)max * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) goto cuda_fail;
    if (cudaMemcpyAsync(out->tsPWS, G.d_out + max, (size_t)max * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) goto cuda_fail;
    if (G.opt_writeback && (tspws->fold || tspws->lrm))
        if (cudaMemcpyAsync(in->sigall, G.d_x, nelem * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) goto cuda_fail;
    if (cudaStreamSynchronize(st) != cudaSuccess) goto cuda_fail;

    /* ---- convergence curves (src/ts_pws1f_lib.c:247-314) ---- */
    if (tspws->convergence) {
        int e = tspws_convergence_run(tspws, out, in, G.d_x, mtr, max, st);
        if (e) return e;
    }
    /* ---- resampling (src/ts_pws1f_lib.c:324-345) ---- */
    if (resampling) {
        int e = tspws_resample_run(tspws, out, in, G.d_x, mtr, max, G.sel,
                                   (G.sel && G.selmtr == mtr) ? G.selM : 0, G.opt_ss_jackknife != 0, st);
        if (e) return e;
    }
    return 0;

cuda_fail:
    {
        cudaError_t e = cudaGetLastError();
        fprintf(stderr, "tspws_b200: CUDA failure in tspws_main (%s)\n", cudaGetErrorString(e));
    }
    return TSPWS_ERR_CUDA;
}

/* ------------------------------------------------------------------------------------
 * Batch of independent stacks (station pairs; SURVEY 8b "tspws_main_batch() for C5").
 * Same results as calling tspws_main(tspws, &outs[i], &ins[i]) for i = 0..npairs-1 in turn.
 * When the pairs share one geometry and ask for the main stack only, the calls are pipelined over
 * two slots (stream + device buffers + plan each): the H2D copy of pair i+1 runs under the kernels
 * of pair i, and the host never waits between pairs.  Anything else runs pair by pair.
 * ------------------------------------------------------------------------------------ */
static int main_batch_locked(t_tsPWS *tspws, t_tsPWS_out *outs, t_data *ins, unsigned int npairs);

int tspws_b200_main_batch(t_tsPWS *tspws, t_tsPWS_out *outs, t_data *ins, unsigned int npairs)
{
    if (tspws == NULL || outs == NULL || ins == NULL) { printf("tspws_main: NULL input\n"); return -1; }
    if (!npairs) return 0;
    LOCK();
    const int rc = main_batch_locked(tspws, outs, ins, npairs);
    UNLOCK();
    return rc;
}

static int main_batch_locked(t_tsPWS *tspws, t_tsPWS_out *outs, t_data *ins, unsigned int npairs)
{

    const int resampling = (tspws->subsmpl_N > 0 && tspws->subsmpl_p > 0) || (tspws->jackknife_n > 0 && tspws->jackknife_d > 0);
    int uniform = !resampling && !tspws->convergence && !G.opt_writeback && !G.opt_fp64_tail && ins[0].hdr.max > 0;
    for (unsigned int i = 0; i < npairs && uniform; i++) {
        const t_hdr *h = &ins[i].hdr;
        if (h->max != ins[0].hdr.max || h->mtr != ins[0].hdr.mtr || h->dt != ins[0].hdr.dt || h->beg != ins[0].hdr.beg ||
            !ins[i].sigall || !outs[i].ls || !outs[i].tsPWS) uniform = 0;
    }
    if (!uniform || npairs == 1) {
        /* every call derives from the caller's parameter block as it was on entry, like a loop of CLI runs */
        const t_tsPWS p0 = *tspws;
        int first = 0;
        for (unsigned int i = 0; i < npairs; i++) {
            t_tsPWS p = p0;
            int e = tspws_main_locked(&p, &outs[i], &ins[i]);
            if (i == 0) *tspws = p;
            if (e && !first) first = e;
        }
        return first;
    }

    /* ---- parameters once (identical for every pair) ---- */
    const t_hdr *hdr = &ins[0].hdr;
    const int max = hdr->max;
    unsigned int mtr = tspws->Nmax ? tspws->Nmax : hdr->mtr;
    if (mtr > hdr->mtr) mtr = hdr->mtr;
    if (tspws->fold && !tspws_fold_allowed(hdr->beg, hdr->dt, max)) {
        printf("Warning: Folding ignored. B = %f, E = %f, nsamp = %u\n", hdr->beg, hdr->beg + (max - 1) * hdr->dt, (unsigned int)max);
        tspws->fold = 0;
    }
    tspws_b200_derive_params(tspws, (unsigned int)max, hdr->dt);
    if (tspws->verbose) tspws_print_params(tspws, (unsigned int)max, hdr->dt);
    int nerr = tspws_check_family(tspws);
    if (nerr) return nerr;
    if (!mtr) return 0;
    nerr = ensure_cuda();
    if (nerr) return nerr;
    cudaGetLastError();                          /* a stale error of an earlier call must not fail this one */

    const size_t nelem = (size_t)mtr * (size_t)max;
    const int two_stage  = tspws->Kmax && tspws->Kmax <= mtr;
    /* main stacks only (no write-back, resampling, convergence): fold / mean removal are applied to the group
     * sums inside the plan, exactly like tspws_main does for such calls -> bit-identical to call-by-call */
    const int exact_prep = 0;
    (void)two_stage;
    t_tsPWS pp = *tspws;

    const int nslot = npairs < TSPWS_NSLOT ? (int)npairs : TSPWS_NSLOT;
    for (int k = 0; k < nslot; k++) {
        struct batch_slot *sl = &S[k];
        if (!sl->st && cudaStreamCreateWithFlags(&sl->st, cudaStreamNonBlocking) != cudaSuccess) goto cuda_fail;
        if (grow(&sl->d_x, &sl->d_x_cap, nelem) || grow(&sl->d_out, &sl->d_out_cap, 2 * (size_t)max)) {
            fprintf(stderr, "tspws_b200: out of device memory\n");
            return TSPWS_ERR_ALLOC;
        }
        if (!sl->plan || !same_plan_key(&sl->key, &pp) || sl->max != max || sl->mtr != mtr) {
            if (sl->plan) { tspws_b200_plan_destroy(sl->plan); sl->plan = NULL; }
            sl->plan = tspws_b200_plan_create(&pp, max, mtr, mtr, 0);
            if (!sl->plan) { fprintf(stderr, "tspws_b200: cannot create the stacking plan\n"); return TSPWS_ERR_ALLOC; }
            sl->key = pp; sl->max = max; sl->mtr = mtr;
        }
    }
    for (unsigned int i = 0; i < npairs; i++) {
        struct batch_slot *sl = &S[i % TSPWS_NSLOT];
        /* stream order makes the slot's buffers safe to reuse: pair i-2 has been queued before */
        if (cudaMemcpyAsync(sl->d_x, ins[i].sigall, nelem * sizeof(float), cudaMemcpyHostToDevice, sl->st) != cudaSuccess) goto cuda_fail;
        if (exact_prep && tspws_cu_prep(sl->d_x, (int)mtr, max, tspws->fold, tspws->lrm, sl->st)) goto cuda_fail;
        if ((nerr = tspws_b200_plan_stack(sl->plan, sl->d_x, sl->d_out, sl->d_out + max, sl->st))) goto cuda_fail;
        if (cudaMemcpyAsync(outs[i].ls, sl->d_out, (size_t)max * sizeof(float), cudaMemcpyDeviceToHost, sl->st) != cudaSuccess) goto cuda_fail;
        if (cudaMemcpyAsync(outs[i].tsPWS, sl->d_out + max, (size_t)max * sizeof(float), cudaMemcpyDeviceToHost, sl->st) != cudaSuccess) goto cuda_fail;
    }
    for (int k = 0; k < nslot; k++)
        if (cudaStreamSynchronize(S[k].st) != cudaSuccess) goto cuda_fail;
    return 0;

cuda_fail:
    {
        cudaError_t e = cudaGetLastError();
        fprintf(stderr, "tspws_b200: CUDA failure in tspws_b200_main_batch (%s)\n", cudaGetErrorString(e));
    }
    return TSPWS_ERR_CUDA;
}
