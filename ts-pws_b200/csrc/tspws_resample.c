/*
 * tspws_resample.c -- resampling (random sub-sampling, day-of-year jackknife) and convergence
 * curves on the GPU.  Host C drivers over the kernels of tspws_kernels.cu.
 *
 *   single-stage sub-sampling  reference tspws_subsmpl_float     src/ts_pws1f_lib.c:501-610
 *   two-stage sub-sampling     reference TwoStage_subsmpl_float  src/ts_pws1f_lib.c:612-709
 *   two-stage jackknife        reference TwoStage_jackknife_float src/ts_pws1f_lib.c:719-831
 *   (single-stage jackknife is an empty stub in the reference, :711-716 -- outputs stay untouched)
 *   convergence                src/ts_pws1f_lib.c:247-314
 *
 * Formulation: a resample is a 0/1 selection row over the traces.
 *   two-stage : the selected traces, by rank, are regrouped into Kmax sums (regroup_kernel, every
 *               trace read once per 8 resamples), then Kmax transforms per resample in the
 *               chunk-accumulating forward kernel, weights with (Kmax, K_m), one inverse each.
 *   single    : PS_m = Sel . Phasor (select_sum_kernel over the per-trace phasor matrix, computed
 *               once), ST_m = analysis(Sel . X) by linearity, ls_m = Sel . X / K in the time domain.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "tspws_internal.h"

#define RS_BATCH 64          /* resamples (or prefixes) processed per batch */
#define TR_BATCH 128         /* traces per batch of per-trace coefficient matrices */

#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "tspws_b200: CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); \
    rc = TSPWS_ERR_CUDA; goto done; } } while (0)
#define CKK(call) do { if ((call)) { fprintf(stderr, "tspws_b200: kernel launch failed at %s:%d\n", __FILE__, __LINE__); \
    rc = TSPWS_ERR_CUDA; goto done; } } while (0)

typedef struct {
    tspws_frame_t   *fr;
    tspws_devframe_t acc;     /* chunk-accumulate item lists */
    tspws_devframe_t pt;      /* per-trace item lists        */
    int have_pt;
} rs_frames_t;

static int frames_open(rs_frames_t *F, const t_tsPWS *p, int max, int need_pt)
{
    memset(F, 0, sizeof(*F));
    F->fr = tspws_b200_frame_create(p->type, p->J, p->V, (unsigned int)max, p->s0, p->b0, p->w0, (int)p->uni);
    if (!F->fr) return TSPWS_ERR_ALLOC;
    if (tspws_devframe_build(&F->acc, F->fr, 0, 0)) return TSPWS_ERR_CUDA;
    if (need_pt) { if (tspws_devframe_build(&F->pt, F->fr, 0, 1)) return TSPWS_ERR_CUDA; F->have_pt = 1; }
    return 0;
}
static void frames_close(rs_frames_t *F)
{
    tspws_devframe_free(&F->acc);
    if (F->have_pt) tspws_devframe_free(&F->pt);
    tspws_b200_frame_destroy(F->fr);
    memset(F, 0, sizeof(*F));
}

/* The frames of the sub-sampling driver are kept between calls: building them (filters on the host, the polyphase
 * tables, the uploads) took ~25 ms of every call at 32768 samples.  Key = everything frame_create depends on. */
static struct { rs_frames_t F; int valid, type, uni, max, need_pt; unsigned int J, V; double s0, b0, w0; } g_fc;
static rs_frames_t *frames_cached(const t_tsPWS *p, int max, int need_pt, int *rc)
{
    if (g_fc.valid && g_fc.type == p->type && g_fc.uni == (int)p->uni && g_fc.max == max && g_fc.J == p->J && g_fc.V == p->V &&
        g_fc.s0 == p->s0 && g_fc.b0 == p->b0 && g_fc.w0 == p->w0 && (g_fc.need_pt || !need_pt)) return &g_fc.F;
    if (g_fc.valid) { frames_close(&g_fc.F); g_fc.valid = 0; }
    if ((*rc = frames_open(&g_fc.F, p, max, need_pt))) { frames_close(&g_fc.F); return NULL; }
    g_fc.valid = 1; g_fc.type = p->type; g_fc.uni = (int)p->uni; g_fc.max = max; g_fc.need_pt = need_pt;
    g_fc.J = p->J; g_fc.V = p->V; g_fc.s0 = p->s0; g_fc.b0 = p->b0; g_fc.w0 = p->w0;
    return &g_fc.F;
}

/* group of the r-th selected trace out of K: src/ts_pws1f_lib.c:648 (unsigned product, double division) */
static int rank_group(unsigned int r, unsigned int Kmax, unsigned int K)
{
    return (int)(unsigned int)floor((double)(r * Kmax) / (double)K);
}

/* ------------------------------------------------------------------------------------ */
/* Row pipe: the stacks of a batch of realisations leave the device through double-buffered device rows and PINNED
 * staging on a copy stream, and are scattered to the caller's (pageable) row pointers by the host while the GPU
 * computes the next batch.  (Round 1 copied every batch synchronously into pageable memory: at configs[3] the
 * 2 x 131 MB of results cost more than all the kernels.) */
typedef struct {
    int B; size_t N;
    float  *d_ls[2], *d_ts[2];
    double *d_sim[2];                    /* [2 B] similarity, misfit (convergence) */
    float  *h_pin[2];                    /* pinned [2][B][N]: ls rows then tsPWS rows */
    double *h_sim[2];
    cudaStream_t cst;
    cudaEvent_t ev_done[2], ev_copied[2];
    int nb[2], with_ls[2], with_ts[2], with_sim[2], used[2];
    unsigned int m0[2];
} rowpipe_t;

static void rp_close(rowpipe_t *rp)
{
    for (int i = 0; i < 2; i++) {
        cudaFree(rp->d_ls[i]); cudaFree(rp->d_ts[i]); cudaFree(rp->d_sim[i]);
        if (rp->h_pin[i]) cudaFreeHost(rp->h_pin[i]);
        if (rp->h_sim[i]) cudaFreeHost(rp->h_sim[i]);
        if (rp->ev_done[i]) cudaEventDestroy(rp->ev_done[i]);
        if (rp->ev_copied[i]) cudaEventDestroy(rp->ev_copied[i]);
    }
    if (rp->cst) cudaStreamDestroy(rp->cst);
    memset(rp, 0, sizeof *rp);
}
static int rp_open(rowpipe_t *rp, int B, size_t N)
{
    memset(rp, 0, sizeof *rp);
    rp->B = B; rp->N = N;
    int ok = cudaStreamCreateWithFlags(&rp->cst, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaMalloc((void **)&rp->d_ls[i], (size_t)B * N * sizeof(float)) == cudaSuccess &&
             cudaMalloc((void **)&rp->d_ts[i], (size_t)B * N * sizeof(float)) == cudaSuccess &&
             cudaMalloc((void **)&rp->d_sim[i], 2 * (size_t)B * sizeof(double)) == cudaSuccess &&
             cudaHostAlloc((void **)&rp->h_pin[i], 2 * (size_t)B * N * sizeof(float), cudaHostAllocDefault) == cudaSuccess &&
             cudaHostAlloc((void **)&rp->h_sim[i], 2 * (size_t)B * sizeof(double), cudaHostAllocDefault) == cudaSuccess &&
             cudaEventCreateWithFlags(&rp->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&rp->ev_copied[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); rp_close(rp); return TSPWS_ERR_ALLOC; }
    return 0;
}
/* One pipe is kept between calls (pinning host memory costs milliseconds); callers hold the library lock. */
static rowpipe_t g_rp;
static int rp_acquire(rowpipe_t **out, int B, size_t N)
{
    if (g_rp.cst && (g_rp.B != B || g_rp.N != N)) rp_close(&g_rp);
    if (!g_rp.cst) { const int e = rp_open(&g_rp, B, N); if (e) return e; }
    g_rp.used[0] = g_rp.used[1] = 0;
    *out = &g_rp;
    return 0;
}
/* Device scratch of the two sub-sampling drivers, kept between calls: at BASELINE configs[3] (2.5 GB in 15 buffers) the
 * cudaMalloc / cudaFree round of every call cost ~30 ms of its 140.  Slot = one buffer of one driver; grows on demand. */
#define RS_SLOTS 28
static struct { void *p; size_t cap; } g_scr[RS_SLOTS];
static cudaError_t rs_scratch(int slot, void **p, size_t bytes)
{
    if (g_scr[slot].cap < bytes) {
        if (g_scr[slot].p) cudaFree(g_scr[slot].p);
        g_scr[slot].p = NULL; g_scr[slot].cap = 0;
        const cudaError_t e = cudaMalloc(&g_scr[slot].p, bytes);
        if (e != cudaSuccess) return e;
        g_scr[slot].cap = bytes;
    }
    *p = g_scr[slot].p;
    return cudaSuccess;
}

void tspws_resample_release(void)
{
    if (g_rp.cst) rp_close(&g_rp);
    for (int i = 0; i < RS_SLOTS; i++) { if (g_scr[i].p) cudaFree(g_scr[i].p); g_scr[i].p = NULL; g_scr[i].cap = 0; }
    if (g_fc.valid) { frames_close(&g_fc.F); g_fc.valid = 0; }
}

/* batch k is about to be computed into buffers k & 1: the copy of batch k - 2 out of them must have finished */
static int rp_begin(rowpipe_t *rp, int k, cudaStream_t st)
{
    const int i = k & 1;
    if (rp->used[i] && cudaStreamWaitEvent(st, rp->ev_copied[i], 0) != cudaSuccess) return TSPWS_ERR_CUDA;
    return 0;
}
/* the kernels of batch k are queued on st: queue its D2H behind them on the copy stream */
static int rp_submit(rowpipe_t *rp, int k, cudaStream_t st, unsigned int m0, int nb, int with_ls, int with_ts, int with_sim)
{
    const int i = k & 1;
    const size_t row = rp->N * sizeof(float);
    if (cudaEventRecord(rp->ev_done[i], st) != cudaSuccess || cudaStreamWaitEvent(rp->cst, rp->ev_done[i], 0) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (with_ls && cudaMemcpyAsync(rp->h_pin[i], rp->d_ls[i], (size_t)nb * row, cudaMemcpyDeviceToHost, rp->cst) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (with_ts && cudaMemcpyAsync(rp->h_pin[i] + (size_t)rp->B * rp->N, rp->d_ts[i], (size_t)nb * row, cudaMemcpyDeviceToHost, rp->cst) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (with_sim && cudaMemcpyAsync(rp->h_sim[i], rp->d_sim[i], 2 * (size_t)rp->B * sizeof(double), cudaMemcpyDeviceToHost, rp->cst) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (cudaEventRecord(rp->ev_copied[i], rp->cst) != cudaSuccess) return TSPWS_ERR_CUDA;
    rp->nb[i] = nb; rp->m0[i] = m0; rp->with_ls[i] = with_ls; rp->with_ts[i] = with_ts; rp->with_sim[i] = with_sim; rp->used[i] = 1;
    return 0;
}
/* host side of batch k: wait for its copy, scatter the rows */
static int rp_collect(rowpipe_t *rp, int k, float **ls_rows, float **ts_rows, double *sim_out, double *mis_out)
{
    const int i = k & 1;
    if (!rp->used[i]) return 0;
    if (cudaEventSynchronize(rp->ev_copied[i]) != cudaSuccess) return TSPWS_ERR_CUDA;
    const size_t row = rp->N * sizeof(float);
    for (int b = 0; b < rp->nb[i]; b++) {
        if (rp->with_ls[i] && ls_rows && ls_rows[rp->m0[i] + b]) memcpy(ls_rows[rp->m0[i] + b], rp->h_pin[i] + (size_t)b * rp->N, row);
        if (rp->with_ts[i] && ts_rows && ts_rows[rp->m0[i] + b]) memcpy(ts_rows[rp->m0[i] + b], rp->h_pin[i] + ((size_t)rp->B + b) * rp->N, row);
        if (rp->with_sim[i] && sim_out) { sim_out[rp->m0[i] + b] = rp->h_sim[i][b]; mis_out[rp->m0[i] + b] = rp->h_sim[i][rp->B + b]; }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* two-stage machinery: nres selection rows -> ls / tsPWS rows (host pointers)            */
/* ------------------------------------------------------------------------------------ */
static int two_stage_rows(const t_tsPWS *p, rs_frames_t *F, const float *d_x, unsigned int mtr, int max,
                          const char *sel /*[nres][mtr]*/, unsigned int nres, unsigned int Kmax,
                          float **ls_rows, float **ts_rows, unsigned int *K_out,
                          const float *d_ref, double *sim_out, double *mis_out, cudaStream_t st)
{
    int rc = 0;
    const size_t N = (size_t)max, ncoef = F->fr->ncoef ? F->fr->ncoef : 1;
    const int B = RS_BATCH;
    short *h_gid = NULL, *d_gid = NULL;          /* group of every trace per realisation (-1: not selected); any Kmax < 32768 */
    float *d_gs = NULL;
    void *d_pps = NULL, *d_pst = NULL, *d_zout = NULL, *d_scr = NULL;
    double *d_scale = NULL, *h_scale = NULL;
    tspws_weight_t *d_rules = NULL, *h_rules = NULL;
    unsigned int *d_tk = NULL, *Ks = NULL;
    const int ntk = (max + 255) / 256;
    rowpipe_t *rpp = NULL;
    signed char *h_cnt = NULL, *d_cnt = NULL;
    void *d_ximg = NULL, *d_aimg = NULL;

    if (Kmax > 32767) { fprintf(stderr, "tspws_b200: two-stage resampling supports Kmax <= 32767\n"); return TSPWS_ERR_ARG; }
    for (size_t i = 0; i < (size_t)nres * mtr; i++)
        if (sel[i] > 1) {
            fprintf(stderr, "tspws_b200: count matrices (bootstrap with replacement) are supported for single-stage stacks only\n");
            return TSPWS_ERR_ARG;
        }
    h_gid   = (short *)malloc((size_t)B * mtr * sizeof(short));
    h_scale = (double *)malloc(B * sizeof(double));
    h_rules = (tspws_weight_t *)malloc(B * sizeof(tspws_weight_t));
    Ks      = (unsigned int *)malloc(B * sizeof(unsigned int));
    if (!h_gid || !h_scale || !h_rules || !Ks) { rc = TSPWS_ERR_ALLOC; goto done; }
    if ((rc = rp_acquire(&rpp, B, N))) goto done;
    CKR(rs_scratch(0, (void **)&d_gid, (size_t)B * mtr * sizeof(short)));
    CKR(rs_scratch(1, (void **)&d_gs, (size_t)B * Kmax * N * sizeof(float)));
    CKR(rs_scratch(2, (void **)&d_pps, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(rs_scratch(3, (void **)&d_pst, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(rs_scratch(4, (void **)&d_zout, (size_t)B * ncoef * 4 * sizeof(float)));
    CKR(rs_scratch(5, (void **)&d_scr, (size_t)2 * (F->acc.inv_slots ? F->acc.inv_slots : 1) * B * N * sizeof(float)));
    CKR(rs_scratch(6, (void **)&d_scale, B * sizeof(double)));
    CKR(rs_scratch(7, (void **)&d_rules, B * sizeof(tspws_weight_t)));
    CKR(rs_scratch(8, (void **)&d_tk, (size_t)B * ntk * sizeof(unsigned int)));
    CKR(cudaMemsetAsync(d_tk, 0, (size_t)B * ntk * sizeof(unsigned int), st));
    /* The Kmax group sums of a batch of realisations are a 0/1 matrix times the traces: G[(b, g)][n] = sum_t [trace t is in
     * group g of realisation b] x_t[n] -- on the tensor cores like the single-stage sums (traces as three bf16 terms,
     * fp32-faithful; indicator exact): 0.4 ms per batch of 64 x 10 groups at 2000 x 32768 instead of the 3 ms of
     * regroup_kernel.  TSPWS_RS_REGROUP=1 / odd lengths / very large Kmax keep the SIMT kernel. */
    const int kblocks = ((int)mtr + 63) / 64;
    const int sums_mma = (max % 2 == 0) && (size_t)B * Kmax <= 8192 && !(getenv("TSPWS_RS_REGROUP") && atoi(getenv("TSPWS_RS_REGROUP"))) &&
                         !(getenv("TSPWS_NO_MMA") && atoi(getenv("TSPWS_NO_MMA")));
    if (sums_mma) {
        h_cnt = (signed char *)malloc((size_t)B * Kmax * mtr);
        if (!h_cnt) { rc = TSPWS_ERR_ALLOC; goto done; }
        CKR(rs_scratch(24, &d_ximg, tspws_cu_mma_b_image_bytes(max, (int)mtr, 3)));
        CKR(rs_scratch(26, (void **)&d_cnt, (size_t)B * Kmax * mtr));
        CKR(rs_scratch(27, &d_aimg, tspws_cu_mma_a_image_bytes(B * (int)Kmax, (int)mtr)));
        CKR(cudaMemsetAsync(d_ximg, 0, tspws_cu_mma_b_image_bytes(max, (int)mtr, 3), st));
        for (unsigned int t0 = 0; t0 < mtr; t0 += 4096) {
            const int nt = (int)((mtr - t0 < 4096) ? mtr - t0 : 4096);
            CKK(tspws_cu_phasor_image(d_x + (size_t)t0 * N, max / 2, nt, (int)t0, kblocks, 3, d_ximg, st));
        }
    }

    int k = 0;
    for (unsigned int m0 = 0; m0 < nres; m0 += B, k++) {
        const int nb = (int)((nres - m0 < (unsigned int)B) ? nres - m0 : (unsigned int)B);
        float *d_ls = rpp->d_ls[k & 1], *d_ts = rpp->d_ts[k & 1];
        double *d_sim = rpp->d_sim[k & 1];
        /* (h_gid / h_scale / h_rules are pageable: cudaMemcpyAsync has staged them when it returns, so they can be refilled) */
        for (int b = 0; b < nb; b++) {
            const char *row = sel + (size_t)(m0 + b) * mtr;
            unsigned int K = 0;
            for (unsigned int t = 0; t < mtr; t++) K += (row[t] == 1);
            unsigned int r = 0;
            for (unsigned int t = 0; t < mtr; t++)
                h_gid[(size_t)b * mtr + t] = (row[t] == 1) ? (short)rank_group(r++, Kmax, K) : (short)-1;
            if (sums_mma) {
                memset(h_cnt + (size_t)b * Kmax * mtr, 0, (size_t)Kmax * mtr);
                for (unsigned int t = 0; t < mtr; t++)
                    if (h_gid[(size_t)b * mtr + t] >= 0) h_cnt[((size_t)b * Kmax + (size_t)h_gid[(size_t)b * mtr + t]) * mtr + t] = 1;
            }
            Ks[b] = K;
            h_scale[b] = K ? 1.0 / (double)K : 0.0;                                   /* W/K with W = 1 (:672,:686) */
            h_rules[b] = tspws_weight_rule(p->wu, p->unbiased, Kmax, K ? K : 1);
            if (K_out) K_out[m0 + b] = K;
        }
        CKR(cudaMemcpyAsync(d_gid, h_gid, (size_t)nb * mtr * sizeof(short), cudaMemcpyHostToDevice, st));
        CKR(cudaMemcpyAsync(d_scale, h_scale, nb * sizeof(double), cudaMemcpyHostToDevice, st));
        CKR(cudaMemcpyAsync(d_rules, h_rules, nb * sizeof(tspws_weight_t), cudaMemcpyHostToDevice, st));
        if ((rc = rp_begin(rpp, k, st))) goto done;
        if (sums_mma) {
            CKR(cudaMemcpyAsync(d_cnt, h_cnt, (size_t)nb * Kmax * mtr, cudaMemcpyHostToDevice, st));
            CKK(tspws_cu_count_image(d_cnt, (int)mtr, nb * (int)Kmax, (int)mtr, d_aimg, st));
            CKK(tspws_cu_resample_mma(d_aimg, d_ximg, 3, nb * (int)Kmax, max, kblocks, d_gs, N, st));
        } else {
            CKK(tspws_cu_regroup(d_x, max, (int)mtr, d_gid, NULL, nb, (int)Kmax, d_gs, st));
        }
        if (ls_rows) CKK(tspws_cu_group_total(d_gs, nb, (int)Kmax, max, d_scale, d_ls, st));
        CKK(tspws_cu_forward_fast(&F->acc, d_gs, nb * (int)Kmax, (int)Kmax, d_pps, d_pst, 0, 5, st));
        CKK(tspws_cu_weight(d_pps, d_pst, 0, (int)F->fr->ncoef, nb, d_rules, d_zout, st));
        CKK(tspws_cu_inverse(&F->acc, d_zout, NULL, 0, nb, d_rules, d_scr, d_tk, d_ts, NULL, st, NULL, NULL, NULL));
        if (d_ref && sim_out) CKK(tspws_cu_simil(d_ts, nb, max, d_ref, d_sim, d_sim + B, st));
        if ((rc = rp_submit(rpp, k, st, m0, nb, ls_rows != NULL, ts_rows != NULL, d_ref && sim_out))) goto done;
        if (k > 0 && (rc = rp_collect(rpp, k - 1, ls_rows, ts_rows, sim_out, mis_out))) goto done;
    }
    if (k > 0) rc = rp_collect(rpp, k - 1, ls_rows, ts_rows, sim_out, mis_out);
done:
    cudaStreamSynchronize(st);
    free(h_gid); free(h_scale); free(h_rules); free(Ks); free(h_cnt);
    /* the device scratch stays cached (rs_scratch) */
    if (rpp) cudaStreamSynchronize(rpp->cst);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* single-stage sub-sampling                                                             */
/* ------------------------------------------------------------------------------------ */
static int single_stage_subsample(const t_tsPWS *p, rs_frames_t *F, const float *d_x, unsigned int mtr, int max,
                                  const char *sel, unsigned int M, unsigned int K, unsigned int *k_rows /* out or NULL */,
                                  float **ls_rows, float **ts_rows, cudaStream_t st)
{
    /* sel[m][t] = how many times realisation m stacks trace t: 0 / 1 (sub-sampling, jackknife) or a count (bootstrap
     * with replacement); K_m = sum_t sel[m][t] is written to k_rows[m].  (K: the count every row is expected to have
     * in sub-sampling mode; only used for the weights' K when all rows agree -- the per-row sums are what is applied.)
     * Phase stacks: PS_m = sum_t sel[m][t] phasor(Y_t), the [M x T] . [T x 2 ncoef] contraction, on the tensor cores
     * (tspws_mma.cu) unless TSPWS_NO_MMA=1 (the SIMT kernel it replaced, kept for the before / after profile). */
    int rc = 0;
    const size_t N = (size_t)max, ncoef = F->fr->ncoef ? F->fr->ncoef : 1;
    const int B = RS_BATCH, ntk = (max + 255) / 256;
    const int use_mma = !getenv("TSPWS_NO_MMA") || atoi(getenv("TSPWS_NO_MMA")) == 0;
    const int nterm = (getenv("TSPWS_MMA_TERMS") && atoi(getenv("TSPWS_MMA_TERMS")) == 3) ? 3 : 2;
    const int kblocks = ((int)mtr + 63) / 64;
    const int pertrace_path = F->have_pt && getenv("TSPWS_RS_PERTRACE") && atoi(getenv("TSPWS_RS_PERTRACE"));   /* the previous transform path (A/B) */
    short *h_gid = NULL, *d_gid = NULL;
    signed char *d_sel = NULL;
    float *d_S = NULL;
    void *d_ypart = NULL, *d_P = NULL, *d_ps = NULL, *d_pps = NULL, *d_pst = NULL, *d_zout = NULL, *d_scr = NULL;
    rowpipe_t *rpp = NULL;
    void *d_aimg = NULL, *d_bimg = NULL, *d_ximg = NULL;
    float *d_Sall = NULL;                  /* [M][N] time-domain sums of all realisations (tensor-core path) */
    double *d_scale = NULL, *h_scale = NULL;
    tspws_weight_t *d_rules = NULL, *h_rules = NULL;
    unsigned int *d_tk = NULL, *km_all = NULL;
    (void)K;

    h_gid   = (short *)malloc((size_t)M * mtr * sizeof(short));
    h_scale = (double *)malloc(B * sizeof(double));
    h_rules = (tspws_weight_t *)malloc(B * sizeof(tspws_weight_t));
    km_all  = (unsigned int *)calloc(M ? M : 1, sizeof(unsigned int));
    if (!h_gid || !h_scale || !h_rules || !km_all) { rc = TSPWS_ERR_ALLOC; goto done; }
    if ((rc = rp_acquire(&rpp, B, N))) goto done;
    for (unsigned int m = 0; m < M; m++)
        for (unsigned int t = 0; t < mtr; t++) {
            const int c = sel[(size_t)m * mtr + t];
            h_gid[(size_t)m * mtr + t] = c > 0 ? 0 : -1;
            if (c > 0) km_all[m] += (unsigned int)c;
        }

    CKR(rs_scratch(9, (void **)&d_gid, (size_t)M * mtr * sizeof(short)));
    CKR(rs_scratch(10, (void **)&d_sel, (size_t)M * mtr));
    CKR(rs_scratch(11, (void **)&d_S, (size_t)B * N * sizeof(float)));
    CKR(rs_scratch(12, (void **)&d_ypart, (size_t)TR_BATCH * ((F->have_pt && (size_t)F->pt.ypart_len > ncoef) ? (size_t)F->pt.ypart_len : ncoef) * 2 * sizeof(float)));
    CKR(rs_scratch(13, (void **)&d_P, (size_t)TR_BATCH * ncoef * 2 * sizeof(float)));
    CKR(rs_scratch(14, (void **)&d_ps, (size_t)M * ncoef * 2 * (use_mma ? sizeof(float) : sizeof(double))));
    CKR(rs_scratch(15, (void **)&d_pps, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(rs_scratch(16, (void **)&d_pst, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(rs_scratch(17, (void **)&d_zout, (size_t)B * ncoef * 4 * sizeof(float)));
    CKR(rs_scratch(18, (void **)&d_scr, (size_t)2 * (F->acc.inv_slots ? F->acc.inv_slots : 1) * B * N * sizeof(float)));
    CKR(rs_scratch(19, (void **)&d_scale, B * sizeof(double)));
    CKR(rs_scratch(20, (void **)&d_rules, B * sizeof(tspws_weight_t)));
    CKR(rs_scratch(21, (void **)&d_tk, (size_t)B * ntk * sizeof(unsigned int)));
    CKR(cudaMemsetAsync(d_tk, 0, (size_t)B * ntk * sizeof(unsigned int), st));
    CKR(cudaMemcpyAsync(d_gid, h_gid, (size_t)M * mtr * sizeof(short), cudaMemcpyHostToDevice, st));
    CKR(cudaMemcpyAsync(d_sel, sel, (size_t)M * mtr, cudaMemcpyHostToDevice, st));
    if (use_mma) {
        /* tile images: counts (A) once, unit phasors (B) batch by batch as the traces are transformed */
        CKR(rs_scratch(22, (void **)&d_aimg, tspws_cu_mma_a_image_bytes((int)M, (int)mtr)));
        CKR(rs_scratch(23, (void **)&d_bimg, tspws_cu_mma_b_image_bytes(2 * (int)F->fr->ncoef, (int)mtr, nterm)));
        /* cached buffers: no stale (possibly non-finite) bits in the padding of the phasor image */
        CKR(cudaMemsetAsync(d_bimg, 0, tspws_cu_mma_b_image_bytes(2 * (int)F->fr->ncoef, (int)mtr, nterm), st));
        CKK(tspws_cu_count_image(d_sel, (int)mtr, (int)M, (int)mtr, d_aimg, st));
    } else {
        CKR(cudaMemsetAsync(d_ps, 0, (size_t)M * ncoef * 2 * sizeof(double), st));
    }

    /* every trace is transformed once: per-trace coefficients -> unit phasors -> image (or the SIMT contraction) */
    for (unsigned int t0 = 0; t0 < mtr; t0 += TR_BATCH) {
        const int nt = (int)((mtr - t0 < TR_BATCH) ? mtr - t0 : TR_BATCH);
        if (pertrace_path) {
            CKK(tspws_cu_forward_pertrace(&F->pt, d_x + (size_t)t0 * N, nt, 0, d_ypart, st));
            CKK(tspws_cu_coef_matrix(&F->pt, d_ypart, nt, NULL, d_P, st));
        } else {
            /* the chunk-accumulating kernels with chunks of ONE trace: the "partial phase stack" of a chunk is then the
             * trace's unit phasors (d_P), its "partial linear stack" the coefficients themselves (scratch, unused).
             * The per-trace kernels are built for the ~10 traces of a two-stage tail (phase splits, a combine pass): at
             * 2000 x 32768 they took 55 ms of the call's 130 ms of kernels. */
            CKK(tspws_cu_forward_fast(&F->acc, d_x + (size_t)t0 * N, nt, 1, d_P, d_ypart, 0, 5, st));
        }
        if (use_mma) CKK(tspws_cu_phasor_image(d_P, (int)F->fr->ncoef, nt, (int)t0, kblocks, nterm, d_bimg, st));
        else         CKK(tspws_cu_select_sum(d_P, nt, (int)F->fr->ncoef, d_sel, (int)mtr, (int)t0, (int)M, d_ps, st));
    }
    if (use_mma)
        CKK(tspws_cu_resample_mma(d_aimg, d_bimg, nterm, (int)M, 2 * (int)F->fr->ncoef, kblocks, (float *)d_ps, 2 * ncoef, st));
    /* The time-domain sums S_m = sum_t sel[m][t] x_t (ls, and ST by linearity) are the same contraction with the traces
     * in place of the phasors: [M x T] . [T x N], traces as three bf16 terms (fp32-faithful), counts exact -- one launch
     * for all realisations instead of a 1.9 ms regroup launch per batch of 64 (30 of the call's 130 ms at configs[3]).
     * Needs an even trace length (the image builder takes the samples in pairs). */
    const int sums_mma = use_mma && (max % 2 == 0) && !(getenv("TSPWS_RS_REGROUP") && atoi(getenv("TSPWS_RS_REGROUP")));
    if (sums_mma) {
        CKR(rs_scratch(24, &d_ximg, tspws_cu_mma_b_image_bytes(max, (int)mtr, 3)));
        CKR(rs_scratch(25, (void **)&d_Sall, (size_t)M * N * sizeof(float)));
        CKR(cudaMemsetAsync(d_ximg, 0, tspws_cu_mma_b_image_bytes(max, (int)mtr, 3), st));
        for (unsigned int t0 = 0; t0 < mtr; t0 += 4096) {          /* the builder's grid.x is a count of 64-trace blocks */
            const int nt = (int)((mtr - t0 < 4096) ? mtr - t0 : 4096);
            CKK(tspws_cu_phasor_image(d_x + (size_t)t0 * N, max / 2, nt, (int)t0, kblocks, 3, d_ximg, st));
        }
        CKK(tspws_cu_resample_mma(d_aimg, d_ximg, 3, (int)M, max, kblocks, d_Sall, N, st));
    }

    /* per batch of realisations: time-domain sums -> ls, ST by linearity, weights, inverse */
    int k = 0;
    for (unsigned int m0 = 0; m0 < M; m0 += B, k++) {
        const int nb = (int)((M - m0 < (unsigned int)B) ? M - m0 : (unsigned int)B);
        float *d_ls = rpp->d_ls[k & 1], *d_ts = rpp->d_ts[k & 1];
        for (int b = 0; b < nb; b++) {
            unsigned int km = km_all[m0 + b];
            if (k_rows) k_rows[m0 + b] = km;
            if (!km) km = 1;
            h_scale[b] = 1.0 / (double)km;
            h_rules[b] = tspws_weight_rule(p->wu, p->unbiased, km, km);
        }
        CKR(cudaMemcpyAsync(d_scale, h_scale, B * sizeof(double), cudaMemcpyHostToDevice, st));
        CKR(cudaMemcpyAsync(d_rules, h_rules, B * sizeof(tspws_weight_t), cudaMemcpyHostToDevice, st));
        if ((rc = rp_begin(rpp, k, st))) goto done;
        if (sums_mma) d_S = d_Sall + (size_t)m0 * N;
        else CKK(tspws_cu_regroup(d_x, max, (int)mtr, d_gid + (size_t)m0 * mtr, d_sel + (size_t)m0 * mtr, nb, 1, d_S, st));
        CKK(tspws_cu_group_total(d_S, nb, 1, max, d_scale, d_ls, st));
        CKK(tspws_cu_forward_fast(&F->acc, d_S, nb, 1, d_pps, d_pst, 0, 5, st));
        if (use_mma)
            CKK(tspws_cu_weight_mixed32((char *)d_ps + (size_t)m0 * ncoef * 2 * sizeof(float), d_pst, (int)F->fr->ncoef, nb,
                                        d_rules, d_zout, st));
        else
            CKK(tspws_cu_weight_mixed((char *)d_ps + (size_t)m0 * ncoef * 2 * sizeof(double), d_pst, (int)F->fr->ncoef, nb,
                                      d_rules, d_zout, st));
        CKK(tspws_cu_inverse(&F->acc, d_zout, NULL, 0, nb, d_rules, d_scr, d_tk, d_ts, NULL, st, NULL, NULL, NULL));
        if ((rc = rp_submit(rpp, k, st, m0, nb, 1, 1, 0))) goto done;
        if (k > 0 && (rc = rp_collect(rpp, k - 1, ls_rows, ts_rows, NULL, NULL))) goto done;
    }
    if (k > 0) rc = rp_collect(rpp, k - 1, ls_rows, ts_rows, NULL, NULL);
done:
    cudaStreamSynchronize(st);
    if (rpp) cudaStreamSynchronize(rpp->cst);
    free(h_gid); free(h_scale); free(h_rules); free(km_all);
    /* the device scratch stays cached (rs_scratch) */
    return rc;
}

/* ------------------------------------------------------------------------------------ */
int tspws_resample_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                       unsigned int mtr, int max, const char *sel_override, unsigned int selM,
                       int ss_jackknife, cudaStream_t stream)
{
    int rc = 0;
    const int two_stage = p->Kmax && p->Kmax <= mtr;
    char *sel = NULL;
    if (!out->ls_subsmpl || !out->tsPWS_subsmpl) { printf("tspws_main: NULL resampling outputs\n"); return -1; }
    /* the per-trace frame is only needed by the previous single-stage transform path (TSPWS_RS_PERTRACE=1) */
    const int need_pt = !two_stage && getenv("TSPWS_RS_PERTRACE") && atoi(getenv("TSPWS_RS_PERTRACE"));
    rs_frames_t *Fp = frames_cached(p, max, need_pt, &rc);
    if (!Fp) return rc;
#define F (*Fp)

    /* ---- random sub-sampling (src/ts_pws1f_lib.c:324-333) ---- */
    if (p->subsmpl_N > 0 && p->subsmpl_p > 0) {
        const unsigned int K = (unsigned int)ceil((double)mtr * p->subsmpl_p);
        const unsigned int M = p->subsmpl_N;
        sel = (char *)malloc((size_t)M * mtr);
        if (!sel) { rc = TSPWS_ERR_ALLOC; goto done; }
        if (sel_override && selM == M) memcpy(sel, sel_override, (size_t)M * mtr);
        else for (unsigned int m = 0; m < M; m++) tspws_b200_subsampling_plan(sel + (size_t)m * mtr, mtr, K);
        if (K == 0 || K > mtr) { fprintf(stderr, "tspws_b200: subsmpl_p out of range\n"); rc = TSPWS_ERR_ARG; goto done; }
        if (two_stage) rc = two_stage_rows(p, &F, d_x, mtr, max, sel, M, p->Kmax, out->ls_subsmpl, out->tsPWS_subsmpl,
                                           NULL, NULL, NULL, NULL, stream);
        else           rc = single_stage_subsample(p, &F, d_x, mtr, max, sel, M, K, NULL, out->ls_subsmpl, out->tsPWS_subsmpl, stream);
        free(sel); sel = NULL;
        if (rc) goto done;
    }
    /* ---- delete-d jackknife over day-of-year bins (src/ts_pws1f_lib.c:335-345) ---- */
    if (p->jackknife_n > 0 && p->jackknife_d > 0) {
        const unsigned int Cn = out->M;
        /* single-stage: the reference is an empty stub that leaves the outputs untouched (:711-716);
         * TSPWS_OPT_SS_JACKKNIFE turns the extension of SURVEY 8(f) rank 3 on */
        if (!two_stage && !ss_jackknife) goto done;
        if (!Cn) goto done;
        sel = (char *)calloc((size_t)Cn, mtr);       /* rows beyond C(n,d) (out->M too large) stay empty, not heap garbage */
        if (!sel) { rc = TSPWS_ERR_ALLOC; goto done; }
        if (sel_override && selM == Cn) memcpy(sel, sel_override, (size_t)Cn * mtr);
        else {
            const int e = tspws_b200_jackknife_plans(sel, in->time, mtr, p->jackknife_d, p->jackknife_n, Cn);
            if (e) {   /* the reference goes on with an uninitialised matrix here */
                printf("Warning: jackknife needs the start time of every trace (binary input); skipped.\n");
                goto done;
            }
        }
        if (two_stage) rc = two_stage_rows(p, &F, d_x, mtr, max, sel, Cn, p->Kmax, out->ls_subsmpl, out->tsPWS_subsmpl,
                                           out->mtr_subsmpl, NULL, NULL, NULL, stream);
        else           rc = single_stage_subsample(p, &F, d_x, mtr, max, sel, Cn, 0, out->mtr_subsmpl, out->ls_subsmpl,
                                                   out->tsPWS_subsmpl, stream);
    }
done:
    free(sel);
#undef F
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* convergence curves (src/ts_pws1f_lib.c:247-314)                                        */
/* ------------------------------------------------------------------------------------ */
int tspws_convergence_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                          unsigned int mtr, int max, cudaStream_t st)
{
    int rc = 0;
    const size_t N = (size_t)max;
    const int B = RS_BATCH, ntk = (max + 255) / 256;
    rs_frames_t F;
    float *d_ref = NULL, *d_ts = NULL, *d_steps = NULL, *h_buf = NULL;
    void *d_ypart = NULL, *d_Y = NULL, *d_P = NULL, *d_zout = NULL, *d_scr = NULL, *d_sst = NULL, *d_sps = NULL;
    double *d_sim = NULL, *h_sim = NULL, *d_part = NULL;
    tspws_weight_t *d_rules = NULL, *h_rules = NULL;
    unsigned int *d_tk = NULL;
    char *sel = NULL;
    float **rows = NULL;
    if (!out->tsPWS_sim || !out->tsPWS_misfit || !out->ls_sim || !out->ls_misfit) {
        printf("tspws_main: NULL convergence outputs\n");
        return -1;
    }
    if ((rc = frames_open(&F, p, max, 1))) { frames_close(&F); return rc; }
    const size_t ncoef = F.fr->ncoef ? F.fr->ncoef : 1;
    /* prefixes 1..n_single are plain (single-stage) stacks; longer ones are regrouped two-stage stacks (:262-269) */
    const unsigned int n_single = p->Kmax ? (p->Kmax < mtr ? p->Kmax : mtr) : mtr;

    h_rules = (tspws_weight_t *)malloc(B * sizeof(tspws_weight_t));
    h_sim   = (double *)malloc(2 * B * sizeof(double));
    h_buf   = (float *)malloc((size_t)B * N * sizeof(float));
    if (!h_rules || !h_sim || !h_buf) { rc = TSPWS_ERR_ALLOC; goto done; }
    CKR(cudaMalloc((void **)&d_ref, 2 * N * sizeof(float)));
    CKR(cudaMalloc((void **)&d_ts, (size_t)B * N * sizeof(float)));
    CKR(cudaMalloc(&d_ypart, (size_t)B * (F.pt.ypart_len ? F.pt.ypart_len : 1) * 2 * sizeof(float)));
    CKR(cudaMalloc(&d_Y, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(cudaMalloc(&d_P, (size_t)B * ncoef * 2 * sizeof(float)));
    CKR(cudaMalloc(&d_zout, (size_t)B * ncoef * 4 * sizeof(float)));
    CKR(cudaMalloc(&d_scr, (size_t)2 * (F.acc.inv_slots ? F.acc.inv_slots : 1) * B * N * sizeof(float)));
    CKR(cudaMalloc(&d_sst, ncoef * 2 * sizeof(double)));
    CKR(cudaMalloc(&d_sps, ncoef * 2 * sizeof(double)));
    CKR(cudaMalloc((void **)&d_sim, 2 * B * sizeof(double)));
    CKR(cudaMalloc((void **)&d_rules, B * sizeof(tspws_weight_t)));
    CKR(cudaMalloc((void **)&d_tk, (size_t)B * ntk * sizeof(unsigned int)));
    CKR(cudaMemsetAsync(d_tk, 0, (size_t)B * ntk * sizeof(unsigned int), st));
    CKR(cudaMemsetAsync(d_sst, 0, ncoef * 2 * sizeof(double), st));
    CKR(cudaMemsetAsync(d_sps, 0, ncoef * 2 * sizeof(double), st));
    /* reference traces of the two curves: the final stacks, or the caller's alternative (:277,:298) */
    CKR(cudaMemcpyAsync(d_ref, in->reference ? in->reference : out->tsPWS, N * sizeof(float), cudaMemcpyHostToDevice, st));
    CKR(cudaMemcpyAsync(d_ref + N, in->reference ? in->reference : out->ls, N * sizeof(float), cudaMemcpyHostToDevice, st));

    /* ---- ts-PWS convergence, plain prefixes: running sums of Y and phasors ---- */
    for (unsigned int t0 = 0; t0 < n_single; t0 += B) {
        const int nb = (int)((n_single - t0 < (unsigned int)B) ? n_single - t0 : (unsigned int)B);
        for (int b = 0; b < nb; b++) h_rules[b] = tspws_weight_rule(p->wu, p->unbiased, t0 + b + 1, t0 + b + 1);
        CKR(cudaMemcpyAsync(d_rules, h_rules, nb * sizeof(tspws_weight_t), cudaMemcpyHostToDevice, st));
        CKK(tspws_cu_forward_pertrace(&F.pt, d_x + (size_t)t0 * N, nb, 0, d_ypart, st));
        CKK(tspws_cu_coef_matrix(&F.pt, d_ypart, nb, d_Y, d_P, st));
        CKK(tspws_cu_prefix_weight(d_Y, d_P, (int)F.fr->ncoef, nb, d_sst, d_sps, d_rules, d_zout, st));
        CKK(tspws_cu_inverse(&F.acc, d_zout, NULL, 0, nb, d_rules, d_scr, d_tk, d_ts, NULL, st, NULL, NULL, NULL));
        CKK(tspws_cu_simil(d_ts, nb, max, d_ref, d_sim, d_sim + B, st));
        CKR(cudaMemcpyAsync(h_sim, d_sim, 2 * B * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (out->tsPWS_steps) CKR(cudaMemcpyAsync(out->tsPWS_steps + (size_t)t0 * N, d_ts, (size_t)nb * N * sizeof(float), cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        for (int b = 0; b < nb; b++) { out->tsPWS_sim[t0 + b] = h_sim[b]; out->tsPWS_misfit[t0 + b] = h_sim[B + b]; }
    }
    /* ---- longer prefixes: two-stage stacks of the first Tr traces, regrouped for every Tr ---- */
    if (n_single < mtr) {
        const unsigned int nres = mtr - n_single;
        sel  = (char *)malloc((size_t)B * mtr);
        rows = (float **)malloc(nres * sizeof(float *));
        if (!sel || !rows) { rc = TSPWS_ERR_ALLOC; goto done; }
        for (unsigned int i = 0; i < nres; i++) rows[i] = out->tsPWS_steps ? out->tsPWS_steps + (size_t)(n_single + i) * N : NULL;
        for (unsigned int r0 = 0; r0 < nres; r0 += B) {
            const unsigned int nb = (nres - r0 < (unsigned int)B) ? nres - r0 : (unsigned int)B;
            for (unsigned int b = 0; b < nb; b++) {
                const unsigned int Tr = n_single + r0 + b + 1;
                memset(sel + (size_t)b * mtr, 1, Tr);
                memset(sel + (size_t)b * mtr + Tr, 0, mtr - Tr);
            }
            rc = two_stage_rows(p, &F, d_x, mtr, max, sel, nb, p->Kmax, NULL, rows + r0, NULL, d_ref,
                                out->tsPWS_sim + n_single + r0, out->tsPWS_misfit + n_single + r0, st);
            if (rc) goto done;
        }
    }
    /* ---- linear-stack convergence in the time domain (:288-308) ---- */
    {
        const int nblk = (max + 255) / 256;
        double *d_ls_sim = NULL;
        CKR(cudaMalloc((void **)&d_part, (size_t)mtr * nblk * 4 * sizeof(double)));
        CKR(cudaMalloc((void **)&d_ls_sim, 2 * (size_t)mtr * sizeof(double)));
        if (out->ls_steps) CKR(cudaMalloc((void **)&d_steps, (size_t)mtr * N * sizeof(float)));
        if (tspws_cu_lsconv(d_x, (int)mtr, max, d_ref + N, d_steps, d_part, d_ls_sim, d_ls_sim + mtr, st)) { cudaFree(d_ls_sim); rc = TSPWS_ERR_CUDA; goto done; }
        cudaMemcpyAsync(out->ls_sim, d_ls_sim, mtr * sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out->ls_misfit, d_ls_sim + mtr, mtr * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (out->ls_steps) cudaMemcpyAsync(out->ls_steps, d_steps, (size_t)mtr * N * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(d_ls_sim);
        if (e != cudaSuccess) { rc = TSPWS_ERR_CUDA; goto done; }
    }
done:
    free(h_rules); free(h_sim); free(h_buf); free(sel); free(rows);
    cudaFree(d_ref); cudaFree(d_ts); cudaFree(d_steps); cudaFree(d_ypart); cudaFree(d_Y); cudaFree(d_P); cudaFree(d_zout);
    cudaFree(d_scr); cudaFree(d_sst); cudaFree(d_sps); cudaFree(d_sim); cudaFree(d_rules); cudaFree(d_tk); cudaFree(d_part);
    frames_close(&F);
    return rc;
}
