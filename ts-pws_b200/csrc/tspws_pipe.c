/*
 * Software pipeline over a stream of independent stacks (device-resident traces).
 *
 * One two-stage stack is an HBM-bound pass (the group sums: every trace sample read once) followed by a tail of short,
 * latency-bound kernels (transforms of the Kmax group sums, phase stack, inverse) that leave HBM idle.  A single stack
 * cannot overlap the two -- the tail needs all group sums -- but a stream of stacks (one per station pair / component /
 * resample, the way src/ts_pws1f.c is run over a data set) can: all group-sum passes go back to back on one stream and
 * the tails follow on a second, higher-priority stream, `depth` plans serving as buffers:
 *
 *     H:  sums(0) sums(1) sums(2) sums(3) ...
 *     T:          tail(0) tail(1) tail(2) ...
 *
 * The sums are one direct launch (their input pointer changes from stack to stack); the tail is replayed as a CUDA
 * graph that writes the two stacks into buffers of the pipe, copied to the caller's pointers right behind it (2 x N
 * floats, device to device) -- so no graph is ever re-captured because a pointer changed.  Results of stack i are
 * complete after tspws_b200_pipe_sync (or once the tail stream passes the submit, for callers that order their own
 * work after tspws_b200_pipe_stream(p, 1)).
 * Every stack is computed exactly as tspws_b200_plan_stack computes it (same kernels, same order within the stack).
 *
 * Whether the overlap pays depends on the frame: the tail kernels slow down beside the streaming kernel (they lose L2
 * and DRAM pages to it) and slow it down in turn.  Measured at 8192 x 16384: Morlet, 44 scales 141 -> 115-120 us per
 * stack; Mexican hat, 20 scales (long inverse filters at the fine scales) 133 -> 192 us (inv2 / inv4 take 70-120 us
 * instead of 10-16, colsum2 179 instead of 84).  depth = 1 runs one stack at a time in the plan_stack configuration
 * through the same calls; depth = 0 measures both ways on the first 16 stacks and keeps the faster (AUTO, below).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include "tspws_b200.h"
#include "tspws_internal.h"

/* depth 0 at creation = AUTO: the first AUTO_W stacks run pipelined, the next AUTO_W one at a time (on an extra plan in the
 * plan_stack configuration), both timed with events on the device; whichever was faster serves the rest.  Nothing blocks:
 * the decision is taken at the first submit that finds the second measurement complete (one at a time until then). */
#define AUTO_W 8
enum { MODE_FIXED = 0, MODE_CAL_A, MODE_CAL_B, MODE_CAL_WAIT, MODE_PIPE, MODE_SOLO };

struct tspws_pipe {
    int depth, nograph;
    int nplans;                            /* depth, + 2 in AUTO mode: two plans in the plan_stack configuration, alternating, so that
                                            * a one-at-a-time stack need not wait for the result copies of its predecessor */
    unsigned long long nsolo;
    int mode, cnt;                         /* AUTO state machine: stacks submitted in the current phase */
    cudaEvent_t cal[4];                    /* pipelined phase begin / end, solo phase begin / end */
    unsigned char *used;                   /* [nplans] the slot has carried a stack before */
    float auto_ms[2];                      /* measured ms per stack: pipelined, one at a time (0 until decided) */
    tspws_plan_t **pl;
    cudaStream_t H, T, C;                  /* sums, tails, result copies */
    cudaEvent_t *red, *done, *fin;         /* sums done, tail done, results delivered (plan free again) */
    cudaGraphExec_t *g_fin;
    unsigned long *k_fin;                  /* kernels per replay (launch accounting) */
    float **out;                           /* [depth] device buffers [2][N]: ls | tsPWS of the plan's last stack */
    size_t N;
    unsigned long long n;
};

void tspws_b200_pipe_destroy(tspws_pipe_t *pp)
{
    if (!pp) return;
    if (pp->H) cudaStreamSynchronize(pp->H);
    if (pp->T) cudaStreamSynchronize(pp->T);
    if (pp->C) cudaStreamSynchronize(pp->C);
    for (int i = 0; i < 4; i++) if (pp->cal[i]) cudaEventDestroy(pp->cal[i]);
    for (int k = 0; k < pp->nplans; k++) {
        if (pp->g_fin && pp->g_fin[k]) cudaGraphExecDestroy(pp->g_fin[k]);
        if (pp->red && pp->red[k]) cudaEventDestroy(pp->red[k]);
        if (pp->done && pp->done[k]) cudaEventDestroy(pp->done[k]);
        if (pp->fin && pp->fin[k]) cudaEventDestroy(pp->fin[k]);
        if (pp->pl && pp->pl[k]) tspws_b200_plan_destroy(pp->pl[k]);
        if (pp->out && pp->out[k]) cudaFree(pp->out[k]);
    }
    if (pp->H) cudaStreamDestroy(pp->H);
    if (pp->T) cudaStreamDestroy(pp->T);
    if (pp->C) cudaStreamDestroy(pp->C);
    free(pp->used); free(pp->pl); free(pp->red); free(pp->done); free(pp->fin); free(pp->g_fin); free(pp->k_fin); free(pp->out);
    free(pp);
}

tspws_pipe_t *tspws_b200_pipe_create(const t_tsPWS *p, int max, unsigned int mtr_local, unsigned int mtr_total,
                                     unsigned int trace_offset, int depth)
{
    if (!p || depth < 0 || depth > 8) return NULL;
    /* AUTO needs every rank of a sharded stack to take the same decision: only for unsharded stacks */
    if (depth == 0 && mtr_local != mtr_total) return NULL;
    tspws_pipe_t *pp = (tspws_pipe_t *)calloc(1, sizeof *pp);
    if (!pp) return NULL;
    pp->mode = depth == 0 ? MODE_CAL_A : MODE_FIXED;
    if (depth == 0) depth = 4;
    pp->depth = depth;
    pp->nplans = pp->mode == MODE_FIXED ? depth : depth + 2;
    const int np = pp->nplans;
    { const char *e = getenv("TSPWS_NO_GRAPH"); pp->nograph = e && atoi(e); }
    pp->pl = calloc(np, sizeof *pp->pl); pp->red = calloc(np, sizeof *pp->red); pp->done = calloc(np, sizeof *pp->done); pp->fin = calloc(np, sizeof *pp->fin);
    pp->g_fin = calloc(np, sizeof *pp->g_fin); pp->k_fin = calloc(np, sizeof *pp->k_fin);
    pp->out = calloc(np, sizeof *pp->out); pp->used = calloc(np, 1);
    pp->N = (size_t)max;
    if (!pp->pl || !pp->red || !pp->done || !pp->fin || !pp->g_fin || !pp->k_fin || !pp->out || !pp->used) goto fail;
    if (pp->mode != MODE_FIXED)
        for (int i = 0; i < 4; i++) if (cudaEventCreate(&pp->cal[i]) != cudaSuccess) goto fail;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    /* Tails at the higher priority.  Measured on the headline stack with the per-CTA timeline (tools/pipe_timeline.py,
     * 8192 x 16384, depth 4): tails high 118 us per stack; equal priorities 118-126; sums high 161 (the next group-sum
     * kernel then starts only after the whole tail: fully serial); two alternating tail streams 125-130.  What is left
     * on the table: at every group-sum boundary the ~3400 queued CTAs of the transform kernel take the freed slots and
     * the next group-sum kernel waits ~25 us for that backlog to drain (HBM idle meanwhile).  TSPWS_PIPE_PRIO=sums|equal. */
    { const char *e = getenv("TSPWS_PIPE_PRIO");
      if (e && !strcmp(e, "sums")) { const int t = lo; lo = hi; hi = t; }
      if (e && !strcmp(e, "equal")) hi = lo; }
    if (cudaStreamCreateWithPriority(&pp->H, cudaStreamNonBlocking, lo) != cudaSuccess) goto fail;
    if (cudaStreamCreateWithPriority(&pp->T, cudaStreamNonBlocking, hi) != cudaSuccess) goto fail;
    if (cudaStreamCreateWithPriority(&pp->C, cudaStreamNonBlocking, hi) != cudaSuccess) goto fail;
    for (int k = 0; k < np; k++) {
        /* 2 group-sum CTAs per SM: measured on the headline stack, depth 3: 114.8 us per stack with 296 CTAs, 122.5
         * with 370, 121.7 with 222, no overlap at all (148-152 us) with 444 or 592 -- the tail needs room beside them */
        if (depth > 1 && k < depth) tspws_plan_set_cs2_hint(148 * 2);   /* depth 1 / the solo plan: the plan_stack configuration */
        pp->pl[k] = tspws_b200_plan_create(p, max, mtr_local, mtr_total, trace_offset);
        tspws_plan_set_cs2_hint(0);
        if (!pp->pl[k]) goto fail;
        if (cudaMalloc((void **)&pp->out[k], 2 * pp->N * sizeof(float)) != cudaSuccess) goto fail;
        if (cudaEventCreateWithFlags(&pp->red[k], cudaEventDisableTiming) != cudaSuccess) goto fail;
        if (cudaEventCreateWithFlags(&pp->done[k], cudaEventDisableTiming) != cudaSuccess) goto fail;
        if (cudaEventCreateWithFlags(&pp->fin[k], cudaEventDisableTiming) != cudaSuccess) goto fail;
    }
    return pp;
fail:
    tspws_b200_pipe_destroy(pp);
    return NULL;
}

/* the tail of plan k's stack on stream s: replayed graph (captured on first use) writing into pp->out[k] */
static int tail(tspws_pipe_t *pp, int k, cudaStream_t s)
{
    float *ls = pp->out[k], *ts = pp->out[k] + pp->N;
    if (pp->nograph || tspws_cu_profile_enabled()) return tspws_b200_plan_finish(pp->pl[k], ls, ts, s);
    if (!pp->g_fin[k]) {
        cudaGraph_t g = NULL;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); pp->nograph = 1; return tail(pp, k, s); }
        const unsigned long n0 = tspws_cu_launch_count();
        const int e = tspws_b200_plan_finish(pp->pl[k], ls, ts, s);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        pp->k_fin[k] = tspws_cu_launch_count() - n0;
        tspws_cu_launch_count_add(0ul - pp->k_fin[k]);              /* nothing has run yet */
        if (e || ce != cudaSuccess || !g || cudaGraphInstantiate(&pp->g_fin[k], g, 0) != cudaSuccess) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            pp->g_fin[k] = NULL;
            return e ? e : TSPWS_ERR_CUDA;
        }
        cudaGraphDestroy(g);
    }
    if (cudaGraphLaunch(pp->g_fin[k], s) != cudaSuccess) { cudaGetLastError(); return TSPWS_ERR_CUDA; }
    tspws_cu_launch_count_add(pp->k_fin[k]);
    return 0;
}

int tspws_b200_pipe_submit(tspws_pipe_t *pp, float *d_sigall, float *d_ls, float *d_tspws)
{
    if (!pp || !d_sigall || !d_ls || !d_tspws) return TSPWS_ERR_NULL;
    int e;
    if (pp->mode == MODE_CAL_WAIT && cudaEventQuery(pp->cal[3]) == cudaSuccess) {
        if (cudaEventElapsedTime(&pp->auto_ms[0], pp->cal[0], pp->cal[1]) != cudaSuccess ||
            cudaEventElapsedTime(&pp->auto_ms[1], pp->cal[2], pp->cal[3]) != cudaSuccess) { cudaGetLastError(); pp->auto_ms[0] = 1.f; pp->auto_ms[1] = 0.f; }
        pp->auto_ms[0] /= AUTO_W; pp->auto_ms[1] /= AUTO_W;
        pp->mode = pp->auto_ms[0] < pp->auto_ms[1] ? MODE_PIPE : MODE_SOLO;
    } else if (pp->mode == MODE_CAL_WAIT) {
        cudaGetLastError();                                        /* cudaErrorNotReady is not an error */
    }
    const int solo = pp->mode == MODE_CAL_B || pp->mode == MODE_CAL_WAIT || pp->mode == MODE_SOLO;
    const int k = solo ? pp->depth + (int)(pp->nsolo & 1) : (int)(pp->n % (unsigned long long)pp->depth);
    if (pp->mode == MODE_CAL_A && pp->cnt == 0 && cudaEventRecord(pp->cal[0], pp->H) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (pp->mode == MODE_CAL_B && pp->cnt == 0 &&                  /* the solo phase starts when the pipelined one has drained */
        (cudaStreamWaitEvent(pp->H, pp->cal[1], 0) != cudaSuccess || cudaEventRecord(pp->cal[2], pp->H) != cudaSuccess)) return TSPWS_ERR_CUDA;
    /* the plan's buffers are free again once the results of the stack that used them last have been delivered */
    if (pp->used[k] && cudaStreamWaitEvent(pp->H, pp->fin[k], 0) != cudaSuccess) return TSPWS_ERR_CUDA;
    pp->used[k] = 1;
    if ((e = tspws_b200_plan_reduce(pp->pl[k], d_sigall, pp->H))) return e;
    cudaStream_t T = (pp->depth > 1 && !solo) ? pp->T : pp->H;     /* one at a time: sums and tail in stream order, no overlap */
    if (cudaEventRecord(pp->red[k], pp->H) != cudaSuccess || cudaStreamWaitEvent(T, pp->red[k], 0) != cudaSuccess) return TSPWS_ERR_CUDA;
    if ((e = tail(pp, k, T))) return e;
    /* the two result copies leave the tail stream (it is the critical path of the pipeline): they run on a third one */
    if (cudaEventRecord(pp->done[k], T) != cudaSuccess || cudaStreamWaitEvent(pp->C, pp->done[k], 0) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (cudaMemcpyAsync(d_ls, pp->out[k], pp->N * sizeof(float), cudaMemcpyDeviceToDevice, pp->C) != cudaSuccess ||
        cudaMemcpyAsync(d_tspws, pp->out[k] + pp->N, pp->N * sizeof(float), cudaMemcpyDeviceToDevice, pp->C) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (cudaEventRecord(pp->fin[k], pp->C) != cudaSuccess) return TSPWS_ERR_CUDA;
    if (solo) pp->nsolo++; else pp->n++;
    if (pp->mode == MODE_CAL_A && ++pp->cnt == AUTO_W) {
        if (cudaEventRecord(pp->cal[1], pp->C) != cudaSuccess) return TSPWS_ERR_CUDA;
        pp->mode = MODE_CAL_B; pp->cnt = 0;
    } else if (pp->mode == MODE_CAL_B && ++pp->cnt == AUTO_W) {
        if (cudaEventRecord(pp->cal[3], pp->C) != cudaSuccess) return TSPWS_ERR_CUDA;
        pp->mode = MODE_CAL_WAIT;
    }
    return 0;
}

/* AUTO pipes: 0 = still measuring, 1 = pipelined, 2 = one at a time; ms[0..1] = the two measurements per stack */
int tspws_b200_pipe_mode(tspws_pipe_t *pp, float *ms)
{
    if (!pp) return TSPWS_ERR_NULL;
    if (ms) { ms[0] = pp->auto_ms[0]; ms[1] = pp->auto_ms[1]; }
    if (pp->mode == MODE_FIXED) return pp->depth > 1 ? 1 : 2;
    return pp->mode == MODE_PIPE ? 1 : pp->mode == MODE_SOLO ? 2 : 0;
}

int tspws_b200_pipe_sync(tspws_pipe_t *pp)
{
    if (!pp) return TSPWS_ERR_NULL;
    if (cudaStreamSynchronize(pp->H) != cudaSuccess || cudaStreamSynchronize(pp->T) != cudaSuccess ||
        cudaStreamSynchronize(pp->C) != cudaSuccess) return TSPWS_ERR_CUDA;
    return 0;
}

/* 0: the stream of the sums (work may be ordered BEFORE a submit through it); 1: the stream the results arrive on */
void *tspws_b200_pipe_stream(tspws_pipe_t *pp, int which) { return pp ? (void *)(which ? pp->C : pp->H) : NULL; }
tspws_plan_t *tspws_b200_pipe_plan(tspws_pipe_t *pp, int k) { return (pp && k >= 0 && k < pp->nplans) ? pp->pl[k] : NULL; }
