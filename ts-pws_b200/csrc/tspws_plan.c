/*
 * tspws_plan.c -- device-resident stacking plans (C host code over the CUDA runtime API).
 *
 * A plan owns the frame (filter bank) on the device, the work lists of the kernels and every
 * workspace, so that a stack is a fixed sequence of kernel launches on one stream with no
 * allocation, no host round trip and no synchronisation:
 *
 *   two-stage (Kmax groups, reference src/ts_pws1f_lib.c:212-222):
 *     reduce : colsum (last-block combine)                      => exchange = fp64 group sums
 *     finish : group_post(mean, fold, fp32) -> forward(Kmax traces) -> reduce+weight -> inverse
 *   single-stage (reference src/ts_pws1f_lib.c:209-211):
 *     reduce : prep(fold, rm, in place) -> forward(all traces) -> reduce+weight
 *                                                                => exchange = PS | ST sums
 *     finish : [weight again if ranks were summed] -> inverse
 *
 * Between reduce and finish the exchange block may be all-reduced (sum) across ranks that hold
 * disjoint trace shards of the same stack.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "tspws_internal.h"

struct tspws_plan {
    t_tsPWS      p;
    int          N;
    unsigned int mtr_local, mtr_total, offset;
    int          two_stage;
    unsigned int K;                 /* divisor of the phase stack: Kmax or mtr_total */
    int          acc64;             /* transforms in fp64                            */
    size_t       csz;               /* sizeof(complex of the transform type)         */
    tspws_frame_t    *fr;
    tspws_devframe_t  dev;
    tspws_weight_t    rule, *d_rule;
    /* two-stage */
    int    ncs, ntile;
    tspws_csum_item_t *d_cs;
    int    *d_slot_off, *d_grp_t0;
    /* single-wave group sums (colsum2_kernel): row parts, their slots, slot totals */
    int     cs2, cs2_nparts, cs2_nslots, cs2_pending;
    double *d_gfin;                 /* [K][N] post-processed group sums when the input is the exchange block */
    int    *d_ident;                /* [K+1] 0..K: the exchange block seen as one slot per group */
    void   *d_cs2_parts;
    int    *d_cs2_slot_off;
    double *d_cs2_slots, *d_cs2_tot;
    double *d_part, *d_gsum, *d_gtile;      /* gsum [K][N] and gtile [K][ntile] are one block */
    float  *d_gsum32;
    unsigned int *d_tickets, *d_inv_tickets;
    /* forward / weights / inverse */
    int    chunk, nchunk, ntr_fwd;
    void  *d_pps, *d_pst, *d_exch, *d_zout, *d_invscratch, *d_ypart;
    int    summed_across_ranks;
    int    exch_bound;              /* exchange block belongs to the caller */
    /* the whole stack as one CUDA graph (captured on first use for a given set of pointers) */
    cudaGraphExec_t gexec;
    const void *g_x, *g_ls, *g_ts;
    unsigned long g_kernels;
    int    graph_broken;
    /* side stream: independent pieces of the tail run concurrently (fork / join, also inside the graph) */
    cudaStream_t side;
    cudaEvent_t  ev_fork, ev_join;
    int          mean_split;        /* fp64 transform of the group mean + fp32 transforms of the residuals */
    int          hybrid;            /* TSPWS_F64_MAXL set: long filters transform the unsplit sums instead */
    /* coarse scales of the single-stage forward transform on the tensor cores (tspws_cu_forward_dense) */
    int     dense_on, dense_s0, dense_ci0, dense_ncd, dense_items;
    /* middle octaves as banded GEMMs (tspws_cu_forward_band); td_*: the whole tensor-core range = band + dense scales */
    int     band_n, band_groups, td_s0, td_ci0, td_ncd, simt_lg_end;
    void   *d_bimg; tspws_band_t *d_bands;
    void   *d_wimg, *d_ximg, *d_ysum_ps, *d_ysum_st;
    float  *d_tsum, *d_sumtrace, *d_dpart;
    unsigned int *d_tile_tickets;   /* [ntile16]: fused phasor combine of fwd_tail_kernel (NULL: separate combine launch) */
    double      *d_gmean;           /* [3N]: the group mean, three periods */
    void        *d_ymean;           /* double2 [TSPWS_F64_MAXSPLIT][ncoef] */
    int         *h_slot_off;        /* [K+1] first work item of every group */
    void  *own_gsum, *own_exch;
    /* peer exchange over CUDA IPC (two-stage, sharded): see tspws_b200_plan_peer_export / _import */
    int           peer_on, peer_rank, peer_world, peer_pending, peer_ipc;
    void         *d_peer_region;    /* own: 2 exchange blocks (step parity) + flag words + error word */
    size_t        peer_flag_off, peer_bytes;
    void        **h_peer_ptr;       /* [world] mapped regions (own entry = d_peer_region) */
    void         *d_peer_blocks;    /* device array [world] of const double * */
    void         *d_peer_flags;     /* device array [world] of unsigned int * (every rank's flag words) */
    unsigned int *d_peer_step;
};

#define TSPWS_PEER_MAXW 64          /* flag words per rank region (ranks of one node) */

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "tspws_b200: CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); \
    goto fail; } } while (0)

/* CTA count of the group-sum kernel for plans created by this thread (0: default); see tspws_pipe.c */
static __thread int g_cs2_ctas_hint = 0;
void tspws_plan_set_cs2_hint(int ctas) { g_cs2_ctas_hint = ctas; }

static long env_long(const char *name, long dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atol(v) : dflt;
}

static int pow2_floor(int v) { int r = 1; while (2 * r <= v) r *= 2; return r; }

/* ---- device frame: scale records, filters (fp32 [+fp64]), forward items, inverse groups ---- */
int tspws_devframe_build(tspws_devframe_t *d, const tspws_frame_t *f, int want64, int per_trace)
{
    memset(d, 0, sizeof(*d));
    const int S = (int)f->S;
    d->S = S; d->N = (int)f->N; d->ncoef = (int)f->ncoef; d->ntaps = (int)f->ntaps;
    tspws_scale_t    *sc    = (tspws_scale_t *)calloc(S ? S : 1, sizeof(*sc));
    float            *w32   = (float *)malloc((f->ntaps ? f->ntaps : 1) * 2 * sizeof(float));
    tspws_fwd_item_t *items = NULL, *items_gen = NULL;
    tspws_fwd2_item_t *items2 = NULL;
    tspws_inv_group_t *grp  = NULL;
    float *g32 = NULL;
    int nitems = 0, nitems_gen = 0, nitems2 = 0, ok = 0, ntile64 = 0;
    int *texp = NULL;           /* [ntile64][2]: fp32 items (per trace) and fp64 items that write into a tile */
    size_t ng = 0;
    if (!sc || !w32) goto done;

    const long work = env_long("TSPWS_FWD_WORK", 4096);    /* taps x coefficients per warp item */
    for (int pass = 0; pass < 2; pass++) {
        nitems = 0;
        for (int s = 0; s < S; s++) {
            const int L = (int)f->L[s], Ns = (int)f->Ns[s];
            int G = pow2_floor(L / 8 > 0 ? L / 8 : 1);
            if (G > 32) G = 32;
            const int opp = 32 / G;
            long run = work / L;
            run -= run % opp;
            if (run < opp) run = opp;
            if (!pass) {
                sc[s].L = L; sc[s].c = f->center[s]; sc[s].D = (int)f->D[s]; sc[s].Ns = Ns;
                sc[s].coef_off = (int)f->coef_off[s]; sc[s].tap_off = (int)f->tap_off[s];
                sc[s].G = G;
                sc[s].gain = log(2.) / (2 * f->Cpsi * f->V * f->scale[s]) * (double)f->D[s];
            }
            if (pass) d->h_item_first[s] = nitems;
            for (int m0 = 0; m0 < Ns; m0 += (int)run) {
                if (pass) {
                    items[nitems].s = s; items[nitems].m0 = m0;
                    items[nitems].nout = (Ns - m0 < run) ? Ns - m0 : (int)run;
                }
                nitems++;
            }
        }
        if (!pass) {
            items = (tspws_fwd_item_t *)calloc(nitems ? nitems : 1, sizeof(*items));
            d->h_item_first = (int *)calloc((size_t)S + 1, sizeof(int));
            if (!items || !d->h_item_first) goto done;
        } else d->h_item_first[S] = nitems;
    }
    for (size_t i = 0; i < 2 * f->ntaps; i++) w32[i] = (float)f->w[i];
    /* 64-coefficient tiles of all scales, numbered scale by scale: the units of the fused phasor combine */
    for (int s = 0; s < S; s++) { sc[s].tile0 = ntile64; ntile64 += (sc[s].Ns + 63) / 64; }
    texp = (int *)calloc((size_t)(ntile64 ? ntile64 : 1) * 2, sizeof(int));
    if (!texp) goto done;

    /* ---- polyphase (register-tiled) kernel: which scales it covers, padded filters, items ---- */
    {
        const int N = (int)f->N;
        const int pbmax = (int)env_long("TSPWS_FW_PB", 4);       /* blocks of 32 phases per per-trace item */
        const int nofast = (int)env_long("TSPWS_NO_FAST", 0);
        d->per_trace = per_trace;
        d->fw_smem_per_warp = 0;
        d->max_Ns = 0;
        int ypart = 0;
        for (int s = 0; s < S; s++) {
            const int L = sc[s].L, c = sc[s].c, D = sc[s].D, Ns = sc[s].Ns;
            const int e = (D - c % D) % D;
            sc[s].A  = (c + e) / D;
            sc[s].Qs = (((e + L + D - 1) / D) + 3) & ~3;            /* padded to a multiple of 4 with zero taps */
            sc[s].lg = -1;
            sc[s].nsplit = 1;
            if (Ns > d->max_Ns) d->max_Ns = Ns;
            if (!nofast && D <= N) {
                if (D >= 32) {
                    sc[s].lg = 5;
                    const int need5 = 2 * (sc[s].Qs + 16) * 32;      /* double-buffered window rows: the Qs + 16 the taps reach */
                    if (need5 > d->fw_smem_per_warp) d->fw_smem_per_warp = need5;
                }
                else if ((D & (D - 1)) == 0) {
                    int lg = 0; while ((1 << lg) < D) lg++;
                    const int tiles = 32 >> lg;
                    const long W = ((long)tiles * 16 + sc[s].Qs) * D;        /* what the taps reach */
                    long need = W + (W / (16 * D) + 1) * D + 8;          /* skewed window + copy overrun */
                    need = 2 * ((need + 3) & ~3L);                       /* double buffered, 16 B aligned */
                    if (W <= N && need * 4 * 4 * 3 <= 200 * 1024) {
                        sc[s].lg = lg;
                        if (need > d->fw_smem_per_warp) d->fw_smem_per_warp = (int)need;
                    }
                }
            }
            if (sc[s].lg >= 0) {
                const int gpass = (sc[s].Qs + 8) * (sc[s].lg == 5 ? 32 : D);    /* taps of one pass */
                if (gpass > d->fw_gmax) d->fw_gmax = gpass;
                const int gs = (sc[s].lg == 5) ? ((D + 31) & ~31) : D;     /* phases padded to 32 lanes */
                sc[s].g_off = (int)ng;
                ng += (size_t)(sc[s].Qs + 8) * gs;                       /* + 8 zero taps: unguarded prefetch */
            }
            if (sc[s].lg == 5 && per_trace) {
                const int nblk = (D + 31) / 32;
                int pb = pbmax;
                if ((nblk + pb - 1) / pb > 8) pb = (nblk + 7) / 8;      /* at most 8 partial slots per scale */
                sc[s].nsplit = (nblk + pb - 1) / pb;
            }
            sc[s].ypart_off = ypart;
            ypart += sc[s].nsplit * Ns;
        }
        d->ypart_len = ypart;
        g32 = (float *)calloc((ng ? ng : 1) * 2, sizeof(float));
        if (!g32) goto done;
        for (int s = 0; s < S; s++) {
            if (sc[s].lg < 0) continue;
            const int e = sc[s].A * sc[s].D - sc[s].c;
            const double *w = f->w + 2 * f->tap_off[s];
            float *g = g32 + 2 * (size_t)sc[s].g_off;
            const int Ds = sc[s].D, gs = (sc[s].lg == 5) ? ((Ds + 31) & ~31) : Ds;
            for (int l = 0; l < sc[s].L; l++) {           /* g[q][rho] = conj(w[q D + rho - e]) */
                const int t = e + l, q = t / Ds, rho = t % Ds;
                g[2 * ((size_t)q * gs + rho)]     = (float)w[2 * l];
                g[2 * ((size_t)q * gs + rho) + 1] = (float)(-w[2 * l + 1]);
            }
        }
        for (int pass = 0; pass < 2; pass++) {
            nitems2 = 0; nitems_gen = 0;
            for (int lg = 0; lg <= 5; lg++) {
                d->fw_lg_begin[lg] = nitems2;
                for (int si = 0; si < S; si++) {
                    const int s = S - 1 - si;                  /* coarsest (most phase blocks) first */
                    if (sc[s].lg != lg) continue;
                    const int per_item = (lg < 5) ? (512 >> lg) : 16;
                    const int nblk = (lg < 5) ? 1 : (sc[s].D + 31) / 32;
                    const int nsp = sc[s].nsplit;
                    for (int m0 = 0; m0 < sc[s].Ns; m0 += per_item)
                        for (int sp = 0; sp < nsp; sp++) {
                            if (pass) {
                                tspws_fwd2_item_t *it = &items2[nitems2];
                                it->s = s; it->m0 = m0; it->split = sp;
                                it->pb0 = (lg < 5) ? 0 : (int)((long)nblk * sp / nsp);
                                it->pb1 = (lg < 5) ? 1 : (int)((long)nblk * (sp + 1) / nsp);
                                long long b0 = ((long long)(m0 - sc[s].A) * sc[s].D) % N;
                                it->base = (int)(b0 < 0 ? b0 + N : b0);
                                const int mend = (m0 + per_item < sc[s].Ns ? m0 + per_item : sc[s].Ns) - 1;
                                for (int t = m0 / 64; t <= mend / 64; t++) texp[2 * (sc[s].tile0 + t)]++;
                            }
                            nitems2++;
                        }
                }
            }
            d->fw_lg_begin[6] = nitems2;
            for (int s = 0; s < S; s++) {
                if (sc[s].lg >= 0) continue;
                const int L = sc[s].L, Ns = sc[s].Ns, opp = 32 / sc[s].G;
                long run = work / L; run -= run % opp; if (run < opp) run = opp;
                for (int m0 = 0; m0 < Ns; m0 += (int)run) {
                    if (pass) {
                        items_gen[nitems_gen].s = s; items_gen[nitems_gen].m0 = m0;
                        items_gen[nitems_gen].nout = (Ns - m0 < run) ? Ns - m0 : (int)run;
                    }
                    nitems_gen++;
                }
            }
            if (!pass) {
                items2    = (tspws_fwd2_item_t *)calloc(nitems2 ? nitems2 : 1, sizeof(*items2));
                items_gen = (tspws_fwd_item_t *)calloc(nitems_gen ? nitems_gen : 1, sizeof(*items_gen));
                if (!items2 || !items_gen) goto done;
            }
        }
    }

    /* inverse: split the scales into groups of about equal taps per output sample */
    {
        int want = (int)env_long("TSPWS_INV_GROUPS", 16);
        if (want > S) want = S;
        if (want < 1) want = 1;
        grp = (tspws_inv_group_t *)calloc(want, sizeof(*grp));
        if (!grp) goto done;
        double total = 0, cum = 0;
        for (int s = 0; s < S; s++) total += ceil((double)f->L[s] / f->D[s]);
        int ng = 0, prev = -1;
        for (int s = 0; s < S; s++) {
            int gi = total > 0 ? (int)(cum / total * want) : 0;
            if (gi >= want) gi = want - 1;
            if (gi != prev) { grp[ng].s0 = s; ng++; prev = gi; }
            grp[ng - 1].s1 = s + 1;
            cum += ceil((double)f->L[s] / f->D[s]);
        }
        d->ninv = ng;
    }

    if (cudaMalloc((void **)&d->d_scale, (S ? S : 1) * sizeof(*sc)) != cudaSuccess) goto done;
    if (cudaMalloc(&d->d_w32, (f->ntaps ? f->ntaps : 1) * 2 * sizeof(float)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&d->d_items, (nitems ? nitems : 1) * sizeof(*items)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&d->d_invgrp, (d->ninv ? d->ninv : 1) * sizeof(*grp)) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_w32, w32, f->ntaps * 2 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_items, items, nitems * sizeof(*items), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_invgrp, grp, d->ninv * sizeof(*grp), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (want64) {   /* fp64 forward work lists: warp items for short filters, one CTA per 4 coefficients for long ones */
        /* TSPWS_F64_MAXL (default: off) restricts the fp64 mean transform to filters shorter than that;
         * longer filters then transform the unsplit group sums in fp32.  Measured at the headline
         * configuration: 6 us faster with 160, but the coherence error of single coarse
         * coefficients rises from 4.5e-6 to 1.8e-5, beyond the 1e-5 tolerance -- so every scale is split. */
        const int LONG_T = (int)env_long("TSPWS_F64_LONG", 1024), R64 = 4, MAXL = (int)env_long("TSPWS_F64_MAXL", 0x7fffffff);
        d->f64_maxl = MAXL;
        const long TAPS = env_long("TSPWS_F64_TAPS", 2048);
        int ns = 0, nl = 0;
        for (int s = 0; s < S; s++) {
            if (sc[s].L >= MAXL) continue;
            if (sc[s].L >= LONG_T) { nl += ((sc[s].Ns + R64 - 1) / R64) * TSPWS_F64_NSPLIT(sc[s].L); continue; }
            const int pass = (32 / sc[s].G) * R64;
            long run = TAPS / sc[s].L; run -= run % pass; if (run < pass) run = pass;
            ns += (sc[s].Ns + (int)run - 1) / (int)run;
        }
        tspws_fwd_item_t *its = (tspws_fwd_item_t *)calloc(ns ? ns : 1, sizeof(*its));
        int *lg = (int *)calloc((size_t)(nl ? nl : 1) * 2, sizeof(int));
        if (!its || !lg) { free(its); free(lg); goto done; }
        int ks = 0, kl = 0;
        for (int s = S - 1; s >= 0; s--) {          /* longest filters first */
            if (sc[s].L >= MAXL) continue;
            if (sc[s].L >= LONG_T) {
                for (int m = 0; m < sc[s].Ns; m += R64)
                    for (int q = 0; q < TSPWS_F64_NSPLIT(sc[s].L); q++) {
                        lg[2 * kl] = s | (q << 16); lg[2 * kl + 1] = m; kl++;
                        texp[2 * (sc[s].tile0 + m / 64) + 1]++;
                    }
                continue;
            }
            const int pass = (32 / sc[s].G) * R64;
            long run = TAPS / sc[s].L; run -= run % pass; if (run < pass) run = pass;
            for (int m0 = 0; m0 < sc[s].Ns; m0 += (int)run) {
                its[ks].s = s; its[ks].m0 = m0; its[ks].nout = (sc[s].Ns - m0 < run) ? sc[s].Ns - m0 : (int)run;
                for (int t = m0 / 64; t <= (m0 + its[ks].nout - 1) / 64; t++) texp[2 * (sc[s].tile0 + t) + 1]++;
                ks++;
            }
        }
        int bad = cudaMalloc((void **)&d->d_items_short, (size_t)(ns ? ns : 1) * sizeof(*its)) != cudaSuccess ||
                  cudaMemcpy(d->d_items_short, its, (size_t)ns * sizeof(*its), cudaMemcpyHostToDevice) != cudaSuccess ||
                  cudaMalloc((void **)&d->d_long, (size_t)(nl ? nl : 1) * 2 * sizeof(int)) != cudaSuccess ||
                  cudaMemcpy(d->d_long, lg, (size_t)nl * 2 * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess;
        free(its); free(lg);
        if (bad) goto done;
        d->nitems_short = ns; d->nlong = nl;
    }
    {   /* origin-aligned, zero-padded filters of the 4-sample fp32 inverse */
        size_t nwi = 0;
        for (int s = 0; s < S; s++) {
            const int dl = (4 - sc[s].c % 4) % 4;
            sc[s].ci = sc[s].c + dl;
            sc[s].Li4 = (sc[s].L + dl + 3) & ~3;
            sc[s].inv4 = (f->N % 4 == 0) && (sc[s].D % 4 == 0) && !env_long("TSPWS_NO_INV4", 0);
            sc[s].wi_off = (int)nwi;
            if (sc[s].inv4) nwi += (size_t)sc[s].Li4;
        }
        float *wi = (float *)calloc((nwi ? nwi : 4) * 2, sizeof(float));
        if (!wi) goto done;
        for (int s = 0; s < S; s++) {
            if (!sc[s].inv4) continue;
            const int dl = sc[s].ci - sc[s].c;
            for (int l = 0; l < sc[s].L; l++) {
                wi[2 * ((size_t)sc[s].wi_off + dl + l)]     = w32[2 * (f->tap_off[s] + l)];
                wi[2 * ((size_t)sc[s].wi_off + dl + l) + 1] = w32[2 * (f->tap_off[s] + l) + 1];
            }
        }
        int bad = cudaMalloc(&d->d_wi32, (nwi ? nwi : 4) * 2 * sizeof(float)) != cudaSuccess ||
                  cudaMemcpy(d->d_wi32, wi, (nwi ? nwi : 4) * 2 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess;
        free(wi);
        if (bad) goto done;
    }
    {   /* polyphase inverse (D >= 32, N % D == 0): reversed phase rows + items; the other scales go to
         * the 4-sample kernel, one group each */
        size_t ngi = 0;
        int ni2 = 0, n4 = 0;
        const int noinv2 = (int)env_long("TSPWS_NO_INV2", 0);
        for (int s = 0; s < S; s++) {
            const int D = sc[s].D;
            sc[s].Qi = (sc[s].Li4 + D - 1) / D;
            sc[s].inv2 = !noinv2 && D >= 32 && (f->N % D == 0);
            sc[s].gi_off = (int)ngi;
            if (sc[s].inv2) {
                if (sc[s].Qi > d->inv2_qmax) d->inv2_qmax = sc[s].Qi;
                ngi += (size_t)(sc[s].Qi + 8) * D;
                ni2 += ((sc[s].Ns + 7) / 8) * ((D + 31) / 32);
            } else n4++;
        }
        float *gi = (float *)calloc((ngi ? ngi : 1) * 2, sizeof(float));
        int *it2 = (int *)calloc((size_t)(ni2 ? ni2 : 1) * 4, sizeof(int));
        tspws_inv_group_t *g4 = (tspws_inv_group_t *)calloc(n4 ? n4 : 1, sizeof(*g4));
        if (!gi || !it2 || !g4) { free(gi); free(it2); free(g4); goto done; }
        int k = 0, k4 = 0;
        for (int s = 0; s < S; s++) {
            if (!sc[s].inv2) { g4[k4].s0 = s; g4[k4].s1 = s + 1; k4++; continue; }
            const int D = sc[s].D, Qi = sc[s].Qi, dl = sc[s].ci - sc[s].c;
            for (int t = 0; t < Qi; t++)
                for (int rho = 0; rho < D; rho++) {
                    const int l = rho + (Qi - 1 - t) * D - dl;          /* tap of the original filter */
                    if (l < 0 || l >= sc[s].L) continue;
                    gi[2 * ((size_t)sc[s].gi_off + (size_t)t * D + rho)]     = w32[2 * (f->tap_off[s] + l)];
                    gi[2 * ((size_t)sc[s].gi_off + (size_t)t * D + rho) + 1] = w32[2 * (f->tap_off[s] + l) + 1];
                }
            for (int a0 = 0; a0 < sc[s].Ns; a0 += 8)
                for (int pb = 0; pb < (D + 31) / 32; pb++) { it2[4 * k] = s; it2[4 * k + 1] = a0; it2[4 * k + 2] = pb; k++; }
        }
        int bad = cudaMalloc(&d->d_gi32, (ngi ? ngi : 1) * 2 * sizeof(float)) != cudaSuccess ||
                  cudaMemcpy(d->d_gi32, gi, (ngi ? ngi : 1) * 2 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                  cudaMalloc((void **)&d->d_items_inv2, (size_t)(ni2 ? ni2 : 1) * 4 * sizeof(int)) != cudaSuccess ||
                  cudaMemcpy(d->d_items_inv2, it2, (size_t)ni2 * 4 * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                  cudaMalloc((void **)&d->d_invgrp4, (size_t)(n4 ? n4 : 1) * sizeof(*g4)) != cudaSuccess ||
                  cudaMemcpy(d->d_invgrp4, g4, (size_t)n4 * sizeof(*g4), cudaMemcpyHostToDevice) != cudaSuccess;
        free(gi); free(it2); free(g4);
        if (bad) goto done;
        d->nitems_inv2 = ni2; d->ninv4 = n4;
        d->inv_slots = S > d->ninv ? S : d->ninv;
    }
    {   /* 16-coefficient tiles of all scales (work list of the combine kernel) */
        int nt = 0;
        for (int s = 0; s < S; s++) nt += (sc[s].Ns + 63) / 64;
        int *tl = (int *)malloc((size_t)(nt ? nt : 1) * 4 * sizeof(int));
        if (!tl) goto done;
        int k = 0;
        for (int s = 0; s < S; s++)
            for (int m0 = 0; m0 < sc[s].Ns; m0 += 64) {
                const int nv = sc[s].Ns - m0 < 64 ? sc[s].Ns - m0 : 64;
                tl[4 * k] = sc[s].ypart_off + m0; tl[4 * k + 1] = sc[s].coef_off + m0;
                /* valid coefficients | phase splits of the fp32 partials << 8 | tap splits of the fp64 mean
                 * transform << 16 (0: the mean of this scale is the extra fp32 trace) */
                const int nsp64 = sc[s].L >= (int)env_long("TSPWS_F64_MAXL", 0x7fffffff) ? 0 :
                                  sc[s].L >= (int)env_long("TSPWS_F64_LONG", 1024) ? TSPWS_F64_NSPLIT(sc[s].L) : 1;
                tl[4 * k + 2] = nv | (sc[s].nsplit << 8) | (nsp64 << 16); tl[4 * k + 3] = sc[s].Ns;
                k++;
            }
        int bad = cudaMalloc((void **)&d->d_tile16, (size_t)(nt ? nt : 1) * 4 * sizeof(int)) != cudaSuccess ||
                  cudaMemcpy(d->d_tile16, tl, (size_t)nt * 4 * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                  nt != ntile64 ||
                  cudaMalloc((void **)&d->d_tile_expect, (size_t)(nt ? nt : 1) * 2 * sizeof(int)) != cudaSuccess ||
                  cudaMemcpy(d->d_tile_expect, texp, (size_t)nt * 2 * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess;
        free(tl);
        if (bad) goto done;
        d->ntile16 = nt;
    }
    if (cudaMalloc(&d->d_g32, (ng ? ng : 1) * 2 * sizeof(float)) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_g32, g32, (ng ? ng : 1) * 2 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&d->d_items2, (nitems2 ? nitems2 : 1) * sizeof(*items2)) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_items2, items2, nitems2 * sizeof(*items2), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&d->d_items_gen, (nitems_gen ? nitems_gen : 1) * sizeof(*items_gen)) != cudaSuccess) goto done;
    if (cudaMemcpy(d->d_items_gen, items_gen, nitems_gen * sizeof(*items_gen), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    d->nitems2 = nitems2; d->nitems_gen = nitems_gen;
    if (want64) {
        if (cudaMalloc(&d->d_w64, (f->ntaps ? f->ntaps : 1) * 2 * sizeof(double)) != cudaSuccess) goto done;
        if (cudaMemcpy(d->d_w64, f->w, f->ntaps * 2 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    }
    /* the scale records go up last: every block above fills some of their fields */
    if (cudaMemcpy(d->d_scale, sc, S * sizeof(*sc), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    d->nitems = nitems;
    /* uploads above used the legacy stream; callers launch on non-blocking streams */
    if (cudaDeviceSynchronize() != cudaSuccess) goto done;
    ok = 1;
done:
    free(sc); free(w32); free(items); free(grp); free(items2); free(items_gen); free(g32); free(texp);
    return ok ? 0 : -1;
}

void tspws_devframe_free(tspws_devframe_t *d)
{
    free(d->h_item_first);
    cudaFree(d->d_scale); cudaFree(d->d_w32); cudaFree(d->d_w64);
    cudaFree(d->d_items); cudaFree(d->d_invgrp);
    cudaFree(d->d_g32); cudaFree(d->d_items2); cudaFree(d->d_items_gen); cudaFree(d->d_tile16); cudaFree(d->d_tile_expect); cudaFree(d->d_wi32); cudaFree(d->d_gi32); cudaFree(d->d_items_inv2); cudaFree(d->d_invgrp4); cudaFree(d->d_items_short); cudaFree(d->d_long);
    memset(d, 0, sizeof(*d));
}

/* group of global trace itr: reference src/ts_pws1f_lib.c:876 (unsigned product, double division) */
static unsigned int group_of(unsigned int itr, unsigned int K, unsigned int mtr)
{
    return (unsigned int)floor((double)(itr * K) / (double)mtr);
}

/* ---- two-stage work lists ---- */
static int twostage_build(tspws_plan_t *pl)
{
    const unsigned int K = pl->K, n = pl->mtr_local;
    const int ntiles = (pl->N + 1023) / 1024;
    int target_items = (int)(env_long("TSPWS_CS_CTAS", 148 * 8 * 2) / ntiles);
    if (target_items < 1) target_items = 1;
    int chunk = (int)((n + target_items - 1) / target_items);
    if (chunk < 8) chunk = 8;

    int *grp_t0   = (int *)calloc(K + 1, sizeof(int));
    int *slot_off = (int *)calloc(K + 1, sizeof(int));
    tspws_csum_item_t *it = (tspws_csum_item_t *)calloc((size_t)n / chunk + 2 * K + 2, sizeof(*it));
    int ni = 0, rc = -1;
    if (!grp_t0 || !slot_off || !it) goto done;
    {
        unsigned int t = 0;
        for (unsigned int g = 0; g < K; g++) {
            grp_t0[g] = (int)t;
            slot_off[g] = ni;
            while (t < n && group_of(pl->offset + t, K, pl->mtr_total) == g) t++;
            for (int a = grp_t0[g]; a < (int)t; a += chunk) {
                it[ni].t0 = a; it[ni].t1 = (a + chunk < (int)t) ? a + chunk : (int)t;
                it[ni].group = (int)g; it[ni].slot = ni;
                ni++;
            }
        }
        grp_t0[K] = (int)t; slot_off[K] = ni;
        if (t != n) { fprintf(stderr, "tspws_b200: internal error, group runs do not cover the shard\n"); goto done; }
    }
    pl->ncs = ni;
    pl->ntile = ntiles;
    const size_t N = (size_t)pl->N;
    if (cudaMalloc((void **)&pl->d_cs, (ni ? ni : 1) * sizeof(*it)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&pl->d_slot_off, (K + 1) * sizeof(int)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&pl->d_grp_t0, (K + 1) * sizeof(int)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&pl->d_part, (ni ? ni : 1) * N * sizeof(double)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&pl->d_tickets, (size_t)K * ntiles * sizeof(unsigned int)) != cudaSuccess) goto done;
    if (cudaMemset(pl->d_tickets, 0, (size_t)K * ntiles * sizeof(unsigned int)) != cudaSuccess) goto done;
    if (cudaMalloc((void **)&pl->d_gsum, (K * N + (size_t)K * ntiles) * sizeof(double)) != cudaSuccess) goto done;
    if (cudaMemset(pl->d_gsum, 0, (K * N + (size_t)K * ntiles) * sizeof(double)) != cudaSuccess) goto done;
    pl->d_gtile = pl->d_gsum + K * N;
    if (cudaMalloc((void **)&pl->d_gsum32, 2 * K * N * sizeof(float)) != cudaSuccess) goto done;   /* rows K..2K-1: unsplit sums (mean split) */
    if (cudaMemcpy(pl->d_cs, it, ni * sizeof(*it), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMemcpy(pl->d_slot_off, slot_off, (K + 1) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    if (cudaMemcpy(pl->d_grp_t0, grp_t0, (K + 1) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
    pl->h_slot_off = slot_off; slot_off = NULL;
    {
        int *id = (int *)malloc((K + 1) * sizeof(int));
        if (!id) goto done;
        for (unsigned int k = 0; k <= K; k++) id[k] = (int)k;
        const int bad = cudaMalloc((void **)&pl->d_ident, (K + 1) * sizeof(int)) != cudaSuccess ||
                        cudaMemcpy(pl->d_ident, id, (K + 1) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                        cudaMalloc((void **)&pl->d_gfin, K * N * sizeof(double)) != cudaSuccess;
        free(id);
        if (bad) goto done;
    }

    /* ---- single-wave variant: ntiles x nparts CTAs, all resident at once, equal rows per part ---- */
    pl->cs2 = 0;
    if (pl->N % 4 == 0 && n > 0 && !env_long("TSPWS_NO_CS2", 0)) {
        /* CTAs of the group-sum kernel.  3 per SM by default (measured on the headline stack: 139.8 us against 143.8 with
         * 4 and 141.3 with 2: enough loads in flight either way, less drain at the end); 2 per SM when the plan is a
         * buffer of a stack pipeline (tspws_pipe.c: leaves half the register file to the tail of the previous stack).
         * Handing the parts out through a queue (atomic per chunk, 592-2368 chunks) was tried for the pipeline: 147-160 us
         * per single stack, no better overlap -- static parts stay. */
        int nparts = (int)(env_long("TSPWS_CS2_CTAS", g_cs2_ctas_hint > 0 ? g_cs2_ctas_hint : 148 * 3) / ntiles);
        if (nparts > (int)((n + 3) / 4)) nparts = (int)((n + 3) / 4);
        if (nparts < 1) nparts = 1;
        int *pr = (int *)calloc((size_t)nparts * 4, sizeof(int));
        int *so = (int *)calloc(K + 1, sizeof(int));
        int ok = pr && so, nslots = 0, g = 0;
        for (unsigned int k = 0; k <= K && ok; k++) so[k] = -1;
        for (int q = 0; q < nparts && ok; q++) {
            const int r0 = (int)((unsigned long long)n * q / nparts), r1 = (int)((unsigned long long)n * (q + 1) / nparts);
            while (g < (int)K && grp_t0[g + 1] <= r0) g++;             /* group of row r0 */
            pr[4 * q] = r0; pr[4 * q + 1] = r1; pr[4 * q + 2] = g; pr[4 * q + 3] = nslots;
            /* the kernel's walk: one flush (= one slot) per group the part touches */
            for (int gg = g, r = r0; r < r1; gg++) {
                if (so[gg] < 0) so[gg] = nslots;
                nslots++;
                r = grp_t0[gg + 1] < r1 ? grp_t0[gg + 1] : r1;
            }
        }
        if (ok) {
            /* groups without local rows: empty slot range */
            int next = nslots;
            for (int k = (int)K; k >= 0; k--) { if (k == (int)K || so[k] < 0) so[k] = next; else next = so[k]; }
            ok = cudaMalloc(&pl->d_cs2_parts, (size_t)nparts * 4 * sizeof(int)) == cudaSuccess &&
                 cudaMalloc((void **)&pl->d_cs2_slot_off, (K + 1) * sizeof(int)) == cudaSuccess &&
                 cudaMalloc((void **)&pl->d_cs2_slots, (size_t)(nslots ? nslots : 1) * N * sizeof(double)) == cudaSuccess &&
                 cudaMalloc((void **)&pl->d_cs2_tot, (size_t)(nslots ? nslots : 1) * ntiles * sizeof(double)) == cudaSuccess &&
                 cudaMemcpy(pl->d_cs2_parts, pr, (size_t)nparts * 4 * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemcpy(pl->d_cs2_slot_off, so, (K + 1) * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
            if (ok) { pl->cs2 = 1; pl->cs2_nparts = nparts; pl->cs2_nslots = nslots; }
            else cudaGetLastError();
        }
        free(pr); free(so);
    }
    rc = 0;
done:
    free(grp_t0); free(slot_off); free(it);
    return rc;
}

int tspws_b200_plan_peer_close(tspws_plan_t *pl);

void tspws_b200_plan_destroy(tspws_plan_t *pl)
{
    if (!pl) return;
    if (pl->exch_bound) { pl->d_gsum = (double *)pl->own_gsum; pl->d_exch = pl->own_exch; }
    if (pl->gexec) { cudaGraphExecDestroy(pl->gexec); pl->gexec = NULL; }
    if (pl->side) { cudaStreamDestroy(pl->side); cudaEventDestroy(pl->ev_join); cudaEventDestroy(pl->ev_fork); }
    cudaFree(pl->d_gmean); cudaFree(pl->d_ymean); cudaFree(pl->d_tile_tickets);
    cudaFree(pl->d_bimg); cudaFree(pl->d_bands); cudaFree(pl->d_wimg); cudaFree(pl->d_ximg); cudaFree(pl->d_tsum); cudaFree(pl->d_sumtrace); cudaFree(pl->d_dpart); cudaFree(pl->d_ysum_ps); cudaFree(pl->d_ysum_st);
    free(pl->h_slot_off);
    tspws_devframe_free(&pl->dev);
    tspws_b200_plan_peer_close(pl);
    cudaFree(pl->d_peer_region); cudaFree(pl->d_peer_blocks); cudaFree(pl->d_peer_flags); cudaFree(pl->d_peer_step);
    cudaFree(pl->d_gfin); cudaFree(pl->d_ident);
    cudaFree(pl->d_cs2_parts); cudaFree(pl->d_cs2_slot_off); cudaFree(pl->d_cs2_slots); cudaFree(pl->d_cs2_tot);
    cudaFree(pl->d_rule); cudaFree(pl->d_cs); cudaFree(pl->d_slot_off); cudaFree(pl->d_grp_t0);
    cudaFree(pl->d_part); cudaFree(pl->d_tickets); cudaFree(pl->d_gsum); cudaFree(pl->d_gsum32);
    cudaFree(pl->d_pps); cudaFree(pl->d_pst); cudaFree(pl->d_exch);
    cudaFree(pl->d_zout); cudaFree(pl->d_invscratch); cudaFree(pl->d_ypart); cudaFree(pl->d_inv_tickets);
    tspws_b200_frame_destroy(pl->fr);
    free(pl);
}

static int floor_div(int a, int b) { int q = a / b; if ((a % b) && ((a < 0) != (b < 0))) q--; return q; }

/* Geometry of the banded GEMMs (host only, no GPU): walks down from scale s_dense (the first scale the dense GEMM
 * covers), one descriptor per run of scales with equal decimation D >= band_D.  all_polyphase: every scale with
 * D = 2^lg < 32 is on its polyphase list (the whole list is then skipped by the SIMT launch).  Returns the number of
 * descriptors; *s_first = first scale covered, coef_rel relative to coef_off[*s_first]; *tiles / *groups = B tiles and
 * CTA groups in total.  Exported for the CPU tests as tspws_b200_band_layout. */
static int band_layout(const tspws_frame_t *fr, int N, int s_dense, long band_D, int all_polyphase,
                       tspws_band_t *bands, int maxb, int *tiles_out, int *groups_out, int *s_first)
{
    int nb = 0, tiles = 0, groups = 0, sb = s_dense;
    const int kblocks = N / 64;
    while (band_D > 0 && sb > 0 && nb < maxb) {
        const int D = (int)fr->D[sb - 1];
        if (D < band_D || D < 2 || (N % D)) break;
        if (D < 32 && !all_polyphase) break;
        if (D <= 64 ? (64 % D) : (D % 64)) break;
        int sa = sb - 1;
        while (sa > 0 && (int)fr->D[sa - 1] == D) sa--;
        const int nv = sb - sa, U = D <= 64 ? 64 / D : 1, E = D <= 64 ? 1 : D / 64;
        if (nv > TSPWS_BAND_MAXV || nv * U * 2 > 128) break;
        int Pn = 128 / (nv * U * 2);                     /* as many row positions per group as 128 columns hold */
        if (Pn * E > kblocks) Pn = kblocks / E;
        if (Pn < 1) break;
        int qmin = 1 << 30, qmax = -(1 << 30);
        for (int v = 0; v < nv; v++) {
            const int c = fr->center[sa + v], L = (int)fr->L[sa + v];
            const int a = floor_div(-c, 64), b = floor_div((U - 1) * D - c + L - 1, 64);
            if (a < qmin) qmin = a;
            if (b > qmax) qmax = b;
        }
        tspws_band_t *b = &bands[nb];
        memset(b, 0, sizeof *b);
        b->nj = E * (Pn - 1) + qmax - qmin + 1;
        if (b->nj > kblocks) break;
        b->tile_base = tiles; b->ncols = (nv * Pn * U * 2 + 15) & ~15; b->rows_per_group = E * Pn; b->qmin = qmin;
        b->nv = nv; b->Pn = Pn; b->U = U; b->E = E; b->Ns = (int)fr->Ns[sa]; b->D = D;
        b->ngroups = (b->Ns + Pn * U - 1) / (Pn * U);
        for (int v = 0; v < nv; v++) { b->s_of[v] = sa + v; b->coef_rel[v] = (int)fr->coef_off[sa + v]; }
        tiles += b->nj; groups += b->ngroups; nb++;
        sb = sa;
    }
    const int ci0 = (int)fr->coef_off[sb];
    for (int i = 0; i < nb; i++)
        for (int v = 0; v < bands[i].nv; v++) bands[i].coef_rel[v] -= ci0;
    if (tiles_out) *tiles_out = tiles;
    if (groups_out) *groups_out = groups;
    if (s_first) *s_first = sb;
    return nb;
}

/* host-only view of the above for tests: descriptors as int[28] records; dense_D / band_D as the environment knobs */
int tspws_b200_band_layout(const tspws_frame_t *fr, long dense_D, long band_D, int *out /* [max][28] */, int max, int *s_first)
{
    if (!fr || !out || max <= 0 || fr->N % 64) return TSPWS_ERR_ARG;
    int s0 = -1;
    for (unsigned int s = 0; s < fr->S; s++)
        if ((long)fr->D[s] >= dense_D && fr->D[s] >= 32 && fr->D[s] <= fr->N) { s0 = (int)s; break; }
    if (s0 < 0) return 0;
    tspws_band_t bands[16];
    const int nb = band_layout(fr, (int)fr->N, s0, band_D, 1, bands, max < 16 ? max : 16, NULL, NULL, s_first);
    for (int i = 0; i < nb; i++) memcpy(out + 28 * i, &bands[i], sizeof(tspws_band_t));
    return nb;
}

/* Middle octaves of a single-stage plan as banded GEMMs (TSPWS_BAND_D, default 16 -- measured at 8192 x 16384, w0 = 6:
 * 9.11 ms with 32, 8.55 with 16, 8.66 with 8; 0 = off).  Updates td_* and dense_items. */
static int band_plan(tspws_plan_t *pl, int N)
{
    const tspws_frame_t *fr = pl->fr;
    tspws_band_t bands[16];
    int tiles = 0, groups = 0, sb = pl->dense_s0;
    const int nb = band_layout(fr, N, pl->dense_s0, env_long("TSPWS_BAND_D", 16), pl->dev.nitems_gen == 0, bands, 16, &tiles, &groups, &sb);
    if (!nb) return 0;
    int items = 0;
    int lg_end = 5;
    for (unsigned int s = (unsigned int)sb; s < fr->S; s++) {
        if (fr->D[s] >= 32) items += ((int)fr->Ns[s] + 15) / 16;
        else { int lg = 0; while ((1u << lg) < fr->D[s]) lg++; if (lg < lg_end) lg_end = lg; }   /* whole list of D = 2^lg skipped */
    }
    pl->simt_lg_end = lg_end;
    if (cudaMalloc((void **)&pl->d_bands, nb * sizeof(tspws_band_t)) != cudaSuccess) return 1;
    if (cudaMemcpy(pl->d_bands, bands, nb * sizeof(tspws_band_t), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
    if (cudaMalloc(&pl->d_bimg, (size_t)tiles * 3 * 16384) != cudaSuccess) return 1;
    if (tspws_cu_band_wimage(&pl->dev, pl->d_bands, nb, tiles, pl->d_bimg, NULL)) return 1;
    pl->band_n = nb; pl->band_groups = groups;
    pl->td_s0 = sb; pl->td_ci0 = (int)fr->coef_off[sb]; pl->td_ncd = (int)(fr->ncoef - fr->coef_off[sb]);
    pl->dense_items = items;
    return 0;
}

tspws_plan_t *tspws_b200_plan_create(const t_tsPWS *p, int max, unsigned int mtr_local,
                                     unsigned int mtr_total, unsigned int trace_offset)
{
    if (!p || max <= 0 || !mtr_total || mtr_local > mtr_total) return NULL;
    tspws_plan_t *pl = (tspws_plan_t *)calloc(1, sizeof(*pl));
    if (!pl) return NULL;
    pl->p = *p; pl->N = max;
    pl->mtr_local = mtr_local; pl->mtr_total = mtr_total; pl->offset = trace_offset;
    pl->two_stage = (p->Kmax && p->Kmax <= mtr_total);            /* src/ts_pws1f_lib.c:209 */
    pl->K     = pl->two_stage ? p->Kmax : mtr_total;
    /* TSPWS_OPT_FP64_TAIL: 1 = the Kmax transforms of a two-stage stack in fp64; 2 = also the per-trace transforms
     * of a single-stage stack (generic kernels, ~30x slower: a diagnostic that separates fp32 conditioning from bugs) */
    pl->acc64 = tspws_b200_get_option(TSPWS_OPT_FP64_TAIL) >= (pl->two_stage ? 1 : 2);
    pl->csz   = pl->acc64 ? 2 * sizeof(double) : 2 * sizeof(float);
    pl->rule  = tspws_weight_rule(p->wu, p->unbiased, pl->K, mtr_total);

    pl->fr = tspws_b200_frame_create(p->type, p->J, p->V, (unsigned int)max, p->s0, p->b0, p->w0, (int)p->uni);
    if (!pl->fr) { fprintf(stderr, "tspws_b200: cannot build the wavelet frame\n"); free(pl); return NULL; }
    for (unsigned int s = 0; s < pl->fr->S; s++)
        if (pl->fr->L[s] > (unsigned int)max) goto fail;          /* cannot happen: L is capped at N */
    pl->mean_split = pl->two_stage && !pl->acc64 && env_long("TSPWS_MEAN_SPLIT", 1);
    pl->hybrid = pl->mean_split && env_long("TSPWS_F64_MAXL", 0x7fffffff) != 0x7fffffff;
    if (tspws_devframe_build(&pl->dev, pl->fr, pl->acc64 || pl->mean_split, pl->two_stage && !pl->acc64)) goto fail;

    CK(cudaMalloc((void **)&pl->d_rule, sizeof(tspws_weight_t)));
    CK(cudaMemcpy(pl->d_rule, &pl->rule, sizeof(tspws_weight_t), cudaMemcpyHostToDevice));

    const size_t ncoef = pl->fr->ncoef ? pl->fr->ncoef : 1;
    if (pl->two_stage) {
        if (twostage_build(pl)) goto fail;
        pl->ntr_fwd = (int)pl->K;
        pl->chunk   = (int)env_long("TSPWS_TAIL_CHUNK", 2);
    } else {
        pl->ntr_fwd = (int)mtr_local;
        long c = env_long("TSPWS_FWD_CHUNK", 0);
        if (c <= 0) { c = (mtr_local + 127) / 128; if (c < 4) c = 4; }
        pl->chunk = (int)c;
    }
    if (pl->chunk < 1) pl->chunk = 1;
    pl->nchunk = (pl->ntr_fwd + pl->chunk - 1) / pl->chunk;
    if (pl->nchunk < 1) pl->nchunk = 1;
    if (pl->dev.per_trace) {
        CK(cudaMalloc(&pl->d_ypart, (size_t)(pl->dev.ypart_len ? pl->dev.ypart_len : 1) * pl->K * 2 * sizeof(float)));
    } else {
        CK(cudaMalloc(&pl->d_pps, (size_t)pl->nchunk * ncoef * pl->csz));
        CK(cudaMalloc(&pl->d_pst, (size_t)pl->nchunk * ncoef * pl->csz));
        /* Dense tensor-core transform of the coarse scales (D >= TSPWS_DENSE_D, default 256; 0 = off): single-stage fp32
         * stacks of at least TSPWS_DENSE_MIN traces (default 500 = the station-pair geometry of BASELINE configs[4]: the products are fp32-faithful, but the phase of a near-zero
         * coefficient of a single trace weighs 1/mtr in the coherence -- measured 1.1e-5 at 256 traces, 2.7-3.3e-6 at 500 (three seeds), 2e-6 from 1024 on) whose length is a multiple of 64.  The scales must be
         * polyphase (D >= 32) items so that the SIMT launch can skip them by item index. */
        const long dD = env_long("TSPWS_DENSE_D", 256);
        if (!pl->two_stage && !pl->acc64 && dD > 0 && max % 64 == 0 && (long)mtr_local >= env_long("TSPWS_DENSE_MIN", 500)) {
            int s0 = -1;
            for (unsigned int s = 0; s < pl->fr->S; s++)
                if ((long)pl->fr->D[s] >= dD && pl->fr->D[s] >= 32 && pl->fr->D[s] <= (unsigned int)max) { s0 = (int)s; break; }
            if (s0 >= 0) {
                int ok = 1, items = 0;
                for (unsigned int s = (unsigned int)s0; s < pl->fr->S; s++) {
                    if (pl->fr->D[s] < 32 || pl->fr->D[s] > (unsigned int)max) ok = 0;
                    items += ((int)pl->fr->Ns[s] + 15) / 16;
                }
                if (ok) {
                    pl->dense_s0 = s0; pl->dense_ci0 = (int)pl->fr->coef_off[s0];
                    pl->dense_ncd = (int)(pl->fr->ncoef - pl->fr->coef_off[s0]);
                    pl->dense_items = items;
                    const int mtiles = ((int)mtr_local + 127) / 128;
                    CK(cudaMalloc(&pl->d_wimg, tspws_cu_dense_wimg_bytes(2 * pl->dense_ncd, max)));
                    CK(cudaMalloc(&pl->d_ximg, tspws_cu_dense_ximg_bytes((int)mtr_local, max)));
                    CK(cudaMalloc((void **)&pl->d_tsum, (size_t)mtiles * max * sizeof(float)));
                    CK(cudaMalloc((void **)&pl->d_sumtrace, (size_t)max * sizeof(float)));
                    pl->td_s0 = s0; pl->td_ci0 = pl->dense_ci0; pl->td_ncd = pl->dense_ncd; pl->simt_lg_end = 5;
                    if (band_plan(pl, (int)max)) goto fail;
                    CK(cudaMalloc((void **)&pl->d_dpart, (size_t)mtiles * 2 * pl->td_ncd * sizeof(float)));
                    CK(cudaMemset(pl->d_dpart, 0, (size_t)mtiles * 2 * pl->td_ncd * sizeof(float)));
                    CK(cudaMalloc(&pl->d_ysum_ps, ncoef * 2 * sizeof(float)));
                    CK(cudaMalloc(&pl->d_ysum_st, ncoef * 2 * sizeof(float)));
                    if (tspws_cu_dense_wimage(&pl->dev, s0, pl->dense_ci0, pl->dense_ncd, pl->d_wimg, NULL)) goto fail;
                    /* the SIMT kernels never write the dense coefficients: their chunk partials stay zero */
                    CK(cudaMemset(pl->d_pps, 0, (size_t)pl->nchunk * ncoef * pl->csz));
                    CK(cudaMemset(pl->d_pst, 0, (size_t)pl->nchunk * ncoef * pl->csz));
                    pl->dense_on = 1;
                }
            }
        }
    }
    CK(cudaMalloc(&pl->d_exch, 2 * ncoef * pl->csz));
    if (pl->mean_split) {
        CK(cudaMalloc((void **)&pl->d_gmean, (size_t)max * 3 * sizeof(double)));      /* period-padded, see fwd_f64_kernel */
        CK(cudaMalloc(&pl->d_ymean, (size_t)TSPWS_F64_MAXSPLIT * ncoef * 2 * sizeof(double)));   /* [split][ncoef] */
        CK(cudaMemset(pl->d_ymean, 0, (size_t)TSPWS_F64_MAXSPLIT * ncoef * 2 * sizeof(double)));
        if (env_long("TSPWS_FUSED_COMBINE", 0)) {     /* experiment (slower): combine fused into fwd_tail_kernel by tile tickets */
            const size_t nt = pl->dev.ntile16 ? pl->dev.ntile16 : 1;
            CK(cudaMalloc((void **)&pl->d_tile_tickets, nt * sizeof(unsigned int)));
            CK(cudaMemset(pl->d_tile_tickets, 0, nt * sizeof(unsigned int)));
        }
    }
    CK(cudaMalloc(&pl->d_zout, ncoef * (pl->acc64 ? 2 * sizeof(double) : 4 * sizeof(float))));
    CK(cudaMalloc(&pl->d_invscratch, (size_t)2 * (pl->dev.inv_slots ? pl->dev.inv_slots : 1) * max * (pl->csz / 2)));
    CK(cudaMemset(pl->d_exch, 0, 2 * ncoef * pl->csz));
    if (!pl->acc64 && env_long("TSPWS_SIDE_STREAM", 1)) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&pl->side, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
    }
    CK(cudaMalloc((void **)&pl->d_inv_tickets, (size_t)((max + 255) / 256) * sizeof(unsigned int)));
    CK(cudaMemset(pl->d_inv_tickets, 0, (size_t)((max + 255) / 256) * sizeof(unsigned int)));
    /* the uploads / memsets above went through the legacy stream; the plan runs on caller streams that do
     * not synchronise with it (cudaStreamNonBlocking), so finish them here, once */
    CK(cudaDeviceSynchronize());
    return pl;
fail:
    tspws_b200_plan_destroy(pl);
    return NULL;
}

static void *plan_ps(tspws_plan_t *pl) { return pl->d_exch; }
static void *plan_st(tspws_plan_t *pl) { return (char *)pl->d_exch + (size_t)pl->fr->ncoef * pl->csz; }

static int single_stage_close(tspws_plan_t *pl, void *stream);

int tspws_b200_plan_reduce(tspws_plan_t *pl, float *d_sigall, void *stream)
{
    if (!pl || (!d_sigall && pl->mtr_local)) return TSPWS_ERR_NULL;
    int e;
    pl->summed_across_ranks = 0;
    tspws_cu_profile_mark(stream);
    pl->cs2_pending = 0;
    pl->peer_pending = 0;
    if (pl->peer_on && ((size_t)d_sigall & 15)) return TSPWS_ERR_ARG;    /* peers would wait for a block that never comes */
    if (pl->two_stage && pl->cs2 && !((size_t)d_sigall & 15)) {
        if ((e = tspws_cu_colsum2(d_sigall, pl->N, pl->d_cs2_parts, pl->cs2_nparts, pl->d_grp_t0, pl->d_cs2_slots,
                                  pl->d_cs2_tot, pl->ntile, stream))) return TSPWS_ERR_CUDA;
        if (pl->peer_on) {
            /* push this rank's block into every rank's region (slot [parity][rank]) and tell every peer */
            unsigned int *flags = (unsigned int *)((char *)pl->d_peer_region + pl->peer_flag_off);
            if ((e = tspws_cu_group_post2_peer(1, pl->d_cs2_slots, pl->d_cs2_tot, pl->d_cs2_slot_off, pl->N, (int)pl->K, pl->ntile,
                                               0, 0, 0, NULL, NULL, NULL, pl->d_peer_blocks, (double *)pl->d_peer_region, flags,
                                               pl->d_peer_step, flags + TSPWS_PEER_MAXW, pl->d_peer_flags, pl->peer_rank, pl->peer_world, stream))) return TSPWS_ERR_CUDA;
            pl->peer_pending = 1;
        } else if (pl->mtr_local != pl->mtr_total || pl->exch_bound) {
            /* sharded: the exchange block (group sums + tile totals) has to be complete before the all-reduce */
            if ((e = tspws_cu_group_post2(pl->d_cs2_slots, pl->d_cs2_tot, pl->d_cs2_slot_off, pl->N, (int)pl->K, pl->ntile,
                                          0, 0, 0, 0, pl->d_gsum, pl->d_gtile, NULL, NULL, stream))) return TSPWS_ERR_CUDA;
        } else {
            pl->cs2_pending = 1;                   /* the slots are added up inside group_post2 (plan_finish) */
        }
        return 0;
    }
    if (pl->two_stage) {
        if (pl->mtr_local != pl->mtr_total)   /* groups without local traces must contribute zeros */
            if (cudaMemsetAsync(pl->d_gsum, 0, tspws_b200_plan_exchange_bytes(pl), (cudaStream_t)stream) != cudaSuccess) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_colsum(d_sigall, pl->N, pl->d_cs, pl->ncs, pl->d_part, pl->d_slot_off, pl->d_tickets,
                                 pl->d_gsum, pl->d_gtile, stream))) return TSPWS_ERR_CUDA;
        return 0;
    }
    if (tspws_b200_plan_part_align(pl)) {
        /* fp32 single-stage: the whole shard as one part, then the closing step */
        if ((e = tspws_b200_plan_reduce_part(pl, d_sigall, 0, pl->mtr_local, stream))) return e;
        return tspws_b200_plan_reduce_done(pl, stream);
    }
    if ((e = tspws_cu_prep(d_sigall, (int)pl->mtr_local, pl->N, pl->p.fold, pl->p.lrm, stream))) return TSPWS_ERR_CUDA;
    if (pl->acc64) {
        if ((e = tspws_cu_forward(&pl->dev, d_sigall, 0, 1, pl->ntr_fwd, pl->chunk, pl->d_pps, pl->d_pst, stream))) return TSPWS_ERR_CUDA;
    } else {
        if (pl->dense_on &&
            (e = tspws_cu_forward_dense(d_sigall, pl->ntr_fwd, pl->N, pl->d_wimg, pl->dense_ncd, pl->d_ximg, pl->d_tsum, pl->d_sumtrace,
                                        pl->d_dpart, 2 * pl->td_ncd, 2 * (pl->dense_ci0 - pl->td_ci0), stream))) return TSPWS_ERR_CUDA;
        if (pl->band_n &&
            (e = tspws_cu_forward_band(pl->d_ximg, pl->ntr_fwd, pl->N, pl->d_bimg, pl->d_bands, pl->band_n, pl->band_groups,
                                       pl->d_dpart, 2 * pl->td_ncd, stream))) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_forward_fast(&pl->dev, d_sigall, pl->ntr_fwd, pl->chunk, pl->d_pps, pl->d_pst,
                                       pl->dense_on ? pl->dense_items : 0, pl->dense_on ? pl->simt_lg_end : 5, stream))) return TSPWS_ERR_CUDA;
        if (pl->dense_on) {
            /* linear stack of the dense coefficients = transform of the summed trace (generic kernel, one trace) */
            if ((e = tspws_cu_forward_one(&pl->dev, pl->d_sumtrace, pl->dev.h_item_first[pl->td_s0], pl->d_ysum_ps, pl->d_ysum_st, stream))) return TSPWS_ERR_CUDA;
            if ((e = tspws_cu_dense_finish(pl->d_dpart, pl->ntr_fwd, pl->td_ncd, pl->d_ysum_st, pl->td_ci0, pl->d_pps, pl->d_pst, stream))) return TSPWS_ERR_CUDA;
        }
    }
    return single_stage_close(pl, stream);
}

/* reduce_weight + the push of the sums to the peers: the end of the reduce step of a single-stage plan */
static int single_stage_close(tspws_plan_t *pl, void *stream)
{
    int e;
    if ((e = tspws_cu_reduce_weight(pl->d_pps, pl->d_pst, pl->acc64, pl->dev.ncoef, 1, pl->nchunk, pl->d_rule,
                                    plan_ps(pl), plan_st(pl), pl->d_zout, stream))) return TSPWS_ERR_CUDA;
    if (pl->peer_on && !pl->acc64) {
        /* push this rank's PS | ST sums into every rank's region; plan_finish adds the blocks of all ranks */
        unsigned int *flags = (unsigned int *)((char *)pl->d_peer_region + pl->peer_flag_off);
        const size_t blkd = ((tspws_b200_plan_exchange_bytes(pl) + 15) & ~(size_t)15) / sizeof(double);
        if ((e = tspws_cu_peer_f32(1, (float *)pl->d_exch, (long long)(4 * pl->fr->ncoef), (long long)blkd, pl->d_peer_blocks,
                                   (double *)pl->d_peer_region, flags, pl->d_peer_step, flags + TSPWS_PEER_MAXW, pl->d_peer_flags,
                                   pl->peer_rank, pl->peer_world, stream))) return TSPWS_ERR_CUDA;
        pl->peer_pending = 1;
    }
    return 0;
}

/* ---- the reduce step of an fp32 single-stage plan in PARTS of the shard: lets a caller that still copies the traces to
 * the device transform the part that has arrived while the next one is on the bus (tspws_main does).  Parts start at
 * multiples of tspws_b200_plan_part_align() traces (0: the plan does not take parts -- two-stage, fp64, non-persistent
 * tensor-core path) and every part but the last is a multiple of it long; after the last part call ..._reduce_done. ---- */
unsigned int tspws_b200_plan_part_align(tspws_plan_t *pl)
{
    if (!pl || pl->two_stage || pl->acc64 || pl->dev.per_trace) return 0;
    if (pl->dense_on && !env_long("TSPWS_MMA_PERSIST", 1)) return 0;
    unsigned int a = (unsigned int)pl->chunk;
    if (pl->dense_on) { unsigned int x = a, y = 128; while (y) { const unsigned int t = x % y; x = y; y = t; } a = a / x * 128; }   /* lcm(chunk, 128) */
    return a;
}

int tspws_b200_plan_reduce_part(tspws_plan_t *pl, float *d_part, unsigned int t0, unsigned int nt, void *stream)
{
    const unsigned int al = tspws_b200_plan_part_align(pl);
    if (!pl || (!d_part && nt)) return TSPWS_ERR_NULL;
    if (!al || t0 % al || t0 + nt > pl->mtr_local || ((nt % al) && t0 + nt != pl->mtr_local)) return TSPWS_ERR_ARG;
    if (!nt) return 0;
    int e;
    if (t0 == 0) { pl->summed_across_ranks = 0; pl->peer_pending = 0; tspws_cu_profile_mark(stream); }
    const size_t ncoef = pl->fr->ncoef;
    if ((e = tspws_cu_prep(d_part, (int)nt, pl->N, pl->p.fold, pl->p.lrm, stream))) return TSPWS_ERR_CUDA;
    if (pl->dense_on) {
        const size_t mt0 = t0 / 128;
        char *ximg = (char *)pl->d_ximg + mt0 * tspws_cu_dense_ximg_bytes(128, pl->N);
        if ((e = tspws_cu_forward_image(d_part, (int)nt, pl->N, ximg, pl->d_tsum + mt0 * pl->N, NULL, stream))) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_forward_mma_all(ximg, (int)nt, pl->N, pl->d_wimg, pl->dense_ncd, pl->d_bimg, pl->d_bands, pl->band_n,
                                          pl->band_groups, pl->d_dpart + mt0 * 2 * (size_t)pl->td_ncd, 2 * pl->td_ncd,
                                          2 * (pl->dense_ci0 - pl->td_ci0), stream))) return TSPWS_ERR_CUDA;
    }
    const size_t c0 = t0 / (unsigned int)pl->chunk;
    if ((e = tspws_cu_forward_fast(&pl->dev, d_part, (int)nt, pl->chunk, (char *)pl->d_pps + c0 * ncoef * pl->csz,
                                   (char *)pl->d_pst + c0 * ncoef * pl->csz, pl->dense_on ? pl->dense_items : 0,
                                   pl->dense_on ? pl->simt_lg_end : 5, stream))) return TSPWS_ERR_CUDA;
    return 0;
}

int tspws_b200_plan_reduce_done(tspws_plan_t *pl, void *stream)
{
    if (!pl) return TSPWS_ERR_NULL;
    if (!tspws_b200_plan_part_align(pl)) return TSPWS_ERR_ARG;
    int e;
    if (pl->dense_on) {
        /* linear stack of the tensor-core coefficients = transform of the summed trace (generic kernel, one trace) */
        if ((e = tspws_cu_tile_sum(pl->d_tsum, pl->ntr_fwd, pl->N, pl->d_sumtrace, stream))) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_forward_one(&pl->dev, pl->d_sumtrace, pl->dev.h_item_first[pl->td_s0], pl->d_ysum_ps, pl->d_ysum_st, stream))) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_dense_finish(pl->d_dpart, pl->ntr_fwd, pl->td_ncd, pl->d_ysum_st, pl->td_ci0, pl->d_pps, pl->d_pst, stream))) return TSPWS_ERR_CUDA;
    }
    return single_stage_close(pl, stream);
}

void *tspws_b200_plan_exchange_ptr(tspws_plan_t *pl)
{
    if (!pl) return NULL;
    return pl->two_stage ? (void *)pl->d_gsum : pl->d_exch;
}

/* Use a caller-owned device buffer (>= exchange_bytes, 16 B aligned) as the exchange block, e.g. a
 * torch tensor that torch.distributed can all-reduce in place. */
int tspws_b200_plan_bind_exchange(tspws_plan_t *pl, void *d_buf)
{
    if (!pl || !d_buf) return TSPWS_ERR_NULL;
    if ((size_t)d_buf & 15) return TSPWS_ERR_ARG;
    if (!pl->exch_bound) { pl->own_gsum = pl->d_gsum; pl->own_exch = pl->d_exch; }
    if (pl->two_stage) {
        pl->d_gsum  = (double *)d_buf;
        pl->d_gtile = pl->d_gsum + (size_t)pl->K * pl->N;
    } else {
        pl->d_exch = d_buf;
    }
    pl->exch_bound = 1;
    return 0;
}

/* The exchange block now holds the sum over all ranks (call between reduce and finish). */
int tspws_b200_plan_mark_exchanged(tspws_plan_t *pl)
{
    if (!pl) return TSPWS_ERR_NULL;
    pl->summed_across_ranks = 1;
    return 0;
}
size_t tspws_b200_plan_exchange_bytes(tspws_plan_t *pl)
{
    if (!pl) return 0;
    return pl->two_stage ? ((size_t)pl->K * pl->N + (size_t)pl->K * pl->ntile) * sizeof(double)
                         : 2 * pl->fr->ncoef * pl->csz;
}
int tspws_b200_plan_exchange_is_f64(tspws_plan_t *pl) { return pl && pl->two_stage; }

/* ---- peer exchange: the ranks of one node map each other's exchange blocks (CUDA IPC) and the sum over
 * ranks is read straight from peer memory inside group_post2_kernel<true>; no collective call ---- */
static int peer_region_alloc(tspws_plan_t *pl, int world)
{
    /* two-stage: fp64 group sums (needs the single-wave group-sum kernel); single-stage: the fp32 PS | ST sums */
    if ((pl->two_stage && !pl->cs2) || pl->acc64 || world < 1 || world > TSPWS_PEER_MAXW) return TSPWS_ERR_ARG;
    if (pl->d_peer_region && pl->peer_world != world) return TSPWS_ERR_ARG;
    if (!pl->d_peer_region) {
        const size_t blk = pl->two_stage ? tspws_b200_plan_exchange_bytes(pl) + pl->K * sizeof(double)   /* + presence marks */
                                         : ((tspws_b200_plan_exchange_bytes(pl) + 15) & ~(size_t)15);
        pl->peer_world = world;
        pl->peer_flag_off = (2 * (size_t)world * blk + 255) & ~(size_t)255;     /* [parity][rank][block] */
        pl->peer_bytes = pl->peer_flag_off + (TSPWS_PEER_MAXW + 1) * sizeof(unsigned int);
        if (cudaMalloc(&pl->d_peer_region, pl->peer_bytes) != cudaSuccess ||
            cudaMemset(pl->d_peer_region, 0, pl->peer_bytes) != cudaSuccess ||
            cudaMalloc((void **)&pl->d_peer_step, 2 * sizeof(unsigned int)) != cudaSuccess ||      /* step, CTA ticket */
            cudaMemset(pl->d_peer_step, 0, 2 * sizeof(unsigned int)) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return TSPWS_ERR_ALLOC; }
    }
    return 0;
}

int tspws_b200_plan_peer_export(tspws_plan_t *pl, int world, void *handle64)
{
    if (!pl || !handle64) return TSPWS_ERR_NULL;
    if (sizeof(cudaIpcMemHandle_t) != 64) return TSPWS_ERR_ARG;
    const int e = peer_region_alloc(pl, world);
    if (e) return e;
    if (cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, pl->d_peer_region) != cudaSuccess) { cudaGetLastError(); return TSPWS_ERR_CUDA; }
    return 0;
}

/* Same-process variant (several plans of one process, on one GPU or on GPUs with peer access enabled): the
 * region's device pointer instead of an IPC handle.  NULL when the plan cannot take part (not two-stage,
 * N % 4 != 0, no local traces). */
void *tspws_b200_plan_peer_region(tspws_plan_t *pl, int world)
{
    if (!pl || peer_region_alloc(pl, world)) return NULL;
    return pl->d_peer_region;
}

/* ptr[r]: mapped region of rank r (own entry ignored); takes ownership of ptr on success */
static int peer_attach(tspws_plan_t *pl, int rank, int world, void **ptr, int ipc)
{
    void **blocks = (void **)calloc(world, sizeof(void *)), **flags = (void **)calloc(world, sizeof(void *));
    int rc = TSPWS_ERR_CUDA;
    if (!blocks || !flags) { rc = TSPWS_ERR_ALLOC; goto out; }
    ptr[rank] = pl->d_peer_region;
    for (int r = 0; r < world; r++) {
        blocks[r] = ptr[r];
        flags[r]  = (char *)ptr[r] + pl->peer_flag_off;
    }
    if (cudaMalloc(&pl->d_peer_blocks, world * sizeof(void *)) != cudaSuccess ||
        cudaMalloc(&pl->d_peer_flags, world * sizeof(void *)) != cudaSuccess ||
        cudaMemcpy(pl->d_peer_blocks, blocks, world * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(pl->d_peer_flags, flags, world * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); goto out; }
    pl->h_peer_ptr = ptr;
    pl->peer_ipc = ipc;
    pl->peer_rank = rank; pl->peer_world = world; pl->peer_on = 1;
    if (pl->gexec) { cudaGraphExecDestroy(pl->gexec); pl->gexec = NULL; }
    rc = 0;
out:
    free(blocks); free(flags);
    return rc;
}

int tspws_b200_plan_peer_import(tspws_plan_t *pl, int rank, int world, const void *handles)
{
    if (!pl || !handles) return TSPWS_ERR_NULL;
    if (!pl->d_peer_region || world != pl->peer_world || rank < 0 || rank >= world || pl->h_peer_ptr) return TSPWS_ERR_ARG;
    void **ptr = (void **)calloc(world, sizeof(void *));
    if (!ptr) return TSPWS_ERR_ALLOC;
    int rc = TSPWS_ERR_CUDA;
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        if (cudaIpcOpenMemHandle(&ptr[r], ((const cudaIpcMemHandle_t *)handles)[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            fprintf(stderr, "tspws_b200: cannot map the exchange block of rank %d (%s)\n", r, cudaGetErrorString(cudaGetLastError()));
            ptr[r] = NULL;
            goto out;
        }
    }
    rc = peer_attach(pl, rank, world, ptr, 1);
out:
    if (rc) {                                        /* failed half way: unmap what was mapped */
        for (int r = 0; r < world; r++)
            if (r != rank && ptr[r]) cudaIpcCloseMemHandle(ptr[r]);
        cudaGetLastError();
        free(ptr);
    }
    return rc;
}

int tspws_b200_plan_peer_import_ptrs(tspws_plan_t *pl, int rank, int world, void *const *regions)
{
    if (!pl || !regions) return TSPWS_ERR_NULL;
    if (!pl->d_peer_region || world != pl->peer_world || rank < 0 || rank >= world || pl->h_peer_ptr) return TSPWS_ERR_ARG;
    void **ptr = (void **)calloc(world, sizeof(void *));
    if (!ptr) return TSPWS_ERR_ALLOC;
    for (int r = 0; r < world; r++) {
        ptr[r] = regions[r];
        if (r != rank && !ptr[r]) { free(ptr); return TSPWS_ERR_ARG; }
    }
    const int rc = peer_attach(pl, rank, world, ptr, 0);
    if (rc) free(ptr);
    return rc;
}

/* Unmap the other ranks' regions (the own region stays allocated until plan_destroy).  CUDA requires that an
 * exported allocation is not freed while another process still maps it: ranks call this, synchronise among
 * themselves (barrier), and only then destroy their plans.  Peer mode is off afterwards; the plan is only
 * good for plan_destroy (or the NCCL-style reduce / exchange / finish sequence). */
int tspws_b200_plan_peer_close(tspws_plan_t *pl)
{
    if (!pl) return TSPWS_ERR_NULL;
    if (pl->h_peer_ptr) {
        cudaDeviceSynchronize();
        for (int r = 0; r < pl->peer_world; r++)
            if (pl->peer_ipc && r != pl->peer_rank && pl->h_peer_ptr[r]) cudaIpcCloseMemHandle(pl->h_peer_ptr[r]);
        free(pl->h_peer_ptr);
        pl->h_peer_ptr = NULL;
        pl->peer_on = 0;
        if (pl->gexec) { cudaGraphExecDestroy(pl->gexec); pl->gexec = NULL; }
    }
    return 0;
}

/* 1 if a peer wait timed out since the plan was created (a rank died or never arrived) */
int tspws_b200_plan_peer_error(tspws_plan_t *pl)
{
    if (!pl || !pl->d_peer_region) return 0;
    unsigned int e = 0;
    if (cudaMemcpy(&e, (char *)pl->d_peer_region + pl->peer_flag_off + TSPWS_PEER_MAXW * sizeof(unsigned int), sizeof e,
                   cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 1; }
    return e != 0;
}

int tspws_b200_plan_finish(tspws_plan_t *pl, float *d_ls, float *d_tspws, void *stream)
{
    if (!pl || !d_ls || !d_tspws) return TSPWS_ERR_NULL;
    int e;
    const long stop = env_long("TSPWS_DEBUG_STOP", 0);   /* timing experiments only: cut the tail after stage n */
    if (stop == 1) return 0;
    tspws_cu_profile_mark(stream);
    if (pl->two_stage) {
        /* post-processed fp64 group sums: from the slots of colsum2 (unsharded), or from the exchange block
         * (sharded / bound / N % 4 != 0) seen as one slot per group, written to a separate buffer */
        const double *gfin = pl->d_gsum;
        if (pl->peer_pending) {
            /* the sum over ranks is fused in: the blocks every rank pushed into this region, in rank order */
            unsigned int *flags = (unsigned int *)((char *)pl->d_peer_region + pl->peer_flag_off);
            if ((e = tspws_cu_group_post2_peer(0, NULL, NULL, NULL, pl->N, (int)pl->K, pl->ntile, pl->p.lrm, pl->p.fold,
                                               pl->mean_split ? 1 + pl->hybrid : 0, pl->d_gfin, pl->d_gsum32,
                                               pl->mean_split ? pl->d_gmean : NULL, pl->d_peer_blocks, (double *)pl->d_peer_region, flags,
                                               pl->d_peer_step, flags + TSPWS_PEER_MAXW, pl->d_peer_flags, pl->peer_rank, pl->peer_world, stream))) return TSPWS_ERR_CUDA;
            gfin = pl->d_gfin;
        } else if (pl->cs2_pending) {
            if ((e = tspws_cu_group_post2(pl->d_cs2_slots, pl->d_cs2_tot, pl->d_cs2_slot_off, pl->N, (int)pl->K, pl->ntile,
                                          1, pl->p.lrm, pl->p.fold, pl->mean_split ? 1 + pl->hybrid : 0, pl->d_gsum, pl->d_gtile, pl->d_gsum32,
                                          pl->mean_split ? pl->d_gmean : NULL, stream))) return TSPWS_ERR_CUDA;
        } else {
            if ((e = tspws_cu_group_post2(pl->d_gsum, pl->d_gtile, pl->d_ident, pl->N, (int)pl->K, pl->ntile,
                                          1, pl->p.lrm, pl->p.fold, pl->mean_split ? 1 + pl->hybrid : 0, pl->d_gfin, NULL, pl->d_gsum32,
                                          pl->mean_split ? pl->d_gmean : NULL, stream))) return TSPWS_ERR_CUDA;
            gfin = pl->d_gfin;
        }
        if (stop == 2) return 0;
        if (pl->dev.per_trace) {
            const int long_row0 = pl->mean_split && pl->hybrid ? pl->ntr_fwd : 0;
            int fused_combine = 0;
            if (pl->mean_split && pl->dev.nitems_gen == 0 && pl->dev.nitems2 > 0 && env_long("TSPWS_FUSED_TAIL", 1)) {
                /* one launch: fp64 transform of the group mean + fp32 transforms of the residuals */
                fused_combine = pl->d_tile_tickets != NULL;
                if ((e = tspws_cu_forward_tail(&pl->dev, pl->d_gsum32, pl->ntr_fwd, long_row0, pl->d_ypart, pl->d_gmean, pl->d_ymean,
                                               pl->d_tile_tickets, pl->d_rule, plan_ps(pl), plan_st(pl), pl->d_zout, stream))) return TSPWS_ERR_CUDA;
            } else {
            /* fork: fp64 transform of the group mean on the side stream, fp32 transforms of the residuals here */
            cudaStream_t A = (cudaStream_t)stream;
            cudaStream_t B = (pl->side && !tspws_cu_profile_enabled() && pl->side != A) ? pl->side : A;
            if (pl->mean_split) {
                if (B != A && (cudaEventRecord(pl->ev_fork, A) != cudaSuccess || cudaStreamWaitEvent(B, pl->ev_fork, 0) != cudaSuccess)) return TSPWS_ERR_CUDA;
                if ((e = tspws_cu_forward_f64(&pl->dev, pl->d_gmean, 1, pl->d_ymean, B))) return TSPWS_ERR_CUDA;
            }
            if ((e = tspws_cu_forward_pertrace(&pl->dev, pl->d_gsum32, pl->ntr_fwd, long_row0, pl->d_ypart, stream))) return TSPWS_ERR_CUDA;
            if (pl->mean_split && B != A &&
                (cudaEventRecord(pl->ev_join, B) != cudaSuccess || cudaStreamWaitEvent(A, pl->ev_join, 0) != cudaSuccess)) return TSPWS_ERR_CUDA;
            }
            if (stop == 3) return 0;
            if (!fused_combine &&
                (e = tspws_cu_phasor_combine(&pl->dev, pl->d_ypart, 1, pl->ntr_fwd, pl->d_rule, pl->mean_split ? pl->d_ymean : NULL,
                                             plan_ps(pl), plan_st(pl), pl->d_zout, stream))) return TSPWS_ERR_CUDA;
        } else {
            if ((e = tspws_cu_forward(&pl->dev, gfin, 1, 1, pl->ntr_fwd, pl->chunk, pl->d_pps, pl->d_pst, stream))) return TSPWS_ERR_CUDA;
            if ((e = tspws_cu_reduce_weight(pl->d_pps, pl->d_pst, 1, pl->dev.ncoef, 1, pl->nchunk, pl->d_rule,
                                            plan_ps(pl), plan_st(pl), pl->d_zout, stream))) return TSPWS_ERR_CUDA;
        }
    } else if (pl->peer_pending) {
        unsigned int *flags = (unsigned int *)((char *)pl->d_peer_region + pl->peer_flag_off);
        const size_t blkd = ((tspws_b200_plan_exchange_bytes(pl) + 15) & ~(size_t)15) / sizeof(double);
        if ((e = tspws_cu_peer_f32(0, (float *)pl->d_exch, (long long)(4 * pl->fr->ncoef), (long long)blkd, pl->d_peer_blocks,
                                   (double *)pl->d_peer_region, flags, pl->d_peer_step, flags + TSPWS_PEER_MAXW, pl->d_peer_flags,
                                   pl->peer_rank, pl->peer_world, stream))) return TSPWS_ERR_CUDA;
        if ((e = tspws_cu_weight(plan_ps(pl), plan_st(pl), pl->acc64, pl->dev.ncoef, 1, pl->d_rule, pl->d_zout, stream))) return TSPWS_ERR_CUDA;
    } else if (pl->summed_across_ranks) {
        if ((e = tspws_cu_weight(plan_ps(pl), plan_st(pl), pl->acc64, pl->dev.ncoef, 1, pl->d_rule, pl->d_zout, stream))) return TSPWS_ERR_CUDA;
    }
    if (stop == 4) return 0;
    e = tspws_cu_inverse(&pl->dev, pl->d_zout, plan_st(pl), pl->acc64, 1, pl->d_rule, pl->d_invscratch,
                         pl->d_inv_tickets, d_tspws, d_ls, stream, pl->side, pl->ev_fork, pl->ev_join);
    return e ? TSPWS_ERR_CUDA : 0;
}

static int plan_stack_direct(tspws_plan_t *pl, float *d_sigall, float *d_ls, float *d_tspws, void *stream)
{
    int e = tspws_b200_plan_reduce(pl, d_sigall, stream);
    if (e) return e;
    return tspws_b200_plan_finish(pl, d_ls, d_tspws, stream);
}

/* One stack = a fixed sequence of launches with fixed arguments, so it is replayed as a CUDA graph
 * (cuts the launch gaps between the short tail kernels).  Capture needs a non-legacy stream; when it
 * is not possible (legacy default stream, profiling on, TSPWS_NO_GRAPH=1) the launches go directly. */
int tspws_b200_plan_stack(tspws_plan_t *pl, float *d_sigall, float *d_ls, float *d_tspws, void *stream)
{
    if (!pl) return TSPWS_ERR_NULL;
    if (pl->graph_broken || !stream || tspws_cu_profile_enabled() || env_long("TSPWS_NO_GRAPH", 0))
        return plan_stack_direct(pl, d_sigall, d_ls, d_tspws, stream);
    cudaStream_t s = (cudaStream_t)stream;
    if (pl->gexec && (pl->g_x != d_sigall || pl->g_ls != d_ls || pl->g_ts != d_tspws)) {
        cudaGraphExecDestroy(pl->gexec);
        pl->gexec = NULL;
    }
    if (!pl->gexec) {
        cudaGraph_t g = NULL;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            pl->graph_broken = 1;
            return plan_stack_direct(pl, d_sigall, d_ls, d_tspws, stream);
        }
        const unsigned long n0 = tspws_cu_launch_count();
        const int e = plan_stack_direct(pl, d_sigall, d_ls, d_tspws, stream);
        const cudaError_t ce = cudaStreamEndCapture(s, &g);
        pl->g_kernels = tspws_cu_launch_count() - n0;
        tspws_cu_launch_count_add(0ul - pl->g_kernels);            /* nothing has run yet */
        if (e || ce != cudaSuccess || !g || cudaGraphInstantiate(&pl->gexec, g, 0) != cudaSuccess) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            pl->gexec = NULL;
            pl->graph_broken = 1;
            return plan_stack_direct(pl, d_sigall, d_ls, d_tspws, stream);
        }
        cudaGraphDestroy(g);
        pl->g_x = d_sigall; pl->g_ls = d_ls; pl->g_ts = d_tspws;
    }
    if (cudaGraphLaunch(pl->gexec, s) != cudaSuccess) { cudaGetLastError(); return TSPWS_ERR_CUDA; }
    tspws_cu_launch_count_add(pl->g_kernels);
    return 0;
}

int tspws_b200_plan_stacks_ptr(tspws_plan_t *pl, const void **d_st, const void **d_ps, unsigned int *K, int *is_f64)
{
    if (!pl) return TSPWS_ERR_NULL;
    if (d_st) *d_st = plan_st(pl);
    if (d_ps) *d_ps = plan_ps(pl);
    if (K) *K = pl->K;
    if (is_f64) *is_f64 = pl->acc64;
    return 0;
}

size_t tspws_b200_plan_ncoef(tspws_plan_t *pl) { return pl ? pl->fr->ncoef : 0; }

/* Host copies (double, interleaved re,im) of the last ST / PS of a plan; returns ncoef. */
long tspws_b200_plan_get_stacks(tspws_plan_t *pl, double *st, double *ps, size_t ncoef)
{
    if (!pl) return TSPWS_ERR_NULL;
    const size_t n = pl->fr->ncoef;
    if (!st && !ps) return (long)n;
    if (ncoef < n) return TSPWS_ERR_ARG;
    const size_t esz = pl->acc64 ? sizeof(double) : sizeof(float);
    void *tmp = malloc(2 * n * esz + 8);
    if (!tmp) return TSPWS_ERR_ALLOC;
    for (int which = 0; which < 2; which++) {
        double *dst = which ? ps : st;
        if (!dst) continue;
        if (cudaMemcpy(tmp, which ? plan_ps(pl) : plan_st(pl), 2 * n * esz, cudaMemcpyDeviceToHost) != cudaSuccess) {
            free(tmp);
            return TSPWS_ERR_CUDA;
        }
        if (pl->acc64) memcpy(dst, tmp, 2 * n * sizeof(double));
        else for (size_t i = 0; i < 2 * n; i++) dst[i] = (double)((float *)tmp)[i];
    }
    free(tmp);
    return (long)n;
}
unsigned int tspws_b200_plan_K(tspws_plan_t *pl) { return pl ? pl->K : 0; }

/* internal accessors used by tspws_main.c */
const tspws_frame_t *tspws_plan_frame(tspws_plan_t *pl) { return pl->fr; }
const tspws_devframe_t *tspws_plan_devframe(tspws_plan_t *pl) { return &pl->dev; }
