/*
 * tspws_internal.h -- shared between the C host code and the CUDA translation unit.
 * Nothing here is part of the public ABI (that is include/tspws_b200.h).
 */
#ifndef TSPWS_INTERNAL_H
#define TSPWS_INTERNAL_H

#include "tspws_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TSPWS_PI 3.14159265358979328   /* the value the reference compiles with (src/ts_pws1f_lib.h:15) */

/* ---- per-scale record uploaded to the device ---- */
typedef struct {
    int   L;         /* filter taps                                   */
    int   c;         /* filter origin                                 */
    int   D;         /* decimation                                    */
    int   Ns;        /* coefficients = ceil(N/D)                      */
    int   coef_off;  /* offset in a ragged coefficient array          */
    int   tap_off;   /* offset in the tap array                       */
    int   G;         /* generic kernel: lanes cooperating on one coefficient (pow2) */
    /* polyphase (register-tiled) kernel: x[(m-A+q)D + rho] * g[qD + rho], q < Qs, rho < D */
    int   lg;        /* log2(lanes over phases): 0..4 when D = 2^lg < 32, 5 when D >= 32, -1: generic kernel */
    int   A;         /* ceil(c / D)                                   */
    int   Qs;        /* taps per phase = ceil((A*D - c + L) / D)      */
    int   g_off;     /* offset of the padded conjugated filter g      */
    int   nsplit;    /* partial-Y slots (phase splits) in per-trace mode */
    int   ypart_off; /* offset of this scale in the partial-Y scratch (per unit of K traces) */
    /* 4-samples-per-thread inverse (D % 4 == 0, N % 4 == 0): filter shifted so that its origin is a
     * multiple of 4 and zero-padded to a multiple of 4 taps */
    int   inv4;      /* 1: eligible                                   */
    int   ci;        /* shifted origin c' = c + (4 - c % 4) % 4        */
    int   Li4;       /* padded length                                  */
    int   wi_off;    /* offset (multiple of 4) in the inverse filter array */
    /* polyphase (register-tiled) inverse, D >= 32 and N % D == 0:
     *   x^[(a D + rho - c') mod N] = gain * sum_t Re( Z[(a - Qi + 1 + t) mod Ns] * gi[t][rho] ),
     *   gi[t][rho] = w'[rho + (Qi - 1 - t) D]  (time-reversed phase rows) */
    int   inv2;      /* 1: eligible                                   */
    int   Qi;        /* taps per phase = ceil(Li4 / D)                 */
    int   gi_off;    /* offset of gi in the reversed filter array      */
    int   tile0;     /* first 64-coefficient tile of this scale in the tile list (tickets of the fused combine) */
    int   pad0;
    double gain;     /* synthesis gain ln2/(2 Cpsi V scale) * D       */
} tspws_scale_t;

/* fp64 mean transform: filters on the long list are split over CTAs of at most 4096 taps; the partial
 * coefficients land in separate slots [split][ncoef] that phasor_combine_kernel adds in slot order */
#define TSPWS_F64_MAXSPLIT 8
#define TSPWS_F64_NSPLIT(L) (((L) + 4095) / 4096 < TSPWS_F64_MAXSPLIT ? ((L) + 4095) / 4096 : TSPWS_F64_MAXSPLIT)

/* ---- forward-transform work item of the generic kernel: a run of coefficients of one scale ---- */
typedef struct {
    int s;           /* scale                           */
    int m0;          /* first coefficient of the run    */
    int nout;        /* coefficients in the run         */
    int pad;
} tspws_fwd_item_t;

/* ---- work item of the polyphase kernel: one warp, one register tile ---- */
typedef struct {
    int s;           /* scale                                              */
    int m0;          /* first coefficient                                  */
    int pb0, pb1;    /* blocks of 32 phases [pb0,pb1) (D >= 32), else 0,1  */
    int split;       /* partial-Y slot written in per-trace mode           */
    int base;        /* ((m0 - A) D) mod N: sample of (tile 0, tap 0, phase 0), so the kernel needs no 64-bit modulo */
    int pad[2];
} tspws_fwd2_item_t;

/* ---- inverse-transform work split: scales [s0,s1) ---- */
typedef struct { int s0, s1; } tspws_inv_group_t;

/* ---- column-sum work item: local traces [t0,t1), all in group `group`, partial slot `slot` ---- */
typedef struct {
    int t0, t1;
    int group;
    int slot;
} tspws_csum_item_t;

/* ---- device-side frame ---- */
typedef struct {
    int S, N, ncoef, ntaps;
    tspws_scale_t     *d_scale;   /* [S]            */
    void              *d_w32;     /* float2[ntaps]  */
    void              *d_w64;     /* double2[ntaps] (NULL unless the fp64 tail is enabled) */
    void              *d_g32;     /* float2[ng]: conjugated, phase-padded filters of the polyphase kernel */
    void              *d_wi32;    /* float2[]: origin-aligned, zero-padded filters of the 4-sample inverse */
    tspws_fwd_item_t  *d_items;   /* [nitems]  generic kernel: all scales   */
    int                nitems;
    int               *h_item_first; /* host: [S+1] first generic item of every scale */
    tspws_fwd_item_t  *d_items_gen; /* [nitems_gen] generic kernel: only scales with lg < 0 */
    int                nitems_gen;
    int                f64_maxl;  /* mean split: filters at least this long are transformed unsplit in fp32 */
    tspws_fwd_item_t  *d_items_short; /* fp64 mean transform: warp items of the scales with L < 1024 */
    int                nitems_short;
    int               *d_long;    /* [nlong][2] (scale, coefficient) of the scales with L >= 1024   */
    int                nlong;
    tspws_fwd2_item_t *d_items2;  /* [nitems2] polyphase kernel              */
    int                nitems2;
    int                fw_lg_begin[7]; /* items2 are sorted by lg: items of lg are [begin[lg], begin[lg+1]) */
    int                fw_smem_per_warp; /* floats of warp-private staging smem of the polyphase kernel */
    int                fw_gmax;   /* float2: most filter taps one pass of the polyphase kernel reads (CTA-shared copy) */
    int                per_trace; /* item list built for per-trace (partial-Y) mode */
    int                ypart_len; /* partial-Y scratch length per trace (complex)   */
    int                max_Ns;
    int               *d_tile16;  /* [ntile16][2]: (scale, first coefficient) of every 16-coefficient tile */
    int                ntile16;
    int               *d_tile_expect; /* [ntile16][2]: writers of a tile in fwd_tail_kernel: fp32 items (x traces), fp64 items */
    tspws_inv_group_t *d_invgrp;  /* [ninv]  scale groups of the fp64 inverse                       */
    int                ninv;
    tspws_inv_group_t *d_invgrp4; /* [ninv4] one group per scale the polyphase inverse does not cover */
    int                ninv4;
    void              *d_gi32;    /* float2[]: reversed, phase-padded filters of the polyphase inverse */
    int               *d_items_inv2; /* [nitems_inv2][4]: (scale, first a, phase block, -)         */
    int                nitems_inv2;
    int                inv2_qmax; /* largest Qi among the scales of the polyphase inverse (shared-memory sizing) */
    int                inv_slots; /* scratch slots of the inverse: max(S, ninv)                         */
} tspws_devframe_t;

/* ---- weighting rule (src/ts_pws1f_lib.c:226-228, :909-984) ---- */
typedef struct {
    int    mode;      /* 0: nu==2 biased, 1: nu==1, 2: general nu, 3: unbiased (nu==2, K>1) */
    double wu;
    double K;         /* Kmax (two-stage) or number of stacked traces */
    double T;         /* mtr (divisor of the linear stack)            */
    /* host-side reciprocals (always filled by tspws_weight_rule): no fp64 division on the device */
    double inv_k;     /* 1 / K                                                                 */
    double inv_f;     /* mode 0: 1/(K K T), mode 1: 1/(K T), mode 2: 1/T, mode 3: 1/((K-1) T)    */
} tspws_weight_t;

/* ---------------------------------------------------------------------------------- */
/* host helpers (C)                                                                    */
/* ---------------------------------------------------------------------------------- */
void tspws_print_params(const t_tsPWS *p, unsigned int nsamp, float dt);
int  tspws_fold_allowed(float beg, float dt, int max);
tspws_weight_t tspws_weight_rule(double wu, int unbiased, unsigned int K, unsigned int T);
unsigned long tspws_binomial(unsigned int n, unsigned int k);

/* ---------------------------------------------------------------------------------- */
/* CUDA launchers (tspws_kernels.cu).  All return cudaError_t as int; stream is a      */
/* cudaStream_t.  "T" variants: is64 selects the double instantiation.                */
/* ---------------------------------------------------------------------------------- */
typedef void *tspws_stream_t;

/* fold (src/ts_pws1f_lib.c:76-86) + mean removal (:160-168) in place, bit-faithful fp32. */
int tspws_cu_prep(float *d_x, int ntr, int N, int do_fold, int do_rm, tspws_stream_t st);

/* Two-stage group sums (src/ts_pws1f_lib.c:866-881): fp32 traces -> fp64 [G][N] group sums
 * (per-item partial slots [nslots][N] combined by the last CTA of each (group, tile), in slot
 * order) and per-(group, tile) totals gtile [G][ntile], ntile = ceil(N/1024).
 * d_tickets: [G][ntile] zero-initialised counters (self re-arming). */
int tspws_cu_colsum(const float *d_x, int N, const tspws_csum_item_t *d_items, int nitems,
                    double *d_part, const int *d_slot_off, unsigned int *d_tickets,
                    double *d_gsum, double *d_gtile, tspws_stream_t st);
/* Mean removal + fold of the group sums (both linear) and the fp32 copy for the transform. */
int tspws_cu_colsum2(const float *d_x, int N, const void *d_parts /*int4[nparts]*/, int nparts, const int *d_grp_t0,
                     double *d_slots, double *d_slot_tot, int ntile, tspws_stream_t st);
int tspws_cu_group_post2(const double *d_slots, const double *d_slot_tot, const int *d_slot_off, int N, int K,
                         int ntile, int mode, int do_rm, int do_fold, int split, double *d_gsum, double *d_gtile,
                         float *d_out32, double *d_gmean, tspws_stream_t st);
/* fp64 transform (generic kernel) of ntr double traces -> [ntr][ncoef] double2 */
int tspws_cu_forward_f64(const tspws_devframe_t *f, const double *d_x, int ntr, void *d_y, tspws_stream_t st);

/* Forward frame transform + phasor / linear accumulation over chunks of traces (generic kernel,
 * any geometry, float or double).
 *   x: [ntr][N] (float, or double when in64), chunk c covers traces [c*chunk, min(ntr,(c+1)*chunk))
 *   part_ps / part_st: [nchunk][ncoef] complex (float2, or double2 when acc64). */
int tspws_cu_forward(const tspws_devframe_t *f, const void *d_x, int in64, int acc64,
                     int ntr, int chunk, void *d_part_ps, void *d_part_st, tspws_stream_t st);
/* fp32 fast path: polyphase register-tiled kernel for the scales it covers + generic kernel for the
 * rest.  Chunk-accumulate mode (f->per_trace == 0): same outputs as tspws_cu_forward. */
/* skip5: the first skip5 items of the D >= 32 list (coarsest scales first) are left out -- the tensor-core
 * transforms cover them (tspws_cu_forward_dense / _band); lg_end: the lists of D = 2^lg < 32 run only for lg < lg_end (5 = all) */
/* generic kernel, ONE fp32 trace, only the items [item0, nitems) (= the scales from some scale on): Y -> d_st, phasor -> d_ps */
int tspws_cu_forward_one(const tspws_devframe_t *f, const float *d_x, int item0, void *d_ps, void *d_st, tspws_stream_t st);
int tspws_cu_forward_fast(const tspws_devframe_t *f, const float *d_x, int ntr, int chunk,
                          void *d_part_ps, void *d_part_st, int skip5, int lg_end, tspws_stream_t st);
/* ---- coarse scales of the single-stage forward transform as one dense GEMM on tcgen05 (tspws_mma.cu) ---- */
size_t tspws_cu_dense_ximg_bytes(int ntr, int N);
size_t tspws_cu_dense_wimg_bytes(int ncol, int N);
int tspws_cu_dense_wimage(const tspws_devframe_t *f, int s0, int ci0, int ncd, void *d_img, tspws_stream_t st);
int tspws_cu_forward_dense(const float *d_x, int ntr, int N, const void *d_wimg, int ncd, void *d_ximg,
                           float *d_tsum, float *d_sumtrace, float *d_part, int ld_part, int col0, tspws_stream_t st);
/* ---- middle octaves (32 <= D < dense D) as BANDED GEMMs on tcgen05: one octave (the voices sharing a decimation D)
 * per descriptor.  With the trace cut into rows of 64 samples, the coefficient m = U P + u (U = 64/D sub-positions per
 * row when D <= 64; one coefficient every E = D/64 rows otherwise) is
 *     Y[m] = sum_q  x_row[(P E') + q] . B_q[(u)]      B_q[r] = conj(w[64 q + r - u D + c])
 * so a group of Pn consecutive row positions x U sub-positions x nv voices x (re|im) = ncols columns is the product of
 * nj = E (Pn - 1) + qmax - qmin + 1 consecutive 64-sample row blocks of the traces with nj small B tiles that are THE
 * SAME for every group of the octave (shift invariance) -- [128 traces x 64] . [64 x ncols], accumulated over the nj
 * row blocks in TMEM.  The A tiles are the dense path's trace image; the epilogue is the dense path's. */
#define TSPWS_BAND_MAXV 8
typedef struct {
    int tile_base;       /* first B tile (unit: one 3-term tile triple) of this octave                  */
    int nj;              /* row blocks per group                                                        */
    int ncols;           /* UMMA N: nv * Pn * U * 2 rounded up to a multiple of 16, <= 128              */
    int rows_per_group;  /* E * Pn: row blocks between consecutive groups                               */
    int qmin;            /* first row block of group 0 (negative: wraps)                                */
    int ngroups;
    int nv, Pn, U, E;    /* column = ((v * Pn + p') * U + u) * 2 + (re | im)                            */
    int Ns, D;
    int coef_rel[TSPWS_BAND_MAXV];   /* coefficient offset of voice v relative to the first band coefficient */
    int s_of[TSPWS_BAND_MAXV];       /* scale index of voice v                                          */
} tspws_band_t;
int tspws_cu_forward_image(const float *d_x, int ntr, int N, void *d_ximg, float *d_tsum, float *d_sumtrace, tspws_stream_t st);
int tspws_cu_tile_sum(const float *d_tsum, int ntr, int N, float *d_sumtrace, tspws_stream_t st);
int tspws_cu_forward_mma_all(const void *d_ximg, int ntr, int N, const void *d_wimg, int ncd_dense, const void *d_bimg,
                             const tspws_band_t *d_bands, int nbands, int ngroups_total, float *d_part, int ld_part,
                             int col0_dense, tspws_stream_t st);
int tspws_cu_band_wimage(const tspws_devframe_t *f, const tspws_band_t *d_bands, int nbands, int ntiles, void *d_img, tspws_stream_t st);
int tspws_cu_forward_band(const void *d_ximg, int ntr, int N, const void *d_bimg, const tspws_band_t *d_bands, int nbands,
                          int ngroups_total, float *d_part, int ld_part, tspws_stream_t st);
int tspws_cu_dense_finish(const float *d_part, int ntr, int ncd, const void *d_ysum, int ci0, void *d_pps0, void *d_pst0,
                          tspws_stream_t st);
/* Per-trace mode (f->per_trace == 1): raw coefficients of each of the ntr traces (split over
 * phase ranges for coarse scales) go to d_ypart [f->ypart_len * ntr] complex fp32 ... */
/* peer exchange (CUDA IPC mappings of the other ranks' exchange blocks), see group_post2_kernel<true> */
int tspws_cu_group_post2_peer(int publish, const double *d_slots, const double *d_slot_tot, const int *d_slot_off, int N, int K,
                              int ntile, int do_rm, int do_fold, int split, double *d_gsum,
                              float *d_out32, double *d_gmean, void *d_regions, double *d_own, const unsigned int *d_flags,
                              unsigned int *d_step /* [2]: step, CTA ticket */, unsigned int *d_err, void *d_peer_flags,
                              int rank, int world, tspws_stream_t st);
/* d_tickets != NULL ([f->ntile16] zeroed counters): the phasor combine is fused in (PS, ST, Zout written by
 * the warp that completes a coefficient tile); NULL: per-trace coefficients only, tspws_cu_phasor_combine follows */
/* single-stage peer exchange of a flat fp32 block (PS | ST sums): publish into every region / wait + sum in rank order */
int tspws_cu_peer_f32(int publish, float *d_buf, long long n, long long block_len_doubles, void *d_regions, double *d_own,
                      const unsigned int *d_flags, unsigned int *d_step, unsigned int *d_err, void *d_peer_flags,
                      int rank, int world, tspws_stream_t st);
int tspws_cu_forward_tail(const tspws_devframe_t *f, const float *d_x, int ntr, int long_row0, void *d_ypart,
                          const double *d_x64, void *d_y64, unsigned int *d_tickets, const tspws_weight_t *d_rule,
                          void *d_ps, void *d_st, void *d_zout, tspws_stream_t st);
/* long_row0 > 0: scales with L >= f->f64_maxl read their traces from row long_row0 on (mean split) */
int tspws_cu_forward_pertrace(const tspws_devframe_t *f, const float *d_x, int ntr, int long_row0, void *d_ypart,
                              tspws_stream_t st);
/* ... and are combined here: for segment g (ntr = nseg*seg traces), PS = sum phasor(Y), ST = sum Y,
 * Zout = ST * weight(PS).  Outputs [nseg][ncoef] complex fp32. */
int tspws_cu_phasor_combine(const tspws_devframe_t *f, const void *d_ypart, int nseg, int seg,
                            const tspws_weight_t *d_rules, const void *d_ymean /*double2 [nseg][ncoef] or NULL*/,
                            void *d_ps, void *d_st, void *d_zout, tspws_stream_t st);
/* Sum the chunk partials in segments of `seg` chunks -> PS, ST per segment, and the weighted
 * coefficients Zout = ST * weight(PS).  nseg segments; outputs [nseg][ncoef] complex. */
int tspws_cu_reduce_weight(const void *d_part_ps, const void *d_part_st, int acc64, int ncoef,
                           int nseg, int seg, const tspws_weight_t *d_rules /*[nseg] device*/,
                           void *d_ps, void *d_st, void *d_zout, tspws_stream_t st);
/* d_zout layout: fp32 -> float4 per coefficient (Zout.re, Zout.im, ST.re, ST.im) ("zs", what the
 * fp32 inverse reads with one 16-byte load); fp64 -> double2 Zout only. */
/* Weighting only (PS, ST already reduced, e.g. after an all-reduce). */
int tspws_cu_weight(const void *d_ps, const void *d_st, int acc64, int ncoef, int nseg,
                    const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st);
/* Inverse frame transform of nseg coefficient sets (two per set: Zout -> tsPWS, ST -> ls).
 *   d_out_ts/d_out_ls: float [nseg][N];  ls is divided by rules[seg].T like src/ts_pws1f_lib.c:234.
 *   d_st may be NULL (no linear stack wanted).  d_scratch: [2][ninv][nseg][N] of T;
 *   d_tickets: [nseg][ceil(N/256)] zero-initialised counters (self re-arming). */
/* side / ev_fork / ev_join (cudaStream_t, cudaEvent_t x 2; all NULL: one stream): the two independent halves
 * of the fp32 inverse run concurrently on `side` (fork / join, also inside a graph capture). */
int tspws_cu_inverse(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int acc64,
                     int nseg, const tspws_weight_t *d_rules, void *d_scratch, unsigned int *d_tickets,
                     float *d_out_ts, float *d_out_ls, tspws_stream_t st, tspws_stream_t side, void *ev_fork, void *ev_join);

/* device frame construction (tspws_plan.c) */
int  tspws_devframe_build(tspws_devframe_t *d, const tspws_frame_t *f, int want64, int per_trace);
void tspws_devframe_free(tspws_devframe_t *d);

/* ---- resampling / convergence support ---- */
/* d_cnt (optional, [M][ntr] like d_gid): every selected trace is added d_cnt times (bootstrap with replacement) */
int tspws_cu_regroup(const float *d_x, int N, int ntr, const short *d_gid /*[M][ntr]*/, const signed char *d_cnt, int M, int G,
                     float *d_out /*[M][G][N]*/, tspws_stream_t st);
int tspws_cu_group_total(const float *d_gs, int M, int G, int N, const double *d_scale /*[M]*/, float *d_out /*[M][N]*/,
                         tspws_stream_t st);
int tspws_cu_coef_matrix(const tspws_devframe_t *f, const void *d_ypart, int ntr, void *d_Y, void *d_P, tspws_stream_t st);
int tspws_cu_select_sum(const void *d_P, int T, int ncoef, const signed char *d_sel, int sel_ld, int t0, int M,
                        void *d_acc /*double2 [M][ncoef]*/, tspws_stream_t st);
int tspws_cu_weight_mixed(const void *d_ps64, const void *d_st32, int ncoef, int nseg, const tspws_weight_t *d_rules,
                          void *d_zout, tspws_stream_t st);
/* the same with PS in float2 (the tensor-core contraction's output) */
int tspws_cu_weight_mixed32(const void *d_ps32, const void *d_st32, int ncoef, int nseg, const tspws_weight_t *d_rules,
                            void *d_zout, tspws_stream_t st);
/* ---- tensor-core resampling contraction (tspws_mma.cu): tile images + tcgen05 GEMM ---- */
size_t tspws_cu_mma_a_image_bytes(int M, int T);
size_t tspws_cu_mma_b_image_bytes(int N, int T, int nterm);
/* counts [M][ld] (signed char; > 0 counts, anything else 0), traces [0, T) -> bf16 A image */
int tspws_cu_count_image(const signed char *d_cnt, int ld, int M, int T, void *d_img, tspws_stream_t st);
/* phasors P [nt][ncoef] float2 of traces [t0, t0 + nt), t0 % 64 == 0 -> their k blocks of the bf16 B image (nterm = 2 or 3) */
int tspws_cu_phasor_image(const void *d_P, int ncoef, int nt, int t0, int kblocks_total, int nterm, void *d_img, tspws_stream_t st);
/* out [M][ld_out] fp32 (columns [0, N)) = A image . B image^T over kblocks k blocks of 64 */
int tspws_cu_resample_mma(const void *d_a_img, const void *d_b_img, int nterm, int M, int N, int kblocks,
                          float *d_out, size_t ld_out, tspws_stream_t st);
int tspws_cu_mma_selftest(const signed char *h_cnt, const float *h_P, int M, int T, int ncoef, int nterm, float *h_out);
int tspws_cu_prefix_weight(const void *d_Y, const void *d_P, int ncoef, int B, void *d_st_state, void *d_ps_state,
                           const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st);
int tspws_cu_simil(const float *d_cand, int B, int N, const float *d_ref, double *d_sim, double *d_mis, tspws_stream_t st);
int tspws_cu_lsconv(const float *d_x, int ntr, int N, const float *d_ref, float *d_steps, double *d_part,
                    double *d_sim, double *d_mis, tspws_stream_t st);

unsigned long tspws_cu_launch_count(void);
void tspws_cu_profile_mark(tspws_stream_t st);
void tspws_cu_profile_launched(const char *name, tspws_stream_t st);
int  tspws_cu_profile_enabled(void);
void tspws_cu_launch_count_add(unsigned long n);
int  tspws_cu_profile_enable(int on);
int  tspws_cu_profile_reset(void);
int  tspws_cu_profile_read(char *buf, int cap);

#ifdef __cplusplus
}
#endif
#endif

void tspws_plan_set_cs2_hint(int ctas);
