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
    int   G;         /* lanes cooperating on one coefficient (pow2)   */
    int   pad;
    double gain;     /* synthesis gain ln2/(2 Cpsi V scale) * D       */
} tspws_scale_t;

/* ---- forward-transform work item: a run of coefficients of one scale ---- */
typedef struct {
    int s;           /* scale                           */
    int m0;          /* first coefficient of the run    */
    int nout;        /* coefficients in the run         */
    int pad;
} tspws_fwd_item_t;

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
    tspws_fwd_item_t  *d_items;   /* [nitems]       */
    int                nitems;
    tspws_inv_group_t *d_invgrp;  /* [ninv]         */
    int                ninv;
} tspws_devframe_t;

/* ---- weighting rule (src/ts_pws1f_lib.c:226-228, :909-984) ---- */
typedef struct {
    int    mode;      /* 0: nu==2 biased, 1: nu==1, 2: general nu, 3: unbiased (nu==2, K>1) */
    double wu;
    double K;         /* Kmax (two-stage) or number of stacked traces */
    double T;         /* mtr (divisor of the linear stack)            */
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

/* Two-stage column sums (src/ts_pws1f_lib.c:866-881): per-item partial sums of fp32 traces into
 * fp64 slots [nslots][N]; optional per-(trace,tile) row partial sums for the mean. */
int tspws_cu_colsum(const float *d_x, int N, const tspws_csum_item_t *d_items, int nitems,
                    double *d_part, float *d_rowpart, int rowpart_ld, tspws_stream_t st);
/* Combine partial slots into group sums and subtract the per-group mean correction
 * (sum over the group's local traces of the float-cast trace mean) when d_rowpart != NULL. */
int tspws_cu_colsum_finalize(const double *d_part, const int *d_slot_off /*[G+1]*/, int G, int N,
                             const float *d_rowpart, int rowpart_ld, int rowpart_n,
                             const int *d_grp_t0 /*[G+1] local trace range of each group*/,
                             double *d_meancorr /*[G] scratch*/, double *d_gsum /*[G][N]*/,
                             tspws_stream_t st);
/* Fold the group sums (linear) and emit the fp32 copy read by the fp32 transform. */
int tspws_cu_group_post(double *d_gsum, float *d_gsum32, int G, int N, int do_fold, tspws_stream_t st);

/* Forward frame transform + phasor / linear accumulation over chunks of traces.
 *   x: [ntr][N] (float, or double when in64), chunk c covers traces [c*chunk, min(ntr,(c+1)*chunk))
 *   part_ps / part_st: [nchunk][ncoef] complex (float2, or double2 when acc64). */
int tspws_cu_forward(const tspws_devframe_t *f, const void *d_x, int in64, int acc64,
                     int ntr, int chunk, void *d_part_ps, void *d_part_st, tspws_stream_t st);
/* Sum the chunk partials in segments of `seg` chunks -> PS, ST per segment, and the weighted
 * coefficients Zout = ST * weight(PS).  nseg segments; outputs [nseg][ncoef] complex. */
int tspws_cu_reduce_weight(const void *d_part_ps, const void *d_part_st, int acc64, int ncoef,
                           int nseg, int seg, const tspws_weight_t *d_rules /*[nseg] device*/,
                           void *d_ps, void *d_st, void *d_zout, tspws_stream_t st);
/* Weighting only (PS, ST already reduced, e.g. after an all-reduce). */
int tspws_cu_weight(const void *d_ps, const void *d_st, int acc64, int ncoef, int nseg,
                    const tspws_weight_t *d_rules, void *d_zout, tspws_stream_t st);
/* Inverse frame transform of nseg coefficient sets (two per set: Zout -> tsPWS, ST -> ls).
 *   d_out_ts/d_out_ls: float [nseg][N];  ls is divided by rules[seg].T like src/ts_pws1f_lib.c:234.
 *   d_st may be NULL (no linear stack wanted).  d_scratch: [2][ninv][nseg][N] of T. */
int tspws_cu_inverse(const tspws_devframe_t *f, const void *d_zout, const void *d_st, int acc64,
                     int nseg, const tspws_weight_t *d_rules, void *d_scratch,
                     float *d_out_ts, float *d_out_ls, tspws_stream_t st);

unsigned long tspws_cu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
