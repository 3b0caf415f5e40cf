/*
 * tspws_b200.h -- C ABI of libtspws_b200.so, the B200-native ts-PWS stacking path.
 *
 * PART 1 is the drop-in boundary.  A program compiled against the reference's
 * own header (reference src/ts_pws1f_lib.h) links against this library
 * unchanged: the three parameter blocks below have the same x86-64 layout as
 * the reference's t_tsPWS (src/ts_pws1f_lib.h:19-57, 184 B), t_hdr (:59-77,
 * 108 B), t_tsPWS_out (:79-94, 104 B) and t_data (:96-102, 144 B) -- checked by
 * the _Static_asserts at the bottom -- and the one entry point has the same
 * name and prototype as src/ts_pws1f_lib.h:104.
 *
 * PART 2 are additive exports that the reference does not have (coherence
 * map, explicit resampling plans, device-resident planning API used by the
 * benchmark and the multi-GPU driver).  Plain pointers and sizes only.
 *
 * There is no CPU fallback: every compute entry point returns
 * TSPWS_ERR_CUDA when no sm_100 device / driver is usable.
 */
#ifndef TSPWS_B200_H
#define TSPWS_B200_H

#include <stddef.h>
#include <stdint.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* PART 1 -- drop-in boundary (layout = reference src/ts_pws1f_lib.h:19-102)  */
/* ------------------------------------------------------------------------- */
#ifndef TSPWS_LIB /* the reference header defines the same tags; do not clash */

typedef struct {
    int          type;        /* -1 Morlet, -2 exact Morlet, -3 complex Mexican hat            */
    unsigned int uni;         /* 1: no decimation (uniform sampling of every scale)            */
    unsigned int J;           /* octaves (0 = derive)                                          */
    unsigned int V;           /* voices per octave                                             */
    double       s0;          /* finest scale                                                  */
    double       b0;          /* translation step at scale 1                                   */
    double       w0;          /* mother-wavelet centre frequency                               */
    double       wu;          /* PWS power nu                                                  */
    double       fmin;        /* lowest analysed frequency in Hz (0 = derive from J)           */
    double       Q;           /* quality factor, used when w0set == 1                          */
    double       cycle;       /* number of cycles, used when w0set == 2                        */
    int          w0set;       /* 0 none, 1 Q, 2 cycles, 3 cyc, 4 w0                            */
    int          lrm;         /* remove the mean of every trace                                */
    int          bin;         /* (front-end) input is a sac2bin file                           */
    int          lkinst;      /* (front-end)                                                   */
    int          lVfix;       /* V was set by the user: keep it                                */
    int          ls0fix;      /* s0 was set by the user: keep it                               */
    int          lb0fix;      /* b0 was set by the user: keep it                               */
    int          verbose;
    int          fold;        /* average causal and acausal lags                               */
    int          unbiased;    /* unbiased coherence (honoured for wu == 2 only)                */
    int          convergence;
    unsigned int subsmpl_N;   /* number of random sub-samples                                  */
    double       subsmpl_p;   /* fraction of traces per sub-sample                             */
    unsigned int jackknife_n; /* day-of-year bins                                              */
    unsigned int jackknife_d; /* bins deleted per jackknife realisation                        */
    unsigned int obin;        /* (front-end)                                                   */
    int          AllSteps;
    unsigned int Nmax;        /* use the first Nmax traces only (0 = all)                      */
    unsigned int Kmax;        /* two-stage stack with Kmax groups (0 = single stage)           */
    char        *kinst;       /* (front-end)                                                   */
    char        *filein;      /* (front-end)                                                   */
    char        *fileout;     /* (front-end)                                                   */
    char        *fileconv;    /* (front-end)                                                   */
} t_tsPWS;

typedef struct {
    int          max;         /* samples per trace    */
    unsigned int mtr;         /* number of traces     */
    float        evla, evlo, stla, stlo, stel;
    float        dt;          /* sampling period (s)  */
    float        beg;         /* time of sample 0 (s) */
    char         net1[9], sta1[9], loc1[9], chn1[9];
    char         net2[9], sta2[9], loc2[9], chn2[9];
} t_hdr;

typedef struct {
    float        *ls;            /* [max] linear stack (through the frame)     */
    float        *tsPWS;         /* [max] time-scale phase-weighted stack      */
    double       *ls_sim;        /* [mtr] convergence curves (optional)        */
    double       *tsPWS_sim;
    double       *ls_misfit;
    double       *tsPWS_misfit;
    float        *ls_steps;      /* [mtr][max] (optional)                      */
    float        *tsPWS_steps;
    float       **ls_subsmpl;    /* [M] rows of [max] (resampling, optional)   */
    float       **tsPWS_subsmpl;
    unsigned int *mtr_subsmpl;   /* [M] traces used by each realisation        */
    unsigned int  M;
    unsigned int  N;
    unsigned int  mtr;
} t_tsPWS_out;

typedef struct {
    float  *sigall;              /* [mtr][max] row-major traces                */
    time_t *time;                /* [mtr] start times (jackknife), may be NULL */
    float  *lag0;
    t_hdr   hdr;
    float  *reference;           /* optional convergence reference [max]       */
} t_data;

#endif /* TSPWS_LIB */

/* Replaces reference tspws_main (src/ts_pws1f_lib.h:104, body src/ts_pws1f_lib.c:48-352).
 * Same arguments, same mutation of *tspws (derived w0,V,b0,s0,J,fmin,fold), same outputs
 * and return convention (0 ok, -1 NULL argument, 4 allocation failure or unusable wavelet family);
 * additionally TSPWS_ERR_CUDA (77) when the GPU path is unusable. Host buffers in, host buffers out.
 * Wavelet types: the complex families -1, -2, -3 (all the reference's front-ends reach); others return 4. */
int tspws_main(t_tsPWS *tspws, t_tsPWS_out *out, t_data *in);

/* Additive (BASELINE configs[4]): npairs independent stacks (station pairs) with one
 * parameter block; same results as tspws_main(tspws, &outs[i], &ins[i]) called in turn -- what the
 * reference's users do with a shell loop over ts_pws (examples/example.sh).  Pairs that share one
 * geometry (max, mtr, dt, beg) and ask for the main stack only are pipelined over two device slots
 * so that the host-to-device copy of pair i+1 runs under the kernels of pair i.  *tspws is mutated
 * as by the first call; returns the first non-zero code. */
int tspws_b200_main_batch(t_tsPWS *tspws, t_tsPWS_out *outs, t_data *ins, unsigned int npairs);

/* ------------------------------------------------------------------------- */
/* PART 2 -- additive exports                                                 */
/* ------------------------------------------------------------------------- */
#define TSPWS_ERR_NULL   (-1)
#define TSPWS_ERR_ALLOC  4
#define TSPWS_ERR_ARG    5
#define TSPWS_ERR_CUDA   77

/* Frame geometry as the reference's CreateWaveletFamily would build it
 * (src/FWTa/wavelet_def_v7.c:204-339).  Host-only, needs no GPU. */
typedef struct {
    int          type;
    unsigned int S;            /* number of scales J*V                              */
    unsigned int N;            /* trace length                                      */
    unsigned int V;
    double       Cpsi;
    size_t       ncoef;        /* sum_s Ns[s]                                       */
    size_t       ntaps;        /* sum_s L[s]                                        */
    double      *scale;        /* [S]                                               */
    unsigned int *L;           /* [S] filter length                                 */
    int          *center;      /* [S] filter origin                                 */
    unsigned int *D;           /* [S] decimation                                    */
    unsigned int *Ns;          /* [S] coefficients per scale = ceil(N/D)            */
    size_t       *coef_off;    /* [S+1] offset of scale s in a ragged coef array    */
    size_t       *tap_off;     /* [S+1] offset of scale s in the tap array          */
    double       *w;           /* [2*ntaps] analysis filters, interleaved re,im     */
} tspws_frame_t;

/* Derive V,b0,s0,J,w0,fmin exactly like src/ts_pws1f_lib.c:91-124 (mutates *p). */
int  tspws_b200_derive_params(t_tsPWS *p, unsigned int nsamp, float dt);
tspws_frame_t *tspws_b200_frame_create(int type, unsigned int J, unsigned int V, unsigned int N,
                                       double s0, double b0, double w0, int uni);
void tspws_b200_frame_destroy(tspws_frame_t *f);

/* Geometry of the banded tensor-core transform of the middle octaves, host only: one record of 28 ints
 * per octave = { tile_base, nj, ncols, rows_per_group, qmin, ngroups, nv, Pn, U, E, Ns, D, coef_rel[8], s_of[8] }
 * (csrc/tspws_internal.h: tspws_band_t).  For the CPU tests, which rebuild the B tiles in numpy and compare the banded
 * product with the oracle's analysis.  Returns the number of records. */
int tspws_b200_band_layout(const tspws_frame_t *fr, long dense_D, long band_D, int *out, int max, int *s_first);

/* Resampling plans (host, integer work).
 * tspws_b200_subsampling_plan: reference SubsamplingPlan (src/ts_pws1f_lib.c:355-383), uses rand().
 * tspws_b200_jackknife_plans : reference JackknifePlans  (src/ts_pws1f_lib.c:385-430). */
int tspws_b200_subsampling_plan(char *sel, unsigned int J, unsigned int K);
int tspws_b200_jackknife_plans(char *sel /*[C][mtr]*/, const time_t *time, unsigned int mtr,
                               unsigned int d, unsigned int n, unsigned int C);

/* After a successful tspws_main: the phase coherence map |PS|/K of the main stack,
 * ragged [sum_s Ns], scale-major (finest first) = order of the reference's t_CWTvar.d[0]
 * (src/FWTa/wavelet_mem_v7.c:78-106).  Returns the number of coefficients (call with
 * coh == NULL to query), or a negative error. */
long tspws_b200_get_coherence(double *coh, size_t n);
/* Same for the complex linear (ST) and phase (PS) stacks, interleaved re,im, [2*ncoef]. */
long tspws_b200_get_stacks(double *st, double *ps, size_t ncoef);

/* Explicit selection matrix for the next resampling call(s) of tspws_main
 * (sel[M][mtr] of 0/1, copied).  Lets the CPU oracle and the GPU path use the SAME
 * realisations (the reference draws them from rand(), src/ts_pws1f_lib.c:522,640).
 * Pass NULL to return to rand()-driven plans.
 * Entries > 1 are COUNTS (bootstrap with replacement, an extension: SURVEY 8f rank 3): realisation m stacks trace t
 * sel[m][t] times, K_m = sum_t sel[m][t]; single-stage stacks only (a two-stage call with counts returns
 * TSPWS_ERR_ARG).  Counts up to 127 per entry. */
int tspws_b200_set_selection(const char *sel, unsigned int M, unsigned int mtr);

/* Options (all default 0):
 *  TSPWS_OPT_WRITEBACK  1: copy the folded / mean-removed traces back into in->sigall
 *                          like the reference's in-place mutation (src/ts_pws1f_lib.c:77-86,167).
 *  TSPWS_OPT_FP64_TAIL  1: run the two-stage transforms in fp64 instead of fp32; 2: also the per-trace transforms of a
 *                       single-stage stack (generic kernels, slow: a diagnostic for fp32 conditioning).
 *  TSPWS_OPT_DEVICE     n: CUDA device ordinal to use.
 *  TSPWS_OPT_SS_JACKKNIFE 1: EXTENSION: compute the single-stage delete-d jackknife,
 *                          which the reference leaves as an empty stub (src/ts_pws1f_lib.c:711-716): realisation
 *                          c keeps the K_c traces outside its d deleted bins, weights with (K_c, K_c) like
 *                          tspws_subsmpl_float, ls_c in the time domain; writes mtr_subsmpl[c] = K_c. */
enum { TSPWS_OPT_WRITEBACK = 1, TSPWS_OPT_FP64_TAIL = 2, TSPWS_OPT_DEVICE = 3, TSPWS_OPT_SS_JACKKNIFE = 4 };
int tspws_b200_set_option(int option, long value);
long tspws_b200_get_option(int option);

/* Library state. */
int  tspws_b200_device_count(void);          /* <0: CUDA unusable                       */
void tspws_b200_shutdown(void);              /* frees cached plans / device workspaces  */
const char *tspws_b200_version(void);
unsigned long tspws_b200_kernel_launches(void); /* kernels launched by this library so far */
/* Optional in-stream timing of every kernel with CUDA events (used by bench.py for the roofline
 * line): enable, run, then read "kernel_name launches total_ms" lines.  Off by default. */
int tspws_b200_profile(int enable);
int tspws_b200_profile_reset(void);
int tspws_b200_profile_read(char *buf, int cap);

/* ---- data formats on either side of the path, no sacio.a needed ----
 * tspws_b200_load_sac_list : reference ReadData, SAC-list branch (src/ts_pws1f.c:609-722)
 * tspws_b200_load_bin      : reference ReadData, `bin` branch (src/ts_pws1f.c:586-608,664-675)
 * tspws_b200_save_bin      : the container the reference's sac2bin writes (src/sac2bin.c:192-195)
 * tspws_b200_write_sac     : reference wrsac (src/ts_pws1f.c:725-763)
 * Traces are placed in page-locked memory when `pinned` and CUDA are available. */
int  tspws_b200_load_sac_list(const char *listfile, t_data *in, int pinned);
int  tspws_b200_load_bin(const char *binfile, t_data *in, int pinned);
int  tspws_b200_save_bin(const char *binfile, const t_data *in, const char *method);
/* resampled stacks as the reference's wrbin writes them (src/ts_pws1f.c:766-823): header | time_t[M] holding the
 * traces behind every row | data[M][npts]; rows[m] + n0 is the first sample of row m */
int  tspws_b200_save_rows_bin(const char *binfile, float *const *rows, unsigned int M, int n0, int npts, const t_hdr *hdr,
                              float beg, const char *method, const unsigned int *traces_per_row);
int  tspws_b200_write_sac(const char *path, const float *y, const t_hdr *hdr, int npts, float beg,
                          const char *kinst, float user0);
long tspws_b200_read_sac(const char *path, float *data, long maxpts, float *dt, float *beg, void *header632);
void tspws_b200_free_data(t_data *in);

/* ---- device-resident API (inputs already in HBM; used by bench.py and multi-GPU) ---- */
typedef struct tspws_plan tspws_plan_t;

/* Builds the frame, uploads the filter bank, allocates workspaces for stacks of up to
 * `mtr_local` traces of `max` samples resident on the current device.  `p` must already be
 * derived (tspws_b200_derive_params).  mtr_total is the global number of traces of the
 * stack (== mtr_local on one GPU; > mtr_local when traces are sharded across ranks) and
 * trace_offset the global index of this rank's first trace. */
tspws_plan_t *tspws_b200_plan_create(const t_tsPWS *p, int max, unsigned int mtr_local,
                                     unsigned int mtr_total, unsigned int trace_offset);
void tspws_b200_plan_destroy(tspws_plan_t *plan);

/* Phase A (per rank): reduce the local traces to the small "exchange" block
 *   two-stage : fp64 group sums      [Kmax][max]            (+ per-trace mean correction)
 *   single    : fp32 phasor sums     [2*ncoef]  + fp64 linear time-domain sum [max]
 * d_sigall: device pointer [mtr_local][max] fp32; on the single-stage path it is folded /
 * mean-removed IN PLACE like the reference mutates in->sigall (src/ts_pws1f_lib.c:77-86,167).
 * stream: cudaStream_t as void*. */
int tspws_b200_plan_reduce(tspws_plan_t *plan, float *d_sigall, void *stream);
/* The exchange block to all-reduce (sum) across ranks between phase A and B. */
void  *tspws_b200_plan_exchange_ptr(tspws_plan_t *plan);
size_t tspws_b200_plan_exchange_bytes(tspws_plan_t *plan);
int    tspws_b200_plan_exchange_is_f64(tspws_plan_t *plan);
/* Use a caller-owned device buffer (>= exchange_bytes) as the exchange block, so that e.g.
 * torch.distributed / NCCL can all-reduce it in place; then tell the plan it has been summed. */
int    tspws_b200_plan_bind_exchange(tspws_plan_t *plan, void *d_buf);
int    tspws_b200_plan_mark_exchanged(tspws_plan_t *plan);
/* Phase B (replicated): transforms of the group sums (two-stage), weighting, two inverse
 * frame transforms.  d_ls / d_tspws: device pointers [max] fp32. */
/* Peer exchange instead of a collective (two-stage plans, ranks of ONE node, one process per GPU): every
 * rank exports the CUDA IPC handle (64 bytes) of its exchange region, the handles are gathered by any
 * means (bench.py: torch.distributed.all_gather) and imported; from then on plan_reduce publishes the rank's
 * block into every rank's region (stores over NVLink / NVSwitch) and plan_finish sums the blocks of all
 * ranks inside its first kernel -- no all-reduce call, no plan_mark_exchanged, and tspws_b200_plan_stack
 * replays the whole multi-GPU step as one CUDA graph.  Ranks must call the same sequence of steps.
 * plan_peer_error: 1 if a wait for a peer ever timed out (~4 s); the stacks of that step are then NaN (never
 * a silent partial sum). */
int tspws_b200_plan_peer_export(tspws_plan_t *plan, int world, void *handle64);
int tspws_b200_plan_peer_import(tspws_plan_t *plan, int rank, int world, const void *handles /* world x 64 bytes */);
/* Same-process variant (several plans in one process -- e.g. a test that runs all the shards of a node on one
 * GPU, or one host thread per GPU with peer access enabled): device pointers instead of IPC handles.
 * plan_peer_region returns the plan's exchange region (NULL when the plan cannot take part: not two-stage,
 * max % 4 != 0 or no local traces -- decide collectively and fall back to the all-reduce path together);
 * plan_peer_import_ptrs takes the `world` regions in rank order. */
void *tspws_b200_plan_peer_region(tspws_plan_t *plan, int world);
int   tspws_b200_plan_peer_import_ptrs(tspws_plan_t *plan, int rank, int world, void *const *regions);
int tspws_b200_plan_peer_error(tspws_plan_t *plan);
/* Teardown order (CUDA IPC rule: never free an exported allocation that another process still maps):
 * every rank calls plan_peer_close, the ranks synchronise (barrier), then every rank calls plan_destroy. */
int tspws_b200_plan_peer_close(tspws_plan_t *plan);
/* The reduce step of an fp32 single-stage plan in PARTS of the shard, for callers that are still copying the traces to
 * the device: transform the part that has arrived while the next one is on the bus (tspws_main does exactly that).
 * plan_part_align: parts start at multiples of this many traces and all but the last are a multiple of it long
 * (0: the plan does not take parts -- two-stage or fp64 plans).  d_part = address of trace t0.  After the last part:
 * plan_reduce_done, then the exchange (if any) and plan_finish as usual.  plan_reduce == one part + done. */
unsigned int tspws_b200_plan_part_align(tspws_plan_t *plan);
int tspws_b200_plan_reduce_part(tspws_plan_t *plan, float *d_part, unsigned int t0, unsigned int nt, void *stream);
int tspws_b200_plan_reduce_done(tspws_plan_t *plan, void *stream);
int tspws_b200_plan_finish(tspws_plan_t *plan, float *d_ls, float *d_tspws, void *stream);
/* A+B on one GPU. */
int tspws_b200_plan_stack(tspws_plan_t *plan, float *d_sigall, float *d_ls, float *d_tspws,
                          void *stream);
/* Device pointers to the last ST / PS (complex interleaved [2*ncoef]; fp32, or fp64 when
 * *is_f64), and the divisor K of the coherence |PS|/K. */
int tspws_b200_plan_stacks_ptr(tspws_plan_t *plan, const void **d_st, const void **d_ps,
                               unsigned int *K, int *is_f64);
size_t tspws_b200_plan_ncoef(tspws_plan_t *plan);
/* Host copies (double, interleaved re,im, [2*ncoef]) of the last ST / PS of a plan (either may
 * be NULL); returns ncoef.  K = divisor of the coherence |PS|/K. */
long tspws_b200_plan_get_stacks(tspws_plan_t *plan, double *st, double *ps, size_t ncoef);
unsigned int tspws_b200_plan_K(tspws_plan_t *plan);

/* A stream of independent stacks of one geometry (what a data set run of src/ts_pws1f.c is: one stack per station
 * pair / component), software-pipelined on the GPU: the HBM-bound group sums of stack i+1 run while the latency-bound
 * tail (transforms of the group sums, phase stack, inverse) of stack i finishes.  `depth` plans (3-4) are the buffers;
 * depth = 1: no overlap, one stack at a time as tspws_b200_plan_stack runs it (the overlap does not pay for every
 * frame: +20 % for the 44-scale Morlet headline, a loss for the Mexican hat -- DESIGN.md 4.1b).
 * depth = 0: AUTO (unsharded stacks only): the first 8 stacks run pipelined (4 plans), the next 8 one at a time, both
 * timed on the device, and the faster way serves the rest; tspws_b200_pipe_mode tells which (0 = still measuring,
 * 1 = pipelined, 2 = one at a time; ms[0], ms[1] = the two measurements in ms per stack).
 * submit enqueues one stack (device pointers; any device pointers, they may change from stack to stack) and
 * returns at once; sync waits for everything submitted.  Each stack is computed exactly as tspws_b200_plan_stack
 * computes it.  pipe_stream(p, 0 | 1): the cudaStream_t the sums run on | the results arrive on (to order or time work against).
 * pipe_plan(p, k): plan k, k = (submit index) % depth, for tspws_b200_plan_get_stacks after a sync. */
typedef struct tspws_pipe tspws_pipe_t;
tspws_pipe_t *tspws_b200_pipe_create(const t_tsPWS *p, int max, unsigned int mtr_local, unsigned int mtr_total,
                                     unsigned int trace_offset, int depth);   /* shard arguments as tspws_b200_plan_create */
int   tspws_b200_pipe_submit(tspws_pipe_t *pipe, float *d_sigall, float *d_ls, float *d_tspws);
int   tspws_b200_pipe_sync(tspws_pipe_t *pipe);
void *tspws_b200_pipe_stream(tspws_pipe_t *pipe, int which);
tspws_plan_t *tspws_b200_pipe_plan(tspws_pipe_t *pipe, int k);
int   tspws_b200_pipe_mode(tspws_pipe_t *pipe, float *ms);
void  tspws_b200_pipe_destroy(tspws_pipe_t *pipe);

/* ------------------------------------------------------------------------- */
/* ABI checks against the reference layout (probed with the reference header) */
/* ------------------------------------------------------------------------- */
#if defined(__x86_64__) && !defined(__cplusplus) && !defined(TSPWS_LIB)
_Static_assert(sizeof(t_tsPWS) == 184, "t_tsPWS must be 184 B");
_Static_assert(offsetof(t_tsPWS, s0) == 16 && offsetof(t_tsPWS, wu) == 40, "t_tsPWS doubles");
_Static_assert(offsetof(t_tsPWS, w0set) == 72 && offsetof(t_tsPWS, subsmpl_N) == 116, "t_tsPWS ints");
_Static_assert(offsetof(t_tsPWS, subsmpl_p) == 120 && offsetof(t_tsPWS, Kmax) == 148, "t_tsPWS tail");
_Static_assert(offsetof(t_tsPWS, kinst) == 152 && offsetof(t_tsPWS, fileconv) == 176, "t_tsPWS ptrs");
_Static_assert(sizeof(t_hdr) == 108, "t_hdr must be 108 B");
_Static_assert(sizeof(t_tsPWS_out) == 104 && offsetof(t_tsPWS_out, M) == 88, "t_tsPWS_out");
_Static_assert(sizeof(t_data) == 144 && offsetof(t_data, hdr) == 24 && offsetof(t_data, reference) == 136, "t_data");
#endif

#ifdef __cplusplus
}
#endif
#endif /* TSPWS_B200_H */
