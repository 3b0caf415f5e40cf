/*
 * tspws_params.c -- host-side parameter derivation, frame (filter bank) construction
 * and resampling plans.  Plain C, no GPU needed.
 *
 * Written from the mathematics in SURVEY.md Appendix A; every function names the
 * reference lines whose observable behaviour it reproduces.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tspws_internal.h"

/* ------------------------------------------------------------------------------------ */
/* Parameter derivation -- reference src/ts_pws1f_lib.c:91-124                           */
/* ------------------------------------------------------------------------------------ */
int tspws_b200_derive_params(t_tsPWS *p, unsigned int nsamp, float dt)
{
    if (!p) return TSPWS_ERR_NULL;
    const double ln2 = log(2.);

    if (p->fmin != 0 && p->fmin < 1 / (dt * nsamp)) {     /* float product, as in the reference */
        printf("Warning: fmin is too low. Replaced by the default value.\n");
        p->fmin = 0;
    }
    if (p->w0set == 1)      p->w0 = 2 * sqrt(ln2) * p->Q;
    else if (p->w0set == 2) p->w0 = TSPWS_PI / sqrt(ln2) * p->cycle;

    if (p->type == -1 || p->type == -2) {
        double rel = p->w0 / (TSPWS_PI * sqrt(2 / ln2));   /* w0 relative to the default */
        if (!p->lVfix)  p->V  = (unsigned int)ceil(4. * rel);
        if (!p->lb0fix) p->b0 = (unsigned int)pow(2, round(log2(rel)));
        if (!p->ls0fix) p->s0 = 2.;
    } else if (p->type == -3) {
        p->w0 = sqrt(2);
        if (!p->lVfix)  p->V  = 2;
        if (!p->lb0fix) p->b0 = 0.5;
        if (!p->ls0fix) p->s0 = 1.;
    }

    if (p->fmin) {
        double a = p->w0 / (2 * TSPWS_PI * dt * p->fmin);  /* coarsest scale wanted */
        if (p->J) {
            a /= pow(2, p->J - 1 / (double)p->V);          /* -> finest scale       */
            while (a < p->s0 * 0.9) { a *= 2; p->J--; }
            p->s0 = a;
        } else {
            p->J = (unsigned int)floor(log2(a / p->s0) + 1 / (double)p->V);
        }
    } else if (!p->J) {
        double a = nsamp * p->w0 / (2 * TSPWS_PI * 4. * p->s0);
        p->J = (unsigned int)floor(log2(a) + 1 / (double)p->V);
    }
    return 0;
}

/* The text the reference prints with `verbose` (src/ts_pws1f_lib.c:126-151): users read the
 * derived V/J/b0/s0/w0 from it, so the wording is kept. */
void tspws_print_params(const t_tsPWS *p, unsigned int nsamp, float dt)
{
    printf("Sequence length = %d, dt = %f\n", nsamp, dt);
    printf("  PWS power = %f\n", p->wu);
    printf("  Two-stage ");
    if (p->Kmax) printf("on, %d groups\n", p->Kmax); else printf("off\n");
    printf("  unbiased %s\n", p->unbiased ? "on" : "off");
    printf("  rm %s\n",       p->lrm      ? "on" : "off");
    printf("  fold %s\n",     p->fold     ? "on" : "off");
    printf("Mother wavelet: ");
    if (p->type == -1)      printf("complex Morlet\n");
    else if (p->type == -2) printf("exact complex Morlet\n");
    else if (p->type == -3) printf("complex Mexican hat\n");
    printf("Sampling of the time-frequency domain:\n");
    printf("         V = %d, J = %d, b0 = %f, s0 = %f, w0 = %f\n", p->V, p->J, p->b0, p->s0, p->w0);
    printf("This is: fmax = %f Hz, fmin = %f Hz\n", p->w0 / (2 * TSPWS_PI * p->s0 * dt),
           p->w0 / (2 * TSPWS_PI * dt * p->s0 * pow(2, p->J - 1 / (double)p->V)));
    printf("         Q = %f (equivalently cycles = %f)\n", p->w0 / (2 * sqrt(log(2))),
           p->w0 * sqrt(log(2)) / TSPWS_PI);
    if (p->subsmpl_N > 0 && p->subsmpl_p > 0)
        printf("Subsampling:\n  N = %d, p = %f\n", p->subsmpl_N, p->subsmpl_p);
    if (p->jackknife_n > 0 && p->jackknife_d > 0)
        printf("Jackknife Resampling:\n  n = %d, d = %d\n", p->jackknife_n, p->jackknife_d);
    printf("\n");
}

/* Fold is legal when the lag axis is symmetric: src/ts_pws1f_lib.c:72 (one-sided test). */
int tspws_fold_allowed(float beg, float dt, int max)
{
    return !(2 * beg + (max - 1) * dt > 0.5 * dt);
}

/* Which weighting applies: src/ts_pws1f_lib.c:226-228 with the K==1 fallback of :972. */
tspws_weight_t tspws_weight_rule(double wu, int unbiased, unsigned int K, unsigned int T)
{
    tspws_weight_t r;
    r.wu = wu; r.K = (double)K; r.T = (double)T;
    if (wu == 2 && unbiased && K > 1) r.mode = 3;
    else if (wu == 2)                 r.mode = 0;
    else if (wu == 1)                 r.mode = 1;
    else                              r.mode = 2;
    r.inv_k = 1.0 / r.K;
    r.inv_f = r.mode == 0 ? 1.0 / (r.K * r.K * r.T) : r.mode == 1 ? 1.0 / (r.K * r.T)
            : r.mode == 3 ? 1.0 / ((r.K - 1.0) * r.T) : 1.0 / r.T;
    return r;
}

unsigned long tspws_binomial(unsigned int n, unsigned int k)
{
    if (k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned long r = 1;
    for (unsigned int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

/* ------------------------------------------------------------------------------------ */
/* Frame -- reference CreateWaveletFamily, src/FWTa/wavelet_def_v7.c:204-339             */
/* ------------------------------------------------------------------------------------ */
/* erfi by its Maclaurin series, 500 terms (src/FWTa/wavelet_def_v7.c:104-117). */
static double erfi_series(double z)
{
    double zz = z * z, term = z, sum = z;
    for (unsigned int n = 1; n < 500; n++) {
        term *= zz / n;
        sum  += term / (2 * n + 1);
    }
    return sum * (2 / sqrt(TSPWS_PI));
}

/* Admissibility constant (src/FWTa/wavelet_def_v7.c:133-148): a 0.01-step Riemann sum whose
 * abscissa is accumulated in floating point, reproduced as such. */
static double frame_cpsi(int type, double w0)
{
    if (type == -3) return (4. / 3.) * sqrt(TSPWS_PI);
    double acc = 0;
    for (double w = 0.01; w < 100; w += 0.01) {
        double d = w - w0;
        acc += exp(-d * d) / w;
    }
    return acc * (0.01 * sqrt(TSPWS_PI) / 2);
}

/* One analysis filter, L taps about origin L/2, at `scale`
 * (Morlet :38-52, exact Morlet :71-87, complex Mexican hat :119-131). */
static void frame_fill_filter(double *w /*2L*/, int type, unsigned int L, double scale, double w0)
{
    const int    origin = (int)(L / 2);
    const double inv    = 1 / scale;
    if (type == -3) {
        const double k = 2 / sqrt(3 * sqrt(TSPWS_PI) * scale);
        for (int i = 0; i < (int)L; i++) {
            double u = inv * (i - origin);
            double g = (u * u - 1) * exp(-u * u / 2);
            w[2 * i]     = k * g;
            w[2 * i + 1] = k * (g * erfi_series(u / sqrt(2)) - sqrt(2 / TSPWS_PI) * u);
        }
        return;
    }
    const double k = 1 / sqrt(sqrt(TSPWS_PI) * scale);
    if (type == -2) {                       /* zero-mean ("exact") Morlet */
        const double dc = exp((-w0 * w0) / 2);
        for (int i = 0; i < (int)L; i++) {
            double u   = inv * (i - origin);
            double ph  = w0 * u;
            double env = k * exp(-0.5 * (u * u));
            w[2 * i]     = env * (cos(ph) - dc);
            w[2 * i + 1] = env * sin(ph);
        }
        return;
    }
    for (int i = 0; i < (int)L; i++) {      /* standard Morlet */
        double u  = inv * (i - origin);
        double ph = w0 * u;
        double g  = exp(-0.5 * (u * u));
        w[2 * i]     = (k * cos(ph)) * g;
        w[2 * i + 1] = (k * sin(ph)) * g;
    }
}

void tspws_b200_frame_destroy(tspws_frame_t *f)
{
    if (!f) return;
    free(f->scale); free(f->L); free(f->center); free(f->D); free(f->Ns);
    free(f->coef_off); free(f->tap_off); free(f->w);
    free(f);
}

tspws_frame_t *tspws_b200_frame_create(int type, unsigned int J, unsigned int V, unsigned int N,
                                       double s0, double b0, double w0, int uni)
{
    if (type > -1 || type < -3 || !V || !N) return NULL;   /* complex families only on this path */
    const unsigned int S = J * V;
    tspws_frame_t *f = (tspws_frame_t *)calloc(1, sizeof(*f));
    if (!f) return NULL;
    f->type = type; f->S = S; f->N = N; f->V = V;
    const size_t n1 = S ? S : 1;
    f->scale    = (double *)calloc(n1, sizeof(double));
    f->L        = (unsigned int *)calloc(n1, sizeof(unsigned int));
    f->center   = (int *)calloc(n1, sizeof(int));
    f->D        = (unsigned int *)calloc(n1, sizeof(unsigned int));
    f->Ns       = (unsigned int *)calloc(n1, sizeof(unsigned int));
    f->coef_off = (size_t *)calloc(n1 + 1, sizeof(size_t));
    f->tap_off  = (size_t *)calloc(n1 + 1, sizeof(size_t));
    if (!f->scale || !f->L || !f->center || !f->D || !f->Ns || !f->coef_off || !f->tap_off) {
        tspws_b200_frame_destroy(f);
        return NULL;
    }
    /* scales: geometric, by repeated multiplication (:299-309) */
    {
        double a = s0, step = pow(2., 1 / (double)V);
        for (unsigned int s = 0; s < S; s++) { f->scale[s] = a; a *= step; }
    }
    /* support: +-5 sigma, capped at the trace length (:312-322); origin in the middle */
    for (unsigned int s = 0; s < S; s++) {
        unsigned int len = 2 * (unsigned int)ceil(5 * f->scale[s]) + 1;
        f->L[s]      = len > N ? N : len;
        f->center[s] = (int)(f->L[s] / 2);
    }
    /* decimation: one step per octave, truncated to an integer >= 1 (:325-339) */
    {
        double a = s0 * (uni ? 1. : b0);
        for (unsigned int j = 0; j < J; j++) {
            unsigned int d = (uni || a < 1) ? 1u : (unsigned int)a;
            for (unsigned int v = 0; v < V; v++) f->D[j * V + v] = d;
            a *= 2;
        }
    }
    for (unsigned int s = 0; s < S; s++) {
        f->Ns[s]           = (N + f->D[s] - 1) / f->D[s];
        f->coef_off[s + 1] = f->coef_off[s] + f->Ns[s];
        f->tap_off[s + 1]  = f->tap_off[s] + f->L[s];
    }
    f->ncoef = f->coef_off[S];
    f->ntaps = f->tap_off[S];
    f->w = (double *)malloc((f->ntaps ? f->ntaps : 1) * 2 * sizeof(double));
    if (!f->w) { tspws_b200_frame_destroy(f); return NULL; }
    for (unsigned int s = 0; s < S; s++)
        frame_fill_filter(f->w + 2 * f->tap_off[s], type, f->L[s], f->scale[s], w0);
    f->Cpsi = frame_cpsi(type, w0);
    return f;
}

/* ------------------------------------------------------------------------------------ */
/* Resampling plans                                                                      */
/* ------------------------------------------------------------------------------------ */
/* Exactly K of J flags set, drawn with rand() % J by rejection, flipping the minority
 * (reference SubsamplingPlan, src/ts_pws1f_lib.c:355-383) -- same rand() call sequence, so a
 * process that has consumed the same number of rand() values gets the same plan. */
int tspws_b200_subsampling_plan(char *sel, unsigned int J, unsigned int K)
{
    if (!sel) return 1;
    if (K > J) return 2;
    const int minority_is_one = 2 * K < J;
    unsigned int todo = minority_is_one ? K : J - K;
    memset(sel, minority_is_one ? 0 : 1, J);
    while (todo) {
        unsigned int j = (unsigned int)rand() % J;
        if ((sel[j] != 0) != minority_is_one) { sel[j] = (char)minority_is_one; todo--; }
    }
    return 0;
}

/* Delete-d jackknife over n day-of-year bins (reference JackknifePlans,
 * src/ts_pws1f_lib.c:385-430): bin = floor(yday*n/365); the C(n,d) deleted-bin sets are
 * enumerated in lexicographic order; sel[c][t] = 0 when trace t falls in a deleted bin. */
int tspws_b200_jackknife_plans(char *sel, const time_t *tm, unsigned int mtr,
                               unsigned int d, unsigned int n, unsigned int C)
{
    if (!sel || !tm) return 1;
    if (tm[0] == 0) return -2;
    if (!d || d > n) return 2;
    unsigned int *bin  = (unsigned int *)malloc((mtr ? mtr : 1) * sizeof(unsigned int));
    unsigned int *comb = (unsigned int *)malloc(d * sizeof(unsigned int));
    if (!bin || !comb) { free(bin); free(comb); return -1; }
    for (unsigned int t = 0; t < mtr; t++) {
        struct tm g;
        gmtime_r(tm + t, &g);
        bin[t] = (unsigned int)floor((double)(g.tm_yday * (int)n) / 365.);
    }
    for (unsigned int i = 0; i < d; i++) comb[i] = i;
    for (unsigned int c = 0; c < C; c++) {
        if (c) {                                   /* next combination, lexicographic */
            int i = (int)d - 1;
            while (i >= 0 && comb[i] >= n - d + (unsigned int)i) i--;
            if (i < 0) break;                      /* C exceeded C(n,d): leave the rest untouched */
            comb[i]++;
            for (unsigned int j = (unsigned int)i + 1; j < d; j++) comb[j] = comb[j - 1] + 1;
        }
        char *row = sel + (size_t)c * mtr;
        for (unsigned int t = 0; t < mtr; t++) {
            char keep = 1;
            for (unsigned int i = 0; i < d; i++) if (bin[t] == comb[i]) keep = 0;
            row[t] = keep;
        }
    }
    free(bin); free(comb);
    return 0;
}
