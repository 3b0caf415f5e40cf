/*
 * tspws_cli.c -- SAC-free command line front-end over libtspws_b200.so.
 * Accepts the reference CLI's key=value vocabulary (reference src/ts_pws1f.c:163-215) so that the
 * three commands of examples/example.sh run unchanged:
 *     tspws_b200_cli filelist.txt verbose
 *     tspws_b200_cli filelist.txt osac=example2 wu=2 rm fold fmin=0.004 J=3 verbose
 *     tspws_b200_cli sac2bin filelist.txt ECH.bin          (the reference's sac2bin step)
 *     tspws_b200_cli ECH.bin osac=twostage wu=2 rm bin fold TwoStage=10 unbiased fmin=0.004 J=3
 * Outputs tl_<osac>.sac / ts_pws_<osac>.sac like the reference (folded traces keep the second half,
 * src/ts_pws1f.c:322-328); with convergence[=ref.sac] / AllSteps the raw curve files, with subsmpl_N= +
 * subsmpl_prob= or jackknife_n= + jackknife_d= the resampled stacks as SAC files or (obin) as two .bin
 * containers, named as in src/ts_pws1f.c:354-418.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tspws_b200.h"

int  tspws_b200_load_sac_list(const char *listfile, t_data *in, int pinned);
int  tspws_b200_load_bin(const char *binfile, t_data *in, int pinned);
int  tspws_b200_save_bin(const char *binfile, const t_data *in, const char *method);
int  tspws_b200_write_sac(const char *path, const float *y, const t_hdr *hd, int npts, float beg, const char *kinst, float user0);
void tspws_b200_free_data(t_data *in);
int  tspws_b200_save_rows_bin(const char *binfile, float *const *rows, unsigned int M, int n0, int npts, const t_hdr *hd,
                              float beg, const char *method, const unsigned int *traces_per_row);
long tspws_b200_read_sac(const char *path, float *data, long maxpts, float *dt, float *beg, void *header632);

static unsigned long n_choose_k(unsigned int n, unsigned int k)
{
    if (k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned long r = 1;
    for (unsigned int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

static void dump(const char *name, const void *p, size_t size, size_t n)
{
    FILE *f = fopen(name, "wb");
    if (!f) { printf("cannot open %s\n", name); return; }
    fwrite(p, size, n, f);
    fclose(f);
}

static int starts(const char *a, const char *key) { return !strncmp(a, key, strlen(key)); }

int main(int argc, char **argv)
{
    if (argc < 2) {
        printf("usage: %s <filelist|file.bin> [key=value ...]   (keys as in the reference's ts_pws)\n"
               "       %s sac2bin <filelist> <out.bin>\n", argv[0], argv[0]);
        return 0;
    }
    if (!strcmp(argv[1], "sac2bin")) {
        if (argc < 4) { printf("sac2bin needs a file list and an output name\n"); return 1; }
        t_data in;
        int e = tspws_b200_load_sac_list(argv[2], &in, 0);
        if (!e) e = tspws_b200_save_bin(argv[3], &in, "");
        tspws_b200_free_data(&in);
        return e;
    }
    t_tsPWS p;
    memset(&p, 0, sizeof(p));
    p.type = -1; p.V = 4; p.s0 = 2.; p.b0 = 1.; p.w0 = 3.14159265358979328 * sqrt(2 / log(2.)); p.wu = 2.; p.cycle = 2.;
    p.filein = argv[1];
    for (int i = 2; i < argc; i++) {
        const char *a = argv[i];
        if      (starts(a, "wu="))            p.wu = atof(a + 3);
        else if (starts(a, "J="))             p.J = (unsigned)atoi(a + 2);
        else if (starts(a, "Nmax="))          p.Nmax = (unsigned)atoi(a + 5);
        else if (starts(a, "fmin="))          p.fmin = atof(a + 5);
        else if (starts(a, "V="))           { p.V = (unsigned)atoi(a + 2); p.lVfix = 1; }
        else if (starts(a, "s0="))          { p.s0 = atof(a + 3); p.ls0fix = 1; }
        else if (starts(a, "b0="))          { p.b0 = atof(a + 3); p.lb0fix = 1; }
        else if (starts(a, "uni"))            p.uni = 1;
        else if (starts(a, "rm"))             p.lrm = 1;
        else if (starts(a, "bin"))            p.bin = 1;
        else if (starts(a, "verbose"))        p.verbose = 1;
        else if (starts(a, "fold"))           p.fold = 1;
        else if (starts(a, "unbiased"))       p.unbiased = 1;
        else if (starts(a, "MexHat"))         p.type = -3;
        else if (starts(a, "AllSteps"))       p.AllSteps = 1;
        else if (starts(a, "subsmpl_N="))     p.subsmpl_N = (unsigned)atoi(a + 10);
        else if (starts(a, "subsmpl_prob="))  p.subsmpl_p = atof(a + 13);
        else if (starts(a, "jackknife_n="))   p.jackknife_n = (unsigned)atoi(a + 12);
        else if (starts(a, "jackknife_d="))   p.jackknife_d = (unsigned)atoi(a + 12);
        else if (starts(a, "obin"))           p.obin = 1;
        else if (starts(a, "convergence"))  { p.convergence = 1; if (starts(a, "convergence=")) p.fileconv = (char *)(a + 12); }
        else if (starts(a, "TwoStage"))     { p.Kmax = 10; if (starts(a, "TwoStage=")) p.Kmax = (unsigned)atoi(a + 9); }
        else if (starts(a, "Q="))           { p.Q = atof(a + 2); if (p.w0set < 1) p.w0set = 1; }
        else if (starts(a, "cycles="))      { p.cycle = atof(a + 7); if (p.w0set < 2) p.w0set = 2; }
        else if (starts(a, "cyc="))         { if (p.w0set < 3) { p.w0set = 3; p.w0 = atof(a + 4) * 3.14159265358979328; } }
        else if (starts(a, "w0="))          { p.w0 = atof(a + 3); if (p.w0set < 4) p.w0set = 4; }
        else if (starts(a, "osac="))        { p.fileout = (char *)(a + 5); if (!*p.fileout) p.fileout = NULL; }
        else if (starts(a, "kinst="))       { p.kinst = (char *)(a + 6); p.lkinst = 1; }
        else printf("ignored argument: %s\n", a);
    }
    t_data in;
    int e = p.bin ? tspws_b200_load_bin(p.filein, &in, 1) : tspws_b200_load_sac_list(p.filein, &in, 1);
    if (e) return e;
    t_tsPWS_out out;
    memset(&out, 0, sizeof(out));
    const int max = in.hdr.max;
    out.N = (unsigned)max; out.mtr = p.Nmax ? p.Nmax : in.hdr.mtr;
    out.ls = (float *)calloc(max, sizeof(float));
    out.tsPWS = (float *)calloc(max, sizeof(float));
    if (!out.ls || !out.tsPWS) return 4;
    /* optional outputs, allocated by the caller like src/ts_pws1f.c:236-277 */
    float *blk_ls = NULL, *blk_ts = NULL;
    const size_t mtr_all = in.hdr.mtr;
    if (p.convergence) {
        out.ls_sim = (double *)calloc(mtr_all, sizeof(double)); out.tsPWS_sim = (double *)calloc(mtr_all, sizeof(double));
        out.ls_misfit = (double *)calloc(mtr_all, sizeof(double)); out.tsPWS_misfit = (double *)calloc(mtr_all, sizeof(double));
        if (!out.ls_sim || !out.tsPWS_sim || !out.ls_misfit || !out.tsPWS_misfit) return 4;
        if (p.AllSteps) {
            out.ls_steps = (float *)calloc(mtr_all * (size_t)max, sizeof(float));
            out.tsPWS_steps = (float *)calloc(mtr_all * (size_t)max, sizeof(float));
            if (!out.ls_steps || !out.tsPWS_steps) return 4;
        }
        if (p.fileconv) {                            /* reference trace for the curves; ignored when unusable */
            in.reference = (float *)calloc(max, sizeof(float));
            float rdt = 0, rbeg = 0;
            const long got = in.reference ? tspws_b200_read_sac(p.fileconv, in.reference, max, &rdt, &rbeg, NULL) : -1;
            if (got != max) {
                if (got >= 0) printf("The reference for convergence has a different length (%d:%ld)\n", max, got);
                free(in.reference); in.reference = NULL; p.fileconv = NULL;
            }
        }
    }
    if (p.jackknife_n) {
        p.subsmpl_N = 0; p.subsmpl_p = 0;
        if (p.jackknife_d == 0 || p.jackknife_d >= p.jackknife_n) { p.jackknife_d = 0; p.jackknife_n = 0; }
        else out.M = (unsigned)n_choose_k(p.jackknife_n, p.jackknife_d);
    }
    if (p.subsmpl_N) {
        if (p.subsmpl_p < 0 || p.subsmpl_p > 1) { p.subsmpl_N = 0; p.subsmpl_p = 0; }
        else out.M = p.subsmpl_N;
    }
    if ((p.jackknife_n || p.subsmpl_N) && out.M) {
        out.mtr_subsmpl = (unsigned int *)calloc(out.M, sizeof(unsigned int));
        out.ls_subsmpl = (float **)calloc(out.M, sizeof(float *));
        out.tsPWS_subsmpl = (float **)calloc(out.M, sizeof(float *));
        blk_ls = (float *)calloc((size_t)out.M * max, sizeof(float));
        blk_ts = (float *)calloc((size_t)out.M * max, sizeof(float));
        if (!out.mtr_subsmpl || !out.ls_subsmpl || !out.tsPWS_subsmpl || !blk_ls || !blk_ts) return 4;
        for (unsigned int m = 0; m < out.M; m++) { out.ls_subsmpl[m] = blk_ls + (size_t)m * max; out.tsPWS_subsmpl[m] = blk_ts + (size_t)m * max; }
        if (p.subsmpl_N) for (unsigned int m = 1; m < out.M; m++) out.mtr_subsmpl[m] = (unsigned int)(out.mtr * p.subsmpl_p);   /* src/ts_pws1f.c:276-277 */
    }
    e = tspws_main(&p, &out, &in);
    if (!e) {
        t_hdr hd = in.hdr;
        int n0 = 0, npts = max;
        float beg = hd.beg;
        if (p.fold && 2 * hd.beg + (max - 1) * hd.dt < 0.5 * hd.dt) { n0 = max / 2; npts = (max + 1) / 2; beg += hd.dt * n0; }
        char name[1024];
        if (p.fileout) snprintf(name, sizeof(name), "tl_%s.sac", p.fileout); else snprintf(name, sizeof(name), "tl.sac");
        tspws_b200_write_sac(name, out.ls + n0, &hd, npts, beg, p.lkinst ? p.kinst : "t-lin", (float)out.mtr);
        if (p.verbose) printf("Output files:\n  Linear stack: %s\n", name);
        if (p.fileout) snprintf(name, sizeof(name), "ts_pws_%s.sac", p.fileout); else snprintf(name, sizeof(name), "ts_pws.sac");
        tspws_b200_write_sac(name, out.tsPWS + n0, &hd, npts, beg, p.lkinst ? p.kinst : "ts_pws", (float)out.mtr);
        if (p.verbose) printf("  ts-PWS:       %s\n", name);
        const char *tag = p.fileout ? p.fileout : "";
        if (p.convergence) {                          /* raw double / float files, src/ts_pws1f.c:354-391 */
            snprintf(name, sizeof(name), "ts_pws_%s_convergence", tag); dump(name, out.tsPWS_sim, sizeof(double), out.mtr);
            snprintf(name, sizeof(name), "tl_%s_convergence", tag);     dump(name, out.ls_sim, sizeof(double), out.mtr);
            snprintf(name, sizeof(name), "ts_pws_%s_misfit", tag);      dump(name, out.tsPWS_misfit, sizeof(double), out.mtr);
            snprintf(name, sizeof(name), "tl_%s_misfit", tag);          dump(name, out.ls_misfit, sizeof(double), out.mtr);
            if (out.tsPWS_steps) { snprintf(name, sizeof(name), "ts_pws_%s_steps", tag); dump(name, out.tsPWS_steps, sizeof(float), (size_t)out.mtr * out.N); }
            if (out.ls_steps)    { snprintf(name, sizeof(name), "tl_%s_steps", tag);     dump(name, out.ls_steps, sizeof(float), (size_t)out.mtr * out.N); }
        }
        if ((p.jackknife_n || p.subsmpl_N) && out.M) {    /* src/ts_pws1f.c:393-418 */
            char sub[900];
            if (p.fileout) snprintf(sub, sizeof(sub), "_%s_subsmpl", p.fileout); else snprintf(sub, sizeof(sub), "_subsmpl");
            if (p.obin) {
                snprintf(name, sizeof(name), "tl%s.bin", sub);
                tspws_b200_save_rows_bin(name, out.ls_subsmpl, out.M, n0, npts, &hd, beg, p.lkinst ? p.kinst : "t-lin", out.mtr_subsmpl);
                snprintf(name, sizeof(name), "ts_pws%s.bin", sub);
                tspws_b200_save_rows_bin(name, out.tsPWS_subsmpl, out.M, n0, npts, &hd, beg, p.lkinst ? p.kinst : "ts_pws", out.mtr_subsmpl);
            } else {
                for (unsigned int m = 0; m < out.M; m++) {
                    snprintf(name, sizeof(name), "tl%s_%u.sac", sub, m);
                    tspws_b200_write_sac(name, out.ls_subsmpl[m] + n0, &hd, npts, beg, p.lkinst ? p.kinst : "t-lin", (float)out.mtr_subsmpl[m]);
                    snprintf(name, sizeof(name), "ts_pws%s_%u.sac", sub, m);
                    tspws_b200_write_sac(name, out.tsPWS_subsmpl[m] + n0, &hd, npts, beg, p.lkinst ? p.kinst : "ts_pws", (float)out.mtr_subsmpl[m]);
                }
            }
        }
    }
    free(out.ls_sim); free(out.tsPWS_sim); free(out.ls_misfit); free(out.tsPWS_misfit); free(out.ls_steps); free(out.tsPWS_steps);
    free(out.mtr_subsmpl); free(out.ls_subsmpl); free(out.tsPWS_subsmpl); free(blk_ls); free(blk_ts); free(in.reference);
    in.reference = NULL;
    free(out.ls); free(out.tsPWS);
    tspws_b200_free_data(&in);
    tspws_b200_shutdown();
    return e;
}
