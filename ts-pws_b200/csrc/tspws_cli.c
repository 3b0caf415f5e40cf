/*
 * tspws_cli.c -- SAC-free command line front-end over libtspws_b200.so.
 * Accepts the reference CLI's key=value vocabulary (reference src/ts_pws1f.c:163-215) so that the
 * three commands of examples/example.sh run unchanged:
 *     tspws_b200_cli filelist.txt verbose
 *     tspws_b200_cli filelist.txt osac=example2 wu=2 rm fold fmin=0.004 J=3 verbose
 *     tspws_b200_cli sac2bin filelist.txt ECH.bin          (the reference's sac2bin step)
 *     tspws_b200_cli ECH.bin osac=twostage wu=2 rm bin fold TwoStage=10 unbiased fmin=0.004 J=3
 * Outputs tl_<osac>.sac / ts_pws_<osac>.sac like the reference (folded traces keep the second half,
 * src/ts_pws1f.c:322-328).
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
    }
    free(out.ls); free(out.tsPWS);
    tspws_b200_free_data(&in);
    tspws_b200_shutdown();
    return e;
}
