/*
 * tspws_io.c -- SURVEY.md §8(f) rank 1: the data formats on either side of the path, without
 * sacio.a: SAC binary files (little-endian, 632-byte header), SAC file lists, and the single-file
 * `.bin` container written by the reference's sac2bin.
 *
 *   .bin layout (reference src/sac2bin.h:6-27, writer src/sac2bin.c:192-195, reader
 *   src/ts_pws1f.c:586-608,664-675):
 *       t_ccheader (116 B) | time_t time[nseq] | float lag0[nseq] | float data[nseq][nlags]
 *   SAC list reader: reference ReadData, src/ts_pws1f.c:609-722 (header of the first file fills
 *   t_hdr; traces whose dt differs by more than 1 % are skipped).
 *
 * Traces are loaded into page-locked memory when CUDA is available, so that tspws_main's H2D
 * copy runs at PCIe rate ("the most demanding operation is conventionally reading the data",
 * reference README.md:16-18).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda_runtime_api.h>

#include "tspws_internal.h"

#define SAC_HDR_BYTES 632
#define SAC_UNDEF (-12345.0f)

/* word indices of the SAC header fields used here */
enum { F_DELTA = 0, F_B = 5, F_O = 7, F_STLA = 31, F_STLO = 32, F_STEL = 33, F_EVLA = 35, F_EVLO = 36, F_USER0 = 40,
       I_NZYEAR = 70, I_NZJDAY = 71, I_NZHOUR = 72, I_NZMIN = 73, I_NZSEC = 74, I_NZMSEC = 75, I_NVHDR = 76,
       I_NPTS = 79, I_IFTYPE = 85, I_IZTYPE = 87, I_LEVEN = 105, I_LPSPOL = 106, I_LOVROK = 107, I_LCALDA = 108 };
/* byte offsets of the string fields */
enum { K_KSTNM = 440, K_KEVNM = 448, K_KHOLE = 464, K_KUSER0 = 576, K_KUSER1 = 584, K_KUSER2 = 592,
       K_KCMPNM = 600, K_KNETWK = 608, K_KINST = 624 };

typedef struct {                      /* t_ccheader, reference src/sac2bin.h:6-27 */
    char     method[8], net1[8], sta1[8], loc1[8], chn1[8], net2[8], sta2[8], loc2[8], chn2[8];
    float    stlat1, stlon1, stel1, stlat2, stlon2, stel2;
    uint32_t nlags, nseq;
    float    tlength, lag1, lag2;
} tspws_ccheader_t;
_Static_assert(sizeof(tspws_ccheader_t) == 116, "t_ccheader is 116 bytes");

static void *host_alloc(size_t bytes, int pinned, int *was_pinned)
{
    void *p = NULL;
    *was_pinned = 0;
    if (pinned && cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess) { *was_pinned = 1; return p; }
    cudaGetLastError();
    return calloc(bytes ? bytes : 1, 1);
}

/* SAC strings are blank padded; "-12345" marks an undefined field */
static void sac_string(char dst[9], const unsigned char *h, int off, int len)
{
    int n = len < 8 ? len : 8;
    memcpy(dst, h + off, n);
    dst[n] = 0;
    for (int i = n - 1; i >= 0 && (dst[i] == ' ' || dst[i] == 0); i--) dst[i] = 0;
    if (!strncmp(dst, "-12345", 6)) dst[0] = 0;
}

static int sac_read_header(FILE *f, unsigned char h[SAC_HDR_BYTES])
{
    if (fread(h, 1, SAC_HDR_BYTES, f) != SAC_HDR_BYTES) return -1;
    int32_t nvhdr;
    memcpy(&nvhdr, h + 4 * I_NVHDR, 4);
    return (nvhdr >= 1 && nvhdr <= 7) ? 0 : -2;              /* byte-swapped or not a SAC file */
}
static float   hf(const unsigned char *h, int w) { float v;   memcpy(&v, h + 4 * w, 4); return v; }
static int32_t hi(const unsigned char *h, int w) { int32_t v; memcpy(&v, h + 4 * w, 4); return v; }

/* Reads one SAC file: up to maxpts samples into data (rest zero-filled), header into h. Returns npts or <0. */
long tspws_b200_read_sac(const char *path, float *data, long maxpts, float *dt, float *beg, void *header632)
{
    unsigned char h[SAC_HDR_BYTES];
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    if (sac_read_header(f, h)) { fclose(f); return -2; }
    const long npts = hi(h, I_NPTS);
    if (dt)  *dt  = hf(h, F_DELTA);
    if (beg) *beg = hf(h, F_B);
    if (header632) memcpy(header632, h, SAC_HDR_BYTES);
    if (data && maxpts > 0) {
        const long n = npts < maxpts ? npts : maxpts;
        const size_t got = fread(data, sizeof(float), (size_t)(n > 0 ? n : 0), f);
        for (long i = (long)got; i < maxpts; i++) data[i] = 0.f;
    }
    fclose(f);
    return npts;
}

void tspws_b200_free_data(t_data *in)
{
    if (!in) return;
    if (in->sigall) { if (cudaFreeHost(in->sigall) != cudaSuccess) { cudaGetLastError(); free(in->sigall); } }
    free(in->time); free(in->lag0); free(in->reference);
    in->sigall = NULL; in->time = NULL; in->lag0 = NULL; in->reference = NULL;
}

static time_t utc_from_header(const unsigned char *h)
{
    const int year = hi(h, I_NZYEAR), yday = hi(h, I_NZJDAY);
    if (year == -12345 || yday == -12345) return 0;
    struct tm tm;
    memset(&tm, 0, sizeof(tm));
    tm.tm_year = year - 1900; tm.tm_mon = 0; tm.tm_mday = yday;     /* day of year via mday, like src/sac2bin.c:146-153 */
    tm.tm_hour = hi(h, I_NZHOUR) == -12345 ? 0 : hi(h, I_NZHOUR);
    tm.tm_min  = hi(h, I_NZMIN)  == -12345 ? 0 : hi(h, I_NZMIN);
    tm.tm_sec  = hi(h, I_NZSEC)  == -12345 ? 0 : hi(h, I_NZSEC);
    return timegm(&tm);
}

/* List of SAC files (one name per line) -> t_data.  Replaces ReadData's SAC branch. */
int tspws_b200_load_sac_list(const char *listfile, t_data *in, int pinned)
{
    if (!listfile || !in) return TSPWS_ERR_NULL;
    memset(in, 0, sizeof(*in));
    FILE *fl = fopen(listfile, "r");
    if (!fl) { printf("tspws_main: cannot open the %s file\n", listfile); return -2; }
    char **names = NULL, line[4096];
    size_t n = 0, cap = 0;
    while (fgets(line, sizeof(line), fl)) {
        size_t L = strlen(line);
        while (L && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ')) line[--L] = 0;
        if (!L) continue;
        if (n == cap) { cap = cap ? 2 * cap : 256; names = (char **)realloc(names, cap * sizeof(char *)); }
        names[n++] = strdup(line);
    }
    fclose(fl);
    int rc = 0, was_pinned = 0;
    if (!n) { printf("tspws_main: nothing to do, %s is empty!\n", listfile); goto done; }

    unsigned char h[SAC_HDR_BYTES];
    {
        FILE *f = fopen(names[0], "rb");
        if (!f || sac_read_header(f, h)) { if (f) fclose(f); printf("tspws_main: cannot read the header of %s\n", names[0]); rc = -2; goto done; }
        fclose(f);
    }
    t_hdr *hd = &in->hdr;
    hd->max = hi(h, I_NPTS); hd->mtr = (unsigned int)n;
    hd->stla = hf(h, F_STLA); hd->stlo = hf(h, F_STLO); hd->stel = hf(h, F_STEL);
    hd->evla = hf(h, F_EVLA); hd->evlo = hf(h, F_EVLO);
    hd->dt = hf(h, F_DELTA); hd->beg = hf(h, F_B);
    sac_string(hd->net2, h, K_KNETWK, 8); sac_string(hd->sta2, h, K_KSTNM, 8);
    sac_string(hd->loc2, h, K_KHOLE, 8);  sac_string(hd->chn2, h, K_KCMPNM, 8);
    sac_string(hd->net1, h, K_KUSER0, 8); sac_string(hd->sta1, h, K_KEVNM, 8);
    sac_string(hd->loc1, h, K_KUSER1, 8); sac_string(hd->chn1, h, K_KUSER2, 8);
    if (hd->max <= 0) { rc = -2; goto done; }

    const size_t max = (size_t)hd->max;
    in->sigall = (float *)host_alloc(n * max * sizeof(float), pinned, &was_pinned);
    in->time   = (time_t *)calloc(n, sizeof(time_t));
    in->lag0   = (float *)calloc(n, sizeof(float));
    if (!in->sigall || !in->time || !in->lag0) { rc = TSPWS_ERR_ALLOC; goto done; }
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) {
        float dt1 = 0, b1 = 0;
        const long npts = tspws_b200_read_sac(names[i], in->sigall + kept * max, (long)max, &dt1, &b1, h);
        if (npts < 0) { printf("tspws_main: Error reading %s file\n", names[i]); rc = -2; goto done; }
        if ((size_t)npts > max) printf("tspws_main: WARNING: using only %zu samples on %zu trace\n", max, i);
        else if ((size_t)npts < max) printf("tspws_main: WARNING: trace %zu has only %ld samples\n", i, npts);
        if (fabsf(dt1 - hd->dt) > hd->dt * 0.01f) {
            printf("tspws_main: WARNING: trace %zu has a different dt !\ntspws_main: WARNING: skipping trace %zu\n", i, i);
            continue;
        }
        in->time[kept] = utc_from_header(h);
        in->lag0[kept] = hf(h, F_O) == SAC_UNDEF ? 0.f : hf(h, F_O);
        kept++;
    }
    hd->mtr = (unsigned int)kept;
done:
    for (size_t i = 0; i < n; i++) free(names[i]);
    free(names);
    if (rc) tspws_b200_free_data(in);
    return rc;
}

/* sac2bin container -> t_data.  Replaces ReadData's `bin` branch (src/ts_pws1f.c:586-608,664-675). */
int tspws_b200_load_bin(const char *binfile, t_data *in, int pinned)
{
    if (!binfile || !in) return TSPWS_ERR_NULL;
    memset(in, 0, sizeof(*in));
    FILE *f = fopen(binfile, "rb");
    if (!f) { printf("tspws_main: cannot open the %s file\n", binfile); return -2; }
    tspws_ccheader_t b;
    int rc = 0, was_pinned = 0;
    if (fread(&b, sizeof(b), 1, f) != 1) { printf("tspws_main: %s is shorter than predicted (header).\n", binfile); fclose(f); return -2; }
    t_hdr *hd = &in->hdr;
    hd->max = (int)b.nlags; hd->mtr = b.nseq;
    hd->evla = b.stlat1; hd->evlo = b.stlon1; hd->stla = b.stlat2; hd->stlo = b.stlon2; hd->stel = b.stel2;
    hd->dt = (b.lag2 - b.lag1) / (float)(b.nlags - 1);
    hd->beg = b.lag1;
    memcpy(hd->net1, b.net1, 8); memcpy(hd->sta1, b.sta1, 8); memcpy(hd->loc1, b.loc1, 8); memcpy(hd->chn1, b.chn1, 8);
    memcpy(hd->net2, b.net2, 8); memcpy(hd->sta2, b.sta2, 8); memcpy(hd->loc2, b.loc2, 8); memcpy(hd->chn2, b.chn2, 8);
    const size_t mtr = b.nseq, max = b.nlags;
    in->sigall = (float *)host_alloc(mtr * max * sizeof(float), pinned, &was_pinned);
    in->time   = (time_t *)calloc(mtr ? mtr : 1, sizeof(time_t));
    in->lag0   = (float *)calloc(mtr ? mtr : 1, sizeof(float));
    if (!in->sigall || !in->time || !in->lag0) { rc = TSPWS_ERR_ALLOC; goto done; }
    if (fread(in->time, sizeof(time_t), mtr, f) != mtr) printf("tspws_main: %s is shorter than predicted (time).\n", binfile);
    if (fread(in->lag0, sizeof(float), mtr, f) != mtr)  printf("tspws_main: %s is shorter than predicted (lag0).\n", binfile);
    if (fread(in->sigall, sizeof(float), mtr * max, f) != mtr * max) printf("tspws_main: %s is shorter than predicted (data).\n", binfile);
done:
    fclose(f);
    if (rc) tspws_b200_free_data(in);
    return rc;
}

/* t_data -> sac2bin container (what the reference's sac2bin writes, src/sac2bin.c:192-195). */
int tspws_b200_save_bin(const char *binfile, const t_data *in, const char *method)
{
    if (!binfile || !in || !in->sigall) return TSPWS_ERR_NULL;
    tspws_ccheader_t b;
    memset(&b, 0, sizeof(b));
    const t_hdr *hd = &in->hdr;
    strncpy(b.method, method ? method : "", 8);
    memcpy(b.net1, hd->net1, 8); memcpy(b.sta1, hd->sta1, 8); memcpy(b.loc1, hd->loc1, 8); memcpy(b.chn1, hd->chn1, 8);
    memcpy(b.net2, hd->net2, 8); memcpy(b.sta2, hd->sta2, 8); memcpy(b.loc2, hd->loc2, 8); memcpy(b.chn2, hd->chn2, 8);
    b.stlat1 = hd->evla; b.stlon1 = hd->evlo; b.stlat2 = hd->stla; b.stlon2 = hd->stlo; b.stel2 = hd->stel;
    b.nlags = (uint32_t)hd->max; b.nseq = hd->mtr;
    b.tlength = hd->dt * (float)(b.nlags - 1);
    b.lag1 = hd->beg; b.lag2 = b.lag1 + b.tlength;
    FILE *f = fopen(binfile, "wb");
    if (!f) return -2;
    const size_t mtr = hd->mtr, max = (size_t)hd->max;
    time_t zero_t = 0; float zero_f = 0;
    fwrite(&b, sizeof(b), 1, f);
    for (size_t i = 0; i < mtr; i++) fwrite(in->time ? &in->time[i] : &zero_t, sizeof(time_t), 1, f);
    for (size_t i = 0; i < mtr; i++) fwrite(in->lag0 ? &in->lag0[i] : &zero_f, sizeof(float), 1, f);
    fwrite(in->sigall, sizeof(float), mtr * max, f);
    fclose(f);
    return 0;
}

/* M stacked rows -> the container the reference's wrbin writes for resampled stacks (src/ts_pws1f.c:766-823):
 * header | time_t[M] (= traces behind every row) | float data[M][npts] -- NO lag0 block, unlike the files
 * sac2bin produces.  rows[m] + n0 is the first sample written of row m. */
int tspws_b200_save_rows_bin(const char *binfile, float *const *rows, unsigned int M, int n0, int npts, const t_hdr *hd,
                             float beg, const char *method, const unsigned int *traces_per_row)
{
    if (!binfile || !rows || !hd || npts <= 0) return TSPWS_ERR_NULL;
    tspws_ccheader_t b;
    memset(&b, 0, sizeof(b));
    strncpy(b.method, method ? method : "", 8);
    memcpy(b.net1, hd->net1, 8); memcpy(b.sta1, hd->sta1, 8); memcpy(b.loc1, hd->loc1, 8); memcpy(b.chn1, hd->chn1, 8);
    memcpy(b.net2, hd->net2, 8); memcpy(b.sta2, hd->sta2, 8); memcpy(b.loc2, hd->loc2, 8); memcpy(b.chn2, hd->chn2, 8);
    b.stlat1 = hd->evla; b.stlon1 = hd->evlo; b.stlat2 = hd->stla; b.stlon2 = hd->stlo; b.stel2 = hd->stel;
    b.nlags = (uint32_t)npts; b.nseq = M;
    b.tlength = hd->dt * (float)(b.nlags - 1);
    b.lag1 = beg; b.lag2 = b.lag1 + b.tlength;
    FILE *f = fopen(binfile, "wb");
    if (!f) return -2;
    fwrite(&b, sizeof(b), 1, f);
    for (unsigned int m = 0; m < M; m++) { const time_t t = (time_t)(traces_per_row ? traces_per_row[m] : 0); fwrite(&t, sizeof(t), 1, f); }
    for (unsigned int m = 0; m < M; m++) fwrite(rows[m] + n0, sizeof(float), (size_t)npts, f);
    fclose(f);
    return 0;
}

/* One evenly sampled trace -> SAC file with the fields the reference's wrsac sets (src/ts_pws1f.c:725-763). */
int tspws_b200_write_sac(const char *path, const float *y, const t_hdr *hd, int npts, float beg, const char *kinst, float user0)
{
    if (!path || !y || !hd) return TSPWS_ERR_NULL;
    unsigned char h[SAC_HDR_BYTES];
    float fu = SAC_UNDEF; int32_t iu = -12345;
    for (int w = 0; w < 70; w++) memcpy(h + 4 * w, &fu, 4);
    for (int w = 70; w < 110; w++) memcpy(h + 4 * w, &iu, 4);
    memset(h + 440, ' ', 192);
    for (int k = 440; k < 632; k += 8) memcpy(h + k, "-12345  ", 8);
    memcpy(h + K_KEVNM, "-12345          ", 16);
#define SETF(w, v) do { float t_ = (v); memcpy(h + 4 * (w), &t_, 4); } while (0)
#define SETI(w, v) do { int32_t t_ = (v); memcpy(h + 4 * (w), &t_, 4); } while (0)
#define SETK(off, str, len) do { memset(h + (off), ' ', (len)); size_t l_ = strlen(str); memcpy(h + (off), (str), l_ < (size_t)(len) ? l_ : (size_t)(len)); } while (0)
    float ymin = y[0], ymax = y[0]; double ymean = 0;
    for (int i = 0; i < npts; i++) { if (y[i] < ymin) ymin = y[i]; if (y[i] > ymax) ymax = y[i]; ymean += y[i]; }
    SETF(F_DELTA, hd->dt); SETF(1, ymin); SETF(2, ymax); SETF(F_B, beg); SETF(6, beg + hd->dt * (float)(npts - 1));
    SETF(F_STLA, hd->stla); SETF(F_STLO, hd->stlo); SETF(F_STEL, hd->stel); SETF(F_EVLA, hd->evla); SETF(F_EVLO, hd->evlo);
    SETF(F_USER0, user0); SETF(56, (float)(ymean / (npts > 0 ? npts : 1)));
    SETI(I_NVHDR, 6); SETI(I_NPTS, npts); SETI(I_IFTYPE, 1); SETI(I_IZTYPE, 11);     /* ITIME, IO */
    SETI(I_LEVEN, 1); SETI(I_LPSPOL, 0); SETI(I_LOVROK, 1); SETI(I_LCALDA, 1);
    SETK(K_KINST, kinst ? kinst : "", 8);
    if (hd->net2[0]) SETK(K_KNETWK, hd->net2, 8);
    if (hd->sta2[0]) SETK(K_KSTNM, hd->sta2, 8);
    if (hd->loc2[0]) SETK(K_KHOLE, hd->loc2, 8);
    if (hd->chn2[0]) SETK(K_KCMPNM, hd->chn2, 8);
    if (hd->net1[0]) SETK(K_KUSER0, hd->net1, 8);
    if (hd->sta1[0]) SETK(K_KEVNM, hd->sta1, 16);
    if (hd->loc1[0]) SETK(K_KUSER1, hd->loc1, 8);
    if (hd->chn1[0]) SETK(K_KUSER2, hd->chn1, 8);
    FILE *f = fopen(path, "wb");
    if (!f) return -2;
    fwrite(h, 1, SAC_HDR_BYTES, f);
    fwrite(y, sizeof(float), (size_t)npts, f);
    fclose(f);
    return 0;
}
