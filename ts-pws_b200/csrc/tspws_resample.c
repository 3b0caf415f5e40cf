/*
 * tspws_resample.c -- resampling (sub-sampling / jackknife) and convergence drivers.
 * (filled in below the main stack)
 */
#include <stdio.h>
#include <cuda_runtime_api.h>
#include "tspws_internal.h"

int tspws_resample_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                       unsigned int mtr, int max, const char *sel_override, unsigned int selM,
                       cudaStream_t stream)
{
    (void)p; (void)out; (void)in; (void)d_x; (void)mtr; (void)max; (void)sel_override; (void)selM; (void)stream;
    fprintf(stderr, "tspws_b200: resampling is not implemented yet\n");
    return TSPWS_ERR_ARG;
}

int tspws_convergence_run(const t_tsPWS *p, t_tsPWS_out *out, const t_data *in, float *d_x,
                          unsigned int mtr, int max, cudaStream_t stream)
{
    (void)p; (void)out; (void)in; (void)d_x; (void)mtr; (void)max; (void)stream;
    fprintf(stderr, "tspws_b200: convergence curves are not implemented yet\n");
    return TSPWS_ERR_ARG;
}
