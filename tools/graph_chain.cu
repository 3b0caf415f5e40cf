// Floor of a dependent kernel chain replayed as a CUDA graph on B200: empty kernels and kernels with one
// L2 round trip, with and without programmatic dependent launch.  (tail latency budget.)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o graph_chain tools/graph_chain.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_empty(float *p, int pdl)
{
    if (pdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
}
__global__ void k_touch(float *p, int pdl)
{
    if (pdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = __ldcg(p + i);
    p[i] = v + 1.0f;
}
template <typename K>
static void launch(K k, int grid, int block, cudaStream_t s, float *p, int pdl)
{
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, p, pdl));
}
int main()
{
    float *p; CK(cudaMalloc(&p, 1 << 24)); CK(cudaMemset(p, 0, 1 << 24));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int touch = 0; touch < 2; touch++)
    for (int grid : {148, 1184})
    for (int pdl = 0; pdl < 2; pdl++)
    for (int n : {1, 6, 12}) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < n; i++) { if (touch) launch(k_touch, grid, 256, s, p, pdl); else launch(k_empty, grid, 256, s, p, pdl); }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int i = 0; i < 5; i++) CK(cudaGraphLaunch(ge, s));
        CK(cudaStreamSynchronize(s));
        float best = 1e9;
        for (int r = 0; r < 20; r++) {
            CK(cudaEventRecord(e0, s));
            for (int i = 0; i < 10; i++) CK(cudaGraphLaunch(ge, s));
            CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("%s grid %4d pdl %d chain %2d: %6.2f us per graph, %5.2f us per kernel\n", touch ? "touch" : "empty", grid, pdl, n, best * 100, best * 100 / n);
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
    return 0;
}
