// Micro-benchmark: cost of a kernel boundary inside a CUDA graph on B200, with and without programmatic dependent launch.
// A chain of N dependent kernels (each ~`work` ns of dependent FMAs on `ctas` CTAs), captured from a stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>   // 0: no griddepcontrol; 1: trigger at entry + wait before work; 2: wait, work, trigger at the end
__global__ void __launch_bounds__(256) work_kernel(float* buf, int iters, int smem_touch) {
    extern __shared__ float sm[];
    if (MODE == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (smem_touch) sm[threadIdx.x] = 0.f;           // "prologue" work that does not depend on the previous kernel
    if (MODE != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    float x = buf[blockIdx.x * blockDim.x + threadIdx.x];
    for (int i = 0; i < iters; ++i) x = fmaf(x, 1.0001f, 0.5f);
    buf[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (MODE == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <int MODE>
void launch(float* buf, int ctas, int iters, size_t smem, cudaStream_t s, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, work_kernel<MODE>, buf, iters, smem ? 1 : 0));
}

template <int MODE>
float run_graph(float* buf, int n, int ctas, int iters, size_t smem, bool pdl) {
    cudaStream_t s; CK(cudaStreamCreate(&s));
    CK(cudaFuncSetAttribute(work_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; ++i) launch<MODE>(buf, ctas, iters, smem, s, pdl);
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(ge, s));
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(s));
    return ms * 1e3f / (10.f * n);
}

int main() {
    float* buf; CK(cudaMalloc(&buf, 4096 * 256 * 4)); CK(cudaMemset(buf, 0, 4096 * 256 * 4));
    const int n = 200;
    for (int ctas : {148, 296, 1184}) {
        for (size_t smem : {(size_t)0, (size_t)100 * 1024, (size_t)200 * 1024}) {
            for (int iters : {0, 500, 2000}) {
                const float a = run_graph<0>(buf, n, ctas, iters, smem, false);
                const float b = run_graph<1>(buf, n, ctas, iters, smem, true);
                const float c = run_graph<2>(buf, n, ctas, iters, smem, true);
                printf("ctas %5d smem %6zu iters %5d : plain %6.2f us/kernel | PDL early trigger %6.2f | PDL late trigger %6.2f\n", ctas, smem, iters, a, b, c);
            }
        }
    }
    return 0;
}
