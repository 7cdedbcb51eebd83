// Micro-benchmark: how fast can ONE resident CTA per SM drain a 128 x BN fp32 accumulator tile to global memory?
// (The tcgen05 GEMM epilogue's situation: few warps per SM, burst of stores from every SM at once.)
// Variants: plain STG.128 with 4/8/16 warps, row pitch of the real output vs contiguous, and TMA bulk stores from smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// each warp instruction writes 4 rows x 128 B (the transposed-epilogue mapping); rows of the tile are `pitch` floats apart
__global__ void stg_kernel(float* out, int pitch, int bn, int rows_per_cta, unsigned long long* t) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    float* base = out + (size_t)blockIdx.x * rows_per_cta * pitch + blockIdx.y * bn;
    __syncthreads();
    const unsigned long long t0 = gtime();
    const float4 v = make_float4(lane, warp, 1.f, 2.f);
    const int chunks = bn / 32, groups = rows_per_cta / 4;         // (chunk, 4-row group) items, dealt round-robin to warps
    for (int item = warp; item < chunks * groups; item += nw) {
        const int c = item / groups, g = item - c * groups;
        *reinterpret_cast<float4*>(base + (size_t)(g * 4 + sub_r) * pitch + c * 32 + c4) = v;
    }
    __syncthreads();
    const unsigned long long t1 = gtime();
    if (threadIdx.x == 0) { t[2 * (blockIdx.y * gridDim.x + blockIdx.x)] = t0; t[2 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = t1; }
}

// per-thread-row mapping (what tcgen05.ld 32x32b hands out): lane = row, 8 x STG.128 along the row per chunk
__global__ void stg_rows_kernel(float* out, int pitch, int bn, int rows_per_cta, unsigned long long* t) {
    const int row = threadIdx.x;                                    // 128 threads
    float* base = out + (size_t)blockIdx.x * rows_per_cta * pitch + blockIdx.y * bn + (size_t)row * pitch;
    __syncthreads();
    const unsigned long long t0 = gtime();
    const float4 v = make_float4(row, 0.f, 1.f, 2.f);
    for (int c = 0; c < bn; c += 4) *reinterpret_cast<float4*>(base + c) = v;
    __syncthreads();
    const unsigned long long t1 = gtime();
    if (threadIdx.x == 0) { t[2 * (blockIdx.y * gridDim.x + blockIdx.x)] = t0; t[2 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = t1; }
}

// TMA: tile staged in smem as [chunk][128 rows][32 floats]; one elected thread issues one 2-D tensor store per chunk
__global__ void tma_kernel(const __grid_constant__ CUtensorMap tm, int bn, int rows_per_cta, unsigned long long* t) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < bn * rows_per_cta; i += blockDim.x) s[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) {
        for (int c = 0; c < bn / 32; ++c) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s + c * 32 * rows_per_cta);
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         :: "l"(&tm), "r"(blockIdx.y * bn + c * 32), "r"(blockIdx.x * rows_per_cta), "r"(sa) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // full completion (writes performed)
    }
    __syncthreads();
    const unsigned long long t1 = gtime();
    if (threadIdx.x == 0) { t[2 * (blockIdx.y * gridDim.x + blockIdx.x)] = t0; t[2 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = t1; }
}

static void report(const char* name, std::vector<unsigned long long>& h, int ctas, double bytes_per_cta) {
    unsigned long long lo = ~0ull, hi = 0; double avg = 0;
    for (int i = 0; i < ctas; ++i) { lo = std::min(lo, h[2 * i]); hi = std::max(hi, h[2 * i + 1]); avg += (double)(h[2 * i + 1] - h[2 * i]); }
    avg /= ctas;
    printf("%-44s per-CTA %7.0f ns  (%5.1f B/ns/SM)   all CTAs %7.0f ns  (%6.0f GB/s chip)\n", name, avg, bytes_per_cta / avg,
           (double)(hi - lo), bytes_per_cta * ctas / (double)(hi - lo));
}

int main() {
    const int M = 8192, N = 320, BN = 160, ROWS = 128;
    float* out; CK(cudaMalloc(&out, (size_t)M * N * 4));
    unsigned long long* t; CK(cudaMalloc(&t, 4096 * 16));
    const dim3 grid(M / ROWS, N / BN);
    const int ctas = grid.x * grid.y;
    std::vector<unsigned long long> h(2 * ctas);
    const double bpc = (double)ROWS * BN * 4;
    CUtensorMap tm;
    {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)N * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)ROWS}; cuuint32_t es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    }
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * BN * 4 + 1024));
    for (int rep = 0; rep < 3; ++rep) {
        printf("--- rep %d (M=%d N=%d fp32, tile %dx%d, %d CTAs)\n", rep, M, N, ROWS, BN, ctas);
        for (int warps : {4, 8, 16}) {
            stg_kernel<<<grid, warps * 32>>>(out, N, BN, ROWS, t); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), t, ctas * 16, cudaMemcpyDeviceToHost));
            char nm[96]; snprintf(nm, 96, "STG.128 4rows x 128B per instr, %2d warps", warps); report(nm, h, ctas, bpc);
        }
        stg_rows_kernel<<<grid, 128>>>(out, N, BN, ROWS, t); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), t, ctas * 16, cudaMemcpyDeviceToHost)); report("STG.128 thread-per-row (32 lines per instr)", h, ctas, bpc);
        tma_kernel<<<grid, 128, ROWS * BN * 4 + 1024>>>(tm, BN, ROWS, t); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), t, ctas * 16, cudaMemcpyDeviceToHost)); report("TMA 2-D tensor store, 5 boxes of 128x32", h, ctas, bpc);
        // same, only 16 CTAs active (is the limit per SM or chip-wide?)
        stg_kernel<<<dim3(8, 2), 128>>>(out, N, BN, ROWS, t); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), t, 16 * 16, cudaMemcpyDeviceToHost)); report("STG.128 4 warps, only 16 CTAs", h, 16, bpc);
        tma_kernel<<<dim3(8, 2), 128, ROWS * BN * 4 + 1024>>>(tm, BN, ROWS, t); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), t, 16 * 16, cudaMemcpyDeviceToHost)); report("TMA store, only 16 CTAs", h, 16, bpc);
    }
    return 0;
}
