// Time-embedding path: sinusoid -> Linear -> SiLU -> Linear, and the 22
// per-ResBlock Linear(SiLU(t_emb)) projections (models/unet/unet.py:182-183,209-220) as ONE batched
// mat-vec over the row-concatenated weight matrix.  n (distinct timesteps) is 1 in sampling and B in
// training-style calls, so this is a weight-streaming (HBM-bound) op: one warp per output row,
// 16-byte loads along K, warp-shuffle reduction, fp32 accumulate.
#define SDK_PDL_CAT 7
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
time_sinusoid_kernel(const long long* __restrict__ t, int n, int dim, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int half = dim >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * half) return;
    const int i = idx / half, j = idx - i * half;
    // fp32 sequence of unet.py:211-215: (-log(10000) * j) / half -> exp -> t.float() * f
    const float f = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)j), (float)half));
    const float x = __fmul_rn((float)t[i], f);
    out[(size_t)i * dim + j] = cosf(x);
    out[(size_t)i * dim + half + j] = sinf(x);
}

__device__ __forceinline__ float act(float v, int code) { return code == 1 ? silu_f(v) : v; }

template <typename TW> struct WVec;
template <> struct WVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float* o) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct WVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* o) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
    }
};

// one warp per output row r; up to NB input vectors per pass (weights are re-streamed for n > NB)
template <typename TW, int NB>
__global__ void __launch_bounds__(256)
gemv_kernel(const TW* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
            int n, int R, int K, int act_in, int act_out) {
    pdl_trigger();
    pdl_wait();
    constexpr int V = WVec<TW>::N;
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const TW* wr = W + (size_t)r * K;
    for (int i0 = 0; i0 < n; i0 += NB) {
        float acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.f;
        for (int k = lane * V; k < K; k += 32 * V) {
            float wv[V];
            WVec<TW>::load(wr + k, wv);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (i0 + b < n) {
                    const float* xr = x + (size_t)(i0 + b) * K + k;
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[b] = fmaf(wv[j], act(__ldg(xr + j), act_in), acc[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float s = warp_sum(acc[b]);
            if (lane == 0 && i0 + b < n) y[(size_t)(i0 + b) * R + r] = act(s + (bias ? __ldg(bias + r) : 0.f), act_out);
        }
    }
}

}  // namespace

extern "C" int sdk_time_sinusoid(const int64_t* t_dev, int n, int dim, float* out, void* stream) {
    SDK_CHECK_ARG(t_dev && out && n > 0 && dim > 0 && dim % 2 == 0, "sdk_time_sinusoid: bad args");
    const int total = n * (dim / 2);
    SDK_CUDA(sdk_launch(time_sinusoid_kernel, dim3((total + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, (const long long*)t_dev, n, dim, out));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

extern "C" int sdk_gemv(const void* W, int w_dtype, const float* bias, const float* x, float* y,
                        int n, int R, int K, int act_in, int act_out, void* stream) {
    SDK_CHECK_ARG(W && x && y && n > 0 && R > 0 && K > 0, "sdk_gemv: bad args");
    SDK_CHECK_ARG(K % 8 == 0, "sdk_gemv: K=%d must be a multiple of 8", K);
    const int grid = (R * 32 + 255) / 256;
    cudaStream_t s = (cudaStream_t)stream;
    if (w_dtype == SDK_F32) {
        if (n == 1) SDK_CUDA(sdk_launch(gemv_kernel<float, 1>, dim3(grid), dim3(256), (size_t)(0), s, (const float*)W, bias, x, y, n, R, K, act_in, act_out));
        else SDK_CUDA(sdk_launch(gemv_kernel<float, 4>, dim3(grid), dim3(256), (size_t)(0), s, (const float*)W, bias, x, y, n, R, K, act_in, act_out));
    } else if (w_dtype == SDK_BF16) {
        if (n == 1) SDK_CUDA(sdk_launch(gemv_kernel<__nv_bfloat16, 1>, dim3(grid), dim3(256), (size_t)(0), s, (const __nv_bfloat16*)W, bias, x, y, n, R, K, act_in, act_out));
        else SDK_CUDA(sdk_launch(gemv_kernel<__nv_bfloat16, 4>, dim3(grid), dim3(256), (size_t)(0), s, (const __nv_bfloat16*)W, bias, x, y, n, R, K, act_in, act_out));
    } else return sdk_fail(SDK_ERR_ARG, "sdk_gemv: w_dtype %d", w_dtype);
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}
