// Shared helpers for the sdb200 kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define SDK_OK 0
#define SDK_ERR_ARG (-1)
#define SDK_ERR_CUDA (-2)
#define SDK_ERR_UNSUPPORTED (-3)

// thread-local last-error text, readable through sdk_last_error()
char* sdk_err_buf();
int sdk_fail(int code, const char* fmt, ...);

#define SDK_CHECK_ARG(cond, ...) \
    do { if (!(cond)) return sdk_fail(SDK_ERR_ARG, __VA_ARGS__); } while (0)

#define SDK_CUDA(call) \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
        return sdk_fail(SDK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

#define SDK_LAUNCH_CHECK() \
    do { cudaError_t e_ = cudaPeekAtLastError(); if (e_ != cudaSuccess) \
        return sdk_fail(SDK_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
static inline int sdk_num_sms() {
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = n[dev & 63];
    if (c == 0) { cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev); if (c <= 0) c = 148; }
    return c;
}

// Opt `fn` in for `bytes` of dynamic shared memory on the CURRENT device (cudaFuncAttributeMaxDynamicSharedMemorySize is
// per device); remembers the largest size set per (device, kernel) so the hot path is one table lookup.
cudaError_t sdk_ensure_dyn_smem(const void* fn, int bytes);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// exact-erf GELU (nn.GELU() default, reference models/activation_fn.py:15)
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// Same function for epilogues whose result is rounded to bf16: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below
// half a bf16 ulp) with MUFU rcp / ex2 -- 13 instructions instead of erff's two-branch polynomial.  The GEGLU epilogue
// evaluates it 84 M times per step at UNet batch 16 and is instruction-bound.
__device__ __forceinline__ float gelu_erf_bf16out(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float erf_abs = fmaf(-poly, e, 1.0f);
    return 0.5f * x + 0.5f * fabsf(x) * erf_abs;        // x * (1 + sign(x) erf|z|) / 2
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- programmatic dependent launch (PDL): every kernel of the step program is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, calls pdl_trigger() first (lets the NEXT kernel's
// CTAs be scheduled as soon as this grid has fully started) and pdl_wait() before its first access to global
// memory (blocks until the PREVIOUS grid has completed and flushed).  Prologues (smem carve-up, mbarrier
// init, TMEM alloc, descriptor prefetch) and launch latency of kernel N+1 thereby overlap kernel N's tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool sdk_pdl_enabled();
// per-source-file category (SDK_PDL_CAT, defined before including this header) so that PDL can be switched off for one family of
// kernels while hunting an ordering bug: SDB200_PDL_MASK is a bit mask of the categories that MAY use PDL (default: all)
#ifndef SDK_PDL_CAT
#define SDK_PDL_CAT 0
#endif
bool sdk_pdl_enabled_cat(int cat);
// one-time per kernel: ask for the max-shared-memory L1 carve-out so that the SM configuration does not flip between the
// big-smem tensor-core kernels and the small elementwise kernels that run in between (a carve-out change drains the SM)
void sdk_prefer_max_smem_once(const void* fn);

template <typename... KArgs, typename... Args>
static inline cudaError_t sdk_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = sdk_pdl_enabled_cat(SDK_PDL_CAT) ? 1 : 0;
    sdk_prefer_max_smem_once(reinterpret_cast<const void*>(kernel));
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// dtype codes used across the C ABI
#define SDK_F32 0
#define SDK_BF16 1
