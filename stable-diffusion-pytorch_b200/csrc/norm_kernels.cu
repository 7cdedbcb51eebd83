// Bandwidth-bound normalisation / layout kernels of the UNet.
//
// Activations are NHWC == [rows = B*H*W][C] with the fp32 "residual stream" as the stored type.
// Each kernel reads its input once with 16-byte vector loads along C (coalesced), reduces with warp
// shuffles, and writes the operand for the following GEMM in that GEMM's operand type (bf16 on the
// tensor-core path, fp32 on the exact path), so cast / SiLU / concat / nearest-upsample never cost a
// pass of their own.
//
//   GroupNorm(32 groups) [+SiLU] over a virtual channel-concat of two sources  (unet.py:66,157,160,343,399)
//   LayerNorm over C                                                             (unet.py:102-108)
//   cast (+ nearest 2x upsample)                                                 (unet.py:250)
//   NCHW -> NHWC for the 4-channel latent                                        (unet.py:256 input)
#define SDK_PDL_CAT 5
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>

namespace {

constexpr int GROUPS = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_CHUNKS = 160;

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics.  grid = (chunks, B), ~2 CTAs per SM over the whole batch.  A CTA owns a run of
// pixels; its 256 threads are laid out as (pixel lanes) x (channel quads) so that a warp reads
// consecutive 16-byte quads of one pixel row (coalesced) and each thread keeps 4-deep loads in flight.
// Per-thread fp32 partials (a few dozen values each) go to shared memory; one thread per group folds
// them in a fixed order in double; the last CTA of a sample (atomic ticket) folds the per-chunk
// partials in chunk order -> deterministic mean / rstd with no fp32 cancellation issue.
// ---------------------------------------------------------------------------------------------
struct GNStatsArgs {
    const float* src0; const float* src1;
    int C0, C1, HW, chunks, pix_per_chunk, px_lanes, q_iters;
    double* partial;        // [B][chunks][GROUPS][2]
    unsigned int* ticket;   // [B], zero on entry, self-resetting
    float* stats;           // [B][GROUPS][2] = mean, rstd
    float eps;
};

__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(GNStatsArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_part[];            // [nq][px_lanes][8] : 4 sums + 4 sums of squares per (quad, pixel lane)
    __shared__ bool s_last;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int C = a.C0 + a.C1, cpg = C / GROUPS, nq = C >> 2;
    const int p0 = chunk * a.pix_per_chunk;
    const int p1 = min(a.HW, p0 + a.pix_per_chunk);
    const int qlanes = GN_THREADS / a.px_lanes;          // threads along the channel-quad axis
    const int pl = threadIdx.x / qlanes, ql = threadIdx.x - pl * qlanes;
    if (pl < a.px_lanes) {
        for (int it = 0; it < a.q_iters; ++it) {
            const int q = ql + it * qlanes;
            if (q >= nq) break;
            const int c = q << 2;
            const float* base; int cs, cl;
            if (c < a.C0) { base = a.src0; cs = a.C0; cl = c; } else { base = a.src1; cs = a.C1; cl = c - a.C0; }
            const float* p = base + ((size_t)b * a.HW) * cs + cl;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
            int pix = p0 + pl;
            const int stride = a.px_lanes;
            for (; pix + 3 * stride < p1; pix += 4 * stride) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + stride) * cs));
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 2 * stride) * cs));
                const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 3 * stride) * cs));
                s[0] += (v0.x + v1.x) + (v2.x + v3.x); s[1] += (v0.y + v1.y) + (v2.y + v3.y);
                s[2] += (v0.z + v1.z) + (v2.z + v3.z); s[3] += (v0.w + v1.w) + (v2.w + v3.w);
                ss[0] += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
                ss[1] += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
                ss[2] += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
                ss[3] += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
            }
            for (; pix < p1; pix += stride) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
            }
            float* dst = s_part + ((size_t)q * a.px_lanes + pl) * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(ss[0], ss[1], ss[2], ss[3]);
        }
    }
    __syncthreads();
    double* part = a.partial + ((size_t)b * a.chunks + chunk) * GROUPS * 2;
    if (threadIdx.x < GROUPS) {
        const int g = threadIdx.x;
        double sum = 0.0, sq = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float* src = s_part + (size_t)(c >> 2) * a.px_lanes * 8 + (c & 3);
            for (int l = 0; l < a.px_lanes; ++l) { sum += (double)src[l * 8]; sq += (double)src[l * 8 + 4]; }
        }
        part[g * 2] = sum; part[g * 2 + 1] = sq;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&a.ticket[b], 1u) == (unsigned)a.chunks - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // last CTA of the sample: 8 slices x 32 groups fold the per-chunk partials (fixed assignment and order)
        __shared__ double s_red[8][GROUPS][2];
        const int g = threadIdx.x & 31, sl = threadIdx.x >> 5;
        const double* pp = a.partial + (size_t)b * a.chunks * GROUPS * 2 + g * 2;
        // all loads of this thread are issued before the first add (one L2 round trip, not chunks/8 of them);
        // the additions still run in index order
        constexpr int MAXL = GN_MAX_CHUNKS / 8;
        double va[MAXL], vb[MAXL];
        const int nl = (a.chunks - sl + 7) / 8;
#pragma unroll
        for (int i = 0; i < MAXL; ++i) {
            if (i < nl) {
                const double2 t = __ldcg(reinterpret_cast<const double2*>(pp + (size_t)(sl + 8 * i) * GROUPS * 2));
                va[i] = t.x; vb[i] = t.y;
            }
        }
        double sum = 0.0, sq = 0.0;
#pragma unroll
        for (int i = 0; i < MAXL; ++i) if (i < nl) { sum += va[i]; sq += vb[i]; }
        s_red[sl][g][0] = sum; s_red[sl][g][1] = sq;
        __syncthreads();
        if (threadIdx.x < GROUPS) {
            sum = 0.0; sq = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { sum += s_red[i][g][0]; sq += s_red[i][g][1]; }
            const double n = (double)cpg * (double)a.HW;
            const double mean = sum / n;
            double var = sq / n - mean * mean;      // biased variance (nn.GroupNorm)
            if (var < 0.0) var = 0.0;
            a.stats[((size_t)b * GROUPS + g) * 2] = (float)mean;
            a.stats[((size_t)b * GROUPS + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)a.eps));
        }
    }
    if (threadIdx.x == 0) a.ticket[b] = 0u;
}

// GroupNorm apply (+SiLU) (+ raw cast copy).  One thread per channel quad per pixel; grid-stride over rows.
struct GNApplyArgs {
    const float* src0; const float* src1;
    int C0, C1, HW, B;
    const float* stats; const float* gamma; const float* beta;
    void* out; void* raw_out;
    int silu;
    // alternative to `stats`: per-channel (sum, sum of squares) of each source in double, [B][C0] and [B][C1], written by the kernels
    // that produced the sources (sdk_tc_gemm_set_stats / sdk_channel_stats); the group fold happens in this kernel's prologue
    const double2* cs0; const double2* cs1; float eps;
};

template <typename TOut>
__device__ __forceinline__ void store4(TOut* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u; u.x = *reinterpret_cast<unsigned*>(&lo); u.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

// grid = (pixel chunks, B).  Threads are laid out (pixel lanes) x (channel quads) as in the statistics kernel; a thread
// keeps gamma/beta and the (mean, rstd) of its 4 channels in registers and walks its pixels with 4 loads in flight --
// no integer division and no table lookups in the inner loop.
template <typename TOut>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(GNApplyArgs a, int pix_per_chunk, int px_lanes, int q_iters) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int C = a.C0 + a.C1, cpg = C / GROUPS, nq = C >> 2;
    const int p0 = blockIdx.x * pix_per_chunk;
    const int p1 = min(a.HW, p0 + pix_per_chunk);
    const int qlanes = GN_THREADS / px_lanes;
    const int pl = threadIdx.x / qlanes, ql = threadIdx.x - pl * qlanes;
    __shared__ float2 s_gstat[GROUPS];
    if (a.cs0) {
        // 8 threads per group: channel sums -> (mean, rstd), fixed order (double accumulation, as gn_stats_kernel's fold)
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        double sum = 0.0, sq = 0.0;
        const double inv_n = 1.0 / ((double)cpg * (double)a.HW);      // the one double division: independent of the loads below
        // all table loads of a thread in flight at once (up to 8 per batch = 64 channels per group and batch): the fold is the head
        // of every launch's dependency chain, and one L2 round trip per loop iteration was most of a small launch's run time
        for (int c0 = g * cpg + sub; c0 < (g + 1) * cpg; c0 += 64) {
            double2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = c0 + 8 * i;
                v[i] = c >= (g + 1) * cpg ? make_double2(0.0, 0.0)
                     : (c < a.C0 ? __ldg(a.cs0 + (size_t)b * a.C0 + c) : __ldg(a.cs1 + (size_t)b * a.C1 + (c - a.C0)));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { sum += v[i].x; sq += v[i].y; }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
        if (sub == 0) {
            // sums stay in double (cancellation in E[x^2] - mean^2); the reciprocal square root is taken in fp32 like
            // nn.GroupNorm's own rstd -- a double division + square root here was ~0.5 us at the head of every launch
            const double mean = sum * inv_n;
            double var = sq * inv_n - mean * mean;      // biased variance (nn.GroupNorm)
            if (var < 0.0) var = 0.0;
            s_gstat[g] = make_float2((float)mean, 1.0f / sqrtf((float)var + a.eps));
        }
        __syncthreads();
    }
    if (pl >= px_lanes) return;
    TOut* out = reinterpret_cast<TOut*>(a.out) + (size_t)b * a.HW * C;
    TOut* raw = a.raw_out ? reinterpret_cast<TOut*>(a.raw_out) + (size_t)b * a.HW * C : nullptr;
    for (int it = 0; it < q_iters; ++it) {
        const int q = ql + it * qlanes;
        if (q >= nq) break;
        const int c = q << 2;
        const float* base; int cs, cl;
        if (c < a.C0) { base = a.src0; cs = a.C0; cl = c; } else { base = a.src1; cs = a.C1; cl = c - a.C0; }
        const float* p = base + ((size_t)b * a.HW) * cs + cl;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
        float sc[4], sh[4];                                  // y = x * sc + sh
        {
            const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 st = a.cs0 ? s_gstat[(c + j) / cpg] : __ldg(reinterpret_cast<const float2*>(a.stats) + (size_t)b * GROUPS + (c + j) / cpg);
                sc[j] = st.y * gs[j];
                sh[j] = bs[j] - st.x * sc[j];
            }
        }
        auto emit = [&](int pix, const float4& v) {
            float y0 = fmaf(v.x, sc[0], sh[0]), y1 = fmaf(v.y, sc[1], sh[1]), y2 = fmaf(v.z, sc[2], sh[2]), y3 = fmaf(v.w, sc[3], sh[3]);
            if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
            store4<TOut>(out + (size_t)pix * C + c, y0, y1, y2, y3);
            if (raw) store4<TOut>(raw + (size_t)pix * C + c, v.x, v.y, v.z, v.w);
        };
        int pix = p0 + pl;
        const int stride = px_lanes;
        for (; pix + 3 * stride < p1; pix += 4 * stride) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + stride) * cs));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 2 * stride) * cs));
            const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 3 * stride) * cs));
            emit(pix, v0); emit(pix + stride, v1); emit(pix + 2 * stride, v2); emit(pix + 3 * stride, v3);
        }
        for (; pix < p1; pix += stride) emit(pix, __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs)));
    }
}

// ---------------------------------------------------------------------------------------------
// Fused GroupNorm: statistics + apply in ONE cooperative launch (one grid barrier), deterministic.
//   phase 1: work item (sample, chunk<=32) -> per-group (sum, sumsq) partials in double, as gn_stats_kernel
//   grid barrier (cooperative groups)
//   phase 2: every CTA folds the <=32 partials of the sample(s) it touches in fixed order (16 KiB of L2
//            reads) and normalises its share of rows.  Saves a launch and the serialized "last CTA" tail.
// ---------------------------------------------------------------------------------------------
struct GNFusedArgs {
    const float* src0; const float* src1;
    int C0, C1, HW, B, chunks, pix_per_chunk, px_lanes, q_iters;
    double* partial;        // [B][chunks][GROUPS][2]
    const float* gamma; const float* beta;
    void* out; void* raw_out;
    float eps; int silu;
};

template <typename TOut>
__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(GNFusedArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_part[];            // phase 1: [nq][px_lanes][8]; phase 2: stats [GROUPS][2]
    const int C = a.C0 + a.C1, cpg = C / GROUPS, nq = C >> 2;
    const int qlanes = GN_THREADS / a.px_lanes;
    const int pl = threadIdx.x / qlanes, ql = threadIdx.x - pl * qlanes;
    // ---------------- phase 1
    for (int w = blockIdx.x; w < a.B * a.chunks; w += gridDim.x) {
        const int b = w / a.chunks, chunk = w - b * a.chunks;
        const int p0 = chunk * a.pix_per_chunk;
        const int p1 = min(a.HW, p0 + a.pix_per_chunk);
        if (pl < a.px_lanes) {
            for (int it = 0; it < a.q_iters; ++it) {
                const int q = ql + it * qlanes;
                if (q >= nq) break;
                const int c = q << 2;
                const float* base; int cs, cl;
                if (c < a.C0) { base = a.src0; cs = a.C0; cl = c; } else { base = a.src1; cs = a.C1; cl = c - a.C0; }
                const float* p = base + ((size_t)b * a.HW) * cs + cl;
                float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
                int pix = p0 + pl;
                const int stride = a.px_lanes;
                for (; pix + 3 * stride < p1; pix += 4 * stride) {
                    const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                    const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + stride) * cs));
                    const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 2 * stride) * cs));
                    const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 3 * stride) * cs));
                    s[0] += (v0.x + v1.x) + (v2.x + v3.x); s[1] += (v0.y + v1.y) + (v2.y + v3.y);
                    s[2] += (v0.z + v1.z) + (v2.z + v3.z); s[3] += (v0.w + v1.w) + (v2.w + v3.w);
                    ss[0] += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
                    ss[1] += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
                    ss[2] += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
                    ss[3] += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
                }
                for (; pix < p1; pix += stride) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                    ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
                }
                float* dst = s_part + ((size_t)q * a.px_lanes + pl) * 8;
                *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(ss[0], ss[1], ss[2], ss[3]);
            }
        }
        __syncthreads();
        if (threadIdx.x < GROUPS) {
            const int g = threadIdx.x;
            double sum = 0.0, sq = 0.0;
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float* src = s_part + (size_t)(c >> 2) * a.px_lanes * 8 + (c & 3);
                for (int l = 0; l < a.px_lanes; ++l) { sum += (double)src[l * 8]; sq += (double)src[l * 8 + 4]; }
            }
            double* part = a.partial + ((size_t)b * a.chunks + chunk) * GROUPS * 2;
            part[g * 2] = sum; part[g * 2 + 1] = sq;
        }
        __syncthreads();
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    // ---------------- phase 2
    float2* s_stats = reinterpret_cast<float2*>(s_part);          // [GROUPS] (mean, rstd) of the cached sample
    const long long rows = (long long)a.B * a.HW;
    const long long r_lo = rows * blockIdx.x / gridDim.x, r_hi = rows * (blockIdx.x + 1) / gridDim.x;
    int cached_b = -1;
    for (long long r0 = r_lo; r0 < r_hi;) {
        const int b = (int)(r0 / a.HW);
        const long long r1 = min(r_hi, (long long)(b + 1) * a.HW);
        if (b != cached_b) {
            __syncthreads();
            {
                // 8 slices x 32 groups: each thread loads <= 4 partials (one L2 round trip), then a fixed-order fold
                double* s_red = reinterpret_cast<double*>(s_part) + 64;          // [8][GROUPS][2] after the stats slots
                const int g = threadIdx.x & 31, sl = threadIdx.x >> 5;
                const double* pp = a.partial + (size_t)b * a.chunks * GROUPS * 2 + g * 2;
                double2 t[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = sl + 8 * i;
                    t[i] = k < a.chunks ? __ldcg(reinterpret_cast<const double2*>(pp + (size_t)k * GROUPS * 2)) : make_double2(0.0, 0.0);
                }
                s_red[(sl * GROUPS + g) * 2] = ((t[0].x + t[1].x) + t[2].x) + t[3].x;
                s_red[(sl * GROUPS + g) * 2 + 1] = ((t[0].y + t[1].y) + t[2].y) + t[3].y;
                __syncthreads();
                if (threadIdx.x < GROUPS) {
                    double sum = 0.0, sq = 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { sum += s_red[(i * GROUPS + g) * 2]; sq += s_red[(i * GROUPS + g) * 2 + 1]; }
                    const double n = (double)cpg * (double)a.HW;
                    const double mean = sum / n;
                    double var = sq / n - mean * mean;
                    if (var < 0.0) var = 0.0;
                    s_stats[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)a.eps)));
                }
            }
            __syncthreads();
            cached_b = b;
        }
        const long long items = (r1 - r0) * nq;
        for (long long i = threadIdx.x; i < items; i += GN_THREADS) {
            const int q = (int)(i % nq);
            const long long row = r0 + i / nq;
            const int c = q << 2;
            const float* p = (c < a.C0) ? a.src0 + row * a.C0 + c : a.src1 + row * a.C1 + (c - a.C0);
            const float4 v = __ldg(reinterpret_cast<const float4*>(p));
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
            const float xs[4] = {v.x, v.y, v.z, v.w}, gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 st = s_stats[(c + j) / cpg];
                const float y = (xs[j] - st.x) * st.y * gs[j] + bs[j];
                o[j] = a.silu ? silu_f(y) : y;
            }
            store4<TOut>(reinterpret_cast<TOut*>(a.out) + row * C + c, o[0], o[1], o[2], o[3]);
            if (a.raw_out) store4<TOut>(reinterpret_cast<TOut*>(a.raw_out) + row * C + c, xs[0], xs[1], xs[2], xs[3]);
        }
        r0 = r1;
    }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm in ONE launch with a thread-block cluster per sample (the default in the step program).
//   grid = (CS, B), cluster = (CS,1,1), 1024 threads per CTA.  CTA r owns a contiguous run of the sample's pixels.
//   pass 1: per-group (sum, sumsq) of the CTA's pixels -> 32 doubles x 2 in its shared memory
//   cluster barrier; every CTA reads all CS partials through distributed shared memory in rank order
//   (deterministic), derives mean / rstd, then pass 2 normalises its own pixels (second read hits L2).
// No global atomics, tickets, fences or second launch: at UNet batch 2 a GroupNorm is latency-, not
// bandwidth-bound, and this removes a launch plus the serialized last-CTA tail of the two-kernel form.
// ---------------------------------------------------------------------------------------------
constexpr int GNC_THREADS = 1024;

__device__ __forceinline__ double dsmem_ld_f64(const double* local, unsigned rank) {
    unsigned laddr = (unsigned)__cvta_generic_to_shared(local), raddr;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(raddr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <typename TOut>
__global__ void __launch_bounds__(GNC_THREADS, 1)
gn_cluster_kernel(GNApplyArgs a, float eps, int CS, int px_lanes, int q_iters) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_part[];                       // [nq][px_lanes][8]
    __shared__ double s_cl[GROUPS][2];                      // this CTA's per-group partial (read by the whole cluster)
    __shared__ float2 s_stats[GROUPS];
    const int b = blockIdx.y, rank = blockIdx.x;            // cluster spans blockIdx.x
    const int C = a.C0 + a.C1, cpg = C / GROUPS, nq = C >> 2;
    const int ppc = (a.HW + CS - 1) / CS;
    const int p0 = rank * ppc, p1 = min(a.HW, p0 + ppc);
    const int qlanes = GNC_THREADS / px_lanes;
    const int pl = threadIdx.x / qlanes, ql = threadIdx.x - pl * qlanes;
    const bool active = pl < px_lanes;
    // ---- pass 1
    if (active) {
        for (int it = 0; it < q_iters; ++it) {
            const int q = ql + it * qlanes;
            if (q >= nq) break;
            const int c = q << 2;
            const float* base; int cs, cl;
            if (c < a.C0) { base = a.src0; cs = a.C0; cl = c; } else { base = a.src1; cs = a.C1; cl = c - a.C0; }
            const float* p = base + ((size_t)b * a.HW) * cs + cl;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
            int pix = p0 + pl;
            const int stride = px_lanes;
            for (; pix + 3 * stride < p1; pix += 4 * stride) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + stride) * cs));
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 2 * stride) * cs));
                const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 3 * stride) * cs));
                s[0] += (v0.x + v1.x) + (v2.x + v3.x); s[1] += (v0.y + v1.y) + (v2.y + v3.y);
                s[2] += (v0.z + v1.z) + (v2.z + v3.z); s[3] += (v0.w + v1.w) + (v2.w + v3.w);
                ss[0] += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
                ss[1] += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
                ss[2] += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
                ss[3] += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
            }
            for (; pix < p1; pix += stride) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
            }
            float* dst = s_part + ((size_t)q * px_lanes + pl) * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(ss[0], ss[1], ss[2], ss[3]);
        }
    }
    __syncthreads();
    if (threadIdx.x < GROUPS) {
        const int g = threadIdx.x;
        double sum = 0.0, sq = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float* src = s_part + (size_t)(c >> 2) * px_lanes * 8 + (c & 3);
            for (int l = 0; l < px_lanes; ++l) { sum += (double)src[l * 8]; sq += (double)src[l * 8 + 4]; }
        }
        s_cl[g][0] = sum; s_cl[g][1] = sq;
    }
    cluster_arrive();
    cluster_wait();
    if (threadIdx.x < GROUPS) {
        const int g = threadIdx.x;
        double sum = 0.0, sq = 0.0;
        for (int r = 0; r < CS; ++r) { sum += dsmem_ld_f64(&s_cl[g][0], r); sq += dsmem_ld_f64(&s_cl[g][1], r); }
        const double n = (double)cpg * (double)a.HW;
        const double mean = sum / n;
        double var = sq / n - mean * mean;                  // biased variance (nn.GroupNorm)
        if (var < 0.0) var = 0.0;
        s_stats[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    cluster_arrive();                                       // our remote reads are done; waited on just before exit
    __syncthreads();
    // ---- pass 2
    if (active) {
        TOut* out = reinterpret_cast<TOut*>(a.out) + (size_t)b * a.HW * C;
        TOut* raw = a.raw_out ? reinterpret_cast<TOut*>(a.raw_out) + (size_t)b * a.HW * C : nullptr;
        for (int it = 0; it < q_iters; ++it) {
            const int q = ql + it * qlanes;
            if (q >= nq) break;
            const int c = q << 2;
            const float* base; int cs, cl;
            if (c < a.C0) { base = a.src0; cs = a.C0; cl = c; } else { base = a.src1; cs = a.C1; cl = c - a.C0; }
            const float* p = base + ((size_t)b * a.HW) * cs + cl;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
            const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
            float sc[4], sh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 st = s_stats[(c + j) / cpg];
                sc[j] = st.y * gs[j];
                sh[j] = bs[j] - st.x * sc[j];
            }
            auto emit = [&](int pix, const float4& v) {
                float y0 = fmaf(v.x, sc[0], sh[0]), y1 = fmaf(v.y, sc[1], sh[1]), y2 = fmaf(v.z, sc[2], sh[2]), y3 = fmaf(v.w, sc[3], sh[3]);
                if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
                store4<TOut>(out + (size_t)pix * C + c, y0, y1, y2, y3);
                if (raw) store4<TOut>(raw + (size_t)pix * C + c, v.x, v.y, v.z, v.w);
            };
            int pix = p0 + pl;
            const int stride = px_lanes;
            for (; pix + 3 * stride < p1; pix += 4 * stride) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + stride) * cs));
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 2 * stride) * cs));
                const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (size_t)(pix + 3 * stride) * cs));
                emit(pix, v0); emit(pix + stride, v1); emit(pix + 2 * stride, v2); emit(pix + 3 * stride, v3);
            }
            for (; pix < p1; pix += stride) emit(pix, __ldg(reinterpret_cast<const float4*>(p + (size_t)pix * cs)));
        }
    }
    cluster_wait();                                         // nobody may exit while a peer still reads its s_cl
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C <= 2048), exact two-pass statistics.
// ---------------------------------------------------------------------------------------------
template <typename TOut, int MAXQ>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, TOut* __restrict__ out, int C, long long rows) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nquads = C >> 2;
    for (long long row = warp; row < rows; row += nwarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + row * C);
        float4 v[MAXQ];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int q = lane + k * 32;
            if (q < nquads) { v[k] = __ldg(xr + q); s += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
        }
        const float mean = warp_sum(s) / (float)C;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int q = lane + k * 32;
            if (q < nquads) {
                const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
                ss += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int q = lane + k * 32;
            if (q < nquads) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + q);
                store4<TOut>(out + row * C + (q << 2),
                             (v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y,
                             (v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w);
            }
        }
    }
}

// cast fp32 NHWC -> TOut NHWC with optional nearest 2x upsample (dst pixel (y,x) <- src (y/2,x/2)).
template <typename TOut>
__global__ void __launch_bounds__(256)
cast_upsample_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int B, int H, int W, int C, int up) {
    pdl_trigger();
    pdl_wait();
    const int Ho = H * up, Wo = W * up, nquads = C >> 2;
    const long long total = (long long)B * Ho * Wo * nquads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % nquads);
        long long r = i / nquads;
        const int x = (int)(r % Wo); r /= Wo;
        const int y = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (((size_t)b * H + y / up) * W + x / up) * C) + q);
        store4<TOut>(dst + (i << 2), v.x, v.y, v.z, v.w);
    }
}

__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B_src, int B_dst, int C, int HW) {
    pdl_trigger();
    pdl_wait();
    // dst[b][p][c] = src[b % B_src][c][p]   (b % B_src implements latent.repeat(2,1,1,1), diffusion.py:228)
    const long long total = (long long)B_dst * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int p = (int)(r % HW);
        const int b = (int)(r / HW);
        dst[i] = __ldg(src + ((size_t)(b % B_src) * C + c) * HW + p);
    }
}

inline int grid_for(long long items, int threads) {
    long long blocks = (items + threads - 1) / threads;
    long long cap = (long long)sdk_num_sms() * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int64_t sdk_groupnorm_workspace_bytes(int B, int HW) {
    // tickets [B] (256-byte aligned) + partial sums [B][chunks <= 512][32][2] doubles
    return (((int64_t)B * sizeof(unsigned int) + 255) / 256) * 256 + (int64_t)B * GN_MAX_CHUNKS * GROUPS * 2 * sizeof(double);
}

// stats [B][32][2] fp32 (mean, rstd).  workspace must be zero-initialised ONCE (tickets self-reset).
extern "C" int sdk_groupnorm_stats(const float* src0, int C0, const float* src1, int C1, int B, int HW, float eps,
                                   float* stats, void* workspace, void* stream) {
    SDK_CHECK_ARG(src0 && stats && workspace, "sdk_groupnorm_stats: null pointer");
    const int C = C0 + C1;
    SDK_CHECK_ARG(C0 > 0 && C1 >= 0 && (C1 == 0 || src1), "sdk_groupnorm_stats: bad sources");
    SDK_CHECK_ARG(C % GROUPS == 0 && C0 % 4 == 0 && C1 % 4 == 0, "sdk_groupnorm_stats: C=%d+%d must be a multiple of 32 (quads of 4)", C0, C1);
    SDK_CHECK_ARG(B > 0 && B < 65536 && HW > 0, "sdk_groupnorm_stats: bad B/HW");
    const int nq = C / 4;
    SDK_CHECK_ARG(nq * 32 <= 48 * 1024 / 1 && nq <= 4096, "sdk_groupnorm_stats: C too large");
    int chunks = (sdk_num_sms() * 2 + B - 1) / B;      // ~2 CTAs per SM over the whole batch ...
    const long long bytes_per_sample = (long long)HW * C * 4;
    const int by_bytes = (int)((bytes_per_sample + 49151) / 49152);   // ... but >= 48 KiB of input per CTA: small tensors are
    if (chunks > by_bytes) chunks = by_bytes;                          // latency-bound and every extra partial lengthens the final fold
    if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
    if (chunks > HW) chunks = HW;
    if (chunks < 1) chunks = 1;
    const int ppc = (HW + chunks - 1) / chunks;
    chunks = (HW + ppc - 1) / ppc;
    // thread layout: as many pixel lanes as fit beside the quad axis (never more than the chunk has pixels)
    int px_lanes = GN_THREADS / (nq < GN_THREADS ? nq : GN_THREADS);
    if (px_lanes > ppc) px_lanes = ppc;
    if (px_lanes < 1) px_lanes = 1;
    const int qlanes = GN_THREADS / px_lanes;
    GNStatsArgs a;
    a.src0 = src0; a.src1 = src1; a.C0 = C0; a.C1 = C1; a.HW = HW; a.chunks = chunks; a.pix_per_chunk = ppc;
    a.px_lanes = px_lanes; a.q_iters = (nq + qlanes - 1) / qlanes;
    a.ticket = reinterpret_cast<unsigned int*>(workspace);
    a.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + (((size_t)B * sizeof(unsigned int) + 255) / 256) * 256);
    a.stats = stats; a.eps = eps;
    const size_t smem = (size_t)nq * px_lanes * 8 * sizeof(float);
    SDK_CHECK_ARG(smem <= 48 * 1024, "sdk_groupnorm_stats: shared memory %zu too large", smem);
    SDK_CUDA(sdk_launch(gn_stats_kernel, dim3(dim3(chunks, B)), dim3(GN_THREADS), (size_t)(smem), (cudaStream_t)stream, a));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

namespace {
int launch_gn_apply(GNApplyArgs& a, int out_dtype, cudaStream_t stream) {
    const int C = a.C0 + a.C1, B = a.B, HW = a.HW;
    const int nq = C / 4;
    int px_lanes = GN_THREADS / (nq < GN_THREADS ? nq : GN_THREADS);
    if (px_lanes > HW) px_lanes = HW;
    if (px_lanes < 1) px_lanes = 1;
    const int qlanes = GN_THREADS / px_lanes;
    const int q_iters = (nq + qlanes - 1) / qlanes;
    // ~8 pixels per thread, but never fewer CTAs than ~2 per SM when the tensor is large enough
    static const int ppt = getenv("SDB200_GN_PPT") ? atoi(getenv("SDB200_GN_PPT")) : 16;   // pixels per thread (batch 16: 8 -> 2.08 ms, 16 -> 1.98 ms, 32 -> 2.01 ms of GroupNorm per step)
    int ppc = ppt * px_lanes;
    int chunks = (HW + ppc - 1) / ppc;
    static const int ctas_per_sm = getenv("SDB200_GN_CTAS") ? atoi(getenv("SDB200_GN_CTAS")) : 4;   // small tensors: CTAs per SM (UNet batch 2: 2 -> 0.57, 4 -> 0.51, 8 -> 0.64 ms of GroupNorm per step)
    const int want = (sdk_num_sms() * ctas_per_sm + B - 1) / B;
    if (chunks < want) { chunks = want < HW ? want : HW; ppc = (HW + chunks - 1) / chunks; }
    chunks = (HW + ppc - 1) / ppc;
    SDK_CHECK_ARG(B < 65536, "sdk_groupnorm_apply: batch too large");
    if (out_dtype == SDK_F32) SDK_CUDA(sdk_launch(gn_apply_kernel<float>, dim3(chunks, B), dim3(GN_THREADS), (size_t)0, stream, a, ppc, px_lanes, q_iters));
    else if (out_dtype == SDK_BF16) SDK_CUDA(sdk_launch(gn_apply_kernel<__nv_bfloat16>, dim3(chunks, B), dim3(GN_THREADS), (size_t)0, stream, a, ppc, px_lanes, q_iters));
    else return sdk_fail(SDK_ERR_ARG, "sdk_groupnorm_apply: out_dtype %d", out_dtype);
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}
}  // namespace

extern "C" int sdk_groupnorm_apply(const float* src0, int C0, const float* src1, int C1, int B, int HW,
                                   const float* stats, const float* gamma, const float* beta, int silu,
                                   void* out, void* raw_out, int out_dtype, void* stream) {
    SDK_CHECK_ARG(src0 && stats && gamma && beta && out, "sdk_groupnorm_apply: null pointer");
    const int C = C0 + C1;
    SDK_CHECK_ARG(C % GROUPS == 0 && C0 % 4 == 0 && C1 % 4 == 0 && (C1 == 0 || src1), "sdk_groupnorm_apply: bad channels %d+%d", C0, C1);
    GNApplyArgs a;
    a.src0 = src0; a.src1 = src1; a.C0 = C0; a.C1 = C1; a.HW = HW; a.B = B;
    a.stats = stats; a.gamma = gamma; a.beta = beta; a.out = out; a.raw_out = raw_out; a.silu = silu;
    a.cs0 = nullptr; a.cs1 = nullptr; a.eps = 0.f;
    return launch_gn_apply(a, out_dtype, (cudaStream_t)stream);
}

// GroupNorm apply whose statistics come from the per-channel sums of the sources (cs0 [B][C0][2], cs1 [B][C1][2], double) instead of a
// separate statistics pass over the tensor: the producing GEMM's epilogue already reduced its columns (sdk_tc_gemm_set_stats).
extern "C" int sdk_groupnorm_apply_cs(const float* src0, int C0, const double* cs0, const float* src1, int C1, const double* cs1,
                                      int B, int HW, float eps, const float* gamma, const float* beta, int silu,
                                      void* out, void* raw_out, int out_dtype, void* stream) {
    SDK_CHECK_ARG(src0 && cs0 && gamma && beta && out, "sdk_groupnorm_apply_cs: null pointer");
    const int C = C0 + C1;
    SDK_CHECK_ARG(C % GROUPS == 0 && C0 % 4 == 0 && C1 % 4 == 0 && (C1 == 0 || (src1 && cs1)), "sdk_groupnorm_apply_cs: bad channels %d+%d", C0, C1);
    GNApplyArgs a;
    a.src0 = src0; a.src1 = src1; a.C0 = C0; a.C1 = C1; a.HW = HW; a.B = B;
    a.stats = nullptr; a.gamma = gamma; a.beta = beta; a.out = out; a.raw_out = raw_out; a.silu = silu;
    a.cs0 = reinterpret_cast<const double2*>(cs0); a.cs1 = reinterpret_cast<const double2*>(cs1); a.eps = eps;
    return launch_gn_apply(a, out_dtype, (cudaStream_t)stream);
}

// Per-channel (sum, sum of squares) of an fp32 [B][HW][C] tensor, ACCUMULATED into out [B][C][2] (double; the caller zeroes it,
// as for sdk_tc_gemm_set_stats); for GroupNorm inputs whose producer cannot deliver them (conv_in's FFMA kernel, unusual tilings).
// CTA = (32 channels, up to 256 rows of one sample): shared-memory fold over the 32 row lanes, then 64 double atomics.
namespace {
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256)
channel_stats_kernel(const float* __restrict__ src, int HW, int C, double2* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_red[32][8][8];
    const int b = blockIdx.z, rl = threadIdx.x >> 3, q = threadIdx.x & 7, c = blockIdx.x * 32 + (q << 2);
    const int r0 = blockIdx.y * CS_ROWS, r1 = min(HW, r0 + CS_ROWS);
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const float* p = src + (size_t)b * HW * C + c;
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)r * C));
            sa[0] += v.x; sa[1] += v.y; sa[2] += v.z; sa[3] += v.w;
            qa[0] = fmaf(v.x, v.x, qa[0]); qa[1] = fmaf(v.y, v.y, qa[1]); qa[2] = fmaf(v.z, v.z, qa[2]); qa[3] = fmaf(v.w, v.w, qa[3]);
        }
    }
    *reinterpret_cast<float4*>(&s_red[rl][q][0]) = make_float4(sa[0], sa[1], sa[2], sa[3]);
    *reinterpret_cast<float4*>(&s_red[rl][q][4]) = make_float4(qa[0], qa[1], qa[2], qa[3]);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int col = threadIdx.x & 31, which = threadIdx.x >> 5;        // which: 0 = sum, 1 = sum of squares
        if (blockIdx.x * 32 + col < C) {
            const int qq = col >> 2, j = (col & 3) + 4 * which;
            double acc = 0.0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) acc += (double)s_red[l][qq][j];
            atomicAdd(reinterpret_cast<double*>(out + (size_t)b * C + blockIdx.x * 32 + col) + which, acc);
        }
    }
}
}  // namespace

extern "C" int sdk_channel_stats(const float* src, int B, int HW, int C, double* out, void* stream) {
    SDK_CHECK_ARG(src && out && B > 0 && B < 65536 && HW > 0 && C > 0 && C % 4 == 0, "sdk_channel_stats: bad args");
    SDK_CUDA(sdk_launch(channel_stats_kernel, dim3((C + 31) / 32, (HW + CS_ROWS - 1) / CS_ROWS, B), dim3(256), (size_t)0, (cudaStream_t)stream, src, HW, C,
                        reinterpret_cast<double2*>(out)));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}


// GroupNorm statistics + apply (+SiLU) (+raw cast copy) in one cooperative launch; workspace as for sdk_groupnorm_stats
extern "C" int sdk_groupnorm_fused(const float* src0, int C0, const float* src1, int C1, int B, int HW, float eps,
                                   const float* gamma, const float* beta, int silu, void* out, void* raw_out, int out_dtype,
                                   void* workspace, void* stream) {
    SDK_CHECK_ARG(src0 && gamma && beta && out && workspace, "sdk_groupnorm_fused: null pointer");
    const int C = C0 + C1;
    SDK_CHECK_ARG(C0 > 0 && C1 >= 0 && (C1 == 0 || src1), "sdk_groupnorm_fused: bad sources");
    SDK_CHECK_ARG(C % GROUPS == 0 && C0 % 4 == 0 && C1 % 4 == 0, "sdk_groupnorm_fused: C=%d+%d must be a multiple of 32 (quads of 4)", C0, C1);
    SDK_CHECK_ARG(B > 0 && HW > 0 && (out_dtype == SDK_F32 || out_dtype == SDK_BF16), "sdk_groupnorm_fused: bad args");
    const int nq = C / 4;
    int chunks = HW < 32 ? HW : 32;
    const int ppc = (HW + chunks - 1) / chunks;
    chunks = (HW + ppc - 1) / ppc;
    int px_lanes = GN_THREADS / (nq < GN_THREADS ? nq : GN_THREADS);
    if (px_lanes > ppc) px_lanes = ppc;
    if (px_lanes < 1) px_lanes = 1;
    const int qlanes = GN_THREADS / px_lanes;
    GNFusedArgs a;
    a.src0 = src0; a.src1 = src1; a.C0 = C0; a.C1 = C1; a.HW = HW; a.B = B; a.chunks = chunks; a.pix_per_chunk = ppc;
    a.px_lanes = px_lanes; a.q_iters = (nq + qlanes - 1) / qlanes;
    a.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + (((size_t)B * sizeof(unsigned int) + 255) / 256) * 256);
    a.gamma = gamma; a.beta = beta; a.out = out; a.raw_out = raw_out; a.eps = eps; a.silu = silu;
    size_t smem = (size_t)nq * px_lanes * 8 * sizeof(float);
    if (smem < 512 + 8 * GROUPS * 2 * sizeof(double)) smem = 512 + 8 * GROUPS * 2 * sizeof(double);   // phase 2: stats + fold scratch
    SDK_CHECK_ARG(smem <= 48 * 1024, "sdk_groupnorm_fused: shared memory %zu too large", smem);
    // grid: every CTA must be co-resident (cooperative launch): 2 CTAs of 256 threads per SM, no more than there is work
    long long want = (long long)B * HW * nq / (GN_THREADS * 4);
    if (want < (long long)B * chunks) want = (long long)B * chunks;
    const void* fn = out_dtype == SDK_F32 ? (const void*)gn_fused_kernel<float> : (const void*)gn_fused_kernel<__nv_bfloat16>;
    int occ = 0;
    SDK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, GN_THREADS, smem));
    SDK_CHECK_ARG(occ >= 1, "sdk_groupnorm_fused: kernel does not fit on an SM");
    if (occ > 4) occ = 4;
    int grid = sdk_num_sms() * occ;
    if (want < grid) grid = (int)(want < 1 ? 1 : want);
    void* kargs[] = {&a};
    SDK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(GN_THREADS), kargs, smem, (cudaStream_t)stream));
    return SDK_OK;
}

// GroupNorm (+SiLU, +raw copy) in one launch, one thread-block cluster per sample (see gn_cluster_kernel)
extern "C" int sdk_groupnorm_cluster(const float* src0, int C0, const float* src1, int C1, int B, int HW, float eps,
                                     const float* gamma, const float* beta, int silu, void* out, void* raw_out, int out_dtype,
                                     void* stream) {
    SDK_CHECK_ARG(src0 && gamma && beta && out, "sdk_groupnorm_cluster: null pointer");
    const int C = C0 + C1;
    SDK_CHECK_ARG(C0 > 0 && C1 >= 0 && (C1 == 0 || src1), "sdk_groupnorm_cluster: bad sources");
    SDK_CHECK_ARG(C % GROUPS == 0 && C0 % 4 == 0 && C1 % 4 == 0, "sdk_groupnorm_cluster: C=%d+%d must be a multiple of 32 (quads of 4)", C0, C1);
    SDK_CHECK_ARG(B > 0 && B < 65536 && HW > 0 && (out_dtype == SDK_F32 || out_dtype == SDK_BF16), "sdk_groupnorm_cluster: bad args");
    const int nq = C / 4;
    SDK_CHECK_ARG(nq <= 4096, "sdk_groupnorm_cluster: C too large");
    int CS = 8;
    while (CS > 1 && CS > HW) CS >>= 1;
    const int ppc = (HW + CS - 1) / CS;
    int px_lanes = GNC_THREADS / (nq < GNC_THREADS ? nq : GNC_THREADS);
    if (px_lanes > ppc) px_lanes = ppc;
    if (px_lanes < 1) px_lanes = 1;
    const int qlanes = GNC_THREADS / px_lanes;
    const int q_iters = (nq + qlanes - 1) / qlanes;
    const size_t smem = (size_t)nq * px_lanes * 8 * sizeof(float);
    SDK_CHECK_ARG(smem <= 160 * 1024, "sdk_groupnorm_cluster: shared memory %zu too large", smem);
    GNApplyArgs a;
    a.src0 = src0; a.src1 = src1; a.C0 = C0; a.C1 = C1; a.HW = HW; a.B = B;
    a.stats = nullptr; a.gamma = gamma; a.beta = beta; a.out = out; a.raw_out = raw_out; a.silu = silu;
    void (*fn)(GNApplyArgs, float, int, int, int) = out_dtype == SDK_F32 ? gn_cluster_kernel<float> : gn_cluster_kernel<__nv_bfloat16>;
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(fn), 160 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, B); cfg.blockDim = dim3(GNC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SDK_CUDA(cudaLaunchKernelEx(&cfg, fn, a, eps, CS, px_lanes, q_iters));
    return SDK_OK;
}

extern "C" int sdk_layernorm(const float* x, const float* gamma, const float* beta, float eps, void* out,
                             int out_dtype, int64_t rows, int C, void* stream) {
    SDK_CHECK_ARG(x && gamma && beta && out, "sdk_layernorm: null pointer");
    SDK_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 2048, "sdk_layernorm: C=%d must be a multiple of 4 and <= 2048", C);
    if (rows <= 0) return SDK_OK;
    const int grid = grid_for(rows * 32, 256);
    cudaStream_t s = (cudaStream_t)stream;
#define LN_LAUNCH(T, MAXQ) SDK_CUDA(sdk_launch(layernorm_kernel<T, MAXQ>, dim3(grid), dim3(256), 0, s, x, gamma, beta, eps, reinterpret_cast<T*>(out), C, rows))
    const int nq = (C / 4 + 31) / 32;
    if (out_dtype == SDK_F32) {
        if (nq <= 3) LN_LAUNCH(float, 3); else if (nq <= 5) LN_LAUNCH(float, 5); else if (nq <= 10) LN_LAUNCH(float, 10); else LN_LAUNCH(float, 16);
    } else if (out_dtype == SDK_BF16) {
        if (nq <= 3) LN_LAUNCH(__nv_bfloat16, 3); else if (nq <= 5) LN_LAUNCH(__nv_bfloat16, 5); else if (nq <= 10) LN_LAUNCH(__nv_bfloat16, 10); else LN_LAUNCH(__nv_bfloat16, 16);
    } else return sdk_fail(SDK_ERR_ARG, "sdk_layernorm: out_dtype %d", out_dtype);
#undef LN_LAUNCH
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

// ---------------------------------------------------------------------------------------------
// Row softmax of a materialised score matrix: out[r][:] = softmax(scale * in[r][:]) -- the single-head, head_dim = 512
// attention of the VAE (models/vae/vae.py:55-80), whose scores and P.V products run as tensor-core GEMMs.
// One CTA per row, the row lives in registers (cols <= 256 * 4 * SM_MAXQ), exact max and sum (fp32), exp2 of the pre-scaled logits.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int SM_MAXQ = 16;                             // float4 per thread -> rows of up to 16384 columns
template <typename TOut>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ in, TOut* __restrict__ out, int cols, float scale_log2) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_red[8];
    const long long row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(in + row * cols);
    const int nq = cols >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 v[SM_MAXQ];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < SM_MAXQ; ++k) {
        const int q = threadIdx.x + k * 256;
        if (q < nq) { v[k] = __ldg(src + q); mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w))); }
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
    __syncthreads();
    const float off = mx * scale_log2;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < SM_MAXQ; ++k) {
        const int q = threadIdx.x + k * 256;
        if (q < nq) {
            v[k].x = exp2f(fmaf(v[k].x, scale_log2, -off)); v[k].y = exp2f(fmaf(v[k].y, scale_log2, -off));
            v[k].z = exp2f(fmaf(v[k].z, scale_log2, -off)); v[k].w = exp2f(fmaf(v[k].w, scale_log2, -off));
            sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += s_red[i];
    const float inv = 1.f / sum;
    TOut* dst = out + row * cols;
#pragma unroll
    for (int k = 0; k < SM_MAXQ; ++k) {
        const int q = threadIdx.x + k * 256;
        if (q < nq) store4<TOut>(dst + (q << 2), v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
    }
}
}  // namespace

extern "C" int sdk_softmax_rows(const float* in, void* out, int out_dtype, int64_t rows, int cols, float scale, void* stream) {
    SDK_CHECK_ARG(in && out && rows >= 0 && rows < (1ll << 31), "sdk_softmax_rows: bad args");
    SDK_CHECK_ARG(cols > 0 && cols % 4 == 0 && cols <= 256 * 4 * SM_MAXQ, "sdk_softmax_rows: cols=%d must be a multiple of 4, at most %d", cols, 256 * 4 * SM_MAXQ);
    if (rows == 0) return SDK_OK;
    const float sl2 = scale * 1.4426950408889634f;
    if (out_dtype == SDK_F32) SDK_CUDA(sdk_launch(softmax_rows_kernel<float>, dim3((unsigned)rows), dim3(256), (size_t)0, (cudaStream_t)stream, in, (float*)out, cols, sl2));
    else if (out_dtype == SDK_BF16) SDK_CUDA(sdk_launch(softmax_rows_kernel<__nv_bfloat16>, dim3((unsigned)rows), dim3(256), (size_t)0, (cudaStream_t)stream, in, (__nv_bfloat16*)out, cols, sl2));
    else return sdk_fail(SDK_ERR_ARG, "sdk_softmax_rows: out_dtype %d", out_dtype);
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

// ---------------------------------------------------------------------------------------------
// Text-encoder front end and activation (models/clip/openclip.py:53-71,73-84; clip.py:37-57; activation_fn.py:4-9)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos, float* __restrict__ out,
                    long long rows, int S, int C, int vocab) {
    pdl_trigger();
    pdl_wait();
    const int nq = C >> 2;
    const long long total = rows * nq;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / nq;
        const int q = (int)(i - r * nq);
        long long id = ids[r];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);                 // nn.Embedding would raise; never read out of bounds
        const float4 a = __ldg(reinterpret_cast<const float4*>(tok + (size_t)id * C) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(r % S) * C) + q);
        reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

// kind 1: exact-erf GELU (nn.GELU(), openclip.py:78); kind 2: QuickGELU x * sigmoid(1.702 x) (activation_fn.py:8)
template <typename TOut>
__global__ void __launch_bounds__(256)
activation_kernel(const float* __restrict__ in, TOut* __restrict__ out, long long n4, int kind) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
        float y[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = kind == 1 ? gelu_erf_f(y[j]) : y[j] / (1.0f + expf(-1.702f * y[j]));
        store4<TOut>(out + (i << 2), y[0], y[1], y[2], y[3]);
    }
}
}  // namespace

extern "C" int sdk_embed_tokens(const int64_t* ids, const float* tok_emb, const float* pos_emb, float* out,
                                int64_t rows, int S, int C, int vocab, void* stream) {
    SDK_CHECK_ARG(ids && tok_emb && pos_emb && out && rows >= 0 && S > 0 && C > 0 && C % 4 == 0 && vocab > 0, "sdk_embed_tokens: bad args");
    if (rows == 0) return SDK_OK;
    SDK_CUDA(sdk_launch(embed_tokens_kernel, dim3(grid_for(rows * (C / 4), 256)), dim3(256), (size_t)0, (cudaStream_t)stream,
                        reinterpret_cast<const long long*>(ids), tok_emb, pos_emb, out, (long long)rows, S, C, vocab));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

extern "C" int sdk_activation(const float* in, void* out, int out_dtype, int64_t n, int kind, void* stream) {
    SDK_CHECK_ARG(in && out && n >= 0 && n % 4 == 0 && (kind == 1 || kind == 2), "sdk_activation: bad args (n %% 4 == 0, kind 1 = GELU, 2 = QuickGELU)");
    if (n == 0) return SDK_OK;
    const int grid = grid_for(n / 4, 256);
    if (out_dtype == SDK_F32) SDK_CUDA(sdk_launch(activation_kernel<float>, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)stream, in, (float*)out, (long long)(n / 4), kind));
    else if (out_dtype == SDK_BF16) SDK_CUDA(sdk_launch(activation_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)stream, in, (__nv_bfloat16*)out, (long long)(n / 4), kind));
    else return sdk_fail(SDK_ERR_ARG, "sdk_activation: out_dtype %d", out_dtype);
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

extern "C" int sdk_cast_upsample(const float* src, void* dst, int out_dtype, int B, int H, int W, int C, int up, void* stream) {
    SDK_CHECK_ARG(src && dst && (up == 1 || up == 2) && C % 4 == 0, "sdk_cast_upsample: bad args (C=%d up=%d)", C, up);
    const long long total = (long long)B * H * up * W * up * (C / 4);
    if (total <= 0) return SDK_OK;
    const int grid = grid_for(total, 256);
    if (out_dtype == SDK_F32) SDK_CUDA(sdk_launch(cast_upsample_kernel<float>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, src, (float*)dst, B, H, W, C, up));
    else if (out_dtype == SDK_BF16) SDK_CUDA(sdk_launch(cast_upsample_kernel<__nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, src, (__nv_bfloat16*)dst, B, H, W, C, up));
    else return sdk_fail(SDK_ERR_ARG, "sdk_cast_upsample: out_dtype %d", out_dtype);
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

extern "C" int sdk_nchw_to_nhwc(const float* src, float* dst, int B_src, int B_dst, int C, int HW, void* stream) {
    SDK_CHECK_ARG(src && dst && B_src > 0 && B_dst > 0 && C > 0 && HW > 0, "sdk_nchw_to_nhwc: bad args");
    SDK_CUDA(sdk_launch(nchw_to_nhwc_kernel, dim3(grid_for((long long)B_dst * HW * C, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, src, dst, B_src, B_dst, C, HW));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}
