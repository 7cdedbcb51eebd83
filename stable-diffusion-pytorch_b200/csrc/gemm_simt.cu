// Exact-fp32 implicit-GEMM convolution / linear (FFMA, fp32 accumulate) — the "fp32 mode" of the
// hot path (final-latent rel-L2 <= 1e-4 against the reference) and the on-device cross-check for
// the tcgen05 path.  Also takes the shapes the tensor-core path does not (conv_in, Cin = 4).
//
// Replaces nn.Conv2d / nn.Linear call sites of models/unet/unet.py:67,71,158,161,168,236,246,256,401,
// models/unet/attention.py:19-25 and models/activation_fn.py:14, including the torch.cat of the
// decoder skip (unet.py:343, two A sources) and F.interpolate(nearest) (unet.py:250, folded into
// the gather).
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, A/B staged transposed in
// shared memory (conflict-free float4 fragment reads), register double-buffering of the next tile.
#define SDK_PDL_CAT 1
#include "common.cuh"
#include <stdlib.h>
#include "../../include/sdb200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PADM = 4;

struct RowInfo { int b, oy, ox; bool valid; };

__device__ __forceinline__ float4 load_a_quad(const SdkConvParams& p, const RowInfo& r, int k, int Cin, int K) {
    // 4 consecutive k (same tap, 4 consecutive channels) when Cin % 4 == 0; else element-wise
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!r.valid || k >= K) return v;
    const int up = p.upsample ? 2 : 1;
    const int pad = p.ksize >> 1;
    if ((Cin & 3) == 0 && (p.C0 & 3) == 0) {
        const int tap = k / Cin, c = k - tap * Cin;
        const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
        const int iy = r.oy * p.stride + ky - pad, ix = r.ox * p.stride + kx - pad;
        if (iy < 0 || ix < 0 || iy >= p.Hin * up || ix >= p.Win * up) return v;
        const size_t pix = ((size_t)r.b * p.Hin + iy / up) * p.Win + ix / up;
        const float* src = (c < p.C0) ? (const float*)p.src0 + pix * p.C0 + c
                                      : (const float*)p.src1 + pix * p.C1 + (c - p.C0);
        return __ldg(reinterpret_cast<const float4*>(src));
    }
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = 0.f;
        const int kk = k + j;
        if (kk >= K) continue;
        const int tap = kk / Cin, c = kk - tap * Cin;
        const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
        const int iy = r.oy * p.stride + ky - pad, ix = r.ox * p.stride + kx - pad;
        if (iy < 0 || ix < 0 || iy >= p.Hin * up || ix >= p.Win * up) continue;
        const size_t pix = ((size_t)r.b * p.Hin + iy / up) * p.Win + ix / up;
        e[j] = (c < p.C0) ? __ldg((const float*)p.src0 + pix * p.C0 + c) : __ldg((const float*)p.src1 + pix * p.C1 + (c - p.C0));
    }
    return make_float4(e[0], e[1], e[2], e[3]);
}

__device__ __forceinline__ float4 load_b_quad(const float* __restrict__ w, int n, int N, int k, int K) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= N || k >= K) return v;
    const float* src = w + (size_t)n * K + k;
    if ((K & 3) == 0) return __ldg(reinterpret_cast<const float4*>(src));
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = (k + j < K) ? __ldg(src + j) : 0.f;
    return make_float4(e[0], e[1], e[2], e[3]);
}

__global__ void __launch_bounds__(THREADS)
conv_gemm_f32_kernel(const SdkConvParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][BK][BM + PADM];
    __shared__ __align__(16) float Bs[2][BK][BN + PADM];

    const int Cin = p.C0 + p.C1;
    const int K = p.ksize * p.ksize * Cin;
    const int M = p.B * p.Hout * p.Wout;
    const int N = p.N;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int lrow = tid >> 2, kq = (tid & 3) << 2;      // loader: rows lrow, lrow+64 ; k offset kq

    RowInfo ri[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = m0 + lrow + h * 64;
        ri[h].valid = m < M;
        const int mm = ri[h].valid ? m : 0;
        ri[h].ox = mm % p.Wout;
        const int t = mm / p.Wout;
        ri[h].oy = t % p.Hout;
        ri[h].b = t / p.Hout;
    }
    const float* w = (const float*)p.weight;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int ty = tid >> 4, tx = tid & 15;
    float4 ra[2], rb[2];
    const int ktiles = (K + BK - 1) / BK;

    ra[0] = load_a_quad(p, ri[0], kq, Cin, K); ra[1] = load_a_quad(p, ri[1], kq, Cin, K);
    rb[0] = load_b_quad(w, n0 + lrow, N, kq, K); rb[1] = load_b_quad(w, n0 + lrow + 64, N, kq, K);

    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + h * 64;
            As[buf][kq + 0][r] = ra[h].x; As[buf][kq + 1][r] = ra[h].y; As[buf][kq + 2][r] = ra[h].z; As[buf][kq + 3][r] = ra[h].w;
            Bs[buf][kq + 0][r] = rb[h].x; Bs[buf][kq + 1][r] = rb[h].y; Bs[buf][kq + 2][r] = rb[h].z; Bs[buf][kq + 3][r] = rb[h].w;
        }
        __syncthreads();
        if (kt + 1 < ktiles) {
            const int k = (kt + 1) * BK + kq;
            ra[0] = load_a_quad(p, ri[0], k, Cin, K); ra[1] = load_a_quad(p, ri[1], k, Cin, K);
            rb[0] = load_b_quad(w, n0 + lrow, N, k, K); rb[1] = load_b_quad(w, n0 + lrow + 64, N, k, K);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        // the next iteration writes the other buffer; one barrier per k-tile suffices because a
        // thread can only be one tile ahead (it needs the barrier above to proceed)
    }

    // ---- epilogue
    const int Nout = p.geglu ? (N >> 1) : N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        const int ox = m % p.Wout; const int t = m / p.Wout; const int oy = t % p.Hout; const int b = t / p.Hout;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int nb = n0 + jh * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = nb + j;
                float x = acc[i][jh * 4 + j];
                if (n < N) {
                    if (p.bias) x += __ldg(p.bias + n);
                    if (p.tbias) x += __ldg(p.tbias + (size_t)b * p.tb_stride + n);
                }
                v[j] = x;
            }
            if (p.geglu) {
                // columns (2j, 2j+1) = (value, gate)
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const int no = (nb + j) >> 1;
                    if (nb + j + 1 < N) {
                        float y = v[j] * gelu_erf_f(v[j + 1]);
                        if (p.residual) y += __ldg(p.residual + (size_t)m * Nout + no);
                        if (p.out_dtype == SDK_BF16) ((__nv_bfloat16*)p.out)[(size_t)m * Nout + no] = __float2bfloat16_rn(y);
                        else ((float*)p.out)[(size_t)m * Nout + no] = y;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nb + j;
                    if (n >= N) continue;
                    float y = v[j];
                    if (p.residual) y += __ldg(p.residual + (size_t)m * N + n);
                    const size_t o = p.out_nchw ? (((size_t)b * N + n) * p.Hout + oy) * p.Wout + ox : (size_t)m * N + n;
                    if (p.out_dtype == SDK_BF16) ((__nv_bfloat16*)p.out)[o] = __float2bfloat16_rn(y);
                    else ((float*)p.out)[o] = y;
                }
            }
        }
    }
}

}  // namespace

int sdk_validate_conv(const SdkConvParams* p, const char* who) {
    SDK_CHECK_ARG(p && p->src0 && p->weight && p->out, "%s: null pointer", who);
    SDK_CHECK_ARG(p->C0 > 0 && p->C1 >= 0 && (p->C1 == 0 || p->src1), "%s: bad sources C0=%d C1=%d", who, p->C0, p->C1);
    SDK_CHECK_ARG(p->ksize == 1 || p->ksize == 3, "%s: ksize %d", who, p->ksize);
    SDK_CHECK_ARG(p->stride == 1 || p->stride == 2, "%s: stride %d", who, p->stride);
    SDK_CHECK_ARG(p->B > 0 && p->Hin > 0 && p->Win > 0 && p->Hout > 0 && p->Wout > 0 && p->N > 0, "%s: bad sizes", who);
    const int up = p->upsample ? 2 : 1, pad = p->ksize / 2;
    const int he = (p->Hin * up + 2 * pad - p->ksize) / p->stride + 1, we = (p->Win * up + 2 * pad - p->ksize) / p->stride + 1;
    SDK_CHECK_ARG(he == p->Hout && we == p->Wout, "%s: output %dx%d inconsistent with input %dx%d (k=%d s=%d up=%d)", who,
                  p->Hout, p->Wout, p->Hin, p->Win, p->ksize, p->stride, up);
    SDK_CHECK_ARG(!p->geglu || (p->N % 2 == 0 && !p->out_nchw), "%s: geglu needs even N and NHWC output", who);
    SDK_CHECK_ARG((long long)p->B * p->Hout * p->Wout < (1ll << 31), "%s: too many rows", who);
    return SDK_OK;
}

extern "C" int sdk_conv_gemm_f32(const SdkConvParams* p, void* stream) {
    int rc = sdk_validate_conv(p, "sdk_conv_gemm_f32");
    if (rc) return rc;
    SDK_CHECK_ARG(p->in_dtype == SDK_F32, "sdk_conv_gemm_f32: in_dtype must be fp32");
    SDK_CHECK_ARG(p->out_dtype == SDK_F32 || p->out_dtype == SDK_BF16, "sdk_conv_gemm_f32: out_dtype %d", p->out_dtype);
    const int M = p->B * p->Hout * p->Wout;
    dim3 grid((M + BM - 1) / BM, (p->N + BN - 1) / BN);
    SDK_CHECK_ARG(grid.y < 65536, "sdk_conv_gemm_f32: N too large");
    SDK_CUDA(sdk_launch(conv_gemm_f32_kernel, dim3(grid), dim3(THREADS), (size_t)(0), (cudaStream_t)stream, *p));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}


// -------------------------------------------------------------------------------------------------
// conv_in (unet.py:256): 3x3, pad 1, Cin = 4 -> N channels on the fp32 NHWC latent.  K = 36 is far too short for a GEMM
// pipeline: the whole weight matrix (36 x N floats) sits in shared memory, a thread owns 4 output channels and walks the
// CTA's pixels with the 9 taps (one float4 each) in registers.  Also accumulates the per-channel (sum, sum of squares)
// table of the first GroupNorm (double atomics, as the tensor-core epilogue does).
// -------------------------------------------------------------------------------------------------
namespace {
constexpr int CI_PIX = 16, CI_THREADS = 256;      // small pixel chunks: ~4 CTAs per SM hide the FFMA dependency chains

__global__ void __launch_bounds__(CI_THREADS)
conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /* [3][3][4][N] */, const float* __restrict__ bias, float* __restrict__ out,
               double2* __restrict__ cstat, int B, int H, int W, int N) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    float* w_s = sm;                                   // [36][N]
    float* s_red = sm + 36 * N;                        // [px_lanes][N][2] column partials
    const int quads = N >> 2, px_lanes = CI_THREADS / quads;
    const int pl = threadIdx.x / quads, qd = threadIdx.x - pl * quads;
    for (int i = threadIdx.x; i < 9 * N; i += CI_THREADS)          // w_t is already [36][N]: straight 16-byte copies
        reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    __syncthreads();
    const int HW = H * W;
    const long long p0 = (long long)blockIdx.x * CI_PIX;      // CTA pixel range inside sample blockIdx.y
    const int b = blockIdx.y;
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
    if (pl < px_lanes) {
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + qd) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int pi = pl; pi < CI_PIX; pi += px_lanes) {
            const long long pix = p0 + pi;
            if (pix >= HW) break;
            const int oy = (int)(pix / W), ox = (int)(pix - (long long)oy * W);
            float4 acc = bv;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int iy = oy + t / 3 - 1, ix = ox + t % 3 - 1;
                if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * 4));
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 ww = *reinterpret_cast<const float4*>(w_s + (t * 4 + c) * N + (qd << 2));
                    acc.x = fmaf(vv[c], ww.x, acc.x); acc.y = fmaf(vv[c], ww.y, acc.y);
                    acc.z = fmaf(vv[c], ww.z, acc.z); acc.w = fmaf(vv[c], ww.w, acc.w);
                }
            }
            *reinterpret_cast<float4*>(out + ((size_t)b * HW + pix) * N + (qd << 2)) = acc;
            sa[0] += acc.x; sa[1] += acc.y; sa[2] += acc.z; sa[3] += acc.w;
            qa[0] = fmaf(acc.x, acc.x, qa[0]); qa[1] = fmaf(acc.y, acc.y, qa[1]); qa[2] = fmaf(acc.z, acc.z, qa[2]); qa[3] = fmaf(acc.w, acc.w, qa[3]);
        }
    }
    if (!cstat) return;
    if (pl < px_lanes) {
        float* d = s_red + ((size_t)pl * N + (qd << 2)) * 2;
        *reinterpret_cast<float4*>(d) = make_float4(sa[0], qa[0], sa[1], qa[1]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(sa[2], qa[2], sa[3], qa[3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * N; i += CI_THREADS) {         // i = channel * 2 + which
        double acc = 0.0;
        for (int l = 0; l < px_lanes; ++l) acc += (double)s_red[(size_t)l * N * 2 + i];
        atomicAdd(reinterpret_cast<double*>(cstat + (size_t)b * N) + i, acc);
    }
}

// Register-blocked form for W % 4 == 0: a thread owns FOUR consecutive pixels of a row for its channel quad, so every weight float4
// read from shared memory feeds 16 FMAs instead of 4 -- the per-pixel kernel above is bound by shared-memory bandwidth (one LDS.128
// per 4 FFMA = a quarter of the FP32 rate).  Same (tap, input channel) summation order per output value, so the results are
// bit-identical to the per-pixel kernel (out-of-image taps add 0 * w instead of being skipped).
constexpr int CI4_PIX = 32;                        // pixels per CTA (8 groups of 4)
__global__ void __launch_bounds__(CI_THREADS)
conv_in4_kernel(const float* __restrict__ x, const float* __restrict__ w /* [3][3][4][N] */, const float* __restrict__ bias, float* __restrict__ out,
                double2* __restrict__ cstat, int B, int H, int W, int N) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    float* w_s = sm;                                   // [36][N]
    float* s_red = sm + 36 * N;                        // [px_lanes][N][2] column partials
    const int quads = N >> 2, px_lanes = CI_THREADS / quads;
    const int pl = threadIdx.x / quads, qd = threadIdx.x - pl * quads;
    for (int i = threadIdx.x; i < 9 * N; i += CI_THREADS)
        reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    __syncthreads();
    const int HW = H * W;
    const int p0 = blockIdx.x * CI4_PIX;
    const int b = blockIdx.y;
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
    if (pl < px_lanes) {
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + qd) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int pg = pl; pg < CI4_PIX / 4; pg += px_lanes) {
            const int pix = p0 + 4 * pg;
            if (pix >= HW) break;
            const int oy = pix / W, ox = pix - oy * W;             // W % 4 == 0: the four pixels share the row
            float4 acc[4] = {bv, bv, bv, bv};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int iy = oy + dy - 1;
                float4 v[6];                                       // input pixels ox-1 .. ox+4 of row iy (zeros outside the image)
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int ix = ox + k - 1;
                    v[k] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                               ? __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int t = dy * 3 + kx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 ww = *reinterpret_cast<const float4*>(w_s + (t * 4 + c) * N + (qd << 2));
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp) {
                            const float4 vv = v[pp + kx];
                            const float xv = c == 0 ? vv.x : c == 1 ? vv.y : c == 2 ? vv.z : vv.w;
                            acc[pp].x = fmaf(xv, ww.x, acc[pp].x); acc[pp].y = fmaf(xv, ww.y, acc[pp].y);
                            acc[pp].z = fmaf(xv, ww.z, acc[pp].z); acc[pp].w = fmaf(xv, ww.w, acc[pp].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                *reinterpret_cast<float4*>(out + ((size_t)b * HW + pix + pp) * N + (qd << 2)) = acc[pp];
                sa[0] += acc[pp].x; sa[1] += acc[pp].y; sa[2] += acc[pp].z; sa[3] += acc[pp].w;
                qa[0] = fmaf(acc[pp].x, acc[pp].x, qa[0]); qa[1] = fmaf(acc[pp].y, acc[pp].y, qa[1]);
                qa[2] = fmaf(acc[pp].z, acc[pp].z, qa[2]); qa[3] = fmaf(acc[pp].w, acc[pp].w, qa[3]);
            }
        }
    }
    if (!cstat) return;
    if (pl < px_lanes) {
        float* d = s_red + ((size_t)pl * N + (qd << 2)) * 2;
        *reinterpret_cast<float4*>(d) = make_float4(sa[0], qa[0], sa[1], qa[1]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(sa[2], qa[2], sa[3], qa[3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * N; i += CI_THREADS) {         // i = channel * 2 + which
        double acc = 0.0;
        for (int l = 0; l < px_lanes; ++l) acc += (double)s_red[(size_t)l * N * 2 + i];
        atomicAdd(reinterpret_cast<double*>(cstat + (size_t)b * N) + i, acc);
    }
}
}  // namespace

extern "C" int sdk_conv_in(const float* x, const float* w_t, const float* bias, float* out, double* chan_stats,
                           int B, int H, int W, int N, void* stream) {
    SDK_CHECK_ARG(x && w_t && out && B > 0 && B < 65536 && H > 0 && W > 0, "sdk_conv_in: bad args");
    SDK_CHECK_ARG((((uintptr_t)x | (uintptr_t)w_t | (uintptr_t)out) & 15) == 0, "sdk_conv_in: pointers must be 16-byte aligned");
    SDK_CHECK_ARG(N % 4 == 0 && N >= 4 && N / 4 <= CI_THREADS, "sdk_conv_in: N=%d must be a multiple of 4, at most %d", N, 4 * CI_THREADS);
    const int px_lanes = CI_THREADS / (N / 4);
    const size_t smem = sizeof(float) * ((size_t)36 * N + (size_t)px_lanes * N * 2);
    SDK_CHECK_ARG(smem <= 200 * 1024, "sdk_conv_in: N=%d needs too much shared memory", N);
    static const int blocked = getenv("SDB200_CONV_IN4") ? atoi(getenv("SDB200_CONV_IN4")) : 1;
    if (blocked && W % 4 == 0 && (long long)H * W < (1ll << 30)) {
        SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(conv_in4_kernel), (int)smem));
        const int chunks = (H * W + CI4_PIX - 1) / CI4_PIX;
        SDK_CUDA(sdk_launch(conv_in4_kernel, dim3(chunks, B), dim3(CI_THREADS), smem, (cudaStream_t)stream, x, w_t, bias, out,
                            reinterpret_cast<double2*>(chan_stats), B, H, W, N));
        SDK_LAUNCH_CHECK();
        return SDK_OK;
    }
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(conv_in_kernel), (int)smem));
    const int chunks = (H * W + CI_PIX - 1) / CI_PIX;
    SDK_CUDA(sdk_launch(conv_in_kernel, dim3(chunks, B), dim3(CI_THREADS), smem, (cudaStream_t)stream, x, w_t, bias, out,
                        reinterpret_cast<double2*>(chan_stats), B, H, W, N));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}
