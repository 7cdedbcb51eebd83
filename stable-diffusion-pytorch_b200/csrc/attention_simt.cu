// Exact-fp32 flash-style attention (online softmax, fp32 everywhere) for the fp32 parity mode.
// out = softmax(q k^T * scale) v per (batch, head)   (models/unet/attention.py:29-50).
//
// CTA = 64 queries of one (batch, head); K/V streamed in 32-key tiles through shared memory; four
// threads cooperate on one query row (each owns 8 of the 32 scores and a quarter of the D outputs)
// and exchange running max / sum with warp shuffles.  No score matrix is ever written to memory.
#define SDK_PDL_CAT 2
#include "common.cuh"

namespace {

constexpr int BQ = 64, BKV = 32, THREADS = 256;

template <int D>
__global__ void __launch_bounds__(THREADS)
attention_f32_kernel(const float* __restrict__ q, long long q_row, long long q_batch,
                     const float* __restrict__ k, long long k_row, long long k_batch,
                     const float* __restrict__ v, long long v_row, long long v_batch,
                     float* __restrict__ out, long long o_row, long long o_batch,
                     int heads, int Sq, int Sk, float scale, int causal) {
    pdl_trigger();
    pdl_wait();
    constexpr int DQ = D / 4;                    // output columns per thread
    extern __shared__ float smem[];
    float (*Qs)[D + 1] = reinterpret_cast<float (*)[D + 1]>(smem);
    float (*Ks)[D + 1] = reinterpret_cast<float (*)[D + 1]>(smem + BQ * (D + 1));
    float (*Vs)[D] = reinterpret_cast<float (*)[D]>(smem + (BQ + BKV) * (D + 1));
    float (*Ps)[BKV + 1] = reinterpret_cast<float (*)[BKV + 1]>(smem + (BQ + BKV) * (D + 1) + BKV * D);

    const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
    const int q0 = blockIdx.x * BQ;
    const int tid = threadIdx.x;
    const int qi = tid >> 2, part = tid & 3;

    const float* qb = q + (size_t)b * q_batch + (size_t)h * D;
    const float* kb = k + (size_t)b * k_batch + (size_t)h * D;
    const float* vb = v + (size_t)b * v_batch + (size_t)h * D;

    for (int i = tid; i < BQ * D; i += THREADS) {
        const int r = i / D, c = i - r * D;
        Qs[r][c] = (q0 + r < Sq) ? __ldg(qb + (size_t)(q0 + r) * q_row + c) * scale : 0.f;
    }
    float o[DQ];
#pragma unroll
    for (int d = 0; d < DQ; ++d) o[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int k0 = 0; k0 < Sk; k0 += BKV) {
        __syncthreads();                          // previous tile fully consumed (and Qs visible)
        for (int i = tid; i < BKV * D; i += THREADS) {
            const int r = i / D, c = i - r * D;
            const bool ok = k0 + r < Sk;
            Ks[r][c] = ok ? __ldg(kb + (size_t)(k0 + r) * k_row + c) : 0.f;
            Vs[r][c] = ok ? __ldg(vb + (size_t)(k0 + r) * v_row + c) : 0.f;
        }
        __syncthreads();
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
        for (int d = 0; d < D; ++d) {
            const float qv = Qs[qi][d];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = fmaf(qv, Ks[part + 4 * j][d], s[j]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k0 + part + 4 * j >= Sk || (causal && k0 + part + 4 * j > q0 + qi)) s[j] = -INFINITY;   // causal: keys after the query (clip/attention.py:44)
            mx = fmaxf(mx, s[j]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);      // finite: every tile has at least one valid key
        const float corr = expf(m_run - m_new);
        float ps = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pj = expf(s[j] - m_new);
            ps += pj;
            Ps[qi][part + 4 * j] = pj;
        }
        ps += __shfl_xor_sync(0xffffffffu, ps, 1);
        ps += __shfl_xor_sync(0xffffffffu, ps, 2);
        l_run = l_run * corr + ps;
        m_run = m_new;
        __syncwarp();                             // the 4 threads of a row are in one warp
#pragma unroll
        for (int d = 0; d < DQ; ++d) o[d] *= corr;
        for (int j = 0; j < BKV; ++j) {
            const float pj = Ps[qi][j];
#pragma unroll
            for (int d = 0; d < DQ; ++d) o[d] = fmaf(pj, Vs[j][part * DQ + d], o[d]);
        }
    }
    if (q0 + qi < Sq) {
        const float inv = 1.f / l_run;
        float* op = out + (size_t)b * o_batch + (size_t)(q0 + qi) * o_row + (size_t)h * D + part * DQ;
#pragma unroll
        for (int d = 0; d < DQ; ++d) op[d] = o[d] * inv;
    }
}

template <int D>
int launch(const float* q, long long q_row, long long q_batch, const float* k, long long k_row, long long k_batch,
           const float* v, long long v_row, long long v_batch, float* out, long long o_row, long long o_batch,
           int B, int heads, int Sq, int Sk, float scale, int causal, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)(BQ + BKV) * (D + 1) + (size_t)BKV * D + (size_t)BQ * (BKV + 1));
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(attention_f32_kernel<D>), (int)smem));
    dim3 grid((Sq + BQ - 1) / BQ, B * heads);
    SDK_CUDA(sdk_launch(attention_f32_kernel<D>, dim3(grid), dim3(THREADS), (size_t)(smem), s, q, q_row, q_batch, k, k_row, k_batch, v, v_row, v_batch,
                                                        out, o_row, o_batch, heads, Sq, Sk, scale, causal));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

}  // namespace

extern "C" int sdk_attention_f32_ex(const float* q, int64_t q_row, int64_t q_batch, const float* k, int64_t k_row, int64_t k_batch,
                                 const float* v, int64_t v_row, int64_t v_batch, float* out, int64_t o_row, int64_t o_batch,
                                 int B, int heads, int Sq, int Sk, int D, float scale, int causal, void* stream) {
    SDK_CHECK_ARG(q && k && v && out, "sdk_attention_f32_ex: null pointer");
    SDK_CHECK_ARG(B > 0 && heads > 0 && Sq > 0 && Sk > 0 && B * heads < 65536, "sdk_attention_f32_ex: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
#define ATT(DD) return launch<DD>(q, q_row, q_batch, k, k_row, k_batch, v, v_row, v_batch, out, o_row, o_batch, B, heads, Sq, Sk, scale, causal, s)
    switch (D) {
        case 40: ATT(40);
        case 64: ATT(64);
        case 80: ATT(80);
        case 160: ATT(160);
        default: return sdk_fail(SDK_ERR_UNSUPPORTED, "sdk_attention_f32_ex: head_dim %d not in {40,64,80,160}", D);
    }
#undef ATT
}

extern "C" int sdk_attention_f32(const float* q, int64_t q_row, int64_t q_batch, const float* k, int64_t k_row, int64_t k_batch,
                                 const float* v, int64_t v_row, int64_t v_batch, float* out, int64_t o_row, int64_t o_batch,
                                 int B, int heads, int Sq, int Sk, int D, float scale, void* stream) {
    return sdk_attention_f32_ex(q, q_row, q_batch, k, k_row, k_batch, v, v_row, v_batch, out, o_row, o_batch, B, heads, Sq, Sk, D, scale, 0, stream);
}
