// bf16 flash attention for the UNet's spatial self-attention and CLIP cross-attention
// (models/unet/attention.py:29-50): out = softmax(q k^T * scale) v per (batch, head), no mask.
//
// Online-softmax, one pass over K/V, nothing but q/k/v/out touches HBM.  bf16 operands, fp32
// scores / statistics / output accumulation.  CTA = 64 queries x one head, 4 warps (16 query rows
// each); K/V stream through shared memory in 64-key tiles with a 2-stage cp.async pipeline;
// QK^T and PV run on mma.sync m16n8k16 with ldmatrix-fed fragments.  Head dims 40/64/80/160 (the
// QK^T reduction is zero-padded to a multiple of 16).  At D=40 this kernel is bound by the exp
// (MUFU) rate, not by the tensor pipe: 4096^2 exps per head against 2*2*4096^2*40 flops.
#define SDK_PDL_CAT 4
#include "common.cuh"

namespace {

constexpr int BQ = 64, BKV = 64, THREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(unsigned& r0, unsigned& r1, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float* c, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {          // MUFU.EX2, flush-to-zero; exp2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}

template <int D>
__global__ void __launch_bounds__(THREADS)
attention_bf16_kernel(const __nv_bfloat16* __restrict__ q, long long q_row, long long q_batch,
                      const __nv_bfloat16* __restrict__ k, long long k_row, long long k_batch,
                      const __nv_bfloat16* __restrict__ v, long long v_row, long long v_batch,
                      __nv_bfloat16* __restrict__ out, long long o_row, long long o_batch,
                      int heads, int Sq, int Sk, float scale_log2) {
    pdl_trigger();
    pdl_wait();
    constexpr int DP = (D + 15) / 16 * 16;          // QK^T reduction length (zero padded)
    constexpr int LD = DP + 8;                      // smem row pitch in elements (+16 B: conflict-free ldmatrix)
    constexpr int KSTEPS = DP / 16;                 // k16 steps of QK^T
    constexpr int NB_O = D / 8;                     // n8 blocks of the output
    constexpr int CHUNKS = D / 8;                   // 16-byte chunks per row

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Ks = Qs + BQ * LD;               // [2][BKV][LD]
    __nv_bfloat16* Vs = Ks + 2 * BKV * LD;          // [2][BKV][LD]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
    const int q0 = blockIdx.x * BQ;
    const __nv_bfloat16* qb = q + (size_t)b * q_batch + (size_t)h * D;
    const __nv_bfloat16* kb = k + (size_t)b * k_batch + (size_t)h * D;
    const __nv_bfloat16* vb = v + (size_t)b * v_batch + (size_t)h * D;

    // zero the K-dim padding columns [D, DP) of Q and both K stages once (cp.async never touches them)
    if constexpr (DP > D) {
        constexpr int PADC = DP - D;
        for (int i = tid; i < (BQ + 2 * BKV) * PADC; i += THREADS) {
            const int r = i / PADC, c = D + i % PADC;
            Qs[r * LD + c] = __float2bfloat16_rn(0.f);       // Qs and Ks are contiguous: rows 0..BQ+2*BKV-1
        }
    }
    // Q tile
    for (int i = tid; i < BQ * CHUNKS; i += THREADS) {
        const int r = i / CHUNKS, c = i % CHUNKS;
        const bool ok = q0 + r < Sq;
        cp_async16(Qs + r * LD + c * 8, qb + (size_t)(ok ? q0 + r : 0) * q_row + c * 8, ok);
    }
    auto load_kv = [&](int tile, int stage) {
        const int k0 = tile * BKV;
        __nv_bfloat16* ks = Ks + stage * BKV * LD;
        __nv_bfloat16* vs = Vs + stage * BKV * LD;
        for (int i = tid; i < BKV * CHUNKS; i += THREADS) {
            const int r = i / CHUNKS, c = i % CHUNKS;
            const bool ok = k0 + r < Sk;
            const size_t row = ok ? (size_t)(k0 + r) : 0;
            cp_async16(ks + r * LD + c * 8, kb + row * k_row + c * 8, ok);
            cp_async16(vs + r * LD + c * 8, vb + row * v_row + c * 8, ok);
        }
    };
    const int ntiles = (Sk + BKV - 1) / BKV;
    load_kv(0, 0);
    cp_async_commit();

    float o[NB_O][4];
#pragma unroll
    for (int i = 0; i < NB_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int g = lane >> 2, tq = lane & 3;

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t & 1;
        if (t + 1 < ntiles) { load_kv(t + 1, stage ^ 1); cp_async_commit(); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const __nv_bfloat16* ks = Ks + stage * BKV * LD;
        const __nv_bfloat16* vs = Vs + stage * BKV * LD;

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks_i = 0; ks_i < KSTEPS; ++ks_i) {
            unsigned a0, a1, a2, a3;
            ldsm_x4(a0, a1, a2, a3, Qs + (warp * 16 + (lane & 15)) * LD + ks_i * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int nb = 0; nb < 8; nb += 2) {
                unsigned b0, b1, b2, b3;
                ldsm_x4(b0, b1, b2, b3, ks + (nb * 8 + (lane & 7) + ((lane >> 4) << 3)) * LD + ks_i * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16(s[nb], a0, a1, a2, a3, b0, b1);
                mma_bf16(s[nb + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        // ---- mask the ragged last tile
        if ((t + 1) * BKV > Sk) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int key = t * BKV + nb * 8 + tq * 2;
                if (key >= Sk) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
                if (key + 1 >= Sk) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
            }
        }
        // ---- online softmax (rows g and g+8 of this warp's 16)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);           // finite: each tile has >= 1 valid key
            corr[r] = fast_exp2((m_run[r] - m_new) * scale_log2);
            m_run[r] = m_new;
            msc[r] = m_new * scale_log2;
        }
        float ps[2] = {0.f, 0.f};
        unsigned pa[4][4];                                      // P as bf16 A fragments: 4 k16 steps over 64 keys
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = fast_exp2(s[nb][0] * scale_log2 - msc[0]);
            const float p1 = fast_exp2(s[nb][1] * scale_log2 - msc[0]);
            const float p2 = fast_exp2(s[nb][2] * scale_log2 - msc[1]);
            const float p3 = fast_exp2(s[nb][3] * scale_log2 - msc[1]);
            ps[0] += p0 + p1; ps[1] += p2 + p3;
            pa[nb >> 1][(nb & 1) * 2] = pack_bf16(p0, p1);
            pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 1);
            ps[r] += __shfl_xor_sync(0xffffffffu, ps[r], 2);
            l_run[r] = l_run[r] * corr[r] + ps[r];
        }
#pragma unroll
        for (int i = 0; i < NB_O; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
            const __nv_bfloat16* vrow = vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD;
#pragma unroll
            for (int nb = 0; nb + 1 < NB_O; nb += 2) {
                unsigned b0, b1, b2, b3;
                ldsm_x4_t(b0, b1, b2, b3, vrow + nb * 8 + (lane >> 4) * 8);
                mma_bf16(o[nb], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
                mma_bf16(o[nb + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
            }
            if (NB_O & 1) {
                unsigned b0, b1;
                ldsm_x2_t(b0, b1, vrow + (NB_O - 1) * 8);
                mma_bf16(o[NB_O - 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
            }
        }
        __syncthreads();                                          // stage is refilled two iterations later
    }
    pdl_trigger();
    // ---- normalise and store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
        if (row < Sq) {
            const float inv = 1.f / l_run[r];
            __nv_bfloat16* op = out + (size_t)b * o_batch + (size_t)row * o_row + (size_t)h * D + tq * 2;
#pragma unroll
            for (int nb = 0; nb < NB_O; ++nb)
                *reinterpret_cast<unsigned*>(op + nb * 8) = pack_bf16(o[nb][r * 2] * inv, o[nb][r * 2 + 1] * inv);
        }
    }
}

template <int D>
int launch(const __nv_bfloat16* q, long long q_row, long long q_batch, const __nv_bfloat16* k, long long k_row, long long k_batch,
           const __nv_bfloat16* v, long long v_row, long long v_batch, __nv_bfloat16* out, long long o_row, long long o_batch,
           int B, int heads, int Sq, int Sk, float scale, cudaStream_t s) {
    constexpr int DP = (D + 15) / 16 * 16, LD = DP + 8;
    const int smem = (BQ + 4 * BKV) * LD * 2;
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(attention_bf16_kernel<D>), (int)smem));
    dim3 grid((Sq + BQ - 1) / BQ, B * heads);
    SDK_CUDA(sdk_launch(attention_bf16_kernel<D>, dim3(grid), dim3(THREADS), (size_t)(smem), s, q, q_row, q_batch, k, k_row, k_batch, v, v_row, v_batch, out, o_row, o_batch,
                                                         heads, Sq, Sk, scale * 1.4426950408889634f));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

}  // namespace

extern "C" int sdk_attention_bf16(const void* q, int64_t q_row, int64_t q_batch, const void* k, int64_t k_row, int64_t k_batch,
                                  const void* v, int64_t v_row, int64_t v_batch, void* out, int64_t o_row, int64_t o_batch,
                                  int B, int heads, int Sq, int Sk, int D, float scale, void* stream) {
    SDK_CHECK_ARG(q && k && v && out, "sdk_attention_bf16: null pointer");
    SDK_CHECK_ARG(B > 0 && heads > 0 && Sq > 0 && Sk > 0 && B * heads < 65536, "sdk_attention_bf16: bad sizes");
    SDK_CHECK_ARG(((q_row | q_batch | k_row | k_batch | v_row | v_batch) % 8) == 0 && (o_row % 2) == 0,
                  "sdk_attention_bf16: strides must keep rows 16-byte aligned");
    SDK_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0 && ((uintptr_t)out & 3) == 0, "sdk_attention_bf16: unaligned pointer");
    cudaStream_t s = (cudaStream_t)stream;
    typedef const __nv_bfloat16* P;
#define ATT(DD) return launch<DD>((P)q, q_row, q_batch, (P)k, k_row, k_batch, (P)v, v_row, v_batch, (__nv_bfloat16*)out, o_row, o_batch, B, heads, Sq, Sk, scale, s)
    switch (D) {
        case 40: ATT(40);
        case 64: ATT(64);
        case 80: ATT(80);
        case 160: ATT(160);
        default: return sdk_fail(SDK_ERR_UNSUPPORTED, "sdk_attention_bf16: head_dim %d not in {40,64,80,160}", D);
    }
#undef ATT
}
