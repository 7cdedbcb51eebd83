// tcgen05 flash attention for head dims 40 / 64 / 80 / 160 (SD-1.5: D = 40 at S = 4096, 80 at 1024, 160 at 256 and 64; SD-2.1: D = 64
// at every level), self-attention and the 77-key cross-attention alike:
// out = softmax(q k^T * scale) v per (batch, head)   (models/unet/attention.py:29-50).
// Head dims above 64 span several 64-wide swizzle atoms: every operand tile is NA = ceil(D/64) atoms side by side (one TMA box
// each, the last one zero-filled past D), Q K^T walks its K steps across the atoms, and P V issues one MMA per atom (N = 64, or
// the 16 / 32 remaining columns), each into its own TMEM column range -- no descriptor spans more than one atom.
//
// CTA = 128 queries x one head; K/V stream in 64-key tiles.  Both contractions run on the 5th-gen tensor cores with
// TMEM accumulators, operands fetched by TMA straight out of the fused [B][S][3C] qkv buffer:
//   * per-head tiles come from 4-D tensor maps (d, head, token, batch) with a 64-wide box: for D = 40 the 24 columns
//     past the head are OUT OF BOUNDS in the innermost dimension and arrive as zeros, so a 128-byte-swizzled
//     [rows][64] tile is exactly the zero-padded operand the MMA needs (no padding pass, no neighbour-head leakage).
//   * S = Q K^T : A = Q (K-major), B = K tile (K-major), 128 x 64 fp32, DOUBLE-BUFFERED in TMEM columns [0,64) / [64,128):
//                 S(t+1) and S(t+2) are computed while the softmax of tile t runs, so the softmax warps never wait for
//                 the tensor core in steady state.
//   * softmax  : a query row = one TMEM lane, handled by TWO threads that run two INDEPENDENT flash streams: thread h of a
//                row owns keys [32h, 32h+32) of every tile (one tcgen05.ld of 32 columns, kept in registers for both the
//                max and the exp pass), with its own running max, row sum and its own P.V accumulator (split-KV inside
//                the CTA).  No cross-thread traffic inside the loop; the streams are merged once at the end
//                (log-sum-exp merge).  P (bf16) goes to shared memory in the K-major SW128 layout = A operand of the
//                second MMA, double-buffered.
//   * PV       : per stream h: A = P_h (smem, 128 x 32 keys), B = V rows [32h, 32h+32) used AS STORED ([key][d], d
//                contiguous) through an MN-major descriptor, ACCUMULATED in TMEM columns [128+64h, 192+64h) over all tiles.
//                The running max is only raised when the tile max exceeds it by more than 2^8 (P stays <= 256: exact in
//                the fp32 row sums, harmless in bf16); a raise rescales the accumulator in TMEM (tcgen05.ld / st), which
//                is rare after the first tiles.  Nothing waits for P.V inside the loop.
// Warp roles: warp 0 TMA producer (3-stage K and V rings), warp 1 MMA issuer + TMEM owner, warps 2..9 softmax (two threads
// per row).  ~98 KiB of smem and 256 TMEM columns per CTA: two CTAs per SM.
#define SDK_PDL_CAT 3
#include "common.cuh"
#include "ptx.cuh"
#include <new>
#include <string.h>

namespace {

constexpr int BQ = 128, BKV = 64, DP = 64, AT_THREADS = 64 + 256;     // TMA warp, MMA warp, 8 softmax warps
constexpr int Q_BYTES = BQ * DP * 2;                 // 16 KiB: one [128][64] bf16 atom of Q
constexpr int KV_BYTES = BKV * DP * 2;               //  8 KiB: one [64][64] bf16 atom of K / V
constexpr int P_BYTES = BQ * BKV * 2;                // 16 KiB: [128][64] bf16

// per-head-dim geometry
template <int D> struct AttnGeo {
    static constexpr int NA = (D + 63) / 64;                       // 64-wide atoms per operand row
    static constexpr int OW = D <= 64 ? 64 : D;                    // TMEM columns of one stream's P.V accumulator
    static constexpr int KV_STAGES = NA >= 3 ? 2 : 3;              // K / V ring depth (shared memory budget)
    static constexpr int TMEM_COLS = 128 + 2 * OW <= 256 ? 256 : 512;
    static constexpr int SMEM = NA * Q_BYTES + 2 * KV_STAGES * NA * KV_BYTES + 2 * P_BYTES + 1024 + 256 + 2 * 128 * 8;
    static constexpr int CTAS = (TMEM_COLS <= 256 && SMEM <= 113 * 1024) ? 2 : 1;
    // columns of the last atom's P.V MMA (a multiple of 16); full atoms use 64
    static constexpr int N_LAST = D <= 64 ? 64 : ((D - 64 * (NA - 1) + 15) / 16) * 16;
};

struct alignas(64) AttnParams {
    CUtensorMap tmQ, tmK, tmV;
    __nv_bfloat16* out; long long o_row, o_batch;
    int heads, Sq, Sk, kv_bcast, causal;
    float scale_log2;
};

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }

// MN-major SW128 operand: tile stored [K rows][64 MN elements] (128 B per K row), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(KV_BYTES >> 4) << 16;            // LBO: next 64-wide MN atom (unused: N = 64 is one atom)
    d |= (uint64_t)(1024 >> 4) << 32;                // SBO: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, AttnGeo<D>::CTAS)
attention_tc_kernel(const __grid_constant__ AttnParams p) {
    using G = AttnGeo<D>;
    constexpr int NA = G::NA, OW = G::OW, KV_STAGES = G::KV_STAGES;
    constexpr int KS = (D + 15) / 16;                                  // K=16 steps of Q K^T (zero padded past D)
    constexpr int KV_STAGE_BYTES = NA * KV_BYTES;
    constexpr uint32_t IDESC_S = ptx::umma_idesc_bf16(128, BKV);       // 128 x 64, both operands K-major
    constexpr uint32_t IDESC_O = ptx::umma_idesc_bf16(128, DP) | (1u << 16);            // 128 x 64, B (= V) MN-major
    constexpr uint32_t IDESC_OL = ptx::umma_idesc_bf16(128, G::N_LAST) | (1u << 16);    // last atom: 128 x (16 | 32 | 64)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                                // [NA] atoms
    uint8_t* sK = sQ + NA * Q_BYTES;                                   // [KV_STAGES][NA]
    uint8_t* sV = sK + KV_STAGES * KV_STAGE_BYTES;                     // [KV_STAGES][NA]
    uint8_t* sP = sV + KV_STAGES * KV_STAGE_BYTES;                     // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 4, *v_full = bars + 7, *v_empty = bars + 10,
             *s_full = bars + 13 /* [2] */, *p_full = bars + 15 /* [2] */, *pv_done = bars + 17 /* [2] */;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
    float2* s_ml = reinterpret_cast<float2*>(bars + 20);              // [2 streams][128 rows] (scaled max, row sum) for the final merge

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
    const int q0 = blockIdx.x * BQ;
    const int kvb = p.kv_bcast ? 0 : b;
    const int ntiles = (p.Sk + BKV - 1) / BKV;

    if (threadIdx.x == 0) {
        ptx::mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1); ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 256); ptx::mbar_init(&pv_done[i], 1); }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&p.tmQ); ptx::prefetch_tmap(&p.tmK); ptx::prefetch_tmap(&p.tmV);
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, G::TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(q_full, NA * Q_BYTES);
#pragma unroll
            for (int a = 0; a < NA; ++a) ptx::tma_load_4d(sQ + a * Q_BYTES, &p.tmQ, q_full, a * DP, h, q0, b);
            int st = 0; uint32_t ph = 0;
            for (int t = 0; t < ntiles; ++t) {
                ptx::mbar_wait(&k_empty[st], ph ^ 1u);
                ptx::mbar_expect_tx(&k_full[st], KV_STAGE_BYTES);
#pragma unroll
                for (int a = 0; a < NA; ++a) ptx::tma_load_4d(sK + st * KV_STAGE_BYTES + a * KV_BYTES, &p.tmK, &k_full[st], a * DP, h, t * BKV, kvb);
                ptx::mbar_wait(&v_empty[st], ph ^ 1u);
                ptx::mbar_expect_tx(&v_full[st], KV_STAGE_BYTES);
#pragma unroll
                for (int a = 0; a < NA; ++a) ptx::tma_load_4d(sV + st * KV_STAGE_BYTES + a * KV_BYTES, &p.tmV, &v_full[st], a * DP, h, t * BKV, kvb);
                if (++st == KV_STAGES) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int kst = 0; uint32_t kph = 0;                               // K ring position of the next Q K^T
            auto issue_qk = [&](int t) {
                ptx::mbar_wait(&k_full[kst], kph);
                ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < KS; ++k) {                          // K step k lives in atom k / 4, +32 B per step inside the atom
                    const uint64_t dq = ptx::umma_smem_desc_sw128(ptx::smem_u32(sQ + (k >> 2) * Q_BYTES));
                    const uint64_t dk = ptx::umma_smem_desc_sw128(ptx::smem_u32(sK + kst * KV_STAGE_BYTES + (k >> 2) * KV_BYTES));
                    ptx::umma_bf16(tmem_base + (uint32_t)(t & 1) * 64u, dq + (uint64_t)((k & 3) * 2), dk + (uint64_t)((k & 3) * 2), IDESC_S, k > 0 ? 1u : 0u);
                }
                ptx::umma_commit(&k_empty[kst]);
                ptx::umma_commit(&s_full[t & 1]);
                if (++kst == KV_STAGES) { kst = 0; kph ^= 1u; }
            };
            ptx::mbar_wait(q_full, 0);
            issue_qk(0);
            if (ntiles > 1) issue_qk(1);
            int vst = 0; uint32_t vph = 0;
            for (int t = 0; t < ntiles; ++t) {
                // ---- O_h += P_h V_h for the two key streams (own accumulators, accumulated over all tiles)
                ptx::mbar_wait(&v_full[vst], vph);
                ptx::mbar_wait(&p_full[t & 1], (uint32_t)(t >> 1) & 1u);
                ptx::tc_fence_after();
                const uint64_t dp = ptx::umma_smem_desc_sw128(ptx::smem_u32(sP + (t & 1) * P_BYTES));
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) {                      // one MMA per 64-wide atom of V (the last one narrower)
                        const uint64_t dv = umma_desc_mn_sw128(ptx::smem_u32(sV + vst * KV_STAGE_BYTES + a * KV_BYTES));
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint64_t da = dp + (uint64_t)((hh * 2 + k) * 2);                       // +32 B per 16 keys inside the 64-key atom
                            const uint64_t db = dv + (uint64_t)((hh * 2 + k) * 16 * 128 >> 4);           // +16 key rows of 128 B
                            ptx::umma_bf16(tmem_base + 128 + hh * OW + a * 64, da, db, a == NA - 1 ? IDESC_OL : IDESC_O, (t > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                }
                ptx::umma_commit(&v_empty[vst]);
                ptx::umma_commit(&pv_done[t & 1]);
                if (++vst == KV_STAGES) { vst = 0; vph ^= 1u; }
                // S buffer (t & 1) has been consumed (p_full(t) implies the softmax threads are done reading it)
                if (t + 2 < ntiles) issue_qk(t + 2);
            }
        }
    } else {
        // ================= softmax / output: TWO threads per query row (8 warps) =================
        // warps 2..5 ("stream 0") own keys [0,32) of every tile and output columns [0,32); warps 6..9 ("stream 1") own keys
        // [32,64) and output columns [32,D).  A warp may only touch TMEM lanes 32*(warp%4)..+31, so the two threads of
        // row r sit in warps with equal warp%4.
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint32_t oaddr = taddr + 128 + half * OW;                  // this stream's accumulator (OW columns, D of them used)
        const int swz = r & 7;
        constexpr float RAISE = 8.f;                                     // raise the running max only beyond 2^8 of head room
        float m_sc = -INFINITY;                                          // running max * scale*log2(e) (the exponent offset in use)
        float l_run = 0.f;
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            ptx::mbar_wait(&s_full[buf], (uint32_t)(t >> 1) & 1u);
            ptx::tc_fence_after();
            const int kbase = t * BKV + half * 32;
            const bool ragged = t * BKV + BKV > p.Sk;
            uint32_t u[32];
            ptx::tmem_ld32(taddr + buf * 64 + half * 32, u);
            ptx::tmem_ld_wait();
            if (ragged) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (kbase + j >= p.Sk) u[j] = 0xff800000u;     // -inf
            }
            if (p.causal && kbase + 31 > q0 + r) {                   // look-ahead mask (clip/attention.py:44): keys after the query
#pragma unroll
                for (int j = 0; j < 32; ++j) if (kbase + j > q0 + r) u[j] = 0xff800000u;
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 2) mx = max3(mx, __uint_as_float(u[j]), __uint_as_float(u[j + 1]));
            const float mx_sc = mx * p.scale_log2;                   // -inf when this stream holds no key of a ragged tile
            const bool raise = mx_sc > m_sc + RAISE || (m_sc == -INFINITY && mx_sc > -INFINITY);
            if (t > 0 && __any_sync(0xffffffffu, raise)) {
                // rescale this stream's accumulator (rows that do not raise use factor 1): needs P.V(t-1) retired
                ptx::mbar_wait(&pv_done[(t - 1) & 1], (uint32_t)((t - 1) >> 1) & 1u);
                ptx::tc_fence_after();
                const float corr = raise ? ex2(m_sc - mx_sc) : 1.f;   // m_sc == -inf (stream empty so far) -> 0
                l_run *= corr;
                uint32_t w[8];
#pragma unroll 1
                for (int c = 0; c < (D <= 64 ? DP : D) / 8; ++c) {   // 8 columns at a time: S stays in registers
                    ptx::tmem_ld8(oaddr + c * 8, w);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(__uint_as_float(w[j]) * corr);
                    ptx::tmem_st8(oaddr + c * 8, w);
                }
                ptx::tmem_st_wait();
            }
            if (raise) m_sc = mx_sc;
            const float msc = m_sc == -INFINITY ? 0.f : m_sc;        // empty stream: exp2(-inf - 0) = 0
            // P buffer `buf` was last read by P.V(t-2)
            if (t >= 2) ptx::mbar_wait(&pv_done[buf], (uint32_t)((t - 2) >> 1) & 1u);
            uint8_t* prow = sP + buf * P_BYTES + r * 128;
            float ps = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                             // 4 chunks of 8 keys = 16 B each
                uint32_t pk[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = i * 8 + jj * 2;
                    const float p0 = ex2(fmaf(__uint_as_float(u[j]), p.scale_log2, -msc));
                    const float p1 = ex2(fmaf(__uint_as_float(u[j + 1]), p.scale_log2, -msc));
                    ps += p0 + p1;                                   // fp32 row sum (rounding of P is zero-mean)
                    __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                    pk[jj] = *reinterpret_cast<uint32_t*>(&pb);
                }
                const int chunk = (half * 4 + i) ^ swz;               // XOR-swizzled 16-byte slot inside the 128-byte row
                *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            l_run += ps;
            ptx::fence_proxy_async();                                // generic-proxy smem writes -> visible to the MMA (async proxy)
            ptx::tc_fence_before();                                  // our TMEM reads of S / writes of O are ordered before the arrive
            ptx::mbar_arrive(&p_full[buf]);
        }
        // merge the two streams of the row: O = (O_0 e_0 + O_1 e_1) / (l_0 e_0 + l_1 e_1), e_h = 2^(m_h - max(m_0, m_1))
        s_ml[half * 128 + r] = make_float2(m_sc, l_run);
        ptx::mbar_wait(&pv_done[(ntiles - 1) & 1], (uint32_t)((ntiles - 1) >> 1) & 1u);
        ptx::tc_fence_after();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float2 other = s_ml[(half ^ 1) * 128 + r];
        const float m_all = fmaxf(m_sc, other.x);                    // finite: stream 0 always holds a key
        const float e_me = ex2(m_sc - m_all), e_ot = ex2(other.x - m_all);
        const float inv = 1.f / (l_run * e_me + other.y * e_ot);
        const float w_me = e_me * inv, w_ot = e_ot * inv;
        if constexpr (D <= 64) {
        constexpr int OD = 32;                                           // output columns per thread (stream 1 stores D - 32 of them)
        const int my_d0 = half * 32;
        const int my_nd = half == 0 ? (D < 32 ? D : 32) : D - 32;
        float o[OD];
        {
            uint32_t u[32];
            ptx::tmem_ld32(oaddr + my_d0, u);                        // own accumulator, my 32 output columns
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < OD; ++j) o[j] = __uint_as_float(u[j]) * w_me;
            ptx::tmem_ld32(taddr + 128 + (half ^ 1) * OW + my_d0, u); // the other stream's accumulator
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < OD; ++j) o[j] = fmaf(__uint_as_float(u[j]), w_ot, o[j]);
        }
        if (q0 + r < p.Sq) {
            __nv_bfloat16* op = p.out + (size_t)b * p.o_batch + (size_t)(q0 + r) * p.o_row + (size_t)h * D + my_d0;
#pragma unroll
            for (int i = 0; i < OD; i += 8) {
                if (i < my_nd) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(o[i], o[i + 1]), b2 = __floats2bfloat162_rn(o[i + 2], o[i + 3]);
                    __nv_bfloat162 c2 = __floats2bfloat162_rn(o[i + 4], o[i + 5]), d2 = __floats2bfloat162_rn(o[i + 6], o[i + 7]);
                    uint4 w;
                    w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b2);
                    w.z = *reinterpret_cast<uint32_t*>(&c2); w.w = *reinterpret_cast<uint32_t*>(&d2);
                    *reinterpret_cast<uint4*>(op + i) = w;
                }
            }
        }
        } else {
            // D = 80 / 160: the two threads of a row each write half of the head's columns, 8 at a time
            constexpr int DH = D / 2;
            static_assert(DH % 8 == 0, "half a head must be a whole number of 16-byte stores");
            const int my_d0 = half * DH;
            const bool row_ok = q0 + r < p.Sq;
            __nv_bfloat16* op = p.out + (size_t)b * p.o_batch + (size_t)(q0 + r) * p.o_row + (size_t)h * D + my_d0;
#pragma unroll 1
            for (int i = 0; i < DH; i += 8) {
                uint32_t u0[8], u1[8];
                ptx::tmem_ld8(oaddr + my_d0 + i, u0);
                ptx::tmem_ld8(taddr + 128 + (half ^ 1) * OW + my_d0 + i, u1);
                ptx::tmem_ld_wait();
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(__uint_as_float(u1[j]), w_ot, __uint_as_float(u0[j]) * w_me);
                if (row_ok) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b2 = __floats2bfloat162_rn(o[2], o[3]);
                    __nv_bfloat162 c2 = __floats2bfloat162_rn(o[4], o[5]), d2 = __floats2bfloat162_rn(o[6], o[7]);
                    uint4 w;
                    w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b2);
                    w.z = *reinterpret_cast<uint32_t*>(&c2); w.w = *reinterpret_cast<uint32_t*>(&d2);
                    *reinterpret_cast<uint4*>(op + i) = w;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, G::TMEM_COLS); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// (d, head, token, batch) view of a [B][S][row] bf16 buffer whose row holds `heads` heads of D elements
int encode_heads(CUtensorMap* m, const void* base, int D, int heads, int S, int B, int64_t row, int64_t batch, int box_tokens) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return sdk_fail(SDK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)heads, (cuuint64_t)S, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)D * 2, (cuuint64_t)row * 2, (cuuint64_t)(batch > 0 ? batch : (int64_t)S * row) * 2};
    const cuuint32_t box[4] = {(cuuint32_t)DP, 1u, (cuuint32_t)box_tokens, 1u};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sdk_fail(SDK_ERR_CUDA, "cuTensorMapEncodeTiled(attention, D=%d heads=%d S=%d) failed with %d", D, heads, S, (int)r);
    return SDK_OK;
}

struct AttnPlan { AttnParams prm; dim3 grid; int D; };

template <int D>
int launch(const AttnPlan* a, cudaStream_t s) {
    constexpr int smem = AttnGeo<D>::SMEM;
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(attention_tc_kernel<D>), (int)smem));
    SDK_CUDA(sdk_launch(attention_tc_kernel<D>, a->grid, dim3(AT_THREADS), (size_t)smem, s, a->prm));
    return SDK_OK;
}

}  // namespace

// Plan-style API (TMA descriptors are built once): strides in ELEMENTS as for sdk_attention_bf16; kv_batch == 0 broadcasts K/V.
extern "C" int sdk_attention_tc_create(const void* q, int64_t q_row, int64_t q_batch, const void* k, int64_t k_row, int64_t k_batch,
                                       const void* v, int64_t v_row, int64_t v_batch, void* out, int64_t o_row, int64_t o_batch,
                                       int B, int heads, int Sq, int Sk, int D, float scale, void** handle) {
    SDK_CHECK_ARG(q && k && v && out && handle, "sdk_attention_tc_create: null pointer");
    SDK_CHECK_ARG(D == 40 || D == 64 || D == 80 || D == 160, "sdk_attention_tc_create: head_dim %d not in {40, 64, 80, 160}", D);
    SDK_CHECK_ARG(B > 0 && heads > 0 && Sq > 0 && Sk > 0 && B * heads < 65536, "sdk_attention_tc_create: bad sizes");
    SDK_CHECK_ARG((q_row % 8) == 0 && (k_row % 8) == 0 && (v_row % 8) == 0 && (q_batch % 8) == 0 && (k_batch % 8) == 0 && (v_batch % 8) == 0 &&
                  (o_row % 8) == 0 && (o_batch % 8) == 0, "sdk_attention_tc_create: strides must keep rows 16-byte aligned");
    SDK_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0, "sdk_attention_tc_create: unaligned pointer");
    SDK_CHECK_ARG((k_batch == 0) == (v_batch == 0), "sdk_attention_tc_create: k and v must broadcast together");
    AttnPlan* a = new (std::nothrow) AttnPlan();
    if (!a) return sdk_fail(SDK_ERR_CUDA, "out of host memory");
    memset(&a->prm, 0, sizeof(a->prm));
    int rc = encode_heads(&a->prm.tmQ, q, D, heads, Sq, B, q_row, q_batch, BQ);
    const int kvB = k_batch == 0 ? 1 : B;
    if (rc == SDK_OK) rc = encode_heads(&a->prm.tmK, k, D, heads, Sk, kvB, k_row, k_batch, BKV);
    if (rc == SDK_OK) rc = encode_heads(&a->prm.tmV, v, D, heads, Sk, kvB, v_row, v_batch, BKV);
    if (rc != SDK_OK) { delete a; return rc; }
    a->prm.out = (__nv_bfloat16*)out; a->prm.o_row = o_row; a->prm.o_batch = o_batch;
    a->prm.heads = heads; a->prm.Sq = Sq; a->prm.Sk = Sk; a->prm.kv_bcast = k_batch == 0;
    a->prm.scale_log2 = scale * 1.4426950408889634f;
    a->grid = dim3((Sq + BQ - 1) / BQ, B * heads);
    a->D = D;
    *handle = a;
    return SDK_OK;
}

// causal = 1: query i attends to keys <= i only (text encoders: models/clip/attention.py:38-45 with lookahead_mask=True)
extern "C" int sdk_attention_tc_set_causal(void* handle, int causal) {
    SDK_CHECK_ARG(handle, "sdk_attention_tc_set_causal: null handle");
    ((AttnPlan*)handle)->prm.causal = causal ? 1 : 0;
    return SDK_OK;
}

extern "C" int sdk_attention_tc_launch(void* handle, void* stream) {
    SDK_CHECK_ARG(handle, "sdk_attention_tc_launch: null handle");
    AttnPlan* a = (AttnPlan*)handle;
    switch (a->D) {
        case 40: return launch<40>(a, (cudaStream_t)stream);
        case 64: return launch<64>(a, (cudaStream_t)stream);
        case 80: return launch<80>(a, (cudaStream_t)stream);
        case 160: return launch<160>(a, (cudaStream_t)stream);
    }
    return sdk_fail(SDK_ERR_ARG, "sdk_attention_tc_launch: head_dim %d", a->D);
}

extern "C" int sdk_attention_tc_destroy(void* handle) {
    delete (AttnPlan*)handle;
    return SDK_OK;
}
