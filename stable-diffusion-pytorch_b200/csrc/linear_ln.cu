// Linear (+ bias + residual) with the LayerNorm of its output rows in the SAME launch -- the three
//   h = proj(x) [+ residual]  ->  LayerNorm(h)
// pairs of every transformer block (reference models/unet/unet.py:86,137-149 and models/unet/attention.py:25,50): conv_input -> layernorm_1,
// attn1.out_proj + residual -> layernorm_2, attn2.out_proj + residual -> layernorm_3.  At UNet batch 2 the stand-alone LayerNorm
// launches (48 per step, ~8 us each for 2-16 MB tensors) are pure latency; here the producing GEMM normalises its own rows.
//
// A LayerNorm row spans N = C = 320 / 640 / 1280 fp32 columns, more than one CTA's accumulator tile (and, above 512, more than a
// CTA's TMEM).  So the N/BN n-tiles of one 128-row block form a THREAD-BLOCK CLUSTER (2, 4 or 8 CTAs):
//   * every CTA runs an ordinary tcgen05 main loop (TMA -> SW128 smem -> tcgen05.mma, fp32 accumulator in TMEM) on its 128 x BN tile;
//   * statistics pass (two groups of four epilogue warps, alternate 32-column chunks): TMEM -> registers -> + bias + residual (the
//     residual tile arrives by TMA under the main loop) -> the finished values go BACK to TMEM (tcgen05.st); each thread keeps the
//     pivot-shifted (sum, sum of squares) of its row, i.e. a (mean, M2) partial over its group's columns, and stores it into every
//     peer's shared memory (st.shared::cluster);
//   * ONE cluster barrier; every CTA combines the 2 x cluster-size partials of a row in the same order (Chan's parallel-variance
//     formula: no cancellation, the statistics equal the two-pass ones to fp32 rounding);
//   * output pass: TMEM -> fp32 chunk -> `out` and (v - mean) * rstd * gamma + beta -> bf16 chunk -> `ln_out` (the A operand of the
//     next GEMM), both by TMA tensor stores from a swizzled staging ring.
// The thread-per-row TMEM layout makes the row reductions thread-local; nothing but 2 x 128 float2 per CTA crosses the cluster.
// Loads of earlier kernels' output go through TMA only (programmatic dependent launch, see gemm_tc.cu); bias / gamma / beta are
// constants of the stream.
#define SDK_PDL_CAT 0
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/sdb200.h"
#include <new>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128, BK = 64, STAGES = 3, MAX_CLUSTER = 8;
constexpr int THREADS = 64 + 256;                    // TMA warp, MMA warp, two epilogue groups of four warps
constexpr int A_STAGE = BM * BK * 2;                 // 16 KiB
constexpr int CHUNK_BYTES = BM * 128;                // one 32-column fp32 chunk of the tile
constexpr int OUT_BUF_BYTES = CHUNK_BYTES + BM * 64; // staging of one chunk: fp32 rows (SWIZZLE_128B) + bf16 rows (SWIZZLE_64B)
constexpr int MAX_NCH = 5;

struct alignas(64) LlnParams {
    CUtensorMap tmA, tmB, tmOut, tmRes, tmLn;
    const float* bias; const float* gamma; const float* beta;
    int has_res, n_kb, nc;
    float eps, inv_n;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// two floats into the same shared-memory offset of CTA `rank` of the cluster
__device__ __forceinline__ void st_cluster_f32x2(float2* local, uint32_t rank, float a, float b) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.v2.f32 [ra], {%2, %3};\n\t"
        "}\n"
        ::"r"(ptx::smem_u32(local)), "r"(rank), "f"(a), "f"(b) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
linear_ln_kernel(const __grid_constant__ LlnParams p) {
    constexpr int B_STAGE = BN * BK * 2;
    constexpr int NCH = BN / 32;
    constexpr uint32_t TMEM_COLS = BN <= 128 ? 128 : 256;
    constexpr uint32_t IDESC = ptx::umma_idesc_bf16(BM, BN);
    static_assert(NCH <= MAX_NCH, "tile too wide");
    static_assert(4 * OUT_BUF_BYTES <= STAGES * (A_STAGE + B_STAGE), "output staging must fit behind the pipeline stages");

    // shared memory: [A stages][B stages][NCH residual chunks][exchange][barriers][bias | gamma | beta]; the output staging
    // (2 groups x 2 buffers) aliases the pipeline stages (idle once the accumulator is complete)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * A_STAGE;
    uint8_t* sRes = sB + STAGES * B_STAGE;
    float2* xch = reinterpret_cast<float2*>(sRes + MAX_NCH * CHUNK_BYTES);    // [2 * MAX_CLUSTER][BM] (mean, M2) partials
    uint64_t* full = reinterpret_cast<uint64_t*>(xch + 2 * MAX_CLUSTER * BM);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint64_t* res_full = tmem_full + 1;                                       // [MAX_NCH]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + MAX_NCH);
    float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);
    float* s_gamma = s_bias + BN;
    float* s_beta = s_gamma + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int nc = p.nc;
    const int m0 = ((int)blockIdx.x / nc) * BM;
    const int n0 = (int)rank * BN;
    const int n_kb = p.n_kb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        for (int i = 0; i < MAX_NCH; ++i) ptx::mbar_init(&res_full[i], 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&p.tmA); ptx::prefetch_tmap(&p.tmB); ptx::prefetch_tmap(&p.tmOut); ptx::prefetch_tmap(&p.tmLn);
        if (p.has_res) ptx::prefetch_tmap(&p.tmRes);
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();                 // barriers / TMEM visible to the CTA, and every peer is running before any DSMEM store
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 0) pdl_wait();               // the producer thread first fetches constant weight tiles (below)

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint32_t stage_bytes = (uint32_t)(A_STAGE + B_STAGE);
            const int pre = n_kb < STAGES ? n_kb : STAGES;
            for (int i = 0; i < pre; ++i) {                 // weights are constants of the stream: fetched before the wait
                ptx::mbar_expect_tx(&full[i], stage_bytes);
                ptx::tma_load_3d(sB + i * B_STAGE, &p.tmB, &full[i], 0, n0, i);
            }
            pdl_wait();
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < n_kb; ++i) {
                if (i >= pre) {
                    ptx::mbar_wait(&empty[s], ph ^ 1u);
                    ptx::mbar_expect_tx(&full[s], stage_bytes);
                    ptx::tma_load_3d(sB + s * B_STAGE, &p.tmB, &full[s], 0, n0, i);
                }
                ptx::tma_load_2d(sA + s * A_STAGE, &p.tmA, &full[s], i * BK, m0);
                if (++s == STAGES) { s = 0; ph ^= 1u; }
                if (i == 0 && p.has_res) {                  // the whole residual tile travels under the main loop
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
                        ptx::mbar_expect_tx(&res_full[c], (uint32_t)CHUNK_BYTES);
                        ptx::tma_load_2d(sRes + c * CHUNK_BYTES, &p.tmRes, &res_full[c], n0 + c * 32, m0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < n_kb; ++i) {
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint64_t da = ptx::umma_smem_desc_sw128(ptx::smem_u32(sA + s * A_STAGE));
                const uint64_t db = ptx::umma_smem_desc_sw128(ptx::smem_u32(sB + s * B_STAGE));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    ptx::umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (i > 0 || k > 0) ? 1u : 0u);
                ptx::umma_commit(&empty[s]);
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
            ptx::umma_commit(tmem_full);
        }
    }

    // ================= epilogue: two groups of four warps (warps 2..5, 6..9), group g owns the chunks g, g+2, ... =================
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                        // tile row == TMEM lane
    const int et = (threadIdx.x - 64) & 127;            // 0..127 within the group
    const int et2 = threadIdx.x - 64;                   // 0..255 over both groups
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t swz = (uint32_t)r & 7u;              // fp32 chunk rows: 128 B, SWIZZLE_128B
    const uint32_t swz2 = ((uint32_t)r >> 1) & 3u;      // bf16 chunk rows: 64 B, SWIZZLE_64B
    if (warp >= 2) {
        for (int j = et2; j < BN; j += 256) {
            s_bias[j] = p.bias ? __ldg(p.bias + n0 + j) : 0.f;
            s_gamma[j] = __ldg(p.gamma + n0 + j);
            s_beta[j] = __ldg(p.beta + n0 + j);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        // ---- statistics pass: finished fp32 values (acc + bias + residual) back into TMEM; shifted sums about a pivot (the row's
        // first value in this group) so that the single-pass variance does not cancel
        float pivot = 0.f, sa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
        int n_mine = 0;
#pragma unroll 1
        for (int c = grp; c < NCH; c += 2, ++n_mine) {
            uint32_t u[32];
            ptx::tmem_ld32(taddr + c * 32, u);
            ptx::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 a = *reinterpret_cast<const float4*>(s_bias + c * 32 + j);       // warp-wide broadcast
                v[j] = __uint_as_float(u[j]) + a.x; v[j + 1] = __uint_as_float(u[j + 1]) + a.y;
                v[j + 2] = __uint_as_float(u[j + 2]) + a.z; v[j + 3] = __uint_as_float(u[j + 3]) + a.w;
            }
            if (p.has_res) {
                ptx::mbar_wait(&res_full[c], 0);
                const uint8_t* rb = sRes + c * CHUNK_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(rb + ((j ^ swz) << 4));
                    v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                }
            }
            if (n_mine == 0) pivot = v[0];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float d = v[j] - pivot;
                sa[j & 3] += d; qa[j & 3] = fmaf(d, d, qa[j & 3]);
                u[j] = __float_as_uint(v[j]);
            }
            ptx::tmem_st32(taddr + c * 32, u);
        }
        ptx::tmem_st_wait();
        if (n_mine > 0) {
            const float cnt = 32.f * (float)n_mine;
            const float s1 = (sa[0] + sa[1]) + (sa[2] + sa[3]), s2 = (qa[0] + qa[1]) + (qa[2] + qa[3]);
            const float part_mean = pivot + s1 / cnt, part_m2 = fmaxf(s2 - s1 * s1 / cnt, 0.f);
            for (int k = 0; k < nc; ++k) st_cluster_f32x2(xch + ((int)rank * 2 + grp) * BM + r, (uint32_t)k, part_mean, part_m2);
        }
    }
    __syncwarp();
    ptx::cluster_sync_all();                 // (release / acquire) every CTA's row partials have landed in every peer
    if (warp >= 2) {
        // ---- combine the 2 * nc partials (counts: group 0 has ceil(NCH/2) chunks, group 1 floor(NCH/2)) in slot order: every CTA of
        // the cluster computes the same mean / rstd
        constexpr float CNT0 = 32.f * (float)((NCH + 1) / 2), CNT1 = 32.f * (float)(NCH / 2);
        float msum = 0.f;
        for (int k = 0; k < nc; ++k) {
            msum = fmaf(CNT0, xch[(2 * k) * BM + r].x, msum);
            if (NCH > 1) msum = fmaf(CNT1, xch[(2 * k + 1) * BM + r].x, msum);
        }
        const float mean = msum * p.inv_n;
        float m2 = 0.f;
        for (int k = 0; k < nc; ++k) {
            const float2 a = xch[(2 * k) * BM + r];
            const float da = a.x - mean;
            m2 += a.y + CNT0 * da * da;
            if (NCH > 1) {
                const float2 b = xch[(2 * k + 1) * BM + r];
                const float db = b.x - mean;
                m2 += b.y + CNT1 * db * db;
            }
        }
        const float rstd = rsqrtf(m2 * p.inv_n + p.eps);
        // ---- output pass: fp32 rows -> `out`, normalised bf16 rows -> `ln_out` (one bulk group per chunk, 2-buffer ring per group)
        uint8_t* my_out = smem + grp * 2 * OUT_BUF_BYTES;
        uint32_t gc = 0;
#pragma unroll 1
        for (int c = grp; c < NCH; c += 2, ++gc) {
            uint32_t u[32];
            ptx::tmem_ld32(taddr + c * 32, u);
            uint8_t* obuf = my_out + (gc & 1u) * OUT_BUF_BYTES;
            uint8_t* ob = obuf + r * 128;
            uint8_t* ob2 = obuf + CHUNK_BYTES + r * 64;
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(ob + ((j ^ swz) << 4)) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[8];
                const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c * 32 + 8 * j), g1 = *reinterpret_cast<const float4*>(s_gamma + c * 32 + 8 * j + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c * 32 + 8 * j), b1 = *reinterpret_cast<const float4*>(s_beta + c * 32 + 8 * j + 4);
                y[0] = (__uint_as_float(u[8 * j]) - mean) * rstd * g0.x + b0.x; y[1] = (__uint_as_float(u[8 * j + 1]) - mean) * rstd * g0.y + b0.y;
                y[2] = (__uint_as_float(u[8 * j + 2]) - mean) * rstd * g0.z + b0.z; y[3] = (__uint_as_float(u[8 * j + 3]) - mean) * rstd * g0.w + b0.w;
                y[4] = (__uint_as_float(u[8 * j + 4]) - mean) * rstd * g1.x + b1.x; y[5] = (__uint_as_float(u[8 * j + 5]) - mean) * rstd * g1.y + b1.y;
                y[6] = (__uint_as_float(u[8 * j + 6]) - mean) * rstd * g1.z + b1.z; y[7] = (__uint_as_float(u[8 * j + 7]) - mean) * rstd * g1.w + b1.w;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(y[0], y[1]), h1 = __floats2bfloat162_rn(y[2], y[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(y[4], y[5]), h3 = __floats2bfloat162_rn(y[6], y[7]);
                uint4 w;
                w.x = *reinterpret_cast<unsigned*>(&h0); w.y = *reinterpret_cast<unsigned*>(&h1);
                w.z = *reinterpret_cast<unsigned*>(&h2); w.w = *reinterpret_cast<unsigned*>(&h3);
                *reinterpret_cast<uint4*>(ob2 + (((uint32_t)j ^ swz2) << 4)) = w;
            }
            ptx::fence_proxy_async();                  // this thread's smem writes -> visible to the TMA engine
            if (et == 0) ptx::bulk_wait_read<0>();     // 2-buffer ring: the previous store (other buffer) has been read before anyone moves on
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et == 0) {
                tma_store_2d(&p.tmOut, obuf, n0 + c * 32, m0);
                tma_store_2d(&p.tmLn, obuf + CHUNK_BYTES, n0 + c * 32, m0);
                ptx::bulk_commit();
            }
        }
        if (et == 0) ptx::bulk_wait_read<0>();          // shared memory must outlive the last store's read
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// densely packed tensor, innermost dimension first
int encode(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, int rank, const uint64_t* dims, const uint32_t* box,
           CUtensorMapSwizzle swz, CUtensorMapL2promotion promo) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return sdk_fail(SDK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5]; cuuint64_t gstride[4]; cuuint32_t bdim[5]; cuuint32_t estr[5];
    uint64_t stride = (uint64_t)esize;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    const CUresult rc = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, bdim, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return sdk_fail(SDK_ERR_CUDA, "sdk_linear_ln: cuTensorMapEncodeTiled failed with %d", (int)rc);
    return SDK_OK;
}

struct LinearLn {
    LlnParams prm;
    int block_n, smem_bytes, grid;
};

int smem_need(int bn) {
    return 1024 + STAGES * (A_STAGE + bn * BK * 2) + MAX_NCH * CHUNK_BYTES + 2 * MAX_CLUSTER * BM * 8 + 256 + 3 * bn * 4 + 32;
}

template <int BN>
int launch_bn(const LinearLn* g, cudaStream_t s) {
    SDK_CUDA(sdk_ensure_dyn_smem(reinterpret_cast<const void*>(linear_ln_kernel<BN>), g->smem_bytes));
    sdk_prefer_max_smem_once(reinterpret_cast<const void*>(linear_ln_kernel<BN>));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g->grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = g->smem_bytes; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = g->prm.nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = sdk_pdl_enabled_cat(SDK_PDL_CAT) ? 2 : 1;
    SDK_CUDA(cudaLaunchKernelEx(&cfg, linear_ln_kernel<BN>, g->prm));
    SDK_LAUNCH_CHECK();
    return SDK_OK;
}

}  // namespace

extern "C" int sdk_linear_ln_create(const SdkLinearLnDesc* d, void** handle) {
    SDK_CHECK_ARG(d && handle, "sdk_linear_ln_create: null pointer");
    SDK_CHECK_ARG(d->a && d->w && d->out && d->ln_out && d->gamma && d->beta, "sdk_linear_ln_create: null tensor");
    SDK_CHECK_ARG(d->M > 0 && d->K > 0 && d->N > 0, "sdk_linear_ln_create: bad sizes");
    if (d->K % BK != 0) return sdk_fail(SDK_ERR_UNSUPPORTED, "sdk_linear_ln_create: K %% 64 != 0 (K=%d)", d->K);
    if (d->M >= (1ll << 31) - BM) return sdk_fail(SDK_ERR_UNSUPPORTED, "sdk_linear_ln_create: too many rows");
    // the n-tiles of a row block are one thread-block cluster (<= 8 CTAs, the portable limit): N = nc * BN
    int bn = 0;
    // both tile widths valid (N = 640): the 128-column tile gives clusters of 5 instead of 4 -- more CTAs per row block, 32 KB instead
    // of 36 KB per k-block and SM, four epilogue chunks instead of five (measured: 0.540 -> 0.533 ms per step for the 48 launches)
    static const int prefer128 = getenv("SDB200_LLN_PREFER128") ? atoi(getenv("SDB200_LLN_PREFER128")) : 1;
    if (prefer128 && d->N % 128 == 0 && d->N / 128 <= MAX_CLUSTER) bn = 128;
    else if (d->N % 160 == 0 && d->N / 160 <= MAX_CLUSTER) bn = 160;
    else if (d->N % 128 == 0 && d->N / 128 <= MAX_CLUSTER) bn = 128;
    if (!bn) return sdk_fail(SDK_ERR_UNSUPPORTED, "sdk_linear_ln_create: N=%d is not 160*k or 128*k with k <= 8", d->N);
    LinearLn* g = new (std::nothrow) LinearLn();
    if (!g) return sdk_fail(SDK_ERR_CUDA, "out of host memory");
    LlnParams& p = g->prm;
    memset(&p, 0, sizeof(p));
    g->block_n = bn;
    p.nc = d->N / bn;
    p.n_kb = d->K / BK;
    p.bias = d->bias; p.gamma = d->gamma; p.beta = d->beta;
    p.has_res = d->residual ? 1 : 0;
    p.eps = d->eps; p.inv_n = 1.0f / (float)d->N;
    const int m_tiles = (int)((d->M + BM - 1) / BM);
    g->grid = m_tiles * p.nc;
    g->smem_bytes = smem_need(bn);
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M}; const uint32_t box[2] = {BK, BM};
        rc = encode(&p.tmA, d->a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, dims, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    }
    if (rc == SDK_OK) {
        const uint64_t dims[3] = {BK, (uint64_t)d->N, (uint64_t)(d->K / BK)}; const uint32_t box[3] = {BK, (uint32_t)bn, 1};
        rc = encode(&p.tmB, d->w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    }
    const uint64_t odims[2] = {(uint64_t)d->N, (uint64_t)d->M}; const uint32_t obox[2] = {32, BM};
    if (rc == SDK_OK) rc = encode(&p.tmOut, d->out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2, odims, obox, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc == SDK_OK && d->residual)
        rc = encode(&p.tmRes, d->residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2, odims, obox, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc == SDK_OK) rc = encode(&p.tmLn, d->ln_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, odims, obox, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc != SDK_OK) { delete g; return rc; }
    *handle = g;
    return SDK_OK;
}

extern "C" int sdk_linear_ln_info(void* handle, int* out, int n) {
    SDK_CHECK_ARG(handle && out && n >= 4, "sdk_linear_ln_info: needs 4 ints");
    const LinearLn* g = static_cast<const LinearLn*>(handle);
    out[0] = g->block_n; out[1] = g->prm.nc; out[2] = g->grid; out[3] = g->smem_bytes;
    return SDK_OK;
}

extern "C" int sdk_linear_ln_launch(void* handle, void* stream) {
    SDK_CHECK_ARG(handle, "sdk_linear_ln_launch: null handle");
    const LinearLn* g = static_cast<const LinearLn*>(handle);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return g->block_n == 160 ? launch_bn<160>(g, s) : launch_bn<128>(g, s);
}

extern "C" int sdk_linear_ln_destroy(void* handle) {
    delete static_cast<LinearLn*>(handle);
    return SDK_OK;
}
