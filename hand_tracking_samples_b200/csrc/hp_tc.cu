// hp_tc.cu -- tensor-core variant of the handposedd forward pass (BASELINE.json north_star:
// "conv and FC layers run as implicit-GEMM contractions on tcgen05 tensor cores, fed by TMA
// into shared memory with TMEM accumulators").
//
// LFull::forward (cnn.h:405-429) for both FC layers is one warp-specialised persistent kernel:
//   warp 0   TMA producer: 128x64 A tiles (activations) and 256x64 B tiles (bf16 shadow of W^T)
//            into a 4-stage 128B-swizzled shared-memory ring, mbarrier full/empty handshakes
//   warp 1   one elected thread issues tcgen05.mma (M=128, N=256, K=16, BF16 x BF16 -> FP32)
//            into one of two 256-column TMEM accumulators; tcgen05.commit frees the smem stage
//   warp 2   TMEM allocation (512 columns = two accumulators, so the epilogue of tile i overlaps
//            the MMAs of tile i+1)
//   warps 4-7 epilogue: tcgen05.ld the accumulator (one row per thread), then
//            fc1: + bias, tanh (LActivation<TanH>, cnn.h:460), -> bf16 activations for fc2
//            fc2: + bias, exp, per-span sums and divide (LSoftMaxChunked::forward, cnn.h:497-511;
//                 every 256-wide N tile is exactly one span of 256 or sixteen spans of 16)
// The bound for this path is 1e-2 max-normalised against the reference (tests/test_gpu_parity.py).
#include "hp_ptx.cuh"
#include "hp_tc.cuh"

#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

namespace hp {

#define LAUNCH_CHECK(net)                \
    do {                                 \
        (net).launches++;                \
        HP_CUDA_TRY(cudaGetLastError()); \
    } while (0)

constexpr int64_t TC_CHUNK = 16384;  // crops per pass of the tensor-core path (activation workspace bound)

// ---- GEMM tile configuration -----------------------------------------------------------
// N tile is a template parameter: 256 for the big-batch GEMMs (and required by the fused chunked softmax), 64 for the
// small-M training GEMMs, where 256-wide tiles would leave most of the 148 SMs without a tile.
constexpr int BM = 128, BK = 64;                // UMMA K = 16 bf16 per instruction
// Depth of the TMA -> smem ring: always 192 KB of operands in flight per SM.  The narrow tiles serve the small-batch
// GEMMs, which are latency-bound streams of the 9.4 MB weight matrix through few CTAs: bytes in flight per SM, not the
// tensor pipe, set their pace (Little's law: 4 x 24 KB in flight gave ~60 GB/s per SM at batch 256).
__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 4 : (bn >= 128 ? 6 : 8); }
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int STG_BYTES = 4096;                 // per-warp output staging tile: 32 rows x 128 B, 16-byte chunks XOR-swizzled by row
constexpr int GEMM_THREADS = 256;
constexpr int gemm_smem(int bn) { return stages_for(bn) * (A_BYTES + bn * BK * 2) + 8 * STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/; }

enum { TC_EPI_TANH_ACT = 0, TC_EPI_SOFTMAX_F32 = 1, TC_EPI_STORE_F32 = 2, TC_EPI_DTANH = 3, TC_EPI_STORE_BF16 = 4, TC_EPI_SOFTMAX_DECODE = 5 };
enum { TC_FLAG_ACCUMULATE = 1, TC_FLAG_ROWS_HWC_TO_CHW = 2 };

struct EpiArgs {
    const float *bias;         // TANH_ACT / SOFTMAX_F32 / STORE_F32 (optional)
    void *out;                 // TANH_ACT: fp16 [M][N]; STORE_BF16: bf16 [M][N]; SOFTMAX_F32 / STORE_F32 / DTANH: fp32 [M][N] (DTANH: may be null)
    __nv_bfloat16 *out2;       // DTANH: bf16 [M][N]
    const act_t *H;            // DTANH: layer output h, fp16 [M][N]: result = (1 - h*h) * acc  (TanH::df, cnn.h:32,467)
    int flags;                 // STORE_F32: TC_FLAG_ACCUMULATE (out += acc), TC_FLAG_ROWS_HWC_TO_CHW (row k' -> (k'&63)*36 + (k'>>6))
    int ksplit;                // STORE_F32 only, 0/1 = off: K is cut into ksplit ranges of whole 64-blocks, every (range, tile) pair is a
                               // work item and range r stores its partial product at out + r*M*N (long-K, small-output weight gradients)
    float *decoded;            // SOFTMAX_DECODE: [M][48] decoded peaks (CNNOutputAnalysis, include/handtrack.h:218-241); `out` (y) may be null
    float *scratch;            // SOFTMAX_DECODE: [gridDim.x][128][256] floats, the CTA's current y tile (stays in L2)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int bind_driver()
{
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HP_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return 3;
    }
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

// 2-D 16-bit row-major [rows][cols] tensor, box = box_rows x 64 columns, 128B swizzle.
static int make_map_16(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, CUtensorMapDataType dt)
{
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols);
        return 3;
    }
    return 0;
}
static int make_map_bf16(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    return make_map_16(m, base, rows, cols, box_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
static int make_map_f16(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    return make_map_16(m, base, rows, cols, box_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

// ============================================================================
// C[M x N] = A[M x K] * Bt[N x K]^T  (+ fused epilogue); A, Bt bf16 K-major via TMA.
// ============================================================================
// OPS_F16: both operands are IEEE half (forward GEMMs); otherwise bf16 (backward GEMMs).
template <int EPI, int BN, bool OPS_F16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiArgs ea, int M, int N, int K)
{
    static_assert((EPI != TC_EPI_SOFTMAX_F32 && EPI != TC_EPI_SOFTMAX_DECODE) || BN == 256, "the fused chunked softmax needs whole 256-wide spans in one tile");
    constexpr int STAGES = stages_for(BN);
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    const float *__restrict__ bias = ea.bias;
    void *__restrict__ out = ea.out;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t *staging = smem + STAGES * STAGE_BYTES;  // [4 epilogue warps][2][STG_BYTES]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(staging + 8 * STG_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;       // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = N / BN;
    const int m_tiles = (M + BM - 1) / BM;
    const int mn_tiles = m_tiles * n_tiles;
    const int kb_total = (K + BK - 1) / BK;     // a ragged last block reads zeros (TMA out-of-bounds fill)
    const int ksplit = ((EPI == TC_EPI_STORE_F32 || EPI == TC_EPI_DTANH) && ea.ksplit > 1) ? ea.ksplit : 1;   // DTANH: the factor (1 - h^2) is applied per partial (linear)
    const int kb_per = (kb_total + ksplit - 1) / ksplit;
    const int num_tiles = mn_tiles * ksplit;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
                const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
                const int kb0 = split * kb_per, kb1 = (kb0 + kb_per < kb_total) ? kb0 + kb_per : kb_total;
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    uint8_t *sa = smem + stage * STAGE_BYTES;
                    ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform control flow, one elected lane issues =====
        constexpr uint32_t idesc = OPS_F16 ? ptx::make_idesc_f16(BM, BN) : ptx::make_idesc_bf16(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            const int split = tile / mn_tiles;
            const int kb0 = split * kb_per, kb1 = (kb0 + kb_per < kb_total) ? kb0 + kb_per : kb_total;
            for (int kb = kb0; kb < kb1; kb++) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = ptx::make_desc_sw128(sa);
                    const uint64_t bdesc = ptx::make_desc_sw128(sa + A_BYTES);
                    // advance 16 bf16 = 32 B along K inside the 128B-swizzled row: +2 in the >>4 address field
                    if (kb == kb0) ptx::umma_f16_c<false>(tmem_d, adesc, bdesc, idesc);
                    else ptx::umma_f16_c<true>(tmem_d, adesc, bdesc, idesc);
                    ptx::umma_f16_c<true>(tmem_d, adesc + 2, bdesc + 2, idesc);
                    ptx::umma_f16_c<true>(tmem_d, adesc + 4, bdesc + 4, idesc);
                    ptx::umma_f16_c<true>(tmem_d, adesc + 6, bdesc + 6, idesc);
                    ptx::umma_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread <-> accumulator row =====
        const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
            const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
            const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_full[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN;
            const float *bptr = bias ? bias + n_blk * BN : nullptr;
            uint8_t *stg = staging + ew * 2 * STG_BYTES;
            // stage one 32-row x 128-byte segment (this thread's row: 8 x 16 B) and write it out coalesced:
            // each store instruction then covers 4 rows x 128 contiguous bytes instead of 32 scattered rows
            auto flush = [&](int buf, const uint4 (&v)[8], uint8_t *gbase /*row 0 of this warp's tile, segment start*/, size_t row_pitch) {
                uint8_t *sb = stg + buf * STG_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++) *reinterpret_cast<uint4 *>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) = v[q];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = i * 4 + (lane >> 3), cq = lane & 7;
                    const uint4 val = *reinterpret_cast<const uint4 *>(sb + r * 128 + ((cq ^ (r & 7)) << 4));
                    if (m_blk * BM + ew * 32 + r < M) *reinterpret_cast<uint4 *>(gbase + (size_t)r * row_pitch + cq * 16) = val;
                }
            };
            // fp32 variant for weight gradients: optional += and optional row permutation (HWC feature order -> .cnnb rows)
            auto flush_grad = [&](int buf, const uint4 (&v)[8], float *gmat, int col0) {
                uint8_t *sb = stg + buf * STG_BYTES;
#pragma unroll
                for (int q = 0; q < 8; q++) *reinterpret_cast<uint4 *>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) = v[q];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = i * 4 + (lane >> 3), cq = lane & 7;
                    float4 val = *reinterpret_cast<const float4 *>(sb + r * 128 + ((cq ^ (r & 7)) << 4));
                    int grow = m_blk * BM + ew * 32 + r;
                    if (grow < M) {
                        if (ea.flags & TC_FLAG_ROWS_HWC_TO_CHW) grow = (grow & 63) * 36 + (grow >> 6);
                        float4 *gp = reinterpret_cast<float4 *>(gmat + (size_t)grow * N + col0 + cq * 4);
                        if (ea.flags & TC_FLAG_ACCUMULATE) {
                            const float4 o = *gp;
                            val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
                        }
                        *gp = val;
                    }
                }
            };
            if (EPI == TC_EPI_TANH_ACT || EPI == TC_EPI_STORE_BF16) {   // STORE_BF16: the plain product, rounded once
                uint8_t *gtile = reinterpret_cast<uint8_t *>(out) + ((size_t)(m_blk * BM + ew * 32) * N + n_blk * BN) * 2;
#pragma unroll 1
                for (int c = 0; c < BN / 64; c++) {   // 64 columns = 128 B of bf16 per row
                    uint4 v[8];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t r[32];
                        ptx::tmem_ld32(taddr + c * 64 + h * 32, r);
                        ptx::tmem_ld_wait();
                        uint32_t packed[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
                            if (EPI == TC_EPI_TANH_ACT) {
                                const float2 bv = *reinterpret_cast<const float2 *>(bptr + c * 64 + h * 32 + j);
                                packed[j >> 1] = pack_act(tanh_tc(v0 + bv.x), tanh_tc(v1 + bv.y));
                            } else {
                                __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                                packed[j >> 1] = *reinterpret_cast<uint32_t *>(&hh);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) v[h * 4 + q] = make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
                    }
                    flush(c & 1, v, gtile + c * 128, (size_t)N * 2);
                }
            } else if (EPI == TC_EPI_SOFTMAX_F32 || EPI == TC_EPI_SOFTMAX_DECODE) {
                uint8_t *gtile = reinterpret_cast<uint8_t *>(out) + ((size_t)(m_blk * BM + ew * 32) * N + n_blk * BN) * 4;
                const bool big = (n_blk * BN) < N_BIG_SPANS * BIG_SPAN;  // one 256-wide span vs sixteen 16-wide spans
                constexpr float LOG2E = 1.4426950408889634f;
                float inv = 0.f;
                if (big) {
                    float sum = 0.f;
#pragma unroll 1
                    for (int c = 0; c < BN / 32; c++) {
                        uint32_t r[32];
                        ptx::tmem_ld32(taddr + c * 32, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4 *>(bptr + c * 32 + j);
                            sum += ex2_fast((__uint_as_float(r[j]) + bv.x) * LOG2E) + ex2_fast((__uint_as_float(r[j + 1]) + bv.y) * LOG2E) +
                                   ex2_fast((__uint_as_float(r[j + 2]) + bv.z) * LOG2E) + ex2_fast((__uint_as_float(r[j + 3]) + bv.w) * LOG2E);
                        }
                    }
                    inv = 1.0f / sum;
                }
                // SOFTMAX_DECODE: the numeric core of CNNOutputAnalysis (include/handtrack.h:218-241) on this thread's own row
                // while it is still on chip.  The row's 256 softmax values go to a per-CTA scratch tile (re-written every
                // tile, so it lives in L2) only because the 3x3 / 5x5 neighbourhoods around a data-dependent peak need
                // indexed access; the running argmax works on the registers: per 32-column chunk its maximum (NaN-ignoring,
                // like `>` in ImageFindMax, misc_image.h:300-304), and the FIRST chunk whose maximum is strictly greater
                // wins, so the first raster-order maximum lies in that chunk and is found by one 32-value scan afterwards.
                const int drow = m_blk * BM + ew * 32 + lane;
                float *srow = (EPI == TC_EPI_SOFTMAX_DECODE) ? ea.scratch + ((size_t)blockIdx.x * BM + ew * 32 + lane) * 256 : nullptr;
                float best = -1.0f;
                int bchunk = 0;
                bool nan00 = false;
                unsigned long long p1d = 0;   // Peaks1D arg-max of the sixteen 16-wide spans, 4 bits each
#pragma unroll 1
                for (int c = 0; c < BN / 32; c++) {   // 32 columns = 128 B of fp32 per row
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4 *>(bptr + c * 32 + j);
                        v[j] = ex2_fast((__uint_as_float(r[j]) + bv.x) * LOG2E);
                        v[j + 1] = ex2_fast((__uint_as_float(r[j + 1]) + bv.y) * LOG2E);
                        v[j + 2] = ex2_fast((__uint_as_float(r[j + 2]) + bv.z) * LOG2E);
                        v[j + 3] = ex2_fast((__uint_as_float(r[j + 3]) + bv.w) * LOG2E);
                    }
                    float i0 = inv, i1 = inv;
                    if (!big) {
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; j++) { s0 += v[j]; s1 += v[16 + j]; }
                        i0 = 1.0f / s0;
                        i1 = 1.0f / s1;
                    }
                    float yv[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) yv[j] = v[j] * ((j < 16) ? i0 : i1);
                    uint4 o[8];
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        o[q] = make_uint4(__float_as_uint(yv[q * 4]), __float_as_uint(yv[q * 4 + 1]), __float_as_uint(yv[q * 4 + 2]), __float_as_uint(yv[q * 4 + 3]));
                    if (EPI == TC_EPI_SOFTMAX_DECODE) {
#pragma unroll
                        for (int q = 0; q < 8; q++) reinterpret_cast<float4 *>(srow + c * 32)[q] = make_float4(yv[q * 4], yv[q * 4 + 1], yv[q * 4 + 2], yv[q * 4 + 3]);
                        if (big) {
                            float mc = fmaxf(yv[0], yv[1]);
#pragma unroll
                            for (int j = 2; j < 32; j++) mc = fmaxf(mc, yv[j]);
                            if (mc > best) { best = mc; bchunk = c; }
                            if (c == 0) nan00 = yv[0] != yv[0];
                        } else {
#pragma unroll
                            for (int sp = 0; sp < 2; sp++) {   // Peaks1D, misc_image.h:389-399: if (r[p] < r[x]) p = x
                                float pv = yv[16 * sp];
                                int pi = 0;
#pragma unroll
                                for (int xx = 1; xx < 16; xx++)
                                    if (pv < yv[16 * sp + xx]) { pv = yv[16 * sp + xx]; pi = xx; }
                                p1d |= (unsigned long long)pi << (4 * (2 * c + sp));
                            }
                        }
                        if (out) flush(c & 1, o, gtile + c * 128, (size_t)N * 4);
                    } else {
                        flush(c & 1, o, gtile + c * 128, (size_t)N * 4);
                    }
                }
                if (EPI == TC_EPI_SOFTMAX_DECODE && drow < M) {
                    // Every read below is this thread's own row of the scratch tile (L2 hits); the loads of a step are issued
                    // together (fully unrolled, predicated) so that the tail costs three L2 round trips, not one per value.
                    float *dec = ea.decoded + (size_t)drow * 48;
                    if (big) {
                        // ImageFindMax: the first element of the winning chunk that equals the maximum; pixel (0,0) NaN keeps (0,0)
                        float cv[32];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 f = reinterpret_cast<const float4 *>(srow + bchunk * 32)[q];
                            cv[4 * q] = f.x; cv[4 * q + 1] = f.y; cv[4 * q + 2] = f.z; cv[4 * q + 3] = f.w;
                        }
                        int bi = bchunk * 32 + 31;
#pragma unroll
                        for (int j = 30; j >= 0; j--)
                            if (cv[j] == best) bi = bchunk * 32 + j;
                        if (nan00) bi = 0;
                        const int bx = bi & 15, by = bi >> 4;
                        float w9[9];
#pragma unroll
                        for (int k = 0; k < 9; k++) {
                            const int sy_ = by - 1 + k / 3, sx = bx - 1 + k % 3;
                            w9[k] = (sy_ >= 0 && sy_ < 16 && sx >= 0 && sx < 16) ? srow[sy_ * 16 + sx] : 0.0f;
                        }
                        const float peak = srow[bi];
                        float wsum = 0.0f, vx = 0.0f, vy = 0.0f;
#pragma unroll
                        for (int k = 0; k < 9; k++) {   // PeakSubPixel, misc_image.h:316-322: rows ascending, columns ascending, clipped
                            const int sy_ = by - 1 + k / 3, sx = bx - 1 + k % 3;
                            if (sy_ >= 0 && sy_ < 16 && sx >= 0 && sx < 16) {
                                vx = __fadd_rn(vx, __fmul_rn((float)sx, w9[k]));
                                vy = __fadd_rn(vy, __fmul_rn((float)sy_, w9[k]));
                                wsum = __fadd_rn(wsum, w9[k]);
                            }
                        }
                        const float px = (wsum == 0) ? (float)bx : __fdiv_rn(vx, wsum);
                        const float py = (wsum == 0) ? (float)by : __fdiv_rn(vy, wsum);
                        // PeakVolume, misc_image.h:330 (float -> int as the pinned x86 build converts it, see hp_post.cu)
                        const float fx = __fadd_rn(px, 0.5f), fy = __fadd_rn(py, 0.5f);
                        const bool x_ok = fx >= -2147483648.0f && fx < 2147483648.0f, y_ok = fy >= -2147483648.0f && fy < 2147483648.0f;
                        const int rx = x_ok ? (int)fx : 0, ry = y_ok ? (int)fy : 0;
                        float vol = 0.0f;
                        if (x_ok && y_ok) {
                            float v9[9];
#pragma unroll
                            for (int k = 0; k < 9; k++) {
                                const int sy_ = ry - 1 + k / 3, sx = rx - 1 + k % 3;
                                v9[k] = (sy_ >= 0 && sy_ < 16 && sx >= 0 && sx < 16) ? srow[sy_ * 16 + sx] : 0.0f;
                            }
#pragma unroll
                            for (int k = 0; k < 9; k++) {
                                const int sy_ = ry - 1 + k / 3, sx = rx - 1 + k % 3;
                                if (sy_ >= 0 && sy_ < 16 && sx >= 0 && sx < 16) vol = __fadd_rn(vol, v9[k]);
                            }
                        }
                        *reinterpret_cast<float4 *>(dec + 4 * n_blk) = make_float4(px, py, vol, peak);
                    } else {
                        float r3[16][3];
#pragma unroll
                        for (int sp = 0; sp < 16; sp++) {
                            const int pi = (int)((p1d >> (4 * sp)) & 15);
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const int i = pi - 1 + k;
                                r3[sp][k] = (i >= 0 && i < 16) ? srow[16 * sp + i] : 0.0f;
                            }
                        }
#pragma unroll
                        for (int sp = 0; sp < 16; sp++) {
                            const int pi = (int)((p1d >> (4 * sp)) & 15);
                            float vv = 0.0f, wsum = 0.0f;
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const int i = pi - 1 + k;
                                if (i >= 0 && i < 16) {
                                    vv = __fadd_rn(vv, __fmul_rn((float)i, r3[sp][k]));
                                    wsum = __fadd_rn(wsum, r3[sp][k]);
                                }
                            }
                            dec[32 + sp] = __fdiv_rn((wsum == 0) ? (float)pi : __fdiv_rn(vv, wsum), 15.0f);
                        }
                    }
                }
            }
            if (EPI == TC_EPI_STORE_F32) {
#pragma unroll 1
                for (int c = 0; c < BN / 32; c++) {
                    uint32_t r[32];
                    ptx::tmem_ld32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
                    uint4 o[8];
                    if (bias && split == 0) {   // LFull::forward: Y = B + x*W (cnn.h:407), logits for the separate softmax / loss kernel (split-K: the bias goes into the first partial)
#pragma unroll
                        for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) + bptr[c * 32 + j]);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) o[q] = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
                    flush_grad(c & 1, o, reinterpret_cast<float *>(out) + (size_t)split * M * N, n_blk * BN + c * 32);
                }
            }
            if (EPI == TC_EPI_DTANH) {
                const int row = m_blk * BM + ew * 32 + lane;
                const int rowc = row < M ? row : M - 1;   // clamp: rows past M are computed but never stored
                const act_t *hrow = ea.H + (size_t)rowc * N + n_blk * BN;
                uint8_t *gt32 = reinterpret_cast<uint8_t *>(out) + ((size_t)split * M * N + (size_t)(m_blk * BM + ew * 32) * N + n_blk * BN) * 4;
                uint8_t *gt16 = reinterpret_cast<uint8_t *>(ea.out2) + ((size_t)(m_blk * BM + ew * 32) * N + n_blk * BN) * 2;
                int fb = 0;
#pragma unroll 1
                for (int c = 0; c < BN / 64; c++) {
                    uint4 ob[8];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t r[32];
                        ptx::tmem_ld32(taddr + c * 64 + h * 32, r);
                        ptx::tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint4 hv = *reinterpret_cast<const uint4 *>(hrow + c * 64 + h * 32 + q * 8);
                            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float2 hf = unpack_act(hw[k]);
                                v[q * 8 + 2 * k] = (1.0f - hf.x * hf.x) * __uint_as_float(r[q * 8 + 2 * k]);
                                v[q * 8 + 2 * k + 1] = (1.0f - hf.y * hf.y) * __uint_as_float(r[q * 8 + 2 * k + 1]);
                            }
                        }
                        if (out) {
                            uint4 o[8];
#pragma unroll
                            for (int q = 0; q < 8; q++)
                                o[q] = make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
                            flush(fb & 1, o, gt32 + (c * 64 + h * 32) * 4, (size_t)N * 4);
                            fb++;
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            uint32_t pk[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                __nv_bfloat162 hh = __floats2bfloat162_rn(v[q * 8 + 2 * k], v[q * 8 + 2 * k + 1]);
                                pk[k] = *reinterpret_cast<uint32_t *>(&hh);
                            }
                            ob[h * 4 + q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    if (ea.out2) {   // (null with split-K: a bf16 copy of a partial is of no use)
                        flush(fb & 1, ob, gt16 + c * 128, (size_t)N * 2);
                        fb++;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// fp32 W[K][N] (row-major, the .cnnb layout of LFull, cnn.h:417) -> fp16 Wt[N][K] (forward B operand)
// HWC: destination k' = pp*64 + co reads source row co*36 + pp (the conv kernel emits its features
// pixel-major, the reference flattens channel-major: x + 6y + 36c).
template <bool HWC>
__global__ void __launch_bounds__(256) transpose_to_act(const float *__restrict__ w, act_t *__restrict__ wt, int K, int N)
{
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int kd = k0 + ty + 8 * i;
        const int ks = HWC ? (kd & 63) * 36 + (kd >> 6) : kd;
        tile[ty + 8 * i][tx] = w[(size_t)ks * N + n0 + tx];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) wt[(size_t)(n0 + ty + 8 * i) * K + k0 + tx] = __float2half_rn(tile[tx][ty + 8 * i]);
}


// fp32 W[K][N] -> bf16 Wb[K'][N] in the stored orientation (the B operand of the dX GEMMs, LFull::backward
// cnn.h:430-437); HWC: destination row k' = pp*64+co reads source row co*36+pp.
template <bool HWC>
__global__ void __launch_bounds__(256) convert_rows_bf16(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wb, int N)
{
    const int kd = blockIdx.x;
    const int ks = HWC ? (kd & 63) * 36 + (kd >> 6) : kd;
    const float4 *src = reinterpret_cast<const float4 *>(w + (size_t)ks * N);
    uint2 *dst = reinterpret_cast<uint2 *>(wb + (size_t)kd * N);
    for (int i = threadIdx.x; i < N / 4; i += 256) {
        const float4 v = src[i];
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        dst[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
}

// One GPU, tensor path: SGD on one FC weight matrix (w <- fma(-alpha, g, w), sgd_kernel's arithmetic) fused with the
// rebuild of both of its 16-bit shadows -- transpose_to_act + convert_rows_bf16 above read the updated fp32 weights twice
// more; here every weight and gradient is read once and the three results are written from registers / one smem tile.
// 75 MB of HBM traffic per FC bucket instead of 113 MB, next to the backward kernels that share the bandwidth.
// One 32x32 tile per CTA.  (A persistent two-CTAs-per-SM variant that leaves room for the GEMM CTAs beside it was tried:
// the co-resident weight-gradient GEMM then ran 25 us instead of 9.  What works is priority: the step's main chain runs on
// a high-priority stream, so its CTAs are dispatched ahead of the pending tiles of this kernel -- hp_api.cu.)
template <bool HWC>
__global__ void __launch_bounds__(256) sgd_refresh_fc(float *__restrict__ w, const float *__restrict__ g, float alpha, act_t *__restrict__ wt,
                                                      __nv_bfloat16 *__restrict__ wb, int K, int N)
{
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int kd = k0 + ty + 8 * i;
        const int ks = HWC ? (kd & 63) * 36 + (kd >> 6) : kd;
        const size_t o = (size_t)ks * N + n0 + tx;
        const float v = fmaf(-alpha, g[o], w[o]);
        w[o] = v;
        wb[(size_t)kd * N + n0 + tx] = __float2bfloat16_rn(v);
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) wt[(size_t)(n0 + ty + 8 * i) * K + k0 + tx] = __float2half_rn(tile[tx][ty + 8 * i]);
}

// 16-bit src[n][C] -> bf16 dst[C][ldk] (k contiguous) with zero fill for n <= k < n_pad: the K-major operands of the
// weight-gradient GEMMs (LFull::update, cnn.h:438-445, is the contraction over the batch).  Both operands of one GEMM
// are transposed by one launch (blockIdx.z picks the job).
struct TransposeJob {
    const void *src;           // 16-bit [n][C]
    __nv_bfloat16 *dst;
    int C;
    int src_f16;               // source is fp16 (forward activations h1, p2): converted to bf16 on the way
};
__global__ void __launch_bounds__(256) transpose_bf16_pair(TransposeJob j0, TransposeJob j1, int n, int n_pad, int ldk)
{
    __shared__ __nv_bfloat16 tile[32][34];
    const TransposeJob j = blockIdx.z ? j1 : j0;
    const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    if (c0 >= j.C) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = k0 + ty + 8 * i;
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (k < n) {
            if (j.src_f16) v = __float2bfloat16_rn(__half2float(reinterpret_cast<const __half *>(j.src)[(size_t)k * j.C + c0 + tx]));
            else v = reinterpret_cast<const __nv_bfloat16 *>(j.src)[(size_t)k * j.C + c0 + tx];
        }
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = k0 + tx;
        if (k < n_pad) j.dst[(size_t)(c0 + ty + 8 * i) * ldk + k] = tile[tx][ty + 8 * i];
    }
}

// ---- conv2 backward on the tensor cores ---------------------------------------------------------------------
// LConv::backward (cnn.h:258-268) and LConv::update (cnn.h:269-279) of the 4x4 16->64 layer as two GEMMs over the
// dense error E[(n,pos)][co] (LMaxPool::backward, cnn.h:149-164, puts the fc1-side gradient at each 2x2 window's
// winner and zero elsewhere):
//     dL/dcol [(n,pos)][k]  = E  x W2[co][k]                        (M = 144 n, N = 256, K = 64)      -> col2im -> dL/dp1
//     dW2^T   [k][co]       = col^T[k][(n,pos)] x E^T[co][(n,pos)]  (M = 256, N = 64, K = 144 n, split-K)
// with k = (ky*4+kx)*16 + ci.  This kernel writes the three bf16 operands for one crop: E rows, and the
// (n,pos)-contiguous transposes E^T and col^T (im2col of p1).  The CTA of the last crop also zeroes the columns up to
// the next multiple of 64 that the last K block of the split-K GEMM reads.
__global__ void __launch_bounds__(256) conv2_bwd_operands(const float *__restrict__ g2_hwc, const uint8_t *__restrict__ idx2, const float *__restrict__ p1,
                                                          __nv_bfloat16 *__restrict__ E, __nv_bfloat16 *__restrict__ ET, __nv_bfloat16 *__restrict__ colT,
                                                          float *__restrict__ db_partial, int n, int64_t ld, const float *__restrict__ g2_hwc_b = nullptr)
{
    // Shared-memory layouts chosen so that every phase reads conflict-free (the first version of this kernel spent its
    // time in 5- to 9-way conflicted loads): the gradient transposed to [co][37], the dense error as bf16 [pos][66]
    // (row pitch 33 words: a column walk over pos pairs is a stride-2 walk over the banks).
    constexpr int GP = 37, EP = 66;
    __shared__ __align__(16) float sgT[C2_CO * GP];
    __shared__ __align__(16) float sp1[P1_N];
    __shared__ __align__(16) uint8_t si[P2_N];
    __shared__ __align__(16) __nv_bfloat16 sE[C2_POS * EP];
    __shared__ __align__(16) uint8_t s_off[C2_POS];   // pos -> y*15 + x      (its patch origin in a p1 plane)
    __shared__ __align__(16) uint8_t s_pp[C2_POS];    // pos -> pooled window (y/2)*6 + x/2
    __shared__ __align__(16) uint8_t s_sub[C2_POS];   // pos -> offset inside the window (y&1)*2 + (x&1)
    const int64_t crop = blockIdx.x;
    const int t = threadIdx.x;
    for (int i = t; i < P2_N / 4; i += 256) {
        float4 v = reinterpret_cast<const float4 *>(g2_hwc + crop * P2_N)[i];   // HWC: index pp*64 + co
        if (g2_hwc_b) {   // the fc1 dX GEMM ran as two K ranges: the gradient is the sum of its two partial products
            const float4 u = reinterpret_cast<const float4 *>(g2_hwc_b + crop * P2_N)[i];
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        const int pp = i >> 4, co0 = (i & 15) * 4;
        sgT[(co0 + 0) * GP + pp] = v.x;
        sgT[(co0 + 1) * GP + pp] = v.y;
        sgT[(co0 + 2) * GP + pp] = v.z;
        sgT[(co0 + 3) * GP + pp] = v.w;
        if (i < P2_N / 16) reinterpret_cast<uint4 *>(si)[i] = reinterpret_cast<const uint4 *>(idx2 + crop * P2_N)[i];
    }
    for (int i = t; i < P1_N / 4; i += 256) reinterpret_cast<float4 *>(sp1)[i] = reinterpret_cast<const float4 *>(p1 + crop * P1_N)[i];
    if (t < C2_POS) {
        const int y = t / C2_W, xx = t - y * C2_W;
        s_off[t] = (uint8_t)(y * P1_W + xx);
        s_pp[t] = (uint8_t)((y >> 1) * P2_W + (xx >> 1));
        s_sub[t] = (uint8_t)((y & 1) * 2 + (xx & 1));
    }
    __syncthreads();
    // conv2 dB partial of this crop (LConv::update, cnn.h:277): sum over the 36 windows
    if (t < C2_CO) {
        float a = 0.f;
#pragma unroll 4
        for (int pp = 0; pp < 36; pp++) a += sgT[t * GP + pp];
        db_partial[crop * C2_CO + t] = a;
    }
    // dense error (LMaxPool::backward, cnn.h:149-164): the window's gradient at its winner, zero elsewhere.  Lanes walk co.
    for (int i = t; i < C2_POS * C2_CO; i += 256) {
        const int pos = i >> 6, co = i & 63;
        const int pp = s_pp[pos];
        const float v = (si[co * 36 + pp] == s_sub[pos]) ? sgT[co * GP + pp] : 0.f;
        sE[pos * EP + co] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    // E: row (crop*144 + pos) = 64 co = 8 x 16 B
    {
        uint4 *dst = reinterpret_cast<uint4 *>(E + crop * (int64_t)(C2_POS * C2_CO));
        for (int i = t; i < C2_POS * 8; i += 256) {
            const uint32_t *r = reinterpret_cast<const uint32_t *>(sE + (i >> 3) * EP + (i & 7) * 8);
            dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
    // E^T: row co, columns crop*144 + pos; a lane writes one pos pair (4 B), a warp 128 contiguous bytes
    for (int i = t; i < C2_CO * (C2_POS / 2); i += 256) {
        const int co = i / (C2_POS / 2), jj = i - co * (C2_POS / 2);
        __nv_bfloat162 v;
        v.x = sE[(2 * jj) * EP + co];
        v.y = sE[(2 * jj + 1) * EP + co];
        *reinterpret_cast<__nv_bfloat162 *>(ET + co * ld + crop * C2_POS + 2 * jj) = v;
    }
    // col^T (im2col of p1): row k = tap*16 + ci, columns crop*144 + pos, again one pos pair per lane
    for (int i = t; i < C2_KDIM * (C2_POS / 2); i += 256) {
        const int k = i / (C2_POS / 2), jj = i - k * (C2_POS / 2);
        const int ci = k & 15, ky = k >> 6, kx = (k >> 4) & 3;
        const float *src = sp1 + ci * (P1_W * P1_H) + ky * P1_W + kx;
        const uchar2 o = *reinterpret_cast<const uchar2 *>(s_off + 2 * jj);
        *reinterpret_cast<__nv_bfloat162 *>(colT + k * ld + crop * C2_POS + 2 * jj) = __floats2bfloat162_rn(src[o.x], src[o.y]);
    }
    if (crop == n - 1) {
        const int64_t R = (int64_t)n * C2_POS, R64 = (R + 63) / 64 * 64;
        const int tail = (int)(R64 - R);
        for (int i = t; i < (C2_CO + C2_KDIM) * tail; i += 256) {
            const int row = i / tail, c = i - row * tail;
            if (row < C2_CO) ET[row * ld + R + c] = __float2bfloat16_rn(0.f);
            else colT[(row - C2_CO) * ld + R + c] = __float2bfloat16_rn(0.f);
        }
    }
}

// col2im of dL/dcol [(n,pos)][k = tap*16 + ci] (LConv::backward, cnn.h:258-268, as a gather) fused with TanH::df of the
// conv1 stage: g1[n][ci][Y][X] = (1 - p1^2) * sum_{ky,kx} dcol[(Y-ky, X-kx)][(ky,kx,ci)].  One crop per CTA; a thread
// owns 4 consecutive ci of one pixel; dL/dcol is stored as bf16 (half the bytes of this memory-bound pair of kernels: the
// GEMM that writes it and this gather), every load is 8 B and a warp reads 8 x 32 contiguous bytes per tap.
__global__ void __launch_bounds__(256) col2im_g1_vec(const __nv_bfloat16 *__restrict__ colgrad, const float *__restrict__ p1, float *__restrict__ g1)
{
    __shared__ float so[P1_N];
    const int64_t crop = blockIdx.x;
    const __nv_bfloat16 *cg = colgrad + crop * (int64_t)(C2_POS * C2_KDIM);
    for (int i = threadIdx.x; i < P1_W * P1_H * 4; i += 256) {
        const int r = i >> 2, c4 = (i & 3) * 4;
        const int Y = r / P1_W, X = r - Y * P1_W;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 4; ky++) {
            const int y = Y - ky;
            if (y < 0 || y >= C2_H) continue;
#pragma unroll
            for (int kx = 0; kx < 4; kx++) {
                const int xx = X - kx;
                if (xx < 0 || xx >= C2_W) continue;
                const uint2 v = *reinterpret_cast<const uint2 *>(cg + (y * C2_W + xx) * C2_KDIM + (ky * 4 + kx) * 16 + c4);
                const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.x));
                const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.y));
                a.x += lo.x; a.y += lo.y; a.z += hi.x; a.w += hi.y;
            }
        }
        so[(c4 + 0) * 225 + r] = a.x;
        so[(c4 + 1) * 225 + r] = a.y;
        so[(c4 + 2) * 225 + r] = a.z;
        so[(c4 + 3) * 225 + r] = a.w;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P1_N / 4; i += 256) {
        const float4 pv = reinterpret_cast<const float4 *>(p1 + crop * P1_N)[i];
        const float4 d = reinterpret_cast<const float4 *>(so)[i];
        reinterpret_cast<float4 *>(g1 + crop * P1_N)[i] =
            make_float4((1.0f - pv.x * pv.x) * d.x, (1.0f - pv.y * pv.y) * d.y, (1.0f - pv.z * pv.z) * d.z, (1.0f - pv.w * pv.w) * d.w);
    }
}

// dW2[co][ci*16+tap] (+)= sum_s partial[s][k = tap*16+ci][co]   (s ascending: deterministic)
__global__ void __launch_bounds__(256) reduce_c2w_t(float *__restrict__ dst, const float *__restrict__ partial, int S, int accumulate)
{
    __shared__ float red[4][64];
    const int k = blockIdx.x, co = threadIdx.x & 63, q = threadIdx.x >> 6;
    float a = 0.f;
    for (int s = q; s < S; s += 4) a += partial[(size_t)s * (C2_KDIM * C2_CO) + k * C2_CO + co];
    red[q][co] = a;
    __syncthreads();
    if (q == 0) {
        const int ci = k & 15, tap = k >> 4;
        float *d = dst + co * C2_KDIM + ci * 16 + tap;
        const float v = (red[0][co] + red[1][co]) + (red[2][co] + red[3][co]);
        *d = accumulate ? *d + v : v;
    }
}

// Train's loss (cnn.h:566-569) and LSoftMaxChunked::backward (cnn.h:512-526) from the softmax OUTPUT y:
// e = y - t, mse = sum e^2 / 2304, per span dp = sum e*y, dlogit = y * (e - dp).  fp32 and bf16 copies.
__global__ void __launch_bounds__(256) loss_from_y(const float *__restrict__ y, const float *__restrict__ t, float *__restrict__ dlog,
                                                   __nv_bfloat16 *__restrict__ dlog_bf, float *__restrict__ mse)
{
    __shared__ float red[8];
    const int64_t crop = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *yy = y + crop * N_OUT, *tt = t + crop * N_OUT;
    float yv[8], e[8], se = 0.f, dp = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        yv[i] = yy[warp * 256 + i * 32 + lane];
        e[i] = yv[i] - tt[warp * 256 + i * 32 + lane];
        se += e[i] * e[i];
        dp += e[i] * yv[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dp += __shfl_xor_sync(0xffffffffu, dp, o);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float d = yv[i] * (e[i] - dp);
        dlog[crop * N_OUT + warp * 256 + i * 32 + lane] = d;
        dlog_bf[crop * N_OUT + warp * 256 + i * 32 + lane] = __float2bfloat16_rn(d);
    }
    const float ys = yy[2048 + tid];
    const float e2 = ys - tt[2048 + tid];
    se += e2 * e2;
    float dp2 = e2 * ys;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dp2 += __shfl_xor_sync(0xffffffffu, dp2, o);
    const float d2 = ys * (e2 - dp2);
    dlog[crop * N_OUT + 2048 + tid] = d2;
    dlog_bf[crop * N_OUT + 2048 + tid] = __float2bfloat16_rn(d2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) red[warp] = se;
    __syncthreads();
    if (tid == 0 && mse) {
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) m += red[k];
        mse[crop] = m / (float)N_OUT;
    }
}

void tc_set_reserved_sms(Net &net, int reserve)
{
    TcState *t = net.tc;
    if (reserve < 0) reserve = 0;
    if (reserve > t->total_sms / 2) reserve = t->total_sms / 2;
    t->num_sms = t->total_sms - reserve;
}

int tc_init(Net &net)
{
    if (int rc = bind_driver()) return rc;
    TcState *t = new TcState;
    net.tc = t;
    cudaDeviceProp prop;
    HP_CUDA_TRY(cudaGetDeviceProperties(&prop, net.device));
    t->num_sms = t->total_sms = prop.multiProcessorCount;
    if (const char *e = getenv("HP_TC_RESERVE_SMS")) {   // experiments: single-GPU runs with the data-parallel SM reservation
        const int r = atoi(e);
        if (r > 0 && r < t->total_sms / 2) t->num_sms = t->total_sms - r;
    }
    HP_CUDA_TRY(cudaMalloc((void **)&t->w1t, (size_t)FC1_OUT * FC1_IN * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->w2t, (size_t)FC2_OUT * FC2_IN * 2));
    if (int rc = make_map_f16(&t->tm_w1t, t->w1t, FC1_OUT, FC1_IN, 256)) return rc;
    if (int rc = make_map_f16(&t->tm_w2t, t->w2t, FC2_OUT, FC2_IN, 256)) return rc;
    if (int rc = make_map_f16(&t->tm_w1t64, t->w1t, FC1_OUT, FC1_IN, 64)) return rc;
    if (int rc = make_map_f16(&t->tm_w2t64, t->w2t, FC2_OUT, FC2_IN, 64)) return rc;
#define HP_GEMM_ATTR(EPI, BN, F16) \
    HP_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel<EPI, BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(BN)))
    HP_GEMM_ATTR(TC_EPI_TANH_ACT, 256, true);
    HP_GEMM_ATTR(TC_EPI_TANH_ACT, 64, true);
    HP_GEMM_ATTR(TC_EPI_SOFTMAX_F32, 256, true);
    HP_GEMM_ATTR(TC_EPI_SOFTMAX_DECODE, 256, true);
    HP_GEMM_ATTR(TC_EPI_STORE_F32, 64, true);      // fc2 logits at small batch
    HP_GEMM_ATTR(TC_EPI_STORE_F32, 256, false);    // weight gradients
    HP_GEMM_ATTR(TC_EPI_STORE_F32, 128, false);
    HP_GEMM_ATTR(TC_EPI_STORE_F32, 64, false);
    HP_GEMM_ATTR(TC_EPI_DTANH, 256, false);
    HP_GEMM_ATTR(TC_EPI_DTANH, 64, false);
    HP_GEMM_ATTR(TC_EPI_STORE_BF16, 256, false);
#undef HP_GEMM_ATTR
    HP_CUDA_TRY(cudaMalloc((void **)&t->w1b, (size_t)FC1_OUT * FC1_IN * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->w2b, (size_t)FC2_OUT * FC2_IN * 2));
    if (int rc = make_map_bf16(&t->tm_w1b, t->w1b, FC1_IN, FC1_OUT, 256)) return rc;
    if (int rc = make_map_bf16(&t->tm_w2b, t->w2b, FC2_IN, FC2_OUT, 256)) return rc;
    if (int rc = make_map_bf16(&t->tm_w1b64, t->w1b, FC1_IN, FC1_OUT, 64)) return rc;
    if (int rc = make_map_bf16(&t->tm_w2b64, t->w2b, FC2_IN, FC2_OUT, 64)) return rc;
    HP_CUDA_TRY(cudaMalloc((void **)&t->w2kt, (size_t)C2_KDIM * C2_CO * 2));
    if (int rc = make_map_bf16(&t->tm_w2kt, t->w2kt, C2_KDIM, C2_CO, 256)) return rc;
    return tc_conv_init(net);
}

void tc_destroy(Net &net)
{
    TcState *t = net.tc;
    if (!t) return;
    if (t->w1t) cudaFree(t->w1t);
    if (t->w2t) cudaFree(t->w2t);
    void *tb[] = {t->w1b, t->w2b, t->dlog_bf, t->da1_bf, t->g2_sink, t->h1T, t->dlogT, t->p2T, t->da1T, t->e2, t->e2T, t->colT, t->w2kt, t->db2_partial};
    for (void *q : tb)
        if (q) cudaFree(q);
    if (t->b1_img) cudaFree(t->b1_img);
    if (t->b2_img) cudaFree(t->b2_img);
    if (t->a2_img) cudaFree(t->a2_img);
    if (t->dec_scratch) cudaFree(t->dec_scratch);
    if (t->p2) cudaFree(t->p2);
    if (t->h1) cudaFree(t->h1);
    delete t;
    net.tc = nullptr;
}

// rebuild the bf16 shadows of one gradient bucket's weights: 0 = fc2 (w2t, w2b), 1 = fc1 (w1t, w1b), 2 = conv images
int tc_refresh_bucket(Net &net, int bucket, cudaStream_t s)
{
    TcState *t = net.tc;
    if (bucket == 0) {
        transpose_to_act<false><<<dim3(FC2_OUT / 32, FC2_IN / 32), 256, 0, s>>>(net.params + OFF_F2W, t->w2t, FC2_IN, FC2_OUT);
        LAUNCH_CHECK(net);
        convert_rows_bf16<false><<<FC2_IN, 256, 0, s>>>(net.params + OFF_F2W, t->w2b, FC2_OUT);
        LAUNCH_CHECK(net);
    } else if (bucket == 1) {
        transpose_to_act<true><<<dim3(FC1_OUT / 32, FC1_IN / 32), 256, 0, s>>>(net.params + OFF_F1W, t->w1t, FC1_IN, FC1_OUT);
        LAUNCH_CHECK(net);
        convert_rows_bf16<true><<<FC1_IN, 256, 0, s>>>(net.params + OFF_F1W, t->w1b, FC1_OUT);
        LAUNCH_CHECK(net);
    } else {
        if (int rc = tc_conv_refresh(net, s)) return rc;   // b1/b2 images and w2kt (the dL/dcol GEMM's B operand)
    }
    return 0;
}

// SGD + shadow refresh of FC bucket 0 (fc2) or 1 (fc1) in one pass over the weights; the bias follows as a plain SGD range
int tc_sgd_refresh_fc(Net &net, int bucket, float alpha, cudaStream_t s)
{
    TcState *t = net.tc;
    if (bucket == 0) {
        sgd_refresh_fc<false><<<dim3(FC2_OUT / 32, FC2_IN / 32), 256, 0, s>>>(net.params + OFF_F2W, net.grads + OFF_F2W, alpha, t->w2t, t->w2b, FC2_IN, FC2_OUT);
        LAUNCH_CHECK(net);
        return sgd_apply_range(net, alpha, OFF_F2B, N_PARAMS - OFF_F2B, s);
    }
    sgd_refresh_fc<true><<<dim3(FC1_OUT / 32, FC1_IN / 32), 256, 0, s>>>(net.params + OFF_F1W, net.grads + OFF_F1W, alpha, t->w1t, t->w1b, FC1_IN, FC1_OUT);
    LAUNCH_CHECK(net);
    return sgd_apply_range(net, alpha, OFF_F1B, OFF_F2W - OFF_F1B, s);
}

int tc_refresh_weights(Net &net, cudaStream_t s)
{
    for (int b = 0; b < 3; b++)
        if (int rc = tc_refresh_bucket(net, b, s)) return rc;
    net.tc_dirty = false;
    return 0;
}

static int tc_ensure(Net &net, int64_t n)
{
    TcState *t = net.tc;
    if (n <= t->cap) return 0;
    int64_t cap = (n + BM - 1) / BM * BM;
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (t->p2) cudaFree(t->p2);
    if (t->h1) cudaFree(t->h1);
    t->p2 = t->h1 = nullptr;
    HP_CUDA_TRY(cudaMalloc((void **)&t->p2, (size_t)cap * FC1_IN * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->h1, (size_t)cap * FC1_OUT * 2));
    // rows past n are never stored by the epilogues, but they are loaded by TMA: keep them finite
    HP_CUDA_TRY(cudaMemset(t->p2, 0, (size_t)cap * FC1_IN * 2));
    HP_CUDA_TRY(cudaMemset(t->h1, 0, (size_t)cap * FC1_OUT * 2));
    if (int rc = make_map_f16(&t->tm_p2, t->p2, cap, FC1_IN, BM)) return rc;
    if (int rc = make_map_f16(&t->tm_h1, t->h1, cap, FC1_OUT, BM)) return rc;
    t->cap = cap;
    net.alloc_epoch++;
    return 0;
}

// x16 != nullptr: the crops arrive as 16-bit depth and include/handtrack.h:700 runs inside the conv kernel's loader
// dec_out != nullptr: the decode runs in the fc2 epilogue (y_out may then be null: the 9.2 KB per crop never reach HBM)
static int tc_forward_impl(Net &net, const float *x, const uint16_t *x16, float depth_scale, float dmin, float dmax, int64_t n, float *y_out,
                           float *dec_out, cudaStream_t s)
{
    TcState *t = net.tc;
    if (dec_out && !t->dec_scratch) HP_CUDA_TRY(cudaMalloc((void **)&t->dec_scratch, (size_t)t->total_sms * BM * 256 * sizeof(float)));
    for (int64_t b = 0; b < n; b += TC_CHUNK) {
        const int64_t m = (n - b < TC_CHUNK) ? n - b : TC_CHUNK;
        if (int rc = tc_ensure(net, m)) return rc;
        {
            StageTimer st(net, 0, s);
            if (x16) {
                if (int rc = tc_conv2_stage_u16(net, x16 + b * N_IN, m, depth_scale, dmin, dmax, t->p2, s)) return rc;
            } else if (int rc = tc_conv_stage(net, x + b * N_IN, m, t->p2, s)) return rc;
        }
        const int m_tiles = (int)((m + BM - 1) / BM);
        {
            StageTimer st(net, 1, s);
            const int tiles = m_tiles * (FC1_OUT / 256);
            const int grid = tiles < t->num_sms ? tiles : t->num_sms;
            EpiArgs ea{net.params + OFF_F1B, t->h1, nullptr, nullptr, 0};
            tc_gemm_kernel<TC_EPI_TANH_ACT, 256, true><<<grid, GEMM_THREADS, gemm_smem(256), s>>>(t->tm_p2, t->tm_w1t, ea, (int)m, FC1_OUT, FC1_IN);
            LAUNCH_CHECK(net);
        }
        {
            StageTimer st(net, 2, s);
            const int tiles = m_tiles * (FC2_OUT / 256);
            const int grid = tiles < t->num_sms ? tiles : t->num_sms;
            EpiArgs ea{net.params + OFF_F2B, y_out ? y_out + b * N_OUT : nullptr, nullptr, nullptr, 0};
            if (dec_out) {
                ea.decoded = dec_out + b * 48;
                ea.scratch = t->dec_scratch;
                tc_gemm_kernel<TC_EPI_SOFTMAX_DECODE, 256, true><<<grid, GEMM_THREADS, gemm_smem(256), s>>>(t->tm_h1, t->tm_w2t, ea, (int)m, FC2_OUT, FC2_IN);
            } else {
                tc_gemm_kernel<TC_EPI_SOFTMAX_F32, 256, true><<<grid, GEMM_THREADS, gemm_smem(256), s>>>(t->tm_h1, t->tm_w2t, ea, (int)m, FC2_OUT, FC2_IN);
            }
            LAUNCH_CHECK(net);
        }
    }
    return 0;
}

int tc_forward(Net &net, const float *x, int64_t n, float *y_out, cudaStream_t s)
{
    return tc_forward_impl(net, x, nullptr, 0.f, 0.f, 1.f, n, y_out, nullptr, s);
}

int tc_forward_u16(Net &net, const uint16_t *depth, float depth_scale, float dmin, float dmax, int64_t n, float *y_out, cudaStream_t s)
{
    return tc_forward_impl(net, nullptr, depth, depth_scale, dmin, dmax, n, y_out, nullptr, s);
}

// Eval + CNNOutputAnalysis decode in one pass (decode inside the fc2 epilogue); x16 or x, y_out optional
int tc_forward_decode(Net &net, const float *x, const uint16_t *x16, float depth_scale, float dmin, float dmax, int64_t n, float *y_out, float *dec_out,
                      cudaStream_t s)
{
    return tc_forward_impl(net, x, x16, depth_scale, dmin, dmax, n, y_out, dec_out, s);
}

template <int EPI, int BN, bool OPS_F16>
static int launch_gemm(Net &net, const CUtensorMap &tmA, const CUtensorMap &tmB, const EpiArgs &ea, int M, int N, int K, cudaStream_t s)
{
    TcState *t = net.tc;
    const int tiles = ((M + BM - 1) / BM) * (N / BN) * (ea.ksplit > 1 ? ea.ksplit : 1);
    const int grid = tiles < t->num_sms ? tiles : t->num_sms;
    tc_gemm_kernel<EPI, BN, OPS_F16><<<grid, GEMM_THREADS, gemm_smem(BN), s>>>(tmA, tmB, ea, M, N, K);
    LAUNCH_CHECK(net);
    return 0;
}

// hp_fp32.cu: softmax + loss + softmax backward from logits, optionally also emitting dlogits as bf16
int fp32_softmax_loss(Net &net, const float *logits, float *y, const float *t, float *dlog, __nv_bfloat16 *dlog_bf, float *mse, int64_t n, cudaStream_t s,
                      const float *logits2 = nullptr);

// hp_fp32.cu
int fp32_conv_stage(Net &net, const float *x, int64_t n, __nv_bfloat16 *p2_bf, cudaStream_t s);
int fp32_conv_backward(Net &net, const float *x, int64_t n, const float *g2_hwc, bool accumulate, cudaStream_t s);
int tc_conv_backward(Net &net, const float *x, int64_t n, const float *g2_hwc, bool accumulate, cudaStream_t s);
int fp32_colsum(Net &net, const float *in, int64_t R, int ncols, float *dst, bool accumulate, cudaStream_t s);

constexpr int64_t TRAIN_CAP = 2048;  // samples per pass (bounded by the FP32 conv-stage workspace)

static int tc_train_ensure(Net &net)
{
    TcState *t = net.tc;
    if (t->dlog_bf) return 0;
    const int64_t cap = TRAIN_CAP;
    HP_CUDA_TRY(cudaMalloc((void **)&t->dlog_bf, (size_t)cap * N_OUT * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->da1_bf, (size_t)cap * FC1_OUT * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->g2_sink, (size_t)cap * FC1_IN * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->h1T, (size_t)FC1_OUT * cap * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->dlogT, (size_t)N_OUT * cap * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->p2T, (size_t)FC1_IN * cap * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->da1T, (size_t)FC1_OUT * cap * 2));
    HP_CUDA_TRY(cudaMemset(t->dlog_bf, 0, (size_t)cap * N_OUT * 2));
    HP_CUDA_TRY(cudaMemset(t->da1_bf, 0, (size_t)cap * FC1_OUT * 2));
    HP_CUDA_TRY(cudaMemset(t->h1T, 0, (size_t)FC1_OUT * cap * 2));
    HP_CUDA_TRY(cudaMemset(t->dlogT, 0, (size_t)N_OUT * cap * 2));
    HP_CUDA_TRY(cudaMemset(t->p2T, 0, (size_t)FC1_IN * cap * 2));
    HP_CUDA_TRY(cudaMemset(t->da1T, 0, (size_t)FC1_OUT * cap * 2));
    if (int rc = make_map_bf16(&t->tm_dlog, t->dlog_bf, cap, N_OUT, BM)) return rc;
    if (int rc = make_map_bf16(&t->tm_da1, t->da1_bf, cap, FC1_OUT, BM)) return rc;
    if (int rc = make_map_bf16(&t->tm_h1T, t->h1T, FC1_OUT, cap, BM)) return rc;
    if (int rc = make_map_bf16(&t->tm_p2T, t->p2T, FC1_IN, cap, BM)) return rc;
    if (int rc = make_map_bf16(&t->tm_dlogT, t->dlogT, N_OUT, cap, 256)) return rc;
    if (int rc = make_map_bf16(&t->tm_da1T, t->da1T, FC1_OUT, cap, 256)) return rc;
    if (int rc = make_map_bf16(&t->tm_dlogT128, t->dlogT, N_OUT, cap, 128)) return rc;
    if (int rc = make_map_bf16(&t->tm_da1T128, t->da1T, FC1_OUT, cap, 128)) return rc;
    // conv2 backward operands (conv2_bwd_operands)
    const int64_t ld = cap * C2_POS;
    HP_CUDA_TRY(cudaMalloc((void **)&t->e2, (size_t)ld * C2_CO * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->e2T, (size_t)C2_CO * ld * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->colT, (size_t)C2_KDIM * ld * 2));
    HP_CUDA_TRY(cudaMalloc((void **)&t->db2_partial, (size_t)cap * C2_CO * 4));
    HP_CUDA_TRY(cudaMemset(t->e2, 0, (size_t)ld * C2_CO * 2));
    HP_CUDA_TRY(cudaMemset(t->e2T, 0, (size_t)C2_CO * ld * 2));
    HP_CUDA_TRY(cudaMemset(t->colT, 0, (size_t)C2_KDIM * ld * 2));
    if (int rc = make_map_bf16(&t->tm_e2, t->e2, ld, C2_CO, BM)) return rc;
    if (int rc = make_map_bf16(&t->tm_e2T, t->e2T, C2_CO, ld, 64)) return rc;
    if (int rc = make_map_bf16(&t->tm_colT, t->colT, C2_KDIM, ld, BM)) return rc;
    net.alloc_epoch++;
    return 0;
}

// hp_fp32.cu kernels reused by the tensor-core conv backward
int fp32_reduce_warp(Net &net, float *dst, const float *partial, int S, int len, bool accumulate, cudaStream_t s);
int fp32_conv1_wgrad(Net &net, const float *x, int64_t n, bool accumulate, cudaStream_t s);

// conv stages backward of the tensor path: conv2 dL/dp1 and dW2 as tcgen05 GEMMs (see conv2_bwd_operands), conv2 dB as a
// column sum, conv1 dW/dB winners-only on FFMA (its contraction has 25 taps x 1 input channel: no GEMM shape worth the name).
// Small batches: the bias-gradient reductions (2-4 us each, no consumer before the bucket is handed to the update) leave
// the GEMM chain for a side stream: fork after their input exists, join before the bucket's event.  Only the scratch-free
// reductions qualify (colsum_direct, reduce_partials_warp on its own partial buffer).
static inline bool side_reductions(int64_t n) { return n <= 512; }
static int side_fork(Net &net, int i, cudaStream_t s)
{
    HP_CUDA_TRY(cudaEventRecord(net.ev_fork[i], s));
    HP_CUDA_TRY(cudaStreamWaitEvent(net.aux_stream, net.ev_fork[i], 0));
    return 0;
}
static int side_done(Net &net, int i)
{
    HP_CUDA_TRY(cudaEventRecord(net.ev_join[i], net.aux_stream));
    return 0;
}
static int side_join(Net &net, int i, cudaStream_t s)
{
    HP_CUDA_TRY(cudaStreamWaitEvent(s, net.ev_join[i], 0));
    return 0;
}

static int tc_conv_backward_gemm(Net &net, const float *x, int64_t n, const float *g2_hwc, bool accumulate, cudaStream_t s, const float *g2_hwc_b = nullptr)
{
    TcState *t = net.tc;
    Workspace &w = net.ws;
    float *G = net.grads;
    const int64_t ld = TRAIN_CAP * C2_POS;
    const int R = (int)(n * C2_POS);
    conv2_bwd_operands<<<(unsigned)n, 256, 0, s>>>(g2_hwc, w.idx2, w.p1, t->e2, t->e2T, t->colT, t->db2_partial, (int)n, ld, g2_hwc_b);
    LAUNCH_CHECK(net);
    const bool side = side_reductions(n);
    if (side) {
        if (int rc = side_fork(net, 2, s)) return rc;
        if (int rc = fp32_reduce_warp(net, G + OFF_C2B, t->db2_partial, (int)n, C2_CO, accumulate, net.aux_stream)) return rc;
        if (int rc = side_done(net, 2)) return rc;
    } else if (int rc = fp32_reduce_warp(net, G + OFF_C2B, t->db2_partial, (int)n, C2_CO, accumulate, s)) return rc;
    // dW2^T partials: split-K over about half the SMs' worth of ranges x 2 M tiles
    const int kb_total = (R + BK - 1) / BK;
    int target = t->num_sms / 2;
    if (target < 1) target = 1;
    const int kb_per = (kb_total + target - 1) / target;
    const int S = (kb_total + kb_per - 1) / kb_per;
    if ((size_t)S * C2_KDIM * C2_CO > w.partial_floats) { set_error("partial buffer too small for %d K ranges", S); return 1; }
    if (int rc = launch_gemm<TC_EPI_STORE_F32, 64, false>(net, t->tm_colT, t->tm_e2T, EpiArgs{nullptr, w.partial, nullptr, nullptr, 0, S}, C2_KDIM, C2_CO, R, s)) return rc;
    reduce_c2w_t<<<C2_KDIM, 256, 0, s>>>(G + OFF_C2W, w.partial, S, accumulate ? 1 : 0);
    LAUNCH_CHECK(net);
    // dL/dcol, then col2im fused with the conv1-stage tanh'
    __nv_bfloat16 *dcol = reinterpret_cast<__nv_bfloat16 *>(w.colgrad);   // the FP32 path's [n*144][256] fp32 buffer, half used
    if (int rc = launch_gemm<TC_EPI_STORE_BF16, 256, false>(net, t->tm_e2, t->tm_w2kt, EpiArgs{nullptr, dcol, nullptr, nullptr, 0, 0}, R, C2_KDIM, C2_CO, s)) return rc;
    col2im_g1_vec<<<(unsigned)n, 256, 0, s>>>(dcol, w.p1, w.g1);
    LAUNCH_CHECK(net);
    if (int rc = fp32_conv1_wgrad(net, x, n, accumulate, s)) return rc;
    if (side) return side_join(net, 2, s);
    return 0;
}

// Forward + backward of one pass (n <= TRAIN_CAP) with every FC contraction on tcgen05:
//   forward   conv stages (tensor, emitting p1 and the pool winners) -> fc1 -> fc2 + softmax (tensor)
//   backward  loss + softmax' -> fc2 dW, dX*tanh' -> fc1 dW, dX*tanh' (tensor) -> conv2 / conv1 backward
//             (winners-only weight gradients and the dL/dp1 route on FFMA)
// Leaves sum_b g_b in net.grads (.cnnb order).  CNN::Train, cnn.h:558-575.
int tc_train_grad(Net &net, const float *x, const float *t_dev, int64_t n, float *mse, bool accumulate, cudaStream_t s)
{
    TcState *t = net.tc;
    Workspace &w = net.ws;
    float *G = net.grads;
    if (net.tc_dirty)
        if (int rc = tc_refresh_weights(net, s)) return rc;
    if (int rc = tc_ensure(net, n)) return rc;
    if (int rc = tc_train_ensure(net)) return rc;
    if (int rc = ensure_workspace(net, n)) return rc;
    const int M = (int)n;
    const int n_pad = (M + BK - 1) / BK * BK;
    const int flags = accumulate ? TC_FLAG_ACCUMULATE : 0;
    // ---- forward
    if (int rc = tc_conv_stage_train(net, x, n, t->p2, s)) return rc;   // p2 bf16 (HWC); p1, idx1, idx2 into the workspace
    // small batches: 64-wide N tiles so that the GEMMs spread over the SMs (256-wide tiles give 2 x 8 CTAs at n = 256)
    const bool narrow = ((M + BM - 1) / BM) * (FC1_OUT / 256) < 74;
    if (narrow) {
        if (int rc = launch_gemm<TC_EPI_TANH_ACT, 64, true>(net, t->tm_p2, t->tm_w1t64, EpiArgs{net.params + OFF_F1B, t->h1, nullptr, nullptr, 0}, M, FC1_OUT, FC1_IN, s)) return rc;
        // fc2 at 64-wide tiles is 2 x 36 CTAs, each bounded by what 192 KB of loads in flight deliver (profiles/r2_train_step.md);
        // when the SMs allow it the K range is cut in two (144 CTAs) and the loss kernel adds the two partial logits
        const int fc2_tiles = ((M + BM - 1) / BM) * (FC2_OUT / 64);
        static const bool no_split2 = getenv("HP_FC2_SPLITK") && getenv("HP_FC2_SPLITK")[0] == '0';   // A/B
        const bool split2 = !no_split2 && 2 * fc2_tiles <= t->num_sms && (size_t)2 * M * N_OUT <= w.partial_floats;
        if (split2) {
            if (int rc = launch_gemm<TC_EPI_STORE_F32, 64, true>(net, t->tm_h1, t->tm_w2t64, EpiArgs{net.params + OFF_F2B, w.partial, nullptr, nullptr, 0, 2}, M, FC2_OUT, FC2_IN, s)) return rc;
            if (int rc = fp32_softmax_loss(net, w.partial, w.y, t_dev, w.dlog, t->dlog_bf, mse, n, s, w.partial + (size_t)M * N_OUT)) return rc;
        } else {
        if (int rc = launch_gemm<TC_EPI_STORE_F32, 64, true>(net, t->tm_h1, t->tm_w2t64, EpiArgs{net.params + OFF_F2B, w.logits, nullptr, nullptr, 0}, M, FC2_OUT, FC2_IN, s)) return rc;
        if (int rc = fp32_softmax_loss(net, w.logits, w.y, t_dev, w.dlog, t->dlog_bf, mse, n, s)) return rc;
        }
    } else {
        if (int rc = launch_gemm<TC_EPI_TANH_ACT, 256, true>(net, t->tm_p2, t->tm_w1t, EpiArgs{net.params + OFF_F1B, t->h1, nullptr, nullptr, 0}, M, FC1_OUT, FC1_IN, s)) return rc;
        if (int rc = launch_gemm<TC_EPI_SOFTMAX_F32, 256, true>(net, t->tm_h1, t->tm_w2t, EpiArgs{net.params + OFF_F2B, w.y, nullptr, nullptr, 0}, M, FC2_OUT, FC2_IN, s)) return rc;
        loss_from_y<<<(unsigned)n, 256, 0, s>>>(w.y, t_dev, w.dlog, t->dlog_bf, mse);
        LAUNCH_CHECK(net);
    }
    // Small batches: the weight-gradient branch of each FC layer (bias column sums, operand transposes, dW GEMM -- nothing
    // downstream in backward reads it) leaves the chain for the side stream; the chain is then loss -> dX2 -> dX1 -> conv
    // backward.  The bucket events are recorded where the branch ends; conv backward joins the side stream at its end.
    const bool side = side_reductions(n);
    cudaStream_t wg = side ? net.aux_stream : s;
    if (side)
        if (int rc = side_fork(net, 0, s)) return rc;
    if (int rc = fp32_colsum(net, w.dlog, n, FC2_OUT, G + OFF_F2B, accumulate, wg)) return rc;
    transpose_bf16_pair<<<dim3(N_OUT / 32, (n_pad + 31) / 32, 2), 256, 0, wg>>>(TransposeJob{t->h1, t->h1T, FC1_OUT, 1}, TransposeJob{t->dlog_bf, t->dlogT, N_OUT, 0}, M,
                                                                                  n_pad, (int)TRAIN_CAP);
    LAUNCH_CHECK(net);
    // dW2[2048][2304] = h1^T * dlog
    // 16 x 9 = 144 tiles of 128x256 are one wave on 148 SMs but two on the 132 left when SMs are reserved for the
    // data-parallel exchange CTAs: 128-wide tiles (288 of them) then waste a fifth of a wave instead of most of one
    const bool half_tiles = t->num_sms < 144;
    if (half_tiles) {
        if (int rc = launch_gemm<TC_EPI_STORE_F32, 128, false>(net, t->tm_h1T, t->tm_dlogT128, EpiArgs{nullptr, G + OFF_F2W, nullptr, nullptr, flags}, FC2_IN, FC2_OUT, n_pad, wg)) return rc;
    } else
    if (int rc = launch_gemm<TC_EPI_STORE_F32, 256, false>(net, t->tm_h1T, t->tm_dlogT, EpiArgs{nullptr, G + OFF_F2W, nullptr, nullptr, flags}, FC2_IN, FC2_OUT, n_pad, wg)) return rc;
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[0], wg));
    // da1 = (dlog * W2^T) .* (1 - h1^2)
    if (narrow) {
        if (int rc = launch_gemm<TC_EPI_DTANH, 64, false>(net, t->tm_dlog, t->tm_w2b64, EpiArgs{nullptr, w.da1, t->da1_bf, t->h1, 0}, M, FC2_IN, FC2_OUT, s)) return rc;
    } else {
        if (int rc = launch_gemm<TC_EPI_DTANH, 256, false>(net, t->tm_dlog, t->tm_w2b, EpiArgs{nullptr, w.da1, t->da1_bf, t->h1, 0}, M, FC2_IN, FC2_OUT, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_dx[0], s));
    // ---- fc1
    if (side)
        if (int rc = side_fork(net, 1, s)) return rc;
    if (int rc = fp32_colsum(net, w.da1, n, FC1_OUT, G + OFF_F1B, accumulate, wg)) return rc;
    transpose_bf16_pair<<<dim3(FC1_IN / 32, (n_pad + 31) / 32, 2), 256, 0, wg>>>(TransposeJob{t->p2, t->p2T, FC1_IN, 1}, TransposeJob{t->da1_bf, t->da1T, FC1_OUT, 0}, M,
                                                                                   n_pad, (int)TRAIN_CAP);
    LAUNCH_CHECK(net);
    // dW1[k'][2048] = p2^T * da1, rows un-permuted from HWC to the reference's CHW flatten on store
    if (half_tiles) {
        if (int rc = launch_gemm<TC_EPI_STORE_F32, 128, false>(net, t->tm_p2T, t->tm_da1T128, EpiArgs{nullptr, G + OFF_F1W, nullptr, nullptr, flags | TC_FLAG_ROWS_HWC_TO_CHW},
                                                        FC1_IN, FC1_OUT, n_pad, wg)) return rc;
    } else
    if (int rc = launch_gemm<TC_EPI_STORE_F32, 256, false>(net, t->tm_p2T, t->tm_da1T, EpiArgs{nullptr, G + OFF_F1W, nullptr, nullptr, flags | TC_FLAG_ROWS_HWC_TO_CHW}, FC1_IN,
                                                    FC1_OUT, n_pad, wg)) return rc;
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[1], wg));
    // g2 = (da1 * W1^T) .* (1 - p2^2), columns in HWC order (the epilogue's bf16 copy is not needed: it goes to a sink of
    // its own -- dlog_bf, the sink before, may still be read by the side stream's transpose)
    // like fc2 forward: at 64-wide tiles this GEMM is 2 x 36 latency-bound CTAs; in two K ranges (144 CTAs) each writes
    // (1 - p2^2) x its partial product and conv2_bwd_operands, the only reader, adds the two
    static const bool ffma_conv_bwd = getenv("HP_CONV_BWD_FFMA") != nullptr;   // A/B switch for the previous FFMA kernels
    static const bool no_dx_split = getenv("HP_DX1_SPLITK") && getenv("HP_DX1_SPLITK")[0] == '0';   // A/B
    const int dx1_tiles = ((M + BM - 1) / BM) * (FC1_IN / 64);
    const bool dx1_split = narrow && !ffma_conv_bwd && !no_dx_split && 2 * dx1_tiles <= t->num_sms && (size_t)2 * M * FC1_IN <= w.partial_floats;
    const float *g2a = w.g2, *g2b = nullptr;
    if (dx1_split) {
        if (int rc = launch_gemm<TC_EPI_DTANH, 64, false>(net, t->tm_da1, t->tm_w1b64, EpiArgs{nullptr, w.partial, nullptr, t->p2, 0, 2}, M, FC1_IN, FC1_OUT, s)) return rc;
        g2a = w.partial;
        g2b = w.partial + (size_t)M * FC1_IN;
    } else if (narrow) {
        if (int rc = launch_gemm<TC_EPI_DTANH, 64, false>(net, t->tm_da1, t->tm_w1b64, EpiArgs{nullptr, w.g2, t->g2_sink, t->p2, 0}, M, FC1_IN, FC1_OUT, s)) return rc;
    } else {
        if (int rc = launch_gemm<TC_EPI_DTANH, 256, false>(net, t->tm_da1, t->tm_w1b, EpiArgs{nullptr, w.g2, t->g2_sink, t->p2, 0}, M, FC1_IN, FC1_OUT, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_dx[1], s));
    // ---- conv stages backward (winners-only weight gradients; FFMA)
    if (ffma_conv_bwd) {
        if (int rc = tc_conv_backward(net, x, n, w.g2, accumulate, s)) return rc;
        if (side) {   // the side stream's weight-gradient branch rejoins the chain here (the GEMM route joins it itself)
            if (int rc = side_done(net, 2)) return rc;
            if (int rc = side_join(net, 2, s)) return rc;
        }
    } else {
        if (int rc = tc_conv_backward_gemm(net, x, n, g2a, accumulate, s, g2b)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[2], s));
    return 0;
}

}  // namespace hp
