// hp_tc_conv.cu -- both convolution stages of handposedd on tcgen05 tensor cores, fused in one
// persistent warp-specialised kernel:  fp32 crop -> [conv1 5x5 + 2x(2x2 max-pool) + tanh] ->
// [conv2 4x4 + tanh + 2x2 max-pool] -> 2304 bf16 features (fc1's A operand).
// Reference layers: LConv::forward (cnn.h:205-257), LActivation<TanH> (cnn.h:460), LMaxPool::forward
// (cnn.h:141-148), instantiated at include/handtrack.h:108-114.
//
// Neither convolution materialises an im2col matrix.  Both A operands are *views* of one small
// shared-memory image expressed through un-swizzled K-major UMMA descriptors, whose address is
// linear in the row index when the 8-row stride (SBO) is 128 B or a multiple of the row pitch:
//
// conv1 as a "pooled-window GEMM".  Row m = pooled pixel (py,px) of the 15x15 grid that survives
//   the two 2x2 pools; K = the 8x8 input patch at stride 4 that pixel depends on; N = 16 window
//   positions x 16 channels (B is the 5x5 kernel embedded at offset (dy,dx) in the 8x8 patch).
//   The 4x4 max-pool then happens inside one thread's registers (all 16 positions of a pooled
//   pixel are columns of the same TMEM lane) and tanh is applied once per pooled value
//   (max and the monotone tanh commute in the forward pass).  A[(py,px')][(r,c)] =
//   img[4py+r][8px'+4e+c]: with the bf16 image stored twice (e = 0: as is, e = 1: shifted by 4
//   pixels) a core matrix is 8 consecutive px' (16 B apart), SBO = 4 image rows, LBO = 1 image row.
// conv2 as 16 shifted taps.  p1 is kept as two planes [pixel q][8 channels] (16 B per row); for
//   tap (ky,kx) the A operand is the same plane pair starting ky*15+kx rows further down, K = 16
//   input channels = one MMA.  Rows whose (y,x) fall outside the 12x12 valid outputs compute
//   garbage that is never read.
//
// Warp roles (640 threads, 1 CTA/SM, crops strided over the grid):
//   warp 0      loads the two 32 KB weight images with cp.async.bulk; allocates TMEM
//   warp 1      conv1 MMA issuer (12 MMAs M128 N128 K16 per crop: all-zero K steps of the embedded 5x5 kernels are skipped)
//   warp 2      conv2 MMA issuer (32 MMAs M128 N64 K16 per crop); two issuers because at 32-64 tensor
//               cycles per instruction a single issuing thread, not the tensor pipe, sets the pace
//   warps 4-11  epilogue 1 (two warpgroups, one per pooled-column parity): TMEM -> running max over window
//               positions -> +bias, tanh -> p1 planes (smem)
//   warps 12-15 epilogue 2: TMEM -> bf16 staging -> 2x2 max, +bias, tanh -> global features
//   warps 16-19 loader: fp32 crop from global -> two bf16 image copies in smem (double-buffered)
#include "hp_ptx.cuh"
#include "hp_tc.cuh"

#include <stdlib.h>

namespace hp {

#define LAUNCH_CHECK(net)                \
    do {                                 \
        (net).launches++;                \
        HP_CUDA_TRY(cudaGetLastError()); \
    } while (0)

namespace cv {
constexpr int THREADS = 640;
constexpr int IMG_COPY = 9216;                 // one bf16 image copy (8 KB) + slack for the pad rows' reads
constexpr int IMG_BUF = 2 * IMG_COPY;          // aligned copy + copy shifted by 4 pixels
constexpr int P1_ROWS = 304;                   // 225 pixels + tap-shift overhang of the second M tile
constexpr int P1_PLANE = P1_ROWS * 16;         // 8 channels x bf16 per row
constexpr int P1_BUF = 2 * P1_PLANE;
constexpr int S_ROWS = 192;                    // conv2 pre-activations of rows q <= 176, 64 ch bf16 = 128 B
constexpr int OFF_B1 = 0;                      // 32 KB, 1024-aligned (128B swizzle)
constexpr int OFF_B2 = 32768;                  // 32 KB
constexpr int OFF_IMG = 65536;                 // 2 x IMG_BUF
constexpr int OFF_P1 = OFF_IMG + 2 * IMG_BUF;  // 2 x P1_BUF
constexpr int OFF_S = OFF_P1 + 2 * P1_BUF;     // S_ROWS x 128
constexpr int OFF_BIAS = OFF_S + S_ROWS * 256; // 16 + 64 floats (S rows: 128 B bf16 for inference, 256 B fp32 for training)
constexpr int OFF_BAR = OFF_BIAS + 512;
constexpr int SMEM = OFF_BAR + 256 + 1024;
// TMEM columns
constexpr int ACC1 = 0;    // two 128-column conv1 accumulators (window-position halves)
constexpr int ACC2 = 256;  // two conv2 accumulator sets of 2 x 64 columns
}  // namespace cv

#ifdef HP_CONV_TRACE
__device__ long long g_conv_trace[64 * 1024];
#define TRACE(role, it, ev)                                                                       \
    do {                                                                                          \
        if (blockIdx.x == 0 && lane == 0 && (it) < 24) g_conv_trace[((role) * 24 + (it)) * 16 + (ev)] = clock64(); \
    } while (0)
#else
#define TRACE(role, it, ev) do {} while (0)
#endif

// TRAIN additionally emits what CNN::Train's backward needs (cnn.h:571-575): the pooled conv1 activations p1
// (fp32, the reference's CHW layout) and the max-pool winners of both stages (LMaxPool::backward, cnn.h:149-164:
// first strict maximum; the window positions are laid out in the hierarchical scan order of the two stacked pools).
template <bool TRAIN>
__global__ void __launch_bounds__(cv::THREADS, 1)
tc_conv_kernel(const float *__restrict__ x, const uint8_t *__restrict__ b1_img, const uint8_t *__restrict__ b2_img,
               const float *__restrict__ params, act_t *__restrict__ p2_out, int n, float *__restrict__ p1_out,
               uint8_t *__restrict__ idx1_out, uint8_t *__restrict__ idx2_out, int tanh_accurate)
{
    using namespace cv;
    const bool acc_tanh = tanh_accurate != 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    float *bias1 = reinterpret_cast<float *>(smem + OFF_BIAS);
    float *bias2 = bias1 + 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint64_t *wgt_full = bars + 0;
    uint64_t *img_full = bars + 1;    // [2]
    uint64_t *img_empty = bars + 3;   // [2]
    uint64_t *acc1_full = bars + 17;  // [4]: one per group g = e*2 + half (each completes once per crop, so the two
                                      //      epilogue-1 warpgroups can wait on plain per-crop parity)
    uint64_t *acc1_empty = bars + 7;  // [2]
    uint64_t *p1_full = bars + 9;     // [2]
    uint64_t *p1_empty = bars + 11;   // [2]
    uint64_t *acc2_full = bars + 13;  // [2]
    uint64_t *acc2_empty = bars + 15; // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_crops = (n > (int)blockIdx.x) ? (n - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        ptx::mbar_init(wgt_full, 1);
        for (int b = 0; b < 2; b++) {
            ptx::mbar_init(&img_full[b], 128);
            ptx::mbar_init(&img_empty[b], 1);
            ptx::mbar_init(&acc1_full[b], 1);
            ptx::mbar_init(&acc1_full[2 + b], 1);
            ptx::mbar_init(&acc1_empty[b], 4);
            ptx::mbar_init(&p1_full[b], 8);
            ptx::mbar_init(&p1_empty[b], 1);
            ptx::mbar_init(&acc2_full[b], 1);
            ptx::mbar_init(&acc2_empty[b], 4);
        }
        ptx::fence_barrier_init();
    }
    // zero the p1 planes once: the overhang rows (225..303) are read by the second conv2 M tile
    for (int i = threadIdx.x; i < 2 * P1_BUF / 16; i += THREADS) reinterpret_cast<uint4 *>(smem + OFF_P1)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * IMG_BUF / 16; i += THREADS) reinterpret_cast<uint4 *>(smem + OFF_IMG)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 16) bias1[threadIdx.x] = params[OFF_C1B + threadIdx.x];
    if (threadIdx.x >= 64 && threadIdx.x < 128) bias2[threadIdx.x - 64] = params[OFF_C2B + threadIdx.x - 64];
    if (warp == 0) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(wgt_full, 65536);
            ptx::bulk_load_1d(smem + OFF_B1, b1_img, 32768, wgt_full);
            ptx::bulk_load_1d(smem + OFF_B2, b2_img, 32768, wgt_full);
        }
    } else if (warp == 1) {
        // ===================== conv1 MMA issuer =====================
        // The whole warp runs the (warp-uniform) control flow so that descriptors stay in uniform
        // registers; one elected lane issues.  16 MMAs per crop, fully unrolled.
        constexpr uint32_t idesc1 = ptx::make_idesc_f16(128, 128);
        ptx::mbar_wait(wgt_full, 0);
        const uint32_t sB1 = ptx::smem_u32(smem + OFF_B1);
        const uint64_t bd0 = ptx::make_desc_sw128(sB1);
        for (int it = 0; it < my_crops; it++) {
            const int ib = it & 1;
            TRACE(0, it, 0);
            ptx::mbar_wait(&img_full[ib], (it >> 1) & 1);
            ptx::tc_fence_after();
            TRACE(0, it, 1);
            const uint64_t ad0 = ptx::make_desc_nosw(ptx::smem_u32(smem + OFF_IMG + ib * IMG_BUF), 128, 512);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int e = g >> 1, half = g & 1;        // e: pooled-column parity (which image copy)
                const uint32_t u = (uint32_t)(it * 2 + e);  // use count of accumulator `half`
                ptx::mbar_wait(&acc1_empty[half], (u & 1) ^ 1);
                ptx::tc_fence_after();
                TRACE(0, it, 2 + 2 * g);
                if (ptx::elect_one()) {
                    const uint32_t d = tmem_base + ACC1 + half * 128;
                    const uint64_t ad = ad0 + ((e * IMG_COPY) >> 4), bd = bd0 + ((half * 16384) >> 4);
                    // K step ks covers patch rows 2ks, 2ks+1.  Half 0 holds the window positions with dy in {0,1}, whose
                    // 5x5 kernels touch patch rows 0..5 only; half 1 (dy in {2,3}) touches rows 2..7: in each half one
                    // of the four K steps multiplies by all-zero weights and is skipped (12 MMAs per crop instead of 16).
                    if (half == 0) {
                        ptx::umma_f16_c<false>(d, ad, bd, idesc1);
                        ptx::umma_f16_c<true>(d, ad + (256 >> 4), bd + 2, idesc1);
                        ptx::umma_f16_c<true>(d, ad + (512 >> 4), bd + 4, idesc1);
                    } else {
                        ptx::umma_f16_c<false>(d, ad + (256 >> 4), bd + 2, idesc1);
                        ptx::umma_f16_c<true>(d, ad + (512 >> 4), bd + 4, idesc1);
                        ptx::umma_f16_c<true>(d, ad + (768 >> 4), bd + 6, idesc1);
                    }
                    ptx::umma_commit(&acc1_full[g]);
                    if (g == 3) ptx::umma_commit(&img_empty[ib]);
                }
                __syncwarp();
                TRACE(0, it, 3 + 2 * g);
            }
        }
    } else if (warp == 2) {
        // ===================== conv2 MMA issuer: 2 M tiles x 16 taps per crop =====================
        constexpr uint32_t idesc2 = ptx::make_idesc_f16(128, 64), idesc2_m64 = ptx::make_idesc_f16(64, 64);
        ptx::mbar_wait(wgt_full, 0);
        const uint64_t bd0 = ptx::make_desc_nosw(ptx::smem_u32(smem + OFF_B2), 1024, 128);
        for (int it = 0; it < my_crops; it++) {
            const int pb = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            TRACE(1, it, 0);
            ptx::mbar_wait(&p1_full[pb], ph);
            TRACE(1, it, 1);
            ptx::mbar_wait(&acc2_empty[pb], ph ^ 1);
            ptx::tc_fence_after();
            TRACE(1, it, 2);
            if (ptx::elect_one()) {
                const uint64_t ad0 = ptx::make_desc_nosw(ptx::smem_u32(smem + OFF_P1 + pb * P1_BUF), P1_PLANE, 128);
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    const uint32_t d = tmem_base + ACC2 + pb * 128 + mt * 64;
                    // rows 0..127 as one M=128 tile; the valid outputs end at row 176, so the second tile is M=64
                    // (rows 128..191): same tensor time, but half the A bytes through the shared-memory port,
                    // which is what paces these N=64 MMAs
                    const uint32_t idesc = mt == 0 ? idesc2 : idesc2_m64;
#pragma unroll
                    for (int tap = 0; tap < 16; tap++) {
                        const int shift = (tap >> 2) * 15 + (tap & 3);   // rows: ky*15 + kx
                        const uint64_t ad = ad0 + (mt * 128 + shift), bd = bd0 + tap * (2048 >> 4);
                        if (tap == 0) ptx::umma_f16_c<false>(d, ad, bd, idesc);
                        else ptx::umma_f16_c<true>(d, ad, bd, idesc);
                    }
                }
                ptx::umma_commit(&acc2_full[pb]);
                ptx::umma_commit(&p1_empty[pb]);
            }
            __syncwarp();
            TRACE(1, it, 3);
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== epilogue 1: conv1 accumulators -> p1 planes =====================
        // Two warpgroups: warps 4-7 drain the even-column tile (e = 0) of every crop, warps 8-11 the odd-column tile
        // (e = 1).  One warpgroup alone needs ~2 k cycles per crop for its four accumulator reads and two tanh/pack
        // passes, which -- not the tensor pipe -- set the pace of the whole kernel.
        const int ew = warp & 3;
        const int my_e = (warp - 4) >> 2;
        const int m = ew * 32 + lane;           // row of the M tile: (py, px')
        const int py = m >> 3, pxh = m & 7;
        for (int it = 0; it < my_crops; it++) {
            const int pb = it & 1;
            uint8_t *planes = smem + OFF_P1 + pb * P1_BUF;
            if (warp == 4) TRACE(2, it, 0);
            ptx::mbar_wait(&p1_empty[pb], ((it >> 1) & 1) ^ 1);
            if (warp == 4) TRACE(2, it, 1);
            {
                const int e = my_e;
                float mx[16];
                int am[16];
                const uint32_t u = (uint32_t)(it * 2 + e);
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    ptx::mbar_wait(&acc1_full[e * 2 + half], it & 1);
                    ptx::tc_fence_after();
                    if (ew == 0 || e == 1) TRACE(e ? 5 + ew : 2, it, 2 + 2 * (e * 2 + half));
                    const uint32_t ta = tmem_base + ((uint32_t)(ew * 32) << 16) + ACC1 + half * 128;
#pragma unroll
                    for (int c = 0; c < 4; c++) {   // 32 columns = 2 window positions x 16 channels
                        uint32_t r[32];
                        ptx::tmem_ld32(ta + c * 32, r);
                        ptx::tmem_ld_wait();
                        if (TRAIN) {
                            const int p0 = half * 8 + 2 * c;
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[16 + j]);
                                if ((half == 0 && c == 0) || v0 > mx[j]) { mx[j] = v0; am[j] = p0; }
                                if (v1 > mx[j]) { mx[j] = v1; am[j] = p0 + 1; }
                            }
                        } else if (half == 0 && c == 0) {
#pragma unroll
                            for (int j = 0; j < 16; j++) mx[j] = fmaxf(__uint_as_float(r[j]), __uint_as_float(r[16 + j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j++) mx[j] = ptx::max3(mx[j], __uint_as_float(r[j]), __uint_as_float(r[16 + j]));
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&acc1_empty[half]);
                    if (ew == 0 || e == 1) TRACE(e ? 5 + ew : 2, it, 3 + 2 * (e * 2 + half));
                }
                const int px = 2 * pxh + e;
                if (py < 15 && px < 15) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        pk[j] = pack_act(tanh_conv(mx[2 * j] + bias1[2 * j], acc_tanh), tanh_conv(mx[2 * j + 1] + bias1[2 * j + 1], acc_tanh));
                    const int q = py * 15 + px;
                    *reinterpret_cast<uint4 *>(planes + q * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4 *>(planes + P1_PLANE + q * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    if (TRAIN) {
                        const int64_t crop = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            // the bf16-rounded value conv2 actually consumed, in the reference's [c][y][x] layout
                            const __half2 h = *reinterpret_cast<const __half2 *>(&pk[j >> 1]);
                            p1_out[crop * P1_N + j * 225 + q] = (j & 1) ? __high2float(h) : __low2float(h);
                            const int blk = am[j] >> 2, sub = am[j] & 3;   // hierarchical position -> (dy, dx)
                            const int dy = 2 * (blk >> 1) + (sub >> 1), dx = 2 * (blk & 1) + (sub & 1);
                            idx1_out[crop * P1_N + j * 225 + q] = (uint8_t)(dy * 4 + dx);
                        }
                    }
                }
            }
            ptx::fence_proxy_async();   // generic-proxy stores -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p1_full[pb]);
            if (ew == 0 || my_e == 1) TRACE(my_e ? 5 + ew : 2, it, 10);
        }
    } else if (warp >= 12 && warp < 16) {
        // ===================== epilogue 2: conv2 accumulators -> pooled features =====================
        const int ew = warp - 12;
        const int t128 = ew * 32 + lane;
        uint8_t *S = smem + OFF_S;
        for (int it = 0; it < my_crops; it++) {
            const int pb = it & 1;
            const int64_t crop = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            if (ew == 0) TRACE(3, it, 0);
            ptx::mbar_wait(&acc2_full[pb], (it >> 1) & 1);
            ptx::tc_fence_after();
            if (ew == 0) TRACE(3, it, 1);
#pragma unroll 1
            for (int mt = 0; mt < 2; mt++) {
                // M=128 tile: accumulator row i sits in TMEM lane i.  M=64 tile: row r sits in lane (r % 16) + 32 * (r / 16),
                // i.e. the lower half of each warp's lane quarter.
                const int q = mt == 0 ? t128 : 128 + ew * 16 + lane;
                const int yy = q / 15, xx = q - yy * 15;
                const bool valid = (yy < 12) && (xx < 12) && (mt == 0 || lane < 16);
                const uint32_t ta = tmem_base + ((uint32_t)(ew * 32) << 16) + ACC2 + pb * 128 + mt * 64;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t r[32];
                    ptx::tmem_ld32(ta + c * 32, r);
                    ptx::tmem_ld_wait();
                    if (valid && !TRAIN) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {   // 16-byte chunk index c*4+k, swizzled by the row to avoid bank conflicts
                            const int chunk = c * 4 + k;
                            *reinterpret_cast<uint4 *>(S + q * 128 + ((chunk ^ (q & 7)) << 4)) =
                                make_uint4(pack_act(__uint_as_float(r[8 * k + 0]), __uint_as_float(r[8 * k + 1])),
                                           pack_act(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3])),
                                           pack_act(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5])),
                                           pack_act(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7])));
                        }
                    }
                    if (valid && TRAIN) {   // fp32 staging: the pool winners are decided on unrounded pre-activations
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int chunk = c * 8 + k;   // 16 chunks of 4 channels per 256-byte row
                            *reinterpret_cast<uint4 *>(S + q * 256 + ((chunk ^ (q & 7)) << 4)) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc2_empty[pb]);
            if (ew == 0) TRACE(3, it, 2);
            ptx::named_bar_sync(1, 128);
            if (!TRAIN) {
                // 36 pooled pixels x 8 chunks of 8 channels
                for (int item = t128; item < 288; item += 128) {
                    const int pp = item >> 3, chunk = item & 7;
                    const int py = pp / 6, px = pp - py * 6;
                    const int q0 = (2 * py) * 15 + 2 * px;
                    auto ld = [&](int q) { return *reinterpret_cast<const uint4 *>(S + q * 128 + ((chunk ^ (q & 7)) << 4)); };
                    const uint4 a = ld(q0), b = ld(q0 + 1), c = ld(q0 + 15), d = ld(q0 + 16);
                    uint32_t o[4];
                    const uint32_t *pa = &a.x, *pb2 = &b.x, *pc = &c.x, *pd = &d.x;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const __half2 m01 = __hmax2(*reinterpret_cast<const __half2 *>(pa + k), *reinterpret_cast<const __half2 *>(pb2 + k));
                        const __half2 m23 = __hmax2(*reinterpret_cast<const __half2 *>(pc + k), *reinterpret_cast<const __half2 *>(pd + k));
                        const float2 f = __half22float2(__hmax2(m01, m23));
                        const int co = chunk * 8 + 2 * k;
                        o[k] = pack_act(tanh_conv(f.x + bias2[co], acc_tanh), tanh_conv(f.y + bias2[co + 1], acc_tanh));
                    }
                    *reinterpret_cast<uint4 *>(p2_out + crop * P2_N + pp * 64 + chunk * 8) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            } else {
                // 36 pooled pixels x 16 chunks of 4 channels; first strict maximum in scan order (0,0),(1,0),(0,1),(1,1)
                for (int item = t128; item < 576; item += 128) {
                    const int pp = item >> 4, chunk = item & 15;
                    const int py = pp / 6, px = pp - py * 6;
                    const int q0 = (2 * py) * 15 + 2 * px;
                    auto ld = [&](int q) { return *reinterpret_cast<const float4 *>(S + q * 256 + ((chunk ^ (q & 7)) << 4)); };
                    const float4 a = ld(q0), b = ld(q0 + 1), c = ld(q0 + 15), d = ld(q0 + 16);
                    const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w}, vc[4] = {c.x, c.y, c.z, c.w}, vd[4] = {d.x, d.y, d.z, d.w};
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float m = va[k];
                        int arg = 0;
                        if (vb[k] > m) { m = vb[k]; arg = 1; }
                        if (vc[k] > m) { m = vc[k]; arg = 2; }
                        if (vd[k] > m) { m = vd[k]; arg = 3; }
                        const int co = chunk * 4 + k;
                        o[k] = tanh_conv(m + bias2[co], acc_tanh);
                        idx2_out[crop * P2_N + co * 36 + pp] = (uint8_t)arg;
                    }
                    *reinterpret_cast<uint2 *>(p2_out + crop * P2_N + pp * 64 + chunk * 4) = make_uint2(pack_act(o[0], o[1]), pack_act(o[2], o[3]));
                }
            }
            ptx::named_bar_sync(1, 128);
            if (ew == 0) TRACE(3, it, 3);
        }
    } else if (warp >= 16) {
        // ===================== loader: fp32 crop -> two bf16 image copies =====================
        const int t = threadIdx.x - 16 * 32;  // 0..127
        for (int it = 0; it < my_crops; it++) {
            const int ib = it & 1;
            const int64_t crop = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const float4 *src = reinterpret_cast<const float4 *>(x + crop * N_IN);
            float4 v[4][3];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = t + 128 * k;  // 8-pixel chunk index
                v[k][0] = __ldg(src + 2 * j);
                v[k][1] = __ldg(src + 2 * j + 1);
                v[k][2] = (j < 511) ? __ldg(src + 2 * j + 2) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (warp == 16) TRACE(4, it, 0);
            ptx::mbar_wait(&img_empty[ib], ((it >> 1) & 1) ^ 1);
            if (warp == 16) TRACE(4, it, 1);
            uint8_t *img = smem + OFF_IMG + ib * IMG_BUF;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = t + 128 * k;
                const uint32_t lo0 = pack_act(v[k][0].x, v[k][0].y), lo1 = pack_act(v[k][0].z, v[k][0].w);
                const uint32_t mi0 = pack_act(v[k][1].x, v[k][1].y), mi1 = pack_act(v[k][1].z, v[k][1].w);
                const uint32_t hi0 = pack_act(v[k][2].x, v[k][2].y), hi1 = pack_act(v[k][2].z, v[k][2].w);
                *reinterpret_cast<uint4 *>(img + j * 16) = make_uint4(lo0, lo1, mi0, mi1);             // pixels 8j .. 8j+7
                *reinterpret_cast<uint4 *>(img + IMG_COPY + j * 16) = make_uint4(mi0, mi1, hi0, hi1);  // pixels 8j+4 .. 8j+11
            }
            ptx::fence_proxy_async();
            ptx::mbar_arrive(&img_full[ib]);
            if (warp == 16) TRACE(4, it, 2);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// Build the two pre-laid-out weight images from the fp32 master weights.
//  b1: row nrow = pos*16 + co (pos = hierarchical pool-scan index of window offset (dy,dx)), 64 k = (r,c) of the 8x8 patch; value = w1[co][r-dy][c-dx]
//      inside the 5x5 support, else 0; 128-byte rows, 16-byte chunk r stored at chunk (r ^ (nrow & 7)).
//  b2: [tap][kchunk][co][8 ci] bf16 (16-byte rows): the un-swizzled K-major core-matrix order.
__device__ __forceinline__ void build_conv_image_elem(int i, const float *__restrict__ params, uint8_t *__restrict__ b1, uint8_t *__restrict__ b2,
                                                      __nv_bfloat16 *__restrict__ w2kt)
{
    if (i < 256 * 64) {
        const int nrow = i >> 6, k = i & 63;
        const int pos = nrow >> 4, co = nrow & 15;
        // window positions in the scan order of two stacked 2x2 pools: outer block (by,bx), inner offset (iy,ix)
        const int blk = pos >> 2, sub = pos & 3, dy = 2 * (blk >> 1) + (sub >> 1), dx = 2 * (blk & 1) + (sub & 1);
        const int r = k >> 3, c = k & 7;
        const int ky = r - dy, kx = c - dx;
        const float v = (ky >= 0 && ky < 5 && kx >= 0 && kx < 5) ? params[OFF_C1W + co * 25 + ky * 5 + kx] : 0.f;
        reinterpret_cast<__half *>(b1 + nrow * 128 + ((r ^ (nrow & 7)) << 4))[c] = __float2half_rn(v);
    }
    if (i < 16 * 2 * 64 * 8) {
        const int ci8 = i & 7, co = (i >> 3) & 63, kc = (i >> 9) & 1, tap = i >> 10;
        const int ci = kc * 8 + ci8;
        reinterpret_cast<__half *>(b2)[i] = __float2half_rn(params[OFF_C2W + co * 256 + ci * 16 + tap]);
    }
    if (i < C2_KDIM * C2_CO) {
        // w2kt[k = tap*16+ci][co] = conv2.W[co][ci][tap]: the K-major B operand of the training dL/dcol GEMM (hp_tc.cu)
        const int co = i & 63, k = i >> 6, ci = k & 15, tap = k >> 4;
        w2kt[i] = __float2bfloat16_rn(params[OFF_C2W + co * C2_KDIM + ci * 16 + tap]);
    }
}

__global__ void __launch_bounds__(256) build_conv_images(const float *__restrict__ params, uint8_t *__restrict__ b1, uint8_t *__restrict__ b2,
                                                         __nv_bfloat16 *__restrict__ w2kt)
{
    build_conv_image_elem(blockIdx.x * 256 + threadIdx.x, params, b1, b2, w2kt);
}

// One GPU, tensor path: SGD on the conv bucket fused with the update of every 16-bit image derived from it (the scatter
// form of build_conv_image_elem / build_conv2_tmem_image: a weight knows where its copies live; the images' structural
// zeros never change).  One kernel at the exposed end of the step instead of three.
__global__ void __launch_bounds__(256) sgd_conv_images(float *__restrict__ params, const float *__restrict__ grads, float alpha, uint8_t *__restrict__ b1,
                                                       uint8_t *__restrict__ b2, __nv_bfloat16 *__restrict__ w2kt, __half *__restrict__ a2)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= OFF_F1W) return;
    const float w = fmaf(-alpha, grads[i], params[i]);   // sgd_kernel's arithmetic
    params[i] = w;
    if (i < OFF_C1B) {
        const int co = i / 25, tap = i - co * 25, ky = tap / 5, kx = tap - ky * 5;
        const __half h = __float2half_rn(w);
#pragma unroll
        for (int pos = 0; pos < 16; pos++) {
            const int blk = pos >> 2, sub = pos & 3, dy = 2 * (blk >> 1) + (sub >> 1), dx = 2 * (blk & 1) + (sub & 1);
            const int nrow = pos * 16 + co, r = ky + dy, c = kx + dx;
            reinterpret_cast<__half *>(b1 + nrow * 128 + ((r ^ (nrow & 7)) << 4))[c] = h;
        }
    } else if (i >= OFF_C2W && i < OFF_C2B) {
        const int j = i - OFF_C2W;
        const int co = j >> 8, ci = (j >> 4) & 15, tap = j & 15, ky = tap >> 2, kx = tap & 3;
        const __half h = __float2half_rn(w);
        reinterpret_cast<__half *>(b2)[((tap * 2 + (ci >> 3)) * 64 + co) * 8 + (ci & 7)] = h;
        w2kt[(tap * 16 + ci) * C2_CO + co] = __float2bfloat16_rn(w);
        a2[(2 * co + (kx >> 1)) * 128 + (ky * 2 + (kx & 1)) * 16 + ci] = h;
    }
}

int tc_sgd_conv_images(Net &net, float alpha, cudaStream_t s)
{
    TcState *t = net.tc;
    sgd_conv_images<<<(OFF_F1W + 255) / 256, 256, 0, s>>>(net.params, net.grads, alpha, t->b1_img, t->b2_img, t->w2kt, reinterpret_cast<__half *>(t->a2_img));
    LAUNCH_CHECK(net);
    return 0;
}

#ifdef HP_CONV_TRACE
extern "C" __attribute__((visibility("default"))) int hp_debug_conv_trace(long long *out, int n)
{
    return (int)cudaMemcpyFromSymbol(out, g_conv_trace, sizeof(long long) * n);
}
#endif

int tc_conv_init(Net &net)
{
    TcState *t = net.tc;
    HP_CUDA_TRY(cudaMalloc((void **)&t->b1_img, 32768));
    HP_CUDA_TRY(cudaMalloc((void **)&t->b2_img, 32768));
    HP_CUDA_TRY(cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cv::SMEM));
    HP_CUDA_TRY(cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cv::SMEM));
    t->conv_v1 = getenv("HP_CONV_V1") != nullptr;
    if (const char *e = getenv("HP_CONV_TANH")) t->conv_tanh_accurate = e[0] == 'a';
    return tc_conv2_init(net);
}

int tc_conv_refresh(Net &net, cudaStream_t s)
{
    TcState *t = net.tc;
    build_conv_images<<<64, 256, 0, s>>>(net.params, t->b1_img, t->b2_img, t->w2kt);
    LAUNCH_CHECK(net);
    return tc_conv2_refresh(net, s);
}

int tc_conv_stage(Net &net, const float *x, int64_t n, act_t *p2_bf, cudaStream_t s)
{
    TcState *t = net.tc;
    if (!t->conv_v1) return tc_conv2_stage(net, x, n, p2_bf, s);
    const int grid = (int)(n < t->num_sms ? n : t->num_sms);
    tc_conv_kernel<false><<<grid, cv::THREADS, cv::SMEM, s>>>(x, t->b1_img, t->b2_img, net.params, p2_bf, (int)n, nullptr, nullptr, nullptr, t->conv_tanh_accurate ? 1 : 0);
    LAUNCH_CHECK(net);
    return 0;
}

// training forward: also writes p1 (fp32 CHW), idx1, idx2 into the FP32 workspace layouts the backward kernels read
int tc_conv_stage_train(Net &net, const float *x, int64_t n, act_t *p2_bf, cudaStream_t s)
{
    TcState *t = net.tc;
    if (!t->conv_v1) return tc_conv2_stage_train(net, x, n, p2_bf, s);
    const int grid = (int)(n < t->num_sms ? n : t->num_sms);
    tc_conv_kernel<true><<<grid, cv::THREADS, cv::SMEM, s>>>(x, t->b1_img, t->b2_img, net.params, p2_bf, (int)n, net.ws.p1, net.ws.idx1, net.ws.idx2, t->conv_tanh_accurate ? 1 : 0);
    LAUNCH_CHECK(net);
    return 0;
}

}  // namespace hp
