// hp_fp32.cu -- the FP32 (FFMA) variant of the handposedd hot path: the "exact
// comparison" arm of BASELINE.json's north_star.  Same arithmetic as the
// reference's third_party/cnn.h (bias first, then products accumulated in
// ascending-k order, the reference's exp-based tanh, unshifted softmax, first
// strict maximum in the pools).  The conv stages use separately rounded multiply
// and add plus a correctly rounded exp, so their outputs and pool winners are
// bit-identical to the pinned reference; the FC contractions use FFMA, the only
// deliberate difference.  Parity bound 1e-5 max-normalised (tests/test_gpu_parity.py).
//
// Kernels
//   conv1_fwd_fp32      LConv 5x5 (cnn.h:205) + TanH (cnn.h:31,460) + 2x LMaxPool (cnn.h:141), fused
//   im2col_p1 + sgemm   LConv 4x4 16->64 (cnn.h:205) as [n*144 x 256] x [256 x 64]
//   tanh_pool2          TanH + LMaxPool (cnn.h:141)
//   sgemm<..>           LFull forward / backward / update contractions (cnn.h:405,430,438)
//   softmax_*           LSoftMaxChunked forward/backward (cnn.h:497,512) + Train's loss (cnn.h:566-569)
//   scatter_e2, col2im_g1, conv1_wgrad   LMaxPool::backward (cnn.h:149), TanH::df (cnn.h:32),
//                       LConv::backward / update (cnn.h:258,269)
//   sgd_kernel          the W -= alpha*g epilogue of every update()
#include "hp_common.cuh"

namespace hp {

#define LAUNCH_CHECK(net)                                   \
    do {                                                    \
        (net).launches++;                                   \
        HP_CUDA_TRY(cudaGetLastError());                    \
    } while (0)

// ============================================================================
// conv1 5x5 (1->16) + tanh + pool 2x2 + pool 2x2, one crop per CTA.
// Thread p < 225 owns one pooled pixel: it holds the 8x8 input patch in
// registers and produces the 4x4 window of conv outputs for each channel.
// tanh is applied to all 16 window values BEFORE the pools (the reference's
// order) so that the hierarchical first-strict-maximum tie-break of two stacked
// LMaxPool::backward calls (cnn.h:157-161) is evaluated on post-tanh values.
// ============================================================================
__device__ __forceinline__ void pool4(float a00, float a10, float a01, float a11, float &val, int &arg)
{
    // forward value: cnn.h:146 chain; winner: cnn.h:157-161 (first strict max, scan (0,0),(1,0),(0,1),(1,1))
    val = std_max(std_max(std_max(a00, a10), a01), a11);
    float cur = a00;
    arg = 0;
    if (a10 > cur) { cur = a10; arg = 1; }
    if (a01 > cur) { cur = a01; arg = 2; }
    if (a11 > cur) { cur = a11; arg = 3; }
}

__global__ void __launch_bounds__(256) conv1_fwd_fp32(const float *__restrict__ x, const float *__restrict__ params,
                                                      float *__restrict__ p1, uint8_t *__restrict__ idx1)
{
    __shared__ __align__(16) float img[N_IN];
    __shared__ float w[400];
    __shared__ float b[16];
    const int64_t crop = blockIdx.x;
    const int tid = threadIdx.x;
    const float4 *src = reinterpret_cast<const float4 *>(x + crop * N_IN);
#pragma unroll
    for (int i = 0; i < 4; i++) reinterpret_cast<float4 *>(img)[tid + 256 * i] = src[tid + 256 * i];
    for (int i = tid; i < 400; i += 256) w[i] = params[OFF_C1W + i];
    if (tid < 16) b[tid] = params[OFF_C1B + tid];
    __syncthreads();
    if (tid >= P1_W * P1_H) return;
    const int py = tid / P1_W, px = tid % P1_W;

    float patch[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float4 lo = *reinterpret_cast<const float4 *>(&img[(4 * py + r) * IN_W + 4 * px]);
        float4 hi = *reinterpret_cast<const float4 *>(&img[(4 * py + r) * IN_W + 4 * px + 4]);
        patch[r][0] = lo.x; patch[r][1] = lo.y; patch[r][2] = lo.z; patch[r][3] = lo.w;
        patch[r][4] = hi.x; patch[r][5] = hi.y; patch[r][6] = hi.z; patch[r][7] = hi.w;
    }
    // blockIdx.y splits the 16 output channels (small batches: 4 CTAs per crop cut the latency of a lone Eval)
    const int co_per = C1_CO / gridDim.y;
    for (int co = blockIdx.y * co_per; co < (int)(blockIdx.y + 1) * co_per; co++) {
        float acc[4][4];
        const float bias = b[co];
#pragma unroll
        for (int oy = 0; oy < 4; oy++)
#pragma unroll
            for (int ox = 0; ox < 4; ox++) acc[oy][ox] = bias;
        // reference accumulation order: ky outer, kx inner (cnn.h:223); separately rounded multiply
        // and add (no FMA) so the pre-activations -- and with them the pool winners -- are
        // bit-identical to the reference built with -ffp-contract=off
#pragma unroll
        for (int ky = 0; ky < 5; ky++)
#pragma unroll
            for (int kx = 0; kx < 5; kx++) {
                const float wv = w[co * 25 + ky * 5 + kx];
#pragma unroll
                for (int oy = 0; oy < 4; oy++)
#pragma unroll
                    for (int ox = 0; ox < 4; ox++) acc[oy][ox] = __fadd_rn(acc[oy][ox], __fmul_rn(patch[oy + ky][ox + kx], wv));
            }
#pragma unroll
        for (int oy = 0; oy < 4; oy++)
#pragma unroll
            for (int ox = 0; ox < 4; ox++) acc[oy][ox] = tanh_ref(acc[oy][ox]);
        float v[4];
        int a[4];
#pragma unroll
        for (int by = 0; by < 2; by++)
#pragma unroll
            for (int bx = 0; bx < 2; bx++)
                pool4(acc[2 * by][2 * bx], acc[2 * by][2 * bx + 1], acc[2 * by + 1][2 * bx], acc[2 * by + 1][2 * bx + 1],
                      v[by * 2 + bx], a[by * 2 + bx]);
        float val;
        int blk;
        pool4(v[0], v[1], v[2], v[3], val, blk);
        const int sub = a[blk];
        const int oy = 2 * (blk >> 1) + (sub >> 1), ox = 2 * (blk & 1) + (sub & 1);
        p1[crop * P1_N + co * (P1_W * P1_H) + tid] = val;
        idx1[crop * P1_N + co * (P1_W * P1_H) + tid] = (uint8_t)(oy * 4 + ox);
    }
}

// conv2.W [co][ci][ky][kx] -> w2p[co][(ky*4+kx)*16+ci]: the k order in which LConv::forward
// accumulates one output element (taps outer, ci inner, cnn.h:223-225).
__global__ void __launch_bounds__(256) permute_c2w(const float *__restrict__ w, float *__restrict__ w2p)
{
    const int co = blockIdx.x, k = threadIdx.x;
    const int ci = k & 15, tap = k >> 4;
    w2p[co * C2_KDIM + k] = w[co * C2_KDIM + ci * 16 + tap];
}
// dst (+)= sum_s partial[s][co][(tap)*16+ci] un-permuted back to .cnnb order [co][ci][tap]
__global__ void __launch_bounds__(256) reduce_c2w(float *__restrict__ dst, const float *__restrict__ src, int S, int accumulate)
{
    const int co = blockIdx.x, k = threadIdx.x;
    const int ci = k & 15, tap = k >> 4;
    float a = accumulate ? dst[co * C2_KDIM + ci * 16 + tap] : 0.f;
    for (int s = 0; s < S; s++) a += src[(size_t)s * (C2_CO * C2_KDIM) + co * C2_KDIM + k];
    dst[co * C2_KDIM + ci * 16 + tap] = a;
}

// im2col of the pooled conv1 stage: col[(n*144+pos)][k], k = (ky*4 + kx)*16 + ci.
__global__ void __launch_bounds__(256) im2col_p1(const float *__restrict__ p1, float *__restrict__ col)
{
    __shared__ float s[P1_N];
    const int64_t crop = blockIdx.x;
    for (int i = threadIdx.x; i < P1_N; i += 256) s[i] = p1[crop * P1_N + i];
    __syncthreads();
    float *dst = col + crop * (int64_t)(C2_POS * C2_KDIM);
    for (int e = threadIdx.x; e < C2_POS * C2_KDIM; e += 256) {
        const int pos = e >> 8, k = e & 255;
        const int ci = k & 15, ky = k >> 6, kx = (k >> 4) & 3;
        const int y = pos / C2_W, xx = pos % C2_W;
        dst[e] = s[ci * (P1_W * P1_H) + (y + ky) * P1_W + xx + kx];
    }
}

// tanh + 2x2 max-pool of conv2's pre-activations c2[(n*144+pos)][co]; writes the
// reference's CHW flatten p2[n][x + 6y + 36c] and the winner offset.
__global__ void __launch_bounds__(256) tanh_pool2(const float *__restrict__ c2, float *__restrict__ p2,
                                                  uint8_t *__restrict__ idx2, __nv_bfloat16 *__restrict__ p2_bf)
{
    __shared__ float sv[P2_N];
    __shared__ uint8_t si[P2_N];
    const int64_t crop = blockIdx.x;
    const float *src = c2 + crop * (int64_t)(C2_POS * C2_CO);
    for (int e = threadIdx.x; e < P2_N; e += 256) {
        const int co = e & 63, pp = e >> 6;
        const int py = pp / P2_W, px = pp % P2_W;
        const int base = ((2 * py) * C2_W + 2 * px) * C2_CO + co;
        float a00 = tanh_ref(src[base]);
        float a10 = tanh_ref(src[base + C2_CO]);
        float a01 = tanh_ref(src[base + C2_W * C2_CO]);
        float a11 = tanh_ref(src[base + C2_W * C2_CO + C2_CO]);
        float val;
        int arg;
        pool4(a00, a10, a01, a11, val, arg);
        sv[co * 36 + pp] = val;
        si[co * 36 + pp] = (uint8_t)arg;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < P2_N; e += 256) {
        p2[crop * P2_N + e] = sv[e];
        idx2[crop * P2_N + e] = si[e];
        // bf16 copy in HWC order (pp*64 + co), the feature order of the tensor-core fc1
        if (p2_bf) p2_bf[crop * P2_N + e] = __float2bfloat16_rn(sv[(e & 63) * 36 + (e >> 6)]);
    }
}

// ============================================================================
// SGEMM  C[M x N] = A[M x K] * B[K x N]  (+ epilogue), FFMA, ascending-k sums.
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register tile, double-buffered
// shared memory with register prefetch.
//   A_KC: A stored [M][K] (k contiguous)   else [K][M] (m contiguous)
//   B_KC: B stored [N][K] (k contiguous)   else [K][N] (n contiguous)
// ============================================================================
enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_BIAS_TANH = 2, EPI_DTANH = 3 };

struct GemmArgs {
    int M, N, K;
    const float *A; int lda;
    const float *B; int ldb;
    float *C; int ldc;
    const float *bias;   // EPI_BIAS*: accumulators start at bias[n] (LFull::forward: Y = B, cnn.h:407)
    const float *H;      // EPI_DTANH: C = (1 - H*H) * acc (TanH::df on the layer OUTPUT, cnn.h:32,467)
    int klen;            // split-K: block z handles k in [z*klen, min(K,(z+1)*klen)) and writes C + z*M*ldc
    int accumulate;      // EPI_STORE: C += acc
};

template <int BN, bool A_KC, bool B_KC, int EPI, bool NOFMA = false>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g)
{
    constexpr int BM = 128, BK = 16, PAD = 4;
    constexpr int TN = BN / 16;  // 8 or 4 columns per thread
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.klen;
    const int kend = min(g.K, kbeg + g.klen);
    float *C = g.C + (size_t)blockIdx.z * g.M * g.ldc;

    float acc[8][TN];
#pragma unroll
    for (int j = 0; j < TN; j++) {
        float bv = 0.f;
        if (EPI == EPI_BIAS || EPI == EPI_BIAS_TANH) {
            const int n = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
            bv = (n < g.N) ? g.bias[n] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i][j] = bv;
    }

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int q = tid + 256 * r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KC) {
                const int m = q >> 2, kq = q & 3;
                if (m0 + m < g.M && k0 + kq * 4 < kend) v = *reinterpret_cast<const float4 *>(g.A + (size_t)(m0 + m) * g.lda + k0 + kq * 4);
            } else {
                const int k = q >> 5, mq = q & 31;
                if (k0 + k < kend && m0 + mq * 4 < g.M) v = *reinterpret_cast<const float4 *>(g.A + (size_t)(k0 + k) * g.lda + m0 + mq * 4);
            }
            ra[r] = v;
        }
#pragma unroll
        for (int r = 0; r < (BN == 128 ? 2 : 1); r++) {
            const int q = tid + 256 * r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KC) {
                const int n = q >> 2, kq = q & 3;
                if (n0 + n < g.N && k0 + kq * 4 < kend) v = *reinterpret_cast<const float4 *>(g.B + (size_t)(n0 + n) * g.ldb + k0 + kq * 4);
            } else {
                const int k = q / (BN / 4), nq = q % (BN / 4);
                if (k0 + k < kend && n0 + nq * 4 < g.N) v = *reinterpret_cast<const float4 *>(g.B + (size_t)(k0 + k) * g.ldb + n0 + nq * 4);
            }
            rb[r] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int q = tid + 256 * r;
            if (A_KC) {
                const int m = q >> 2, kq = q & 3;
                As[buf][kq * 4 + 0][m] = ra[r].x; As[buf][kq * 4 + 1][m] = ra[r].y;
                As[buf][kq * 4 + 2][m] = ra[r].z; As[buf][kq * 4 + 3][m] = ra[r].w;
            } else {
                const int k = q >> 5, mq = q & 31;
                *reinterpret_cast<float4 *>(&As[buf][k][mq * 4]) = ra[r];
            }
        }
#pragma unroll
        for (int r = 0; r < (BN == 128 ? 2 : 1); r++) {
            const int q = tid + 256 * r;
            if (B_KC) {
                const int n = q >> 2, kq = q & 3;
                Bs[buf][kq * 4 + 0][n] = rb[r].x; Bs[buf][kq * 4 + 1][n] = rb[r].y;
                Bs[buf][kq * 4 + 2][n] = rb[r].z; Bs[buf][kq * 4 + 3][n] = rb[r].w;
            } else {
                const int k = q / (BN / 4), nq = q % (BN / 4);
                *reinterpret_cast<float4 *>(&Bs[buf][k][nq * 4]) = rb[r];
            }
        }
    };

    const int nk = (kend - kbeg + BK - 1) / BK;
    if (nk > 0) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    for (int it = 0; it < nk; it++) {
        const int buf = it & 1;
        if (it + 1 < nk) gload(kbeg + (it + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            if (TN == 8) {
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][BN / 2 + tx * 4]);
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < TN; j++)
                    acc[i][j] = NOFMA ? __fadd_rn(acc[i][j], __fmul_rn(a[i], b[j])) : fmaf(a[i], b[j], acc[i][j]);
        }
        if (it + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int h = 0; h < TN / 4; h++) {
            const int n = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
            if (n >= g.N) continue;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            float *cp = C + (size_t)m * g.ldc + n;
            if (EPI == EPI_BIAS_TANH) {
                v.x = tanh_ref(v.x); v.y = tanh_ref(v.y); v.z = tanh_ref(v.z); v.w = tanh_ref(v.w);
            } else if (EPI == EPI_DTANH) {
                const float4 hv = *reinterpret_cast<const float4 *>(g.H + (size_t)m * g.ldc + n);
                v.x = (1.0f - hv.x * hv.x) * v.x; v.y = (1.0f - hv.y * hv.y) * v.y;
                v.z = (1.0f - hv.z * hv.z) * v.z; v.w = (1.0f - hv.w * hv.w) * v.w;
            } else if (EPI == EPI_STORE && g.accumulate) {
                const float4 o = *reinterpret_cast<const float4 *>(cp);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *reinterpret_cast<float4 *>(cp) = v;
        }
    }
}

template <int BN, bool A_KC, bool B_KC, int EPI, bool NOFMA = false>
static int launch_sgemm(Net &net, const GemmArgs &g, int splits, cudaStream_t s)
{
    dim3 grid((g.N + BN - 1) / BN, (g.M + 127) / 128, splits);
    sgemm_kernel<BN, A_KC, B_KC, EPI, NOFMA><<<grid, 256, 0, s>>>(g);
    LAUNCH_CHECK(net);
    return 0;
}

// dst[i] (+)= sum_s src[s*len + i], s ascending (deterministic).
__global__ void reduce_partials(float *__restrict__ dst, const float *__restrict__ src, int S, int len, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    float a = accumulate ? dst[i] : 0.f;
    for (int s = 0; s < S; s++) a += src[(size_t)s * len + i];
    dst[i] = a;
}

// column sums of in[R][ncols] over a row slice: partial[gy][col]
__global__ void __launch_bounds__(256) colsum_partial(const float *__restrict__ in, int64_t R, int ncols, float *__restrict__ partial)
{
    __shared__ float sm[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (col < ncols)
        for (int64_t r = (int64_t)blockIdx.y * 8 + ry; r < R; r += (int64_t)gridDim.y * 8) a += in[r * ncols + col];
    sm[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && col < ncols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += sm[k][cx];
        partial[(size_t)blockIdx.y * ncols + col] = t;
    }
}

// ============================================================================
// LSoftMaxChunked (cnn.h:497-526) + the loss of CNN::Train (cnn.h:566-569).
// One CTA per crop: warp w reduces big span w (256 wide); the 16 small spans
// (16 wide) are reduced inside 16-lane groups.  No max-subtraction, as in the
// reference (overflows above 88.7 just like it).
// ============================================================================
template <bool TRAIN>
__global__ void __launch_bounds__(256) softmax_kernel(const float *__restrict__ logits, float *__restrict__ y,
                                                      const float *__restrict__ t, float *__restrict__ dlog,
                                                      float *__restrict__ mse, __nv_bfloat16 *__restrict__ dlog_bf = nullptr,
                                                      const float *__restrict__ logits2 = nullptr)
{
    const int64_t crop = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *lg = logits + crop * N_OUT;
    const float *lg2 = logits2 ? logits2 + crop * N_OUT : nullptr;   // second split-K half of the logits (tensor path, small batches)
    float ev[8], es;
    // big span `warp`
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ev[i] = exp_cr(lg2 ? lg[warp * 256 + i * 32 + lane] + lg2[warp * 256 + i * 32 + lane] : lg[warp * 256 + i * 32 + lane]);
        sum += ev[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i] = ev[i] / sum;
    // small spans: element 2048 + tid, span = tid / 16
    es = exp_cr(lg2 ? lg[2048 + tid] + lg2[2048 + tid] : lg[2048 + tid]);
    float ssum = es;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    es = es / ssum;

    if (y) {
#pragma unroll
        for (int i = 0; i < 8; i++) y[crop * N_OUT + warp * 256 + i * 32 + lane] = ev[i];
        y[crop * N_OUT + 2048 + tid] = es;
    }
    if (TRAIN) {
        __shared__ float red[8];
        const float *tt = t + crop * N_OUT;
        float e[8], se = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            e[i] = ev[i] - tt[warp * 256 + i * 32 + lane];  // e = y - t (cnn.h:568)
            se += e[i] * e[i];
            dp += e[i] * ev[i];                             // cnn.h:519-520
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dp += __shfl_xor_sync(0xffffffffu, dp, o);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float d = ev[i] * (e[i] - dp);  // cnn.h:522
            dlog[crop * N_OUT + warp * 256 + i * 32 + lane] = d;
            if (dlog_bf) dlog_bf[crop * N_OUT + warp * 256 + i * 32 + lane] = __float2bfloat16_rn(d);
        }
        const float e2 = es - tt[2048 + tid];
        se += e2 * e2;
        float dp2 = e2 * es;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dp2 += __shfl_xor_sync(0xffffffffu, dp2, o);
        dlog[crop * N_OUT + 2048 + tid] = es * (e2 - dp2);
        if (dlog_bf) dlog_bf[crop * N_OUT + 2048 + tid] = __float2bfloat16_rn(es * (e2 - dp2));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (lane == 0) red[warp] = se;
        __syncthreads();
        if (tid == 0 && mse) {
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) m += red[k];
            mse[crop] = m / (float)N_OUT;  // cnn.h:569
        }
    }
}

// LMaxPool::backward (cnn.h:149-164) after TanH::df was folded into g2 by the fc1
// dX epilogue: dense dL/dc2[(n*144+pos)][co] = g2 at the winner, 0 elsewhere.
template <bool G2_HWC>
__global__ void __launch_bounds__(256) scatter_e2(const float *__restrict__ g2, const uint8_t *__restrict__ idx2,
                                                  float *__restrict__ e2)
{
    __shared__ float sg[P2_N];
    __shared__ uint8_t si[P2_N];
    const int64_t crop = blockIdx.x;
    for (int e = threadIdx.x; e < P2_N; e += 256) {
        sg[e] = g2[crop * P2_N + e];
        si[e] = idx2[crop * P2_N + e];
    }
    __syncthreads();
    float *dst = e2 + crop * (int64_t)(C2_POS * C2_CO);
    for (int e = threadIdx.x; e < C2_POS * C2_CO; e += 256) {
        const int co = e & 63, pos = e >> 6;
        const int y = pos / C2_W, xx = pos % C2_W;
        const int pp = (y >> 1) * P2_W + (xx >> 1), off = (y & 1) * 2 + (xx & 1);
        const int j = co * 36 + pp;
        dst[e] = (si[j] == off) ? sg[G2_HWC ? pp * 64 + co : j] : 0.f;
    }
}

// col2im of dL/dcol (LConv::backward, cnn.h:258-268, as a gather) fused with the
// two LMaxPool::backward + TanH::df of the conv1 stage: only the pool winners
// carry gradient, so g1[n][ci][Y][X] = (1 - p1^2) * dL/dp1.
__global__ void __launch_bounds__(256) col2im_g1(const float *__restrict__ colgrad, const float *__restrict__ p1,
                                                 float *__restrict__ g1)
{
    const int64_t crop = blockIdx.x;
    const float *cg = colgrad + crop * (int64_t)(C2_POS * C2_KDIM);
    for (int e = threadIdx.x; e < P1_N; e += 256) {
        const int ci = e / (P1_W * P1_H), r = e % (P1_W * P1_H);
        const int Y = r / P1_W, X = r % P1_W;
        float a = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ky++) {
            const int y = Y - ky;
            if (y < 0 || y >= C2_H) continue;
#pragma unroll
            for (int kx = 0; kx < 4; kx++) {
                const int xx = X - kx;
                if (xx < 0 || xx >= C2_W) continue;
                a += cg[(y * C2_W + xx) * C2_KDIM + (ky * 4 + kx) * 16 + ci];
            }
        }
        const float pv = p1[crop * P1_N + e];
        g1[crop * P1_N + e] = (1.0f - pv * pv) * a;
    }
}

// LConv::update for conv1 (cnn.h:269-279) restricted to the pool winners: every
// other conv1 output has exactly zero gradient.  partial[block][co*25+tap], partial[block][400+co] (bias).
// A warp step covers two pooled pixels of one row, four columns apart, for all 16 channels: lane = (pixel, co).  The 5x5
// patch of a winner starts at (4 py + dy, 4 px + dx) with (dy, dx) the winner's place in its 4x4 window; with the image
// rows padded to 68 floats the bank of a patch element is  const + 4 dy + dx + 16 pixel  -- the 16 (dy, dx) of a pixel
// hit 16 different banks, lanes with the same winner read the same word (broadcast), and the two pixels use the two
// halves of the banks: no conflicts.  (One thread per (co, 16 pooled pixels), as before, ran ~4-way conflicted on 375
// shared loads per thread and crop, which -- not the FFMAs -- was the kernel's time.)
__global__ void __launch_bounds__(256) conv1_wgrad(const float *__restrict__ x, const float *__restrict__ g1,
                                                   const uint8_t *__restrict__ idx1, int64_t n, int per_block,
                                                   float *__restrict__ partial)
{
    constexpr int PITCH = 68;
    __shared__ __align__(16) float img[IN_H * PITCH];
    __shared__ __align__(16) float sg[P1_N];       // the crop's gradients; the cross-warp reduction buffer at the end
    __shared__ __align__(16) uint8_t si[P1_N];
    static_assert(8 * 16 * 26 <= P1_N, "reduction buffer");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sel = lane >> 4, co = lane & 15;
    float acc[26];
#pragma unroll
    for (int k = 0; k < 26; k++) acc[k] = 0.f;
    const int64_t b0 = (int64_t)blockIdx.x * per_block;
    const int64_t b1 = (b0 + per_block < n) ? b0 + per_block : n;
    for (int64_t crop = b0; crop < b1; crop++) {
        __syncthreads();
        const float4 *src = reinterpret_cast<const float4 *>(x + crop * N_IN);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = tid + 256 * i;   // float4 index: row q / 16, columns 4 (q % 16)
            *reinterpret_cast<float4 *>(img + (q >> 4) * PITCH + (q & 15) * 4) = src[q];
        }
        for (int i = tid; i < P1_N / 4; i += 256) reinterpret_cast<float4 *>(sg)[i] = reinterpret_cast<const float4 *>(g1 + crop * P1_N)[i];
        if (tid < P1_N / 16) reinterpret_cast<uint4 *>(si)[tid] = reinterpret_cast<const uint4 *>(idx1 + crop * P1_N)[tid];
        __syncthreads();
        // 8 slots per pooled row: pixel pairs (q, q+4) for q = 0..3, (q+4, q+8) for q = 4..6, and pixel 11 alone
        for (int slot = warp; slot < 8 * P1_H; slot += 8) {
            const int py = slot >> 3, q = slot & 7;
            const bool ok = !(q == 7 && sel == 1);
            const int px = (q < 4 ? q : q + 4) + (ok ? 4 * sel : 0);
            const int pp = py * P1_W + px;
            const float g = ok ? sg[co * (P1_W * P1_H) + pp] : 0.f;
            const int id = si[co * (P1_W * P1_H) + pp];
            const float *ip = img + (4 * py + (id >> 2)) * PITCH + 4 * px + (id & 3);
#pragma unroll
            for (int ky = 0; ky < 5; ky++)
#pragma unroll
                for (int kx = 0; kx < 5; kx++) acc[ky * 5 + kx] = fmaf(ip[ky * PITCH + kx], g, acc[ky * 5 + kx]);
            acc[25] += g;
        }
    }
    // the two pixel halves of a warp, then the 8 warps (fixed order: deterministic)
#pragma unroll
    for (int k = 0; k < 26; k++) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    __syncthreads();
    if (sel == 0) {
#pragma unroll
        for (int k = 0; k < 26; k++) sg[(warp * 16 + co) * 26 + k] = acc[k];
    }
    __syncthreads();
    for (int o = tid; o < 416; o += 256) {
        const int c = o < 400 ? o / 25 : o - 400, k = o < 400 ? o % 25 : 25;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) v += sg[(w * 16 + c) * 26 + k];
        partial[(size_t)blockIdx.x * 416 + o] = v;
    }
}

// LConv::update for conv2 (cnn.h:269-279) restricted to the pool winners: every (sample, pooled pixel, channel)
// contributes g * patch at its winning position.  Thread k = ci*16 + ky*4 + kx holds dW[co][k] for all 64 co;
// the (g, patch offset) pairs are shared by the whole CTA through shared memory.  partial[block][co*256 + k].
// g2 is in HWC order (pp*64 + co), idx2 in CHW order (co*36 + pp), p1 fp32 [ci][15][15].
__global__ void __launch_bounds__(256) conv2_wgrad_sparse(const float *__restrict__ p1, const float *__restrict__ g2,
                                                          const uint8_t *__restrict__ idx2, int64_t n, int per_block, float *__restrict__ partial)
{
    // p1 rows padded to 20 floats and planes to 304: the 16 taps x 2 channels a warp reads per step then fall into
    // 32 distinct banks ({0-3},{20-23},{8-11},{28-31} and the same shifted by 16)
    constexpr int PITCH = 20, PLANE = 304;
    __shared__ float sp1[C2_CI * PLANE];
    __shared__ float2 sgo[P2_N];   // (g, patch base offset as int bits), index pp*64 + co
    const int k = threadIdx.x, ci = k >> 4, tap_off = ((k >> 2) & 3) * PITCH + (k & 3);
    float acc[C2_CO];
#pragma unroll
    for (int co = 0; co < C2_CO; co++) acc[co] = 0.f;
    const int64_t b0 = (int64_t)blockIdx.x * per_block;
    const int64_t b1 = (b0 + per_block < n) ? b0 + per_block : n;
    for (int64_t crop = b0; crop < b1; crop++) {
        __syncthreads();
        for (int i = k; i < P1_N; i += 256) {
            const int c = i / 225, r = i - c * 225, yy = r / P1_W, xx = r - yy * P1_W;
            sp1[c * PLANE + yy * PITCH + xx] = p1[crop * P1_N + i];
        }
        for (int i = k; i < P2_N; i += 256) {
            const int pp = i >> 6, co = i & 63;
            const int a = idx2[crop * P2_N + co * 36 + pp];
            const int py = pp / P2_W, px = pp - py * P2_W;
            const int off = (2 * py + (a >> 1)) * PITCH + 2 * px + (a & 1);
            sgo[i] = make_float2(g2[crop * P2_N + i], __int_as_float(off));
        }
        __syncthreads();
        const float *base = sp1 + ci * PLANE + tap_off;
#pragma unroll 1
        for (int pp = 0; pp < 36; pp++) {
#pragma unroll
            for (int co = 0; co < C2_CO; co++) {
                const float2 go = sgo[pp * 64 + co];
                acc[co] = fmaf(go.x, base[__float_as_int(go.y)], acc[co]);
            }
        }
    }
    float *dst = partial + (size_t)blockIdx.x * (C2_CO * C2_KDIM);
#pragma unroll
    for (int co = 0; co < C2_CO; co++) dst[co * C2_KDIM + k] = acc[co];
}

// LConv::backward for conv2 (cnn.h:258-268) as a register-tiled gather, fused with both LMaxPool::backward calls
// and TanH::df of the conv1 stage: g1[ci][Y][X] = (1 - p1^2) * sum_{co,ky,kx} W[co][ci][ky][kx] * e[co][Y-ky][X-kx],
// where e is dL/dc2 (non-zero only at the pool winners).  One crop per CTA; thread (Y, ci) owns a row of 15 outputs.
__global__ void __launch_bounds__(256) conv2_dx_direct(const float *__restrict__ g2, const uint8_t *__restrict__ idx2,
                                                       const float *__restrict__ params, const float *__restrict__ p1, float *__restrict__ g1)
{
    extern __shared__ __align__(16) float dyn[];
    float *se = dyn;                    // [64 co][12 y][16 x (12 used)]
    float *sw = dyn + 64 * 12 * 16;     // [64 co][4 ky][16 ci][4 kx]
    const int64_t crop = blockIdx.x;
    const int t = threadIdx.x;
    for (int i = t; i < 64 * 12 * 16; i += 256) se[i] = 0.f;
    for (int i = t; i < C2_CO * C2_KDIM; i += 256) {
        const int kx = i & 3, ky = (i >> 2) & 3, ci = (i >> 4) & 15, co = i >> 8;   // source OIHW index
        sw[((co * 4 + ky) * 16 + ci) * 4 + kx] = params[OFF_C2W + i];
    }
    __syncthreads();
    for (int i = t; i < P2_N; i += 256) {
        const int pp = i >> 6, co = i & 63;
        const int a = idx2[crop * P2_N + co * 36 + pp];
        const int py = pp / P2_W, px = pp - py * P2_W;
        se[(co * 12 + 2 * py + (a >> 1)) * 16 + 2 * px + (a & 1)] = g2[crop * P2_N + i];
    }
    __syncthreads();
    float acc[15];
#pragma unroll
    for (int X = 0; X < 15; X++) acc[X] = 0.f;
    const int Y = t >> 4, ci = t & 15;
    if (Y < 15) {
#pragma unroll 1
        for (int co = 0; co < C2_CO; co++) {
#pragma unroll
            for (int ky = 0; ky < 4; ky++) {
                const int y = Y - ky;
                if (y < 0 || y >= C2_H) continue;
                const float4 e0 = *reinterpret_cast<const float4 *>(se + (co * 12 + y) * 16);
                const float4 e1 = *reinterpret_cast<const float4 *>(se + (co * 12 + y) * 16 + 4);
                const float4 e2 = *reinterpret_cast<const float4 *>(se + (co * 12 + y) * 16 + 8);
                const float4 w4 = *reinterpret_cast<const float4 *>(sw + ((co * 4 + ky) * 16 + ci) * 4);
                const float er[12] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w};
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int X = 0; X < 15; X++)
#pragma unroll
                    for (int kx = 0; kx < 4; kx++)
                        if (X - kx >= 0 && X - kx < C2_W) acc[X] = fmaf(wv[kx], er[X - kx], acc[X]);
            }
        }
    }
    __syncthreads();   // se is free: reuse it to stage the outputs for coalesced stores
    if (Y < 15) {
#pragma unroll
        for (int X = 0; X < 15; X++) se[ci * 225 + Y * 15 + X] = acc[X];
    }
    __syncthreads();
    for (int i = t; i < P1_N; i += 256) {
        const float pv = p1[crop * P1_N + i];
        g1[crop * P1_N + i] = (1.0f - pv * pv) * se[i];
    }
}

// dst[i] (+)= sum_s src[s*len + i] for short vectors and many slices: one warp per output element,
// lanes stride over the slices, fixed-shape shuffle tree (deterministic).
__global__ void __launch_bounds__(256) reduce_partials_warp(float *__restrict__ dst, const float *__restrict__ src, int S, int len, int accumulate)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= len) return;
    float a = 0.f;
    for (int s = lane; s < S; s += 32) a += src[(size_t)s * len + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) dst[i] = accumulate ? dst[i] + a : a;
}

// W <- W - alpha * g over the flat .cnnb-ordered stores.
__global__ void __launch_bounds__(256) sgd_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float alpha, int n4)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 w = p[i];
        const float4 d = g[i];
        w.x = fmaf(-alpha, d.x, w.x); w.y = fmaf(-alpha, d.y, w.y);
        w.z = fmaf(-alpha, d.z, w.z); w.w = fmaf(-alpha, d.w, w.w);
        p[i] = w;
    }
}

// ============================================================================
// host-side chains
// ============================================================================
// conv1+tanh+pool+pool and conv2+tanh+pool; optionally also emits the features as bf16
int fp32_conv_stage(Net &net, const float *x, int64_t n, __nv_bfloat16 *p2_bf, cudaStream_t s)
{
    Workspace &w = net.ws;
    const float *P = net.params;
    conv1_fwd_fp32<<<dim3((unsigned)n, n < 148 ? 4 : 1), 256, 0, s>>>(x, P, w.p1, w.idx1);
    LAUNCH_CHECK(net);
    im2col_p1<<<(unsigned)n, 256, 0, s>>>(w.p1, w.col);
    LAUNCH_CHECK(net);
    permute_c2w<<<C2_CO, 256, 0, s>>>(P + OFF_C2W, w.w2p);
    LAUNCH_CHECK(net);
    {   // conv2: [n*144 x 256] x w2p[64][256]^T + bias, un-fused multiply-add in the reference's k order
        GemmArgs g{(int)(n * C2_POS), C2_CO, C2_KDIM, w.col, C2_KDIM, w.w2p, C2_KDIM, w.c2, C2_CO, P + OFF_C2B, nullptr, C2_KDIM, 0};
        if (int rc = launch_sgemm<64, true, true, EPI_BIAS, true>(net, g, 1, s)) return rc;
    }
    tanh_pool2<<<(unsigned)n, 256, 0, s>>>(w.c2, w.p2, w.idx2, p2_bf);
    LAUNCH_CHECK(net);
    return 0;
}

// out[m][j] = act(bias[j] + sum_s partial[s][m][j]), s ascending: the ordered tail of a split-K LFull::forward
template <bool TANH>
__global__ void __launch_bounds__(256) bias_reduce_act(const float *__restrict__ partial, const float *__restrict__ bias, float *__restrict__ out,
                                                       int S, int M, int N)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= M * N) return;
    float a = bias[i % N];
    for (int s = 0; s < S; s++) a += partial[(size_t)s * M * N + i];
    out[i] = TANH ? tanh_ref(a) : a;
}

// LFull::forward for a handful of crops (the reference's own use: one Eval per camera frame): the layer is a
// stream of 18.9 MB of weights, so K is split 8 ways over the grid to get every SM pulling on HBM.
// Summation order: 8 partial sums of ascending k, then bias + partials in order -- differs in the last bits from the
// single ascending-k pass of the large-batch GEMM (both are within 1e-6 of the reference's order).  The choice is made
// per CALL (Net::fp32_small_call: whole call <= 64 crops), never per workspace chunk, so within one call every crop
// gets the same arithmetic wherever it sits in the batch (tests: ragged sizes across 63/64/65 and 2048+3).
constexpr int SMALL_BATCH = 64, SMALL_SPLITS = 8;
template <bool TANH>
static int fc_small(Net &net, const float *x, int M, int K, const float *W, const float *bias, int N, float *out, cudaStream_t s)
{
    const int klen = K / SMALL_SPLITS;   // 288 or 256: multiples of 16
    GemmArgs g{M, N, K, x, K, W, N, net.ws.partial, N, nullptr, nullptr, klen, 0};
    if (int rc = launch_sgemm<128, true, false, EPI_STORE>(net, g, SMALL_SPLITS, s)) return rc;
    bias_reduce_act<TANH><<<(M * N + 255) / 256, 256, 0, s>>>(net.ws.partial, bias, out, SMALL_SPLITS, M, N);
    LAUNCH_CHECK(net);
    return 0;
}

// out[m][i] = (1 - h[m][i]^2) * sum_s partial[s][m][i]: ordered tail of a split-K LFull::backward + TanH::df
__global__ void __launch_bounds__(256) dtanh_reduce(const float *__restrict__ partial, const float *__restrict__ h, float *__restrict__ out, int S,
                                                    int MN)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= MN) return;
    float a = 0.f;
    for (int s = 0; s < S; s++) a += partial[(size_t)s * MN + i];
    const float hv = h[i];
    out[i] = (1.0f - hv * hv) * a;
}
// D[M][Nin] = (E[M][Kout] * W[Nin][Kout]^T) .* (1 - H^2) for a handful of samples, K split over the grid
static int fc_dx_small(Net &net, const float *E, int M, int Kout, const float *W, int Nin, const float *H, float *D, cudaStream_t s)
{
    const int klen = Kout / SMALL_SPLITS;
    GemmArgs g{M, Nin, Kout, E, Kout, W, Kout, net.ws.partial, Nin, nullptr, nullptr, klen, 0};
    if (int rc = launch_sgemm<128, true, true, EPI_STORE>(net, g, SMALL_SPLITS, s)) return rc;
    dtanh_reduce<<<(M * Nin + 255) / 256, 256, 0, s>>>(net.ws.partial, H, D, SMALL_SPLITS, M * Nin);
    LAUNCH_CHECK(net);
    return 0;
}

int fp32_forward(Net &net, const float *x, int64_t n, float *y_out, bool training, cudaStream_t s)
{
    Workspace &w = net.ws;
    const float *P = net.params;
    {
        StageTimer st(net, 0, s);
        if (int rc = fp32_conv_stage(net, x, n, nullptr, s)) return rc;
    }
    if (n <= SMALL_BATCH && net.fp32_small_call) {
        {
            StageTimer st(net, 1, s);
            if (int rc = fc_small<true>(net, w.p2, (int)n, FC1_IN, P + OFF_F1W, P + OFF_F1B, FC1_OUT, w.h1, s)) return rc;
        }
        StageTimer st(net, 2, s);
        if (int rc = fc_small<false>(net, w.h1, (int)n, FC2_IN, P + OFF_F2W, P + OFF_F2B, FC2_OUT, w.logits, s)) return rc;
    } else {
        {   // fc1 + tanh
            StageTimer st(net, 1, s);
            GemmArgs g{(int)n, FC1_OUT, FC1_IN, w.p2, FC1_IN, P + OFF_F1W, FC1_OUT, w.h1, FC1_OUT, P + OFF_F1B, nullptr, FC1_IN, 0};
            if (int rc = launch_sgemm<128, true, false, EPI_BIAS_TANH>(net, g, 1, s)) return rc;
        }
        {   // fc2 logits
            StageTimer st(net, 2, s);
            GemmArgs g{(int)n, FC2_OUT, FC2_IN, w.h1, FC2_IN, P + OFF_F2W, FC2_OUT, w.logits, FC2_OUT, P + OFF_F2B, nullptr, FC2_IN, 0};
            if (int rc = launch_sgemm<128, true, false, EPI_BIAS>(net, g, 1, s)) return rc;
        }
    }
    if (!training) {
        StageTimer st(net, 3, s);
        softmax_kernel<false><<<(unsigned)n, 256, 0, s>>>(w.logits, y_out, nullptr, nullptr, nullptr);
        LAUNCH_CHECK(net);
    }
    return 0;
}

// column sums of in[R][ncols] for short R in ONE launch: thread (cx, ry) adds rows ry, ry+32, ... of its column, the
// 32 row-lanes are combined in a fixed order (deterministic).  Minibatch bias gradients: R = 256 rows.
__global__ void __launch_bounds__(1024) colsum_direct(const float *__restrict__ in, int R, int ncols, float *__restrict__ dst, int accumulate)
{
    __shared__ float sm[32][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (col < ncols) {
        int r = ry;
        for (; r + 96 < R; r += 128) {
            a0 += in[(size_t)r * ncols + col];
            a1 += in[(size_t)(r + 32) * ncols + col];
            a2 += in[(size_t)(r + 64) * ncols + col];
            a3 += in[(size_t)(r + 96) * ncols + col];
        }
        for (; r < R; r += 32) a0 += in[(size_t)r * ncols + col];
    }
    sm[ry][cx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ry == 0 && col < ncols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k++) t += sm[k][cx];
        dst[col] = accumulate ? dst[col] + t : t;
    }
}

static int colsum(Net &net, const float *in, int64_t R, int ncols, float *dst, bool accumulate, cudaStream_t s)
{
    if (R <= 512 && ncols >= 1024) {   // enough columns to fill the machine with one CTA per 32 of them
        colsum_direct<<<(ncols + 31) / 32, 1024, 0, s>>>(in, (int)R, ncols, dst, accumulate ? 1 : 0);
        LAUNCH_CHECK(net);
        return 0;
    }
    int gy = (int)((R + 63) / 64);
    if (gy > 64) gy = 64;
    if (gy < 1) gy = 1;
    dim3 grid((ncols + 31) / 32, gy);
    colsum_partial<<<grid, 256, 0, s>>>(in, R, ncols, net.ws.partial);
    LAUNCH_CHECK(net);
    reduce_partials<<<(ncols + 255) / 256, 256, 0, s>>>(dst, net.ws.partial, gy, ncols, accumulate ? 1 : 0);
    LAUNCH_CHECK(net);
    return 0;
}

// conv2 and conv1 backward from g2 = dL/d(conv2 pre-activation) at the pool winners ([n][2304], CHW or HWC order)
static int fp32_conv_backward_impl(Net &net, const float *x, int64_t n, const float *g2, bool g2_hwc, bool accumulate, cudaStream_t s)
{
    Workspace &w = net.ws;
    float *G = net.grads;
    const int acc = accumulate ? 1 : 0;
    // ---- conv2: dense dL/dc2 (reuses w.c2), dB, dW (split-K over positions), dcol
    if (g2_hwc) scatter_e2<true><<<(unsigned)n, 256, 0, s>>>(g2, w.idx2, w.c2);
    else scatter_e2<false><<<(unsigned)n, 256, 0, s>>>(g2, w.idx2, w.c2);
    LAUNCH_CHECK(net);
    const int64_t R = n * C2_POS;
    if (int rc = colsum(net, w.c2, R, C2_CO, G + OFF_C2B, accumulate, s)) return rc;
    {
        int splits = (int)((R + 287) / 288);  // 2 crops per split: 2 N tiles x splits CTAs should cover the 148 SMs
        if (splits > 128) splits = 128;
        int klen = (int)((R + splits - 1) / splits);
        klen = (klen + 15) / 16 * 16;
        splits = (int)((R + klen - 1) / klen);
        GemmArgs g{C2_CO, C2_KDIM, (int)R, w.c2, C2_CO, w.col, C2_KDIM, w.partial, C2_KDIM, nullptr, nullptr, klen, 0};
        if (int rc = launch_sgemm<128, false, false, EPI_STORE>(net, g, splits, s)) return rc;
        reduce_c2w<<<C2_CO, 256, 0, s>>>(G + OFF_C2W, w.partial, splits, acc);
        LAUNCH_CHECK(net);
    }
    {
        GemmArgs g{(int)R, C2_KDIM, C2_CO, w.c2, C2_CO, w.w2p, C2_KDIM, w.colgrad, C2_KDIM, nullptr, nullptr, C2_CO, 0};
        if (int rc = launch_sgemm<128, true, false, EPI_STORE>(net, g, 1, s)) return rc;
    }
    col2im_g1<<<(unsigned)n, 256, 0, s>>>(w.colgrad, w.p1, w.g1);
    LAUNCH_CHECK(net);
    // ---- conv1 (no dX: CNN::Train never calls layer 0's backward, cnn.h:571)
    {
        int per_block = (int)((n + 147) / 148);
        if (per_block < 1) per_block = 1;
        int blocks = (int)((n + per_block - 1) / per_block);
        conv1_wgrad<<<blocks, 256, 0, s>>>(x, w.g1, w.idx1, n, per_block, w.partial);
        LAUNCH_CHECK(net);
        reduce_partials<<<2, 256, 0, s>>>(G + OFF_C1W, w.partial, blocks, 416, acc);
        LAUNCH_CHECK(net);
    }
    return 0;
}


// Backward + weight-gradient sums into net.grads (W -= alpha*grads is sgd_apply).
// Order fc2 -> fc1 -> conv2 -> conv1 so that the large FC buckets are ready first
// for the data-parallel all-reduce (events ev_bucket[0..2]).
int fp32_backward(Net &net, const float *x, const float *t, int64_t n, float *mse, bool accumulate, cudaStream_t s)
{
    Workspace &w = net.ws;
    const float *P = net.params;
    float *G = net.grads;
    const int acc = accumulate ? 1 : 0;
    softmax_kernel<true><<<(unsigned)n, 256, 0, s>>>(w.logits, w.y, t, w.dlog, mse);
    LAUNCH_CHECK(net);
    // ---- fc2: dB, dW = h1^T * dlog, dX = dlog * W2^T with tanh' of fc1's output
    if (int rc = colsum(net, w.dlog, n, FC2_OUT, G + OFF_F2B, accumulate, s)) return rc;
    {
        GemmArgs g{FC2_IN, FC2_OUT, (int)n, w.h1, FC2_IN, w.dlog, FC2_OUT, G + OFF_F2W, FC2_OUT, nullptr, nullptr, (int)n, acc};
        if (int rc = launch_sgemm<128, false, false, EPI_STORE>(net, g, 1, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[0], s));
    if (n <= SMALL_BATCH && net.fp32_small_call) {
        if (int rc = fc_dx_small(net, w.dlog, (int)n, FC2_OUT, P + OFF_F2W, FC2_IN, w.h1, w.da1, s)) return rc;
    } else {
        GemmArgs g{(int)n, FC2_IN, FC2_OUT, w.dlog, FC2_OUT, P + OFF_F2W, FC2_OUT, w.da1, FC2_IN, nullptr, w.h1, FC2_OUT, 0};
        if (int rc = launch_sgemm<128, true, true, EPI_DTANH>(net, g, 1, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_dx[0], s));
    // ---- fc1
    if (int rc = colsum(net, w.da1, n, FC1_OUT, G + OFF_F1B, accumulate, s)) return rc;
    {
        GemmArgs g{FC1_IN, FC1_OUT, (int)n, w.p2, FC1_IN, w.da1, FC1_OUT, G + OFF_F1W, FC1_OUT, nullptr, nullptr, (int)n, acc};
        if (int rc = launch_sgemm<128, false, false, EPI_STORE>(net, g, 1, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[1], s));
    if (n <= SMALL_BATCH && net.fp32_small_call) {
        if (int rc = fc_dx_small(net, w.da1, (int)n, FC1_OUT, P + OFF_F1W, FC1_IN, w.p2, w.g2, s)) return rc;
    } else {
        GemmArgs g{(int)n, FC1_IN, FC1_OUT, w.da1, FC1_OUT, P + OFF_F1W, FC1_OUT, w.g2, FC1_IN, nullptr, w.p2, FC1_OUT, 0};
        if (int rc = launch_sgemm<128, true, true, EPI_DTANH>(net, g, 1, s)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_dx[1], s));
    if (int rc = fp32_conv_backward_impl(net, x, n, w.g2, false, accumulate, s)) return rc;
    HP_CUDA_TRY(cudaEventRecord(net.ev_bucket[2], s));
    return 0;
}

int fp32_conv_backward(Net &net, const float *x, int64_t n, const float *g2_hwc, bool accumulate, cudaStream_t s)
{
    return fp32_conv_backward_impl(net, x, n, g2_hwc, true, accumulate, s);
}

constexpr int CONV2_DX_SMEM = (64 * 12 * 16 + C2_CO * C2_KDIM) * 4;  // 112 KB

// Conv-stage backward for the tensor-core training path: winners-only conv2 weight gradient (no im2col), a
// direct gather for dL/dp1 (FFMA), winners-only conv1 update.  g2 in HWC order; p1, idx1, idx2 in w.*.
int tc_conv_backward(Net &net, const float *x, int64_t n, const float *g2_hwc, bool accumulate, cudaStream_t s)
{
    Workspace &w = net.ws;
    float *G = net.grads;
    const int acc = accumulate ? 1 : 0;
    // conv2 dB: g2 viewed as [n*36][64]
    if (int rc = colsum(net, g2_hwc, n * 36, C2_CO, G + OFF_C2B, accumulate, s)) return rc;
    {
        int per_block = (int)((n + 147) / 148);
        int blocks = (int)((n + per_block - 1) / per_block);
        if (blocks > 128) { per_block = (int)((n + 127) / 128); blocks = (int)((n + per_block - 1) / per_block); }  // partial buffer holds 128 slices
        conv2_wgrad_sparse<<<blocks, 256, 0, s>>>(w.p1, g2_hwc, w.idx2, n, per_block, w.partial);
        LAUNCH_CHECK(net);
        reduce_partials<<<(C2_CO * C2_KDIM + 255) / 256, 256, 0, s>>>(G + OFF_C2W, w.partial, blocks, C2_CO * C2_KDIM, acc);
        LAUNCH_CHECK(net);
    }
    // dL/dp1 through conv2 and the conv1-stage pools / tanh': direct register-tiled gather
    conv2_dx_direct<<<(unsigned)n, 256, CONV2_DX_SMEM, s>>>(g2_hwc, w.idx2, net.params, w.p1, w.g1);
    LAUNCH_CHECK(net);
    {
        const int blocks = (int)n;   // one crop per CTA: two CTAs per SM hide each other's load latency
        if ((size_t)blocks * 416 > w.partial_floats) { set_error("partial buffer too small for %d crops", blocks); return 1; }
        conv1_wgrad<<<blocks, 256, 0, s>>>(x, w.g1, w.idx1, n, 1, w.partial);
        LAUNCH_CHECK(net);
        reduce_partials_warp<<<(416 + 7) / 8, 256, 0, s>>>(G + OFF_C1W, w.partial, blocks, 416, acc);
        LAUNCH_CHECK(net);
    }
    return 0;
}

int fp32_reduce_warp(Net &net, float *dst, const float *partial, int S, int len, bool accumulate, cudaStream_t s)
{
    reduce_partials_warp<<<(len + 7) / 8, 256, 0, s>>>(dst, partial, S, len, accumulate ? 1 : 0);
    LAUNCH_CHECK(net);
    return 0;
}

int fp32_conv1_wgrad(Net &net, const float *x, int64_t n, bool accumulate, cudaStream_t s)
{
    Workspace &w = net.ws;
    const int blocks = (int)n;   // one crop per CTA: two CTAs per SM hide each other's load latency
    if ((size_t)blocks * 416 > w.partial_floats) { set_error("partial buffer too small for %d crops", blocks); return 1; }
    conv1_wgrad<<<blocks, 256, 0, s>>>(x, w.g1, w.idx1, n, 1, w.partial);
    LAUNCH_CHECK(net);
    reduce_partials_warp<<<(416 + 7) / 8, 256, 0, s>>>(net.grads + OFF_C1W, w.partial, blocks, 416, accumulate ? 1 : 0);
    LAUNCH_CHECK(net);
    return 0;
}

int fp32_init_attributes()
{
    HP_CUDA_TRY(cudaFuncSetAttribute(conv2_dx_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV2_DX_SMEM));
    return 0;
}

int fp32_softmax_loss(Net &net, const float *logits, float *y, const float *t, float *dlog, __nv_bfloat16 *dlog_bf, float *mse, int64_t n, cudaStream_t s,
                      const float *logits2)
{
    softmax_kernel<true><<<(unsigned)n, 256, 0, s>>>(logits, y, t, dlog, mse, dlog_bf, logits2);
    LAUNCH_CHECK(net);
    return 0;
}

int fp32_colsum(Net &net, const float *in, int64_t R, int ncols, float *dst, bool accumulate, cudaStream_t s)
{
    return colsum(net, in, R, ncols, dst, accumulate, s);
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4 *__restrict__ src, uint2 *__restrict__ dst, int n4)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 v = src[i];
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        dst[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
}
__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const uint2 *__restrict__ src, float4 *__restrict__ dst, int n4)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const uint2 v = src[i];
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.y));
        dst[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}
// gradient bucket [off, off+count) -> bf16 wire buffer and back (data-parallel bf16 transport)
int grads_to_bf16(Net &net, int off, int count, cudaStream_t s)
{
    f32_to_bf16_kernel<<<148 * 4, 256, 0, s>>>(reinterpret_cast<const float4 *>(net.grads + off), reinterpret_cast<uint2 *>(net.grads_bf + off), count / 4);
    LAUNCH_CHECK(net);
    return 0;
}
int grads_from_bf16(Net &net, int off, int count, cudaStream_t s)
{
    bf16_to_f32_kernel<<<148 * 4, 256, 0, s>>>(reinterpret_cast<const uint2 *>(net.grads_bf + off), reinterpret_cast<float4 *>(net.grads + off), count / 4);
    LAUNCH_CHECK(net);
    return 0;
}

// SGD over one gradient bucket [off, off+count) (both multiples of 4 floats)
int sgd_apply_range(Net &net, float alpha, int off, int count, cudaStream_t s)
{
    int blocks = (count / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    sgd_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<float4 *>(net.params + off), reinterpret_cast<const float4 *>(net.grads + off), alpha, count / 4);
    LAUNCH_CHECK(net);
    net.tc_dirty = true;
    return 0;
}

int sgd_apply(Net &net, float alpha, cudaStream_t s)
{
    sgd_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<float4 *>(net.params), reinterpret_cast<const float4 *>(net.grads), alpha,
                                       N_PARAMS / 4);
    LAUNCH_CHECK(net);
    net.tc_dirty = true;
    return 0;
}

}  // namespace hp
