// hp_tc.cuh -- state shared by the tensor-core kernel files (hp_tc.cu: FC GEMMs, hp_tc_conv.cu: conv stages).
#pragma once
#include <cuda.h>
#include "hp_common.cuh"

namespace hp {

struct TcState {
    // bf16 shadows of the weights, in the layouts the MMAs consume (rebuilt by tc_refresh_weights)
    __nv_bfloat16 *w1t = nullptr;  // [2048][2304] = fc1.W^T, k contiguous, k in HWC flatten order (pp*64+co)
    __nv_bfloat16 *w2t = nullptr;  // [2304][2048] = fc2.W^T
    uint8_t *b1_img = nullptr;     // 32 KB: conv1 as pooled-window GEMM, B operand [256 (pos,co)][64 (r,c)] bf16, 128B-swizzled image
    uint8_t *b2_img = nullptr;     // 32 KB: conv2 taps, [16 taps][2 k-chunks][64 co][8 ci] bf16 (no-swizzle core matrices)
    __nv_bfloat16 *w1b = nullptr;  // [2304 (k' HWC)][2048] = fc1.W as stored (B operand of the fc1 dX GEMM)
    __nv_bfloat16 *w2b = nullptr;  // [2048][2304] = fc2.W as stored
    // training activations (TRAIN_CAP samples per pass)
    __nv_bfloat16 *dlog_bf = nullptr, *da1_bf = nullptr;            // [cap][2304], [cap][2048]
    __nv_bfloat16 *h1T = nullptr, *dlogT = nullptr, *p2T = nullptr, *da1T = nullptr;  // [features][cap]: batch-contiguous (K-major for dW)
    // conv2 backward as GEMMs: dense error rows / its transpose / transposed im2col of p1, (n,pos) padded to TRAIN_CAP*144
    __nv_bfloat16 *e2 = nullptr, *e2T = nullptr, *colT = nullptr;   // [cap*144][64], [64][cap*144], [256][cap*144]
    float *db2_partial = nullptr;                                    // [cap][64] per-crop conv2 bias-gradient sums
    __nv_bfloat16 *w2kt = nullptr;                                   // [256 k = tap*16+ci][64 co] = conv2.W, K-major over co
    CUtensorMap tm_e2, tm_e2T, tm_colT, tm_w2kt;
    CUtensorMap tm_w1t64, tm_w2t64, tm_w1b64, tm_w2b64;   // the same weights with 64-row boxes (small-batch GEMMs)
    CUtensorMap tm_w1b, tm_w2b, tm_dlog, tm_da1, tm_h1T, tm_p2T, tm_dlogT, tm_da1T;
    CUtensorMap tm_dlogT128, tm_da1T128;                  // 128-row boxes: half-width weight-gradient tiles in data-parallel mode
    // activations
    __nv_bfloat16 *p2 = nullptr;   // [cap][2304] pooled conv2 stage (fc1 input), HWC flatten
    __nv_bfloat16 *h1 = nullptr;   // [cap][2048] tanh(fc1)
    int64_t cap = 0;
    CUtensorMap tm_w1t, tm_w2t, tm_p2, tm_h1;
    int num_sms = 148;     // CTAs of the persistent grids (SM count minus the SMs reserved for NCCL in data-parallel mode)
    int total_sms = 148;
};

int tc_conv_init(Net &net);
void tc_set_reserved_sms(Net &net, int reserve);
int tc_conv_refresh(Net &net, cudaStream_t s);
int tc_train_grad(Net &net, const float *x, const float *t_dev, int64_t n, float *mse, bool accumulate, cudaStream_t s);
int tc_conv_stage_train(Net &net, const float *x, int64_t n, __nv_bfloat16 *p2_bf, cudaStream_t s);
int tc_conv_stage(Net &net, const float *x, int64_t n, __nv_bfloat16 *p2_bf, cudaStream_t s);

}  // namespace hp
