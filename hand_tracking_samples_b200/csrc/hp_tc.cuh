// hp_tc.cuh -- state shared by the tensor-core kernel files (hp_tc.cu: FC GEMMs, hp_tc_conv.cu: conv stages).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "hp_common.cuh"

namespace hp {

// ---- arithmetic of the tensor-core path --------------------------------------------------------------------------
// Forward operands (crop, conv weights, p1, p2, h1, FC weights) are IEEE half: the same tcgen05 rate as bf16 with
// 8x less rounding error (2^-11 vs 2^-8 relative), which is what keeps peaky softmax outputs (logits x30, the regime of
// a trained net) inside the 1e-2 bound of BASELINE.json.  Every one of these values is a tanh output, a [0,1] depth value
// or a weight, so the fp16 range is no constraint; the float -> half conversions saturate to +-65504 (NaN stays NaN).
// Backward operands (dL/dlogits, dL/da1, dense conv2 error) stay bf16: gradients need the exponent range.
typedef __half act_t;

__device__ __forceinline__ uint32_t pack_act(float a, float b)   // a -> low half, b -> high half
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float2 unpack_act(uint32_t v)
{
    return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}
__device__ __forceinline__ float ex2_fast(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// TanH::f (cnn.h:31), (e - 1) / (e + 1) with e = exp(2t), on the MUFU units: ex2.approx (2 ulp) and rcp.approx (1 ulp),
// evaluated as 1 - 2 / (e + 1).  Absolute error ~2e-7 (tanh.approx has 2^-11 RELATIVE error, which at 30x logits is
// what broke the 1e-2 bound on peaky outputs).  The reference's overflow quirk is kept: e = inf for t > 44.36 makes
// (e - 1) / (e + 1) NaN (SURVEY.md 8a note 2); here (e - e) contributes that NaN and is 0 otherwise.  Very negative t
// gives e = 0 and -1, as in the reference.  The .ftz forms matter: without them ptxas wraps every MUFU in a predicated
// denormal-rescue sequence on one predicate register, which serialises the epilogue (measured: fc1 110 -> 140 us).
__device__ __forceinline__ float tanh_tc(float t)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f) + (e - e);
}

// The conv-stage epilogues sit on the critical ring of the fused conv kernel (accumulator drain -> tanh -> p1 -> conv2
// MMAs), where two MUFU ops per value are measurable; their outputs are rounded to fp16 (2^-11) anyway, so there the
// single-MUFU tanh.approx (2^-11 relative) is used, with the reference's overflow-to-NaN quirk kept by a select
// (t > 44.3614: exp(2t) overflows in fp32, cnn.h:31).  `accurate` (HP_CONV_TANH=accurate) switches back for A/B runs.
__device__ __forceinline__ float tanh_conv(float t, bool accurate)
{
    if (accurate) return tanh_tc(t);
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(t));
    return (t > 44.3614f) ? __int_as_float(0x7fc00000) : y;
}

// tanh.approx alone: for callers that have already ruled the overflow quirk out for a whole group of values
__device__ __forceinline__ float tanh_fast(float t)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(t));
    return y;
}

struct TcState {
    // 16-bit shadows of the weights (fp16 for the forward operands, bf16 for the backward ones), in the layouts the MMAs consume (rebuilt by tc_refresh_weights)
    act_t *w1t = nullptr;          // [2048][2304] = fc1.W^T, k contiguous, k in HWC flatten order (pp*64+co)
    act_t *w2t = nullptr;          // [2304][2048] = fc2.W^T
    uint8_t *b1_img = nullptr;     // 32 KB: conv1 as pooled-window GEMM, B operand [256 (pos,co)][64 (r,c)] fp16, 128B-swizzled image
    uint8_t *a2_img = nullptr;     // 32 KB: conv2 weights as the TMEM-resident A operand of the v2 conv kernel, [128 (2co+g)][128 k] fp16
    bool conv_tanh_accurate = false;   // HP_CONV_TANH=accurate: two-MUFU tanh in the conv epilogues too (A/B runs)
    bool conv_serial_drain = false;    // HP_CONV_PIPE=0: v2 conv kernel without the pipelined accumulator drains (A/B runs)
    bool conv_v1 = false;          // HP_CONV_V1=1: run the round-1 conv kernel (hp_tc_conv.cu) instead of hp_tc_conv2.cu
    float *dec_scratch = nullptr;  // [SMs][128][256]: the y tile a fc2 CTA decodes in its epilogue (SOFTMAX_DECODE), L2-resident
    uint8_t *b2_img = nullptr;     // 32 KB: conv2 taps, [16 taps][2 k-chunks][64 co][8 ci] fp16 (no-swizzle core matrices)
    __nv_bfloat16 *w1b = nullptr;  // [2304 (k' HWC)][2048] = fc1.W as stored (B operand of the fc1 dX GEMM)
    __nv_bfloat16 *w2b = nullptr;  // [2048][2304] = fc2.W as stored
    // training activations (TRAIN_CAP samples per pass)
    __nv_bfloat16 *g2_sink = nullptr;                               // [cap][2304]: unused bf16 copy of the fc1 dX epilogue
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
    act_t *p2 = nullptr;           // [cap][2304] pooled conv2 stage (fc1 input), HWC flatten
    act_t *h1 = nullptr;           // [cap][2048] tanh(fc1)
    int64_t cap = 0;
    CUtensorMap tm_w1t, tm_w2t, tm_p2, tm_h1;
    int num_sms = 148;     // CTAs of the persistent grids (SM count minus the SMs reserved for NCCL in data-parallel mode)
    int total_sms = 148;
};

int tc_conv_init(Net &net);
void tc_set_reserved_sms(Net &net, int reserve);
int tc_conv_refresh(Net &net, cudaStream_t s);
int tc_train_grad(Net &net, const float *x, const float *t_dev, int64_t n, float *mse, bool accumulate, cudaStream_t s);
int tc_conv_stage_train(Net &net, const float *x, int64_t n, act_t *p2_bf, cudaStream_t s);
int tc_conv_stage(Net &net, const float *x, int64_t n, act_t *p2_bf, cudaStream_t s);
// hp_tc_conv2.cu: the transposed / tap-paired conv2 formulation (default)
int tc_conv2_init(Net &net);
int tc_conv2_refresh(Net &net, cudaStream_t s);
int tc_conv2_stage(Net &net, const float *x, int64_t n, act_t *p2, cudaStream_t s);
int tc_conv2_stage_u16(Net &net, const uint16_t *depth, int64_t n, float depth_scale, float dmin, float dmax, act_t *p2, cudaStream_t s);
int tc_conv2_stage_train(Net &net, const float *x, int64_t n, act_t *p2, cudaStream_t s);

}  // namespace hp
