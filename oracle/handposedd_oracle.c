/*
 * handposedd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's CNN hot path
 * (IntelRealSense/hand_tracking_samples, third_party/cnn.h) for the one
 * architecture the reference instantiates ("handposedd",
 * include/handtrack.h:108-118).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker.  The shipped CUDA path never calls into it.
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the UNMODIFIED
 * reference header itself, compiled in place from /root/reference by
 * oracle/Makefile into oracle/_ref/libcnnref.so (oracle/ref_shim.cpp), and
 * against the fixtures in tests/golden/ that were produced by that library
 * (tests/golden/make_golden.py).  tests/test_oracle.py demands
 * BIT-EXACT agreement (Eval outputs, per-sample gradients, weights after N
 * Train steps, Init() weights) when both are built with
 * `-O2 -msse2 -ffp-contract=off`.
 *
 * Every function cites the reference lines it follows.  Floating-point
 * operation ORDER is part of the contract: no FMA contraction, the same
 * left-to-right association and the same loop nesting as the reference.
 *
 * Tensor layout is the reference's: planar CHW with x fastest
 * (make_packed_stride, cnn.h:45-47); conv weights index
 * kx + kw*(ky + kh*(ci + cin*co)) (cnn.h:47,201); FC weights j + i*N with
 * i = input, j = output (cnn.h:417,426).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- handposedd geometry (include/handtrack.h:108-118) ------------------ */
enum {
    IN_W = 64, IN_H = 64,
    C1_K = 5, C1_CO = 16, C1_W = 60, C1_H = 60,          /* LConv({64,64,1},{5,5,1,16},{60,60,16}) */
    P1_W = 30, P1_H = 30,                                 /* LMaxPool({60,60,16}) */
    P2_W = 15, P2_H = 15,                                 /* LMaxPool({30,30,16}) */
    C2_K = 4, C2_CI = 16, C2_CO = 64, C2_W = 12, C2_H = 12, /* LConv({15,15,16},{4,4,16,64},{12,12,64}) */
    P3_W = 6, P3_H = 6,                                   /* LMaxPool({12,12,64}) */
    FC1_IN = 2304, FC1_OUT = 2048,                        /* LFull(6*6*64, 16*16*8) */
    FC2_IN = 2048, FC2_OUT = 2304,                        /* LFull(16*16*8, 16*16*8+16*16) */
    N_SPANS = 24,
    N_OUT = 2304, N_IN = 4096
};

/* .cnnb float offsets (cnn.h:97-98,288-289,454-455,590-593) */
enum {
    OFF_C1W = 0,
    OFF_C1B = OFF_C1W + C1_CO * C1_K * C1_K,                 /* 400     */
    OFF_C2W = OFF_C1B + C1_CO,                               /* 416     */
    OFF_C2B = OFF_C2W + C2_CO * C2_CI * C2_K * C2_K,         /* 16800   */
    OFF_F1W = OFF_C2B + C2_CO,                               /* 16864   */
    OFF_F1B = OFF_F1W + FC1_IN * FC1_OUT,                    /* 4735456 */
    OFF_F2W = OFF_F1B + FC1_OUT,                             /* 4737504 */
    OFF_F2B = OFF_F2W + FC2_IN * FC2_OUT,                    /* 9456096 */
    N_PARAMS = OFF_F2B + FC2_OUT                             /* 9458400 */
};

ORC_API int orc_n_params(void) { return N_PARAMS; }

static const int k_spans[N_SPANS] = {256, 256, 256, 256, 256, 256, 256, 256,
                                     16, 16, 16, 16, 16, 16, 16, 16,
                                     16, 16, 16, 16, 16, 16, 16, 16};

/* ---- per-layer restatements -------------------------------------------- */

/* CNN::LConv::forward, cnn.h:205-257: bias fill, then tap-outer accumulation
 * (rect_iteration: kx fastest, geometric.h:24-38), ci then co inside. */
static void conv_forward(const float *in, int iw, int ih, int cin,
                         const float *W, const float *B, int kw, int kh, int cout,
                         float *out, int ow, int oh)
{
    for (int z = 0; z < cout; z++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++)
                out[x + ow * (y + oh * z)] = B[z];
    for (int ky = 0; ky < kh; ky++)
        for (int kx = 0; kx < kw; kx++)
            for (int iz = 0; iz < cin; iz++)
                for (int oz = 0; oz < cout; oz++) {
                    float w = W[kx + kw * (ky + kh * (iz + cin * oz))];
                    float *op = out + oz * ow * oh;
                    for (int y = 0; y < oh; y++) {
                        const float *ip = in + kx + iw * ky + iz * iw * ih + iw * y;
                        for (int x = 0; x < ow; x++) {
                            float prod = ip[x] * w;
                            *op = *op + prod;
                            op++;
                        }
                    }
                }
}

/* CNN::LConv::backward, cnn.h:258-268 with madd cnn.h:70-94:
 * for every output (x fastest, then y, then co): D[patch] += W[co] * er. */
static void conv_backward(const float *E, int ow, int oh, int cout,
                          const float *W, int kw, int kh, int cin,
                          float *D, int iw, int ih)
{
    memset(D, 0, sizeof(float) * (size_t)iw * ih * cin);
    for (int z = 0; z < cout; z++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++) {
                float s = E[x + ow * (y + oh * z)];
                const float *a = W + (size_t)z * kw * kh * cin;
                for (int cz = 0; cz < cin; cz++)
                    for (int cy = 0; cy < kh; cy++)
                        for (int cx = 0; cx < kw; cx++) {
                            float *d = D + (x + cx) + iw * ((y + cy) + ih * cz);
                            float prod = a[cx + kw * (cy + kh * cz)] * s;
                            *d = *d + prod;
                        }
            }
}

/* CNN::LConv::update, cnn.h:269-279: per output position, W[co] += patch *
 * (-alpha*er); B[co] -= er*alpha.  With W,B zeroed and alpha = -1 this yields
 * the reference's own gradient sum (used by orc_grads). */
static void conv_update(const float *X, int iw, int ih, int cin,
                        const float *E, int ow, int oh, int cout,
                        float *W, float *B, int kw, int kh, float alpha)
{
    for (int z = 0; z < cout; z++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++) {
                float er = E[x + ow * (y + oh * z)];
                float s = -alpha * er;
                float *d = W + (size_t)z * kw * kh * cin;
                for (int cz = 0; cz < cin; cz++)
                    for (int cy = 0; cy < kh; cy++)
                        for (int cx = 0; cx < kw; cx++) {
                            float prod = X[(x + cx) + iw * ((y + cy) + ih * cz)] * s;
                            float *dd = d + cx + kw * (cy + kh * cz);
                            *dd = *dd + prod;
                        }
                float be = er * alpha;
                B[z] = B[z] - be;
            }
}

/* TanH::f, cnn.h:31: e = exp(2t); (e-1)/(e+1).  NaN for t >~ 44.4. */
static float tanh_f(float t)
{
    float e = expf(2 * t);
    return (e - 1) / (e + 1);
}
/* TanH::df, cnn.h:32 (takes the OUTPUT value). */
static float tanh_df(float y) { return 1.0f - y * y; }

/* LActivation<TanH>::forward cnn.h:460-463 / backward cnn.h:464-469 */
static void act_forward(const float *in, float *out, int n)
{
    for (int i = 0; i < n; i++) out[i] = tanh_f(in[i]);
}
static void act_backward(const float *Y, const float *E, float *D, int n)
{
    for (int i = 0; i < n; i++) D[i] = tanh_df(Y[i]) * E[i];
}

/* std::max(a,b) == (a<b)?b:a -- matters only for NaN propagation. */
static float std_max(float a, float b) { return (a < b) ? b : a; }

/* CNN::LMaxPool::forward, cnn.h:141-148 */
static void pool_forward(const float *in, int iw, int ih, int c, float *out)
{
    int ow = iw / 2, oh = ih / 2;
    for (int z = 0; z < c; z++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++) {
                const float *p = in + 2 * x + iw * (2 * y + ih * z);
                out[x + ow * (y + oh * z)] =
                    std_max(std_max(std_max(p[0], p[1]), p[iw]), p[iw + 1]);
            }
}
/* CNN::LMaxPool::backward, cnn.h:149-164: first STRICT maximum in scan order
 * (0,0),(1,0),(0,1),(1,1); D zero elsewhere. */
static void pool_backward(const float *X, int iw, int ih, int c, const float *E, float *D)
{
    int ow = iw / 2, oh = ih / 2;
    memset(D, 0, sizeof(float) * (size_t)iw * ih * c);
    for (int z = 0; z < c; z++)
        for (int y = 0; y < oh; y++)
            for (int x = 0; x < ow; x++) {
                int base = 2 * x + iw * (2 * y + ih * z);
                int mx = base;
                for (int vy = 0; vy < 2; vy++)
                    for (int vx = 0; vx < 2; vx++) {
                        int q = base + vx + iw * vy;
                        if (X[q] > X[mx]) mx = q;
                    }
                D[mx] = E[x + ow * (y + oh * z)];
            }
}

/* CNN::LFull::forward, cnn.h:405-429: Y = B; for i: Y[j] += x[i]*W[i*N+j]. */
static void full_forward(const float *x, int M, const float *W, const float *B, int N, float *Y)
{
    for (int j = 0; j < N; j++) Y[j] = B[j];
    const float *w = W;
    for (int i = 0; i < M; i++) {
        float xi = x[i];
        for (int j = 0; j < N; j++) {
            float prod = xi * w[j];
            Y[j] = Y[j] + prod;
        }
        w += N;
    }
}
/* CNN::LFull::backward, cnn.h:430-437 */
static void full_backward(const float *W, const float *E, int M, int N, float *D)
{
    for (int i = 0; i < M; i++) {
        float acc = 0.0f;
        for (int j = 0; j < N; j++) {
            float prod = W[j + (size_t)i * N] * E[j];
            acc = acc + prod;
        }
        D[i] = acc;
    }
}
/* CNN::LFull::update, cnn.h:438-445: (X[i]*E[j])*alpha */
static void full_update(const float *X, const float *E, int M, int N, float *W, float *B, float alpha)
{
    for (int j = 0; j < N; j++) {
        float d = E[j] * alpha;
        B[j] = B[j] - d;
    }
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            float p = X[i] * E[j];
            float d = p * alpha;
            W[(size_t)i * N + j] = W[(size_t)i * N + j] - d;
        }
}

/* CNN::LSoftMaxChunked::forward, cnn.h:497-511: exp WITHOUT max subtraction. */
static void softmax_chunked_forward(const float *in, float *out)
{
    for (int i = 0; i < N_OUT; i++) out[i] = expf(in[i]);
    int base = 0;
    for (int s = 0; s < N_SPANS; s++) {
        float sum = 0.0f;
        for (int i = base; i < base + k_spans[s]; i++) sum += out[i];
        for (int i = base; i < base + k_spans[s]; i++) out[i] /= sum;
        base += k_spans[s];
    }
}
/* CNN::LSoftMaxChunked::backward, cnn.h:512-526 */
static void softmax_chunked_backward(const float *Y, const float *E, float *D)
{
    int base = 0;
    for (int s = 0; s < N_SPANS; s++) {
        float dp = 0.0f;
        for (int i = base; i < base + k_spans[s]; i++) {
            float p = E[i] * Y[i];
            dp += p;
        }
        for (int i = base; i < base + k_spans[s]; i++) D[i] = Y[i] * (E[i] - dp);
        base += k_spans[s];
    }
}

/* ---- whole-net workspace ------------------------------------------------- */
typedef struct {
    float a0[C1_CO * C1_H * C1_W]; /* conv1 out   57600 */
    float a1[C1_CO * C1_H * C1_W]; /* tanh        57600 */
    float a2[C1_CO * P1_H * P1_W]; /* pool        14400 */
    float a3[C1_CO * P2_H * P2_W]; /* pool         3600 */
    float a4[C2_CO * C2_H * C2_W]; /* conv2        9216 */
    float a5[C2_CO * C2_H * C2_W]; /* tanh         9216 */
    float a6[FC1_IN];              /* pool         2304 */
    float a7[FC1_OUT];             /* fc1          2048 */
    float a8[FC1_OUT];             /* tanh         2048 */
    float a9[FC2_OUT];             /* fc2          2304 */
    float a10[N_OUT];              /* softmax      2304 */
    /* errors[i] = dLoss/d(output of layer i), cnn.h:564-572 */
    float e10[N_OUT], e9[N_OUT], e8[FC1_OUT], e7[FC1_OUT], e6[FC1_IN];
    float e5[C2_CO * C2_H * C2_W], e4[C2_CO * C2_H * C2_W];
    float e3[C1_CO * P2_H * P2_W], e2[C1_CO * P1_H * P1_W];
    float e1[C1_CO * C1_H * C1_W], e0[C1_CO * C1_H * C1_W];
} orc_ws;

ORC_API void *orc_ws_create(void) { return calloc(1, sizeof(orc_ws)); }
ORC_API void orc_ws_destroy(void *ws) { free(ws); }

/* CNN::Eval forward chain, cnn.h:550-556 */
static void forward_all(const float *P, const float *x, orc_ws *w)
{
    conv_forward(x, IN_W, IN_H, 1, P + OFF_C1W, P + OFF_C1B, C1_K, C1_K, C1_CO, w->a0, C1_W, C1_H);
    act_forward(w->a0, w->a1, C1_CO * C1_H * C1_W);
    pool_forward(w->a1, C1_W, C1_H, C1_CO, w->a2);
    pool_forward(w->a2, P1_W, P1_H, C1_CO, w->a3);
    conv_forward(w->a3, P2_W, P2_H, C2_CI, P + OFF_C2W, P + OFF_C2B, C2_K, C2_K, C2_CO, w->a4, C2_W, C2_H);
    act_forward(w->a4, w->a5, C2_CO * C2_H * C2_W);
    pool_forward(w->a5, C2_W, C2_H, C2_CO, w->a6);
    full_forward(w->a6, FC1_IN, P + OFF_F1W, P + OFF_F1B, FC1_OUT, w->a7);
    act_forward(w->a7, w->a8, FC1_OUT);
    full_forward(w->a8, FC2_IN, P + OFF_F2W, P + OFF_F2B, FC2_OUT, w->a9);
    softmax_chunked_forward(w->a9, w->a10);
}

/* loss + backward chain, cnn.h:564-572 (layer 0's backward is never called) */
static float backward_all(const float *P, const float *t, orc_ws *w)
{
    float mse = 0;
    for (int i = 0; i < N_OUT; i++) {
        float e = w->a10[i] - t[i];
        mse += e * e;
        w->e10[i] = e;
    }
    mse /= N_OUT;
    softmax_chunked_backward(w->a10, w->e10, w->e9);
    full_backward(P + OFF_F2W, w->e9, FC2_IN, FC2_OUT, w->e8);
    act_backward(w->a8, w->e8, w->e7, FC1_OUT);
    full_backward(P + OFF_F1W, w->e7, FC1_IN, FC1_OUT, w->e6);
    pool_backward(w->a5, C2_W, C2_H, C2_CO, w->e6, w->e5);
    act_backward(w->a5, w->e5, w->e4, C2_CO * C2_H * C2_W);
    conv_backward(w->e4, C2_W, C2_H, C2_CO, P + OFF_C2W, C2_K, C2_K, C2_CI, w->e3, P2_W, P2_H);
    pool_backward(w->a2, P1_W, P1_H, C1_CO, w->e3, w->e2);
    pool_backward(w->a1, C1_W, C1_H, C1_CO, w->e2, w->e1);
    act_backward(w->a1, w->e1, w->e0, C1_CO * C1_H * C1_W);
    return mse;
}

/* update chain, cnn.h:574-575 */
static void update_all(float *P, const float *x, orc_ws *w, float alpha)
{
    conv_update(x, IN_W, IN_H, 1, w->e0, C1_W, C1_H, C1_CO, P + OFF_C1W, P + OFF_C1B, C1_K, C1_K, alpha);
    conv_update(w->a3, P2_W, P2_H, C2_CI, w->e4, C2_W, C2_H, C2_CO, P + OFF_C2W, P + OFF_C2B, C2_K, C2_K, alpha);
    full_update(w->a6, w->e7, FC1_IN, FC1_OUT, P + OFF_F1W, P + OFF_F1B, alpha);
    full_update(w->a8, w->e9, FC2_IN, FC2_OUT, P + OFF_F2W, P + OFF_F2B, alpha);
}

/* ---- public entry points -------------------------------------------------- */

/* CNN::Eval (cnn.h:550) over n crops; params in .cnnb order. */
ORC_API void orc_eval(const float *params, const float *x, long n, float *y, void *ws_)
{
    orc_ws *w = (orc_ws *)ws_;
    for (long b = 0; b < n; b++) {
        forward_all(params, x + b * N_IN, w);
        memcpy(y + b * N_OUT, w->a10, sizeof(float) * N_OUT);
    }
}

/* CNN::Train (cnn.h:558-580), n sequential batch-1 steps exactly as
 * train-cnn.cpp:160 issues them; params updated in place; mse[b] per step. */
ORC_API void orc_train_seq(float *params, const float *x, const float *t, long n, float alpha,
                           float *mse, void *ws_)
{
    orc_ws *w = (orc_ws *)ws_;
    for (long b = 0; b < n; b++) {
        forward_all(params, x + b * N_IN, w);
        float m = backward_all(params, t + b * N_OUT, w);
        update_all(params, x + b * N_IN, w, alpha);
        if (mse) mse[b] = m;
    }
}

/* Per-sample gradient at frozen weights, in .cnnb order: what `update` would
 * subtract per unit alpha.  Obtained the way the reference itself would
 * produce it: run `update` (cnn.h:269-279, 438-445) on zeroed parameters with
 * alpha = -1, so W becomes +sum(X*E) in the reference's own summation order. */
ORC_API float orc_grad_sample(const float *params, const float *x, const float *t, float *grad,
                              void *ws_)
{
    orc_ws *w = (orc_ws *)ws_;
    forward_all(params, x, w);
    float m = backward_all(params, t, w);
    memset(grad, 0, sizeof(float) * N_PARAMS);
    update_all(grad, x, w, -1.0f);
    return m;
}

/* Minibatch step as the new batched entry point defines it:
 * g = sum_b g_b at frozen weights (accumulated here in double, the more
 * accurate reference for a sum), then W -= alpha*g.  grad_sum (double,
 * N_PARAMS) is returned for gradient parity checks; params updated iff
 * apply != 0.  mse[b] per sample. */
ORC_API void orc_train_minibatch(float *params, const float *x, const float *t, long n, float alpha,
                                 double *grad_sum, float *mse, int apply, void *ws_)
{
    float *g = (float *)malloc(sizeof(float) * N_PARAMS);
    memset(grad_sum, 0, sizeof(double) * N_PARAMS);
    for (long b = 0; b < n; b++) {
        float m = orc_grad_sample(params, x + b * N_IN, t + b * N_OUT, g, ws_);
        if (mse) mse[b] = m;
        for (int i = 0; i < N_PARAMS; i++) grad_sum[i] += (double)g[i];
    }
    if (apply)
        for (int i = 0; i < N_PARAMS; i++) params[i] = params[i] - (float)((double)alpha * grad_sum[i]);
    free(g);
}

/* Intermediate activations for per-stage parity tests: which = 3 (pooled
 * conv1 stage, 3600), 6 (pooled conv2 stage, 2304), 8 (fc1+tanh, 2048),
 * 9 (fc2 logits, 2304), 10 (softmax, 2304); errors: 109 (d logits), 107
 * (d fc1 pre-activation), 106 (d flatten), 104 (d conv2 pre-activation),
 * 103 (d pooled conv1), 100 (d conv1 pre-activation). Valid after
 * orc_grad_sample / orc_eval with the same workspace. */
ORC_API int orc_peek(void *ws_, int which, float *out)
{
    orc_ws *w = (orc_ws *)ws_;
    switch (which) {
    case 0: memcpy(out, w->a0, sizeof w->a0); return sizeof w->a0 / 4;
    case 1: memcpy(out, w->a1, sizeof w->a1); return sizeof w->a1 / 4;
    case 3: memcpy(out, w->a3, sizeof w->a3); return sizeof w->a3 / 4;
    case 5: memcpy(out, w->a5, sizeof w->a5); return sizeof w->a5 / 4;
    case 6: memcpy(out, w->a6, sizeof w->a6); return sizeof w->a6 / 4;
    case 8: memcpy(out, w->a8, sizeof w->a8); return sizeof w->a8 / 4;
    case 9: memcpy(out, w->a9, sizeof w->a9); return sizeof w->a9 / 4;
    case 10: memcpy(out, w->a10, sizeof w->a10); return sizeof w->a10 / 4;
    case 109: memcpy(out, w->e9, sizeof w->e9); return sizeof w->e9 / 4;
    case 107: memcpy(out, w->e7, sizeof w->e7); return sizeof w->e7 / 4;
    case 106: memcpy(out, w->e6, sizeof w->e6); return sizeof w->e6 / 4;
    case 104: memcpy(out, w->e4, sizeof w->e4); return sizeof w->e4 / 4;
    case 103: memcpy(out, w->e3, sizeof w->e3); return sizeof w->e3 / 4;
    case 100: memcpy(out, w->e0, sizeof w->e0); return sizeof w->e0 / 4;
    default: return -1;
    }
}

/* CNN::Init, cnn.h:581-586 with LConv::init :280-285 and LFull::init :446-451.
 * std::default_random_engine under libstdc++ is minstd_rand0
 * (x <- 16807*x mod 2^31-1, seed 1); a fresh
 * uniform_real_distribution<float>(-r, r) per weight draws ONE engine value:
 * generate_canonical<float,24> = float(u - 1) / float(2147483646) (clamped
 * below 1), result = canon*(b-a)+a.  One engine is shared across layers;
 * biases stay 0.  Known answers: conv1.W[0] = -0.118816,
 * conv2.W[0] = -0.0156417, fc1.W[0] = -0.0280718, fc2.W[0] = -0.0141025. */
static uint32_t g_rng;
static float next_uniform(float a, float b)
{
    g_rng = (uint32_t)(((uint64_t)g_rng * 16807u) % 2147483647u);
    float sum = (float)(g_rng - 1u);
    float tmp = (float)2147483646.0L;
    float ret = sum / tmp;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret * (b - a) + a;
}
ORC_API void orc_init_xavier(float *params)
{
    memset(params, 0, sizeof(float) * N_PARAMS);
    g_rng = 1u;
    float r;
    r = sqrtf(6.0f / (C1_K * C1_K * 1 + C1_K * C1_K * C1_CO));
    for (int i = 0; i < C1_CO * C1_K * C1_K; i++) params[OFF_C1W + i] = next_uniform(-r, r);
    r = sqrtf(6.0f / (C2_K * C2_K * C2_CI + C2_K * C2_K * C2_CO));
    for (int i = 0; i < C2_CO * C2_CI * C2_K * C2_K; i++) params[OFF_C2W + i] = next_uniform(-r, r);
    r = sqrtf(6.0f / (FC1_IN + FC1_OUT));
    for (int i = 0; i < FC1_IN * FC1_OUT; i++) params[OFF_F1W + i] = next_uniform(-r, r);
    r = sqrtf(6.0f / (FC2_IN + FC2_OUT));
    for (int i = 0; i < FC2_IN * FC2_OUT; i++) params[OFF_F2W + i] = next_uniform(-r, r);
}

/* ---- CNN output decode -------------------------------------------------
 * The numeric core of CNNOutputAnalysis::CNNOutputAnalysis (include/handtrack.h:218-241):
 * per 2-D heatmap ImageFindMax (first maximum in raster order, misc_image.h:298-305), PeakSubPixel
 * (3x3 weighted centroid, misc_image.h:312-324), PeakVolume (3x3 sum around the rounded centroid,
 * misc_image.h:328-336) and the peak value; per 1-D heatmap std::max_element + PeakSubPixel1D
 * (misc_image.h:340-350, 389-399).  out[48] = 8 x (px, py, confidence, peak) then 16 values.
 * Pinned bit-exactly against oracle/_ref/libpostref.so (tests/test_oracle.py). */
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
ORC_API void orc_decode(const float *y, long n, float *out)
{
    for (long b = 0; b < n; b++) {
        const float *yy = y + b * N_OUT;
        float *o = out + b * 48;
        for (int i = 0; i < 8; i++) {
            const float *m = yy + 256 * i;
            int bx = 0, by = 0;
            for (int py = 0; py < 16; py++)
                for (int px = 0; px < 16; px++)
                    if (m[py * 16 + px] > m[by * 16 + bx]) { bx = px; by = py; }
            float wsum = 0.0f, vx = 0.0f, vy = 0.0f;
            for (int sy = imax(0, by - 1); sy < imin(16, by + 2); sy++)
                for (int sx = imax(0, bx - 1); sx < imin(16, bx + 2); sx++) {
                    float w = m[sy * 16 + sx];
                    float tx = (float)sx * w, ty = (float)sy * w;
                    vx = vx + tx;
                    vy = vy + ty;
                    wsum += w;
                }
            float px = (wsum == 0) ? (float)bx : vx / wsum;
            float py = (wsum == 0) ? (float)by : vy / wsum;
            int rx = (int)(px + 0.5f), ry = (int)(py + 0.5f);
            float vol = 0.0f;
            for (int sy = imax(0, ry - 1); sy < imin(16, ry + 2); sy++)
                for (int sx = imax(0, rx - 1); sx < imin(16, rx + 2); sx++) vol += m[sy * 16 + sx];
            o[4 * i + 0] = px;
            o[4 * i + 1] = py;
            o[4 * i + 2] = vol;
            o[4 * i + 3] = m[16 * by + bx];
        }
        for (int row = 0; row < 16; row++) {
            const float *r = yy + 2048 + 16 * row;
            int p = 0;
            for (int x = 1; x < 16; x++)
                if (r[p] < r[x]) p = x; /* std::max_element: first of the largest */
            float v = 0.0f, wsum = 0.0f;
            for (int i = imax(0, p - 1); i < imin(16, p + 2); i++) {
                float w = r[i];
                float tt = (float)i * w;
                v = v + tt;
                wsum += w;
            }
            o[32 + row] = ((wsum == 0) ? (float)p : v / wsum) / (float)(16 - 1);
        }
    }
}

/* ---- depth -> [0,1] crop normalisation, include/handtrack.h:700 ---------
 * (float)clamp(1.0f - (d*depth_scale - dmin) / (dmax - dmin), 0.0f, 1.0f), clamp = min(max(a,mn),mx)
 * (third_party/geometric.h:62). */
ORC_API void orc_normalize_depth(const unsigned short *d, long count, float depth_scale, float dmin, float dmax, float *out)
{
    for (long i = 0; i < count; i++) {
        float z = (float)d[i] * depth_scale;
        float a = 1.0f - (z - dmin) / (dmax - dmin);
        a = a < 0.0f ? 0.0f : a; /* std::max(a, 0) */
        a = 1.0f < a ? 1.0f : a; /* std::min(.., 1) */
        out[i] = a;
    }
}

/* ---- SURVEY.md 8f row 3, second half: SampleD (include/misc_image.h:154-162) ------------------------------------
 * The rotated / scaled point resample at the end of HandSegmentVR (include/handtrack.h:343): for every destination
 * pixel p, v = pose * deprojectz(p, 1) in the source camera's frame, pp = (int2)projectz_src(v) (truncation), and the
 * sampled depth is re-expressed as distance from the DESTINATION image plane: (T)dot(ppdir, deprojectz_src(pp, d)).
 * Arithmetic follows the reference expression by expression (third_party/linalg.h:284-288 for the quaternion rotation,
 * misc_image.h:48-50 for the camera maps), separately rounded, left to right.  The two float -> integer conversions
 * are undefined behaviour in C++ when out of range; the pinned x86 build resolves them with cvttss2si (INT_MIN for NaN
 * and out-of-range values, then the low 16 bits for the unsigned short), which is restated explicitly here. */
static int cvtt_x86(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000u; }
ORC_API void orc_sample_d(const unsigned short *src, int w, int h, float sfx, float sfy, float spx, float spy, int dw, int dh, float dfx,
                          float dfy, float dpx, float dpy, const float *pose7, unsigned short background, unsigned short *out)
{
    const float px = pose7[0], py = pose7[1], pz = pose7[2];
    const float qx = pose7[3], qy = pose7[4], qz = pose7[5], qw = pose7[6];
    /* qxdir / qydir / qzdir, linalg.h:284-286 */
    const float xd[3] = {qw * qw + qx * qx - qy * qy - qz * qz, (qx * qy + qz * qw) * 2, (qz * qx - qy * qw) * 2};
    const float yd[3] = {(qx * qy - qz * qw) * 2, qw * qw - qx * qx + qy * qy - qz * qz, (qy * qz + qx * qw) * 2};
    const float zd[3] = {(qz * qx + qy * qw) * 2, (qy * qz - qx * qw) * 2, qw * qw - qx * qx - qy * qy + qz * qz};
    const float pos[3] = {px, py, pz};
    float ppdir[3];
    {   /* ppdir = pose * deprojectz(principal, 1) */
        const float c[3] = {(dpx - dpx) / dfx * 1.0f, (dpy - dpy) / dfy * 1.0f, 1.0f * 1.0f};
        for (int k = 0; k < 3; k++) ppdir[k] = pos[k] + ((xd[k] * c[0] + yd[k] * c[1]) + zd[k] * c[2]);
    }
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            const float c[3] = {((float)x - dpx) / dfx * 1.0f, ((float)y - dpy) / dfy * 1.0f, 1.0f * 1.0f};
            float v[3];
            for (int k = 0; k < 3; k++) v[k] = pos[k] + ((xd[k] * c[0] + yd[k] * c[1]) + zd[k] * c[2]);
            const int ix = cvtt_x86(v[0] / v[2] * sfx + spx), iy = cvtt_x86(v[1] / v[2] * sfy + spy);
            unsigned short r = background;
            if (ix >= 0 && ix <= w - 1 && iy >= 0 && iy <= h - 1) {
                const float d = (float)src[(long)iy * w + ix];
                const float s[3] = {((float)ix - spx) / sfx * d, ((float)iy - spy) / sfy * d, 1.0f * d};
                r = (unsigned short)(cvtt_x86((ppdir[0] * s[0] + ppdir[1] * s[1]) + ppdir[2] * s[2]) & 0xffff);
            }
            out[y * dw + x] = r;
        }
}

/* ---- label rendering ----------------------------------------------------
 * GatherHandExpectedCNN's label vector (include/handtrack.h:160-173) from 8 image feature points and 16 key
 * values: RenderHeatMap (misc_image.h:259-270: 5x5 window around (int)peak, exp(-d2/(2*0.33)), ToGrayScale =
 * (u8)clamp(255x,0,255), misc_image.h:169), NormalizeHeatMap (integer c*255/sum, misc_image.h:248-257),
 * Render1DHeatMaps (misc_image.h:279-295: exp(-d2/(2*0.5)), per-row integer normalisation), GrayScaleToFloat
 * (c/255.0f, misc_image.h:171).  points[n][8][2], vals[n][16] -> t[n][2304]. */
static unsigned char to_gray(float x)
{
    float y = x * 255.0f;
    y = y < 0.0f ? 0.0f : y;
    y = 255.0f < y ? 255.0f : y;
    return (unsigned char)y;
}
ORC_API void orc_render_labels(const float *points, const float *vals, long n, float *t)
{
    for (long b = 0; b < n; b++) {
        float *o = t + b * N_OUT;
        for (int i = 0; i < 8; i++) {
            unsigned char h[256];
            memset(h, 0, sizeof h);
            float pkx = points[b * 16 + 2 * i], pky = points[b * 16 + 2 * i + 1];
            int hx = (int)pkx, hy = (int)pky;
            for (int py = imax(0, hy - 2); py < imin(16, hy + 3); py++)
                for (int px = imax(0, hx - 2); px < imin(16, hx + 3); px++) {
                    float dx = pkx - (float)px, dy = pky - (float)py;
                    float xx = dx * dx, yy = dy * dy;
                    float d2 = xx + yy;
                    h[py * 16 + px] = to_gray(expf(-d2 / (2.0f * 0.33f)));
                }
            int sum = 0;
            for (int k = 0; k < 256; k++) sum += h[k];
            if (sum)
                for (int k = 0; k < 256; k++) h[k] = (unsigned char)(h[k] * 255 / sum);
            for (int k = 0; k < 256; k++) o[i * 256 + k] = h[k] / 255.0f;
        }
        for (int y = 0; y < 16; y++) {
            unsigned char r[16];
            memset(r, 0, sizeof r);
            float v = vals[b * 16 + y] * (float)(16 - 1);
            int sum = 0;
            for (int x = imax(0, (int)v - 2); x < imin(16, (int)v + 3); x++) {
                float d = (float)x - v;
                float d2 = d * d; /* pow(.,2.0f): exact square rounded once */
                r[x] = to_gray(expf(-d2 / (2.0f * 0.5f)));
                sum += r[x];
            }
            for (int x = imax(0, (int)v - 2); sum && x < imin(16, (int)v + 3); x++) r[x] = (unsigned char)(r[x] * 255 / sum);
            for (int x = 0; x < 16; x++) o[2048 + y * 16 + x] = r[x] / 255.0f;
        }
    }
}
