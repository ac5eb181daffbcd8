// hp_common.cuh -- geometry of the handposedd net, the device weight store and
// the per-net workspace shared by the FP32 and tensor-core kernel files.
//
// Reference: IntelRealSense/hand_tracking_samples include/handtrack.h:108-118
// (layer list) and third_party/cnn.h (layer arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hp {

// ---- geometry (include/handtrack.h:108-118) --------------------------------
constexpr int IN_W = 64, IN_H = 64, N_IN = 4096;
constexpr int C1_K = 5, C1_CO = 16, C1_W = 60, C1_H = 60;
constexpr int P1_W = 15, P1_H = 15, P1_N = C1_CO * P1_W * P1_H;   // 3600: after the two 2x2 pools
constexpr int C2_K = 4, C2_CI = 16, C2_CO = 64, C2_W = 12, C2_H = 12, C2_POS = C2_W * C2_H;  // 144
constexpr int C2_KDIM = C2_CI * C2_K * C2_K;                      // 256
constexpr int P2_W = 6, P2_H = 6, P2_N = C2_CO * P2_W * P2_H;     // 2304
constexpr int FC1_IN = 2304, FC1_OUT = 2048, FC2_IN = 2048, FC2_OUT = 2304;
constexpr int N_OUT = 2304;
constexpr int N_BIG_SPANS = 8, BIG_SPAN = 256, N_SMALL_SPANS = 16, SMALL_SPAN = 16;

// ---- .cnnb float offsets (cnn.h:97-98,288-289,454-455,590-593) --------------
constexpr int OFF_C1W = 0;
constexpr int OFF_C1B = 400;
constexpr int OFF_C2W = 416;
constexpr int OFF_C2B = 16800;
constexpr int OFF_F1W = 16864;
constexpr int OFF_F1B = 4735456;
constexpr int OFF_F2W = 4737504;
constexpr int OFF_F2B = 9456096;
constexpr int N_PARAMS = 9458400;

// ---- workspace: activations kept between forward and backward ---------------
struct Workspace {
    int64_t cap = 0;          // crops this workspace can hold
    // forward (FP32 path; CHW planar like the reference unless noted)
    float *p1 = nullptr;      // [cap][16][15][15] tanh(conv1) after both pools
    uint8_t *idx1 = nullptr;  // [cap][3600] winning offset oy*4+ox inside the 4x4 window (hierarchical tie-break)
    float *col = nullptr;     // [cap*144][256] im2col of p1, k = (ky*4+kx)*16+ci (the reference's accumulation order)
    float *w2p = nullptr;     // [64][256] conv2.W permuted to that k order (refreshed every forward)
    float *c2 = nullptr;      // [cap*144][64] conv2 pre-activation; reused as dense dL/dc2 in backward
    float *p2 = nullptr;      // [cap][2304] tanh(conv2) pooled, index x + 6*y + 36*c (the reference's flatten)
    uint8_t *idx2 = nullptr;  // [cap][2304] winning offset oy*2+ox inside the 2x2 window
    float *h1 = nullptr;      // [cap][2048] tanh(fc1)
    float *logits = nullptr;  // [cap][2304]
    float *y = nullptr;       // [cap][2304] softmax output (training forward)
    // backward
    float *dlog = nullptr;    // [cap][2304] dL/dlogits
    float *da1 = nullptr;     // [cap][2048] dL/d(fc1 pre-activation)
    float *g2 = nullptr;      // [cap][2304] dL/d(conv2 pre-activation) at the pool winners
    float *colgrad = nullptr; // [cap*144][256]
    float *g1 = nullptr;      // [cap][3600] dL/d(conv1 pre-activation) at the pool winners
    float *partial = nullptr; // split-K / column-sum partials
    size_t partial_floats = 0;
    // tensor-core path (bf16 activations)
    __nv_bfloat16 *p2_bf = nullptr;  // [cap][2304]
    __nv_bfloat16 *h1_bf = nullptr;  // [cap][2048]
};

// error plumbing (hp_api.cu)
void set_error(const char *fmt, ...);
#define HP_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            hp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                        \
            return 3; /* HP_ERR_CUDA */                                                     \
        }                                                                                   \
    } while (0)

// TanH::f, cnn.h:31: (exp(2t)-1)/(exp(2t)+1) -- NOT tanhf: NaN above ~44.4 and
// cancellation near 0 are part of the reference's results (SURVEY.md 8a note 2).
// The FP32 path evaluates exp through double precision, (float)exp((double)x): a
// correctly rounded expf, which is what glibc's expf (<= 0.502 ULP) returns for all but a
// ~0.4 % sliver of inputs.  CUDA's own expf (2 ULP) would break the reference's exact
// post-tanh ties differently and so pick different max-pool winners in backward.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float tanh_ref(float t)
{
    float e = exp_cr(2.0f * t);
    return (e - 1.0f) / (e + 1.0f);
}
// std::max(a,b) == (a<b)?b:a (cnn.h:146): only differs from fmaxf for NaN.
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

// ---- FP32 path launchers (hp_fp32.cu); all return an hp_status ---------------
struct Net;
int fp32_forward(Net &net, const float *x, int64_t n, float *y_out, bool training, cudaStream_t s);
int fp32_backward(Net &net, const float *x, const float *t, int64_t n, float *mse, bool accumulate,
                  cudaStream_t s);
int sgd_apply(Net &net, float alpha, cudaStream_t s);
int sgd_apply_range(Net &net, float alpha, int off, int count, cudaStream_t s);
int grads_to_bf16(Net &net, int off, int count, cudaStream_t s);
int grads_from_bf16(Net &net, int off, int count, cudaStream_t s);
int fp32_init_attributes();
// ---- pre/post steps (hp_post.cu) ----------------------------------------------
int post_normalize_depth(Net &net, const uint16_t *d, int64_t n, float depth_scale, float dmin, float dmax, float *x, cudaStream_t s);
int post_decode(Net &net, const float *y, int64_t n, float *out, cudaStream_t s);
int post_sample_d(Net &net, const uint16_t *frames, int w, int h, const float *intr, const int32_t *frame_of_crop, const float *cams, int64_t n,
                  uint16_t background, uint16_t *out, cudaStream_t s);
int post_render_labels(Net &net, const float *points, const float *vals, int64_t n, float *t, cudaStream_t s);
// ---- tensor-core path launchers (hp_tc.cu) -----------------------------------
int tc_forward(Net &net, const float *x, int64_t n, float *y_out, cudaStream_t s);
int tc_forward_u16(Net &net, const uint16_t *depth, float depth_scale, float dmin, float dmax, int64_t n, float *y_out, cudaStream_t s);
int tc_forward_decode(Net &net, const float *x, const uint16_t *x16, float depth_scale, float dmin, float dmax, int64_t n, float *y_out, float *dec_out,
                      cudaStream_t s);
int tc_refresh_weights(Net &net, cudaStream_t s);
int tc_sgd_refresh_fc(Net &net, int bucket, float alpha, cudaStream_t s);   // one GPU: SGD + both shadows of FC bucket 0 / 1 in one pass
int tc_sgd_conv_images(Net &net, float alpha, cudaStream_t s);             // one GPU: SGD on the conv bucket + its 16-bit images
int tc_refresh_bucket(Net &net, int bucket, cudaStream_t s);   // 0: fc2 shadows, 1: fc1 shadows, 2: conv images
int tc_init(Net &net);
void tc_destroy(Net &net);

struct TcState;  // opaque: tensor maps + bf16 shadow weights (hp_tc.cu)
struct PeerState;  // opaque: peers' weight/gradient stores mapped over NVLink (hp_peer.cu)

struct Net {
    int refcount = 1;
    int device = 0;
    cudaStream_t stream = nullptr;        // internal stream for the HOST-buffer entry points
    cudaStream_t comm_stream = nullptr;   // gradient all-reduce stream (H2D copy stream of the host-buffer Eval)
    cudaStream_t d2h_stream = nullptr;    // D2H copy stream of the host-buffer Eval; SGD / shadow-refresh stream of the training tail
    cudaStream_t aux_stream = nullptr;    // small-batch training: the bias-gradient reductions run here, beside the GEMM chain
    cudaEvent_t ev_fork[3] = {nullptr, nullptr, nullptr}, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaStream_t hi_stream = nullptr;     // one GPU, tensor path: the step's forward -> backward chain (priority over the update tail's CTAs)
    cudaEvent_t ev_hi[2] = {nullptr, nullptr};
    float *params = nullptr;              // FP32 master weights, .cnnb order
    float *grads = nullptr;               // FP32 gradient sums, .cnnb order
    Workspace ws;
    TcState *tc = nullptr;
    bool tc_dirty = true;                 // bf16 shadows stale w.r.t. params
    bool fp32_small_call = true;          // FP32 path: the current call holds <= 64 crops in total (split-K FC kernels allowed)
    // pinned staging for HOST entry points
    float *pin_in[2] = {nullptr, nullptr};
    float *pin_out[2] = {nullptr, nullptr};
    float *dev_in[2] = {nullptr, nullptr};
    float *dev_out[2] = {nullptr, nullptr};
    float *dev_t = nullptr, *dev_mse = nullptr;
    float *dev_norm[2] = {nullptr, nullptr};   // normalised crops when the upload is 16-bit depth
    float *dev_dec[2] = {nullptr, nullptr};    // decoded peaks [chunk][48]
    float *pin_dec[2] = {nullptr, nullptr};
    int64_t stage_cap = 0, pin_in_cap = 0, pin_out_cap = 0;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_bucket[3] = {nullptr, nullptr, nullptr},
                ev_comm = nullptr, ev_dx[2] = {nullptr, nullptr}, ev_ar[3] = {nullptr, nullptr, nullptr}, ev_tail = nullptr, ev_start = nullptr;   // ev_dx[b]: the dX GEMM that reads bucket b's weights is done
    // data parallelism
    void *nccl_comm = nullptr;
    PeerState *peer = nullptr;            // NVLink peer-memory exchange (hp_dp_peer_init); takes precedence over NCCL
    int rank = 0, world = 1;
    bool dp_bf16 = false;                 // FC gradient buckets travel as bf16 (tensor-precision steps only)
    __nv_bfloat16 *grads_bf = nullptr;    // wire buffer, .cnnb order
    bool step_timing = false;
    int64_t launches = 0;
    // One optimiser step as a CUDA graph (single GPU, or data parallel over the peer-memory exchange): the ~30 launches + ~15 event hops of a batch-256 step cost more on the
    // host than the kernels run on the device.  Keyed on the call's arguments; captured the second time a key repeats.
    struct StepGraph {
        const void *x = nullptr, *t = nullptr, *mse = nullptr;
        int64_t n = 0;
        float alpha = 0.f;
        int precision = -1;
        cudaStream_t stream = nullptr;
        uint64_t alloc_epoch = 0;     // Net::alloc_epoch the captured pointers belong to
        int seen = 0;                 // consecutive calls with this key
        cudaGraphExec_t exec = nullptr;
        int64_t launches_per_step = 0;
        bool disabled = false;        // HP_NO_GRAPH=1, or a capture failed once
    } step_graph;
    uint64_t alloc_epoch = 0;         // bumped whenever a device buffer a training step touches is (re)allocated: a captured step graph
                                      // holds raw pointers / tensor maps and must not outlive them
    int64_t last_n = 0;
    // per-stage timing (hp_profile): a pool of events, (stage, begin, end) triples
    bool profiling = false;
    static constexpr int PROF_MAX = 8192;
    cudaEvent_t *prof_ev = nullptr;
    int prof_used = 0;            // events consumed
    int prof_stage[PROF_MAX / 2]; // stage id of interval i = events (2i, 2i+1)
};

// record the begin/end of a stage on stream s when profiling is on
struct StageTimer {
    Net &net;
    cudaStream_t s;
    int slot = -1;
    StageTimer(Net &n, int stage, cudaStream_t st) : net(n), s(st)
    {
        if (net.profiling && net.prof_used + 2 <= Net::PROF_MAX) {
            slot = net.prof_used;
            net.prof_used += 2;
            net.prof_stage[slot / 2] = stage;
            cudaEventRecord(net.prof_ev[slot], s);
        }
    }
    ~StageTimer()
    {
        if (slot >= 0) cudaEventRecord(net.prof_ev[slot + 1], s);
    }
};

int ensure_workspace(Net &net, int64_t n);

}  // namespace hp
