// hp_api.cu -- the C ABI of include/handposedd.h: net lifetime, the device-resident
// weight store and its .cnnb (de)serialisation, CNN::Init, the batched Eval / Train
// entry points (host-buffer variants stream through pinned staging buffers), and
// NCCL data parallelism.  Kernels live in hp_fp32.cu and hp_tc.cu.
//
// Reference interfaces replaced: CNN::{Eval,Train,Init,loadb,saveb}
// (third_party/cnn.h:550-593) and PoseInitializerCNN (include/handtrack.h:103-130).
#include "../../include/handposedd.h"
#include "hp_common.cuh"
#include "hp_tc.cuh"
#include "hp_peer.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

namespace hp {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

constexpr int64_t FP32_CHUNK = 2048;   // crops per pass of the FP32 path (im2col workspace bound)
constexpr int64_t STAGE_CHUNK = 8192;  // crops per host<->device staging chunk
// Crops per chunk of the pipelined host-buffer Eval.  The uploads run back to back on PCIe; what the pipeline cannot hide
// is the LAST chunk's compute + download, so the chunk is kept small.  Measured per 65,536 crops, tensor path, pinned
// buffers (tools/dbg/e2e_chunks.py): 8,192 -> 21.96 ms, 4,096 -> 21.23, 2,048 -> 20.78, 1,024 -> 20.63 (the bare 1 GiB upload
// takes 19.3 ms).  2,048 keeps the FP32 path's kernels at the batch size they are tiled for.
constexpr int64_t PIPE_CHUNK = 2048;

static std::mutex g_ref_mutex;

template <class T>
static int dev_alloc(T *&p, size_t count)
{
    if (p) { cudaFree(p); p = nullptr; }
    HP_CUDA_TRY(cudaMalloc((void **)&p, count * sizeof(T)));
    return 0;
}

int ensure_workspace(Net &net, int64_t n)
{
    Workspace &w = net.ws;
    if (n <= w.cap) return 0;
    int64_t cap = std::max<int64_t>(n, std::min<int64_t>(FP32_CHUNK, std::max<int64_t>(2 * w.cap, 64)));
    HP_CUDA_TRY(cudaDeviceSynchronize());   // the workspace may be in use on a caller's stream
    w.cap = 0;                              // a failed reallocation must not leave the old capacity behind
    int rc = 0;
    rc |= dev_alloc(w.p1, cap * P1_N);
    rc |= dev_alloc(w.idx1, cap * P1_N);
    rc |= dev_alloc(w.col, cap * C2_POS * C2_KDIM);
    rc |= dev_alloc(w.c2, cap * C2_POS * C2_CO);
    rc |= dev_alloc(w.p2, cap * P2_N);
    rc |= dev_alloc(w.idx2, cap * P2_N);
    rc |= dev_alloc(w.h1, cap * FC1_OUT);
    rc |= dev_alloc(w.logits, cap * N_OUT);
    rc |= dev_alloc(w.y, cap * N_OUT);
    rc |= dev_alloc(w.dlog, cap * N_OUT);
    rc |= dev_alloc(w.da1, cap * FC1_OUT);
    rc |= dev_alloc(w.g2, cap * P2_N);
    rc |= dev_alloc(w.colgrad, cap * C2_POS * C2_KDIM);
    rc |= dev_alloc(w.g1, cap * P1_N);
    if (!w.w2p) rc |= dev_alloc(w.w2p, (size_t)C2_CO * C2_KDIM);
    if (!w.partial) {
        w.partial_floats = (size_t)2 * 512 * N_OUT;   // >= conv2 split-K partials (128 x 64 x 256); also the two split-K halves of the small-batch fc2 logits
        rc |= dev_alloc(w.partial, w.partial_floats);
    }
    if (rc) return HP_ERR_CUDA;
    w.cap = cap;
    net.alloc_epoch++;
    return 0;
}

static int ensure_staging(Net &net, int64_t n, bool pin_in, bool pin_out)
{
    if (n > net.stage_cap) {
        HP_CUDA_TRY(cudaDeviceSynchronize());
        net.stage_cap = 0;   // committed again only after every allocation below has succeeded
        for (int b = 0; b < 2; b++) {
            if (dev_alloc(net.dev_in[b], n * N_IN)) return HP_ERR_CUDA;
            if (dev_alloc(net.dev_out[b], n * N_OUT)) return HP_ERR_CUDA;
        }
        if (dev_alloc(net.dev_t, n * N_OUT)) return HP_ERR_CUDA;
        if (dev_alloc(net.dev_mse, n)) return HP_ERR_CUDA;
        for (int b = 0; b < 2; b++) {
            if (dev_alloc(net.dev_norm[b], n * N_IN)) return HP_ERR_CUDA;
            if (dev_alloc(net.dev_dec[b], n * 48)) return HP_ERR_CUDA;
            if (net.pin_dec[b]) cudaFreeHost(net.pin_dec[b]);
            net.pin_dec[b] = nullptr;
            HP_CUDA_TRY(cudaMallocHost((void **)&net.pin_dec[b], n * 48 * sizeof(float)));
        }
        net.stage_cap = n;
    }
    // pinned bounce buffers only when the caller's memory is pageable
    if (pin_in && n > net.pin_in_cap) {
        net.pin_in_cap = 0;
        for (int b = 0; b < 2; b++) {
            if (net.pin_in[b]) cudaFreeHost(net.pin_in[b]);
            net.pin_in[b] = nullptr;
            HP_CUDA_TRY(cudaMallocHost((void **)&net.pin_in[b], n * N_IN * sizeof(float)));
        }
        net.pin_in_cap = n;
    }
    if (pin_out && n > net.pin_out_cap) {
        net.pin_out_cap = 0;
        for (int b = 0; b < 2; b++) {
            if (net.pin_out[b]) cudaFreeHost(net.pin_out[b]);
            net.pin_out[b] = nullptr;
            HP_CUDA_TRY(cudaMallocHost((void **)&net.pin_out[b], n * N_OUT * sizeof(float)));
        }
        net.pin_out_cap = n;
    }
    return 0;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// ---- CNN::Init (cnn.h:581-586; LConv::init :280-285; LFull::init :446-451) --------
// std::default_random_engine (libstdc++) == minstd_rand0, default seed 1, ONE engine
// shared by all layers; each weight draws one value through a fresh
// uniform_real_distribution<float>(-r, r): generate_canonical<float,24> =
// float(u - 1) / float(2147483646.0L), result = canon * (b - a) + a.
static void xavier_host(std::vector<float> &p)
{
    std::fill(p.begin(), p.end(), 0.0f);
    uint32_t state = 1u;
    auto draw = [&](float a, float b) {
        state = (uint32_t)(((uint64_t)state * 16807u) % 2147483647u);
        float canon = (float)(state - 1u) / (float)2147483646.0L;
        if (canon >= 1.0f) canon = nextafterf(1.0f, 0.0f);
        return canon * (b - a) + a;
    };
    auto fill = [&](int off, int count, int fan) {
        const float r = sqrtf(6.0f / fan);
        for (int i = 0; i < count; i++) p[off + i] = draw(-r, r);
    };
    fill(OFF_C1W, 400, C1_K * C1_K * 1 + C1_K * C1_K * C1_CO);
    fill(OFF_C2W, C2_CO * C2_KDIM, C2_K * C2_K * C2_CI + C2_K * C2_K * C2_CO);
    fill(OFF_F1W, FC1_IN * FC1_OUT, FC1_IN + FC1_OUT);
    fill(OFF_F2W, FC2_IN * FC2_OUT, FC2_IN + FC2_OUT);
}

// ---- NCCL, bound lazily so that inference has no NCCL dependency ---------------------
struct Id128 { char b[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
static int nccl_bind()
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { set_error("NCCL not found: %s", dlerror()); return HP_ERR_NCCL; }
#define BIND(field, sym)                                                                  \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                                   \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", sym); g_nccl.lib = nullptr; return HP_ERR_NCCL; }
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(AllReduce, "ncclAllReduce");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GetErrorString, "ncclGetErrorString");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
#undef BIND
    return 0;
}
#define HP_NCCL_TRY(expr)                                                               \
    do {                                                                                \
        int _r = (expr);                                                                \
        if (_r != 0) { set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r)); return HP_ERR_NCCL; } \
    } while (0)

// The tail of an optimiser step, pipelined per gradient bucket (fc2 | fc1 | conv, in the order backward produces them)
// on the side stream while the rest of backward still runs on the caller's stream:
//   wait "bucket's weight gradient done"  ->  NCCL all-reduce (data parallel only)
//   -> wait "the dX GEMM that reads these weights is done"  ->  SGD on the bucket  ->  rebuild its bf16 shadows
// The caller's stream then waits once for the whole tail.
static int finish_step(Net &net, float alpha, int precision, cudaStream_t s)
{
    const int off[3] = {OFF_F2W, OFF_F1W, 0};
    const int end[3] = {N_PARAMS, OFF_F2W, OFF_F1W};
    // all-reduces back to back on the (high-priority) comm stream; SGD + shadow refresh of each bucket on a third
    // stream, so that bucket b+1's all-reduce does not queue behind bucket b's update
    cudaStream_t cs = net.comm_stream, us = net.d2h_stream;
    const bool peer = net.world > 1 && net.peer && net.peer->ready;
    for (int b = 0; b < 3; b++) {
        const bool bf16_wire = !peer && net.world > 1 && net.dp_bf16 && precision == HP_PRECISION_TENSOR && b < 2;
        if (peer) {
            // reduce-scatter + SGD + all-gather of the updated weights in one kernel over NVLink peer memory (hp_peer.cu).
            // Peers store into this rank's FP32 master weights, so on the FP32 path (whose dX GEMMs read them) the
            // kernel must also wait for the dX GEMM of the bucket; the tensor path only reads the bf16 shadows.
            HP_CUDA_TRY(cudaStreamWaitEvent(cs, net.ev_bucket[b], 0));
            static const bool after_dx = getenv("HP_DP_EXCH_AFTER_DX") != nullptr;   // experiment: start the exchange after the bucket's dX GEMM
            if (b < 2 && (precision != HP_PRECISION_TENSOR || after_dx)) HP_CUDA_TRY(cudaStreamWaitEvent(cs, net.ev_dx[b], 0));
            if (int rc = peer_sgd_bucket(net, alpha, off[b], end[b] - off[b], cs)) return rc;
            HP_CUDA_TRY(cudaEventRecord(net.ev_ar[b], cs));
            HP_CUDA_TRY(cudaStreamWaitEvent(us, net.ev_ar[b], 0));
            if (b < 2) HP_CUDA_TRY(cudaStreamWaitEvent(us, net.ev_dx[b], 0));
            if (precision == HP_PRECISION_TENSOR)
                if (int rc = tc_refresh_bucket(net, b, us)) return rc;
            continue;
        }
        if (net.world > 1) {
            HP_CUDA_TRY(cudaStreamWaitEvent(cs, net.ev_bucket[b], 0));
            if (bf16_wire) {
                // halve the bytes on NVLink: the local sums go out as bf16 and come back as the bf16 sum over ranks
                if (int rc = grads_to_bf16(net, off[b], end[b] - off[b], cs)) return rc;
                HP_NCCL_TRY(g_nccl.AllReduce(net.grads_bf + off[b], net.grads_bf + off[b], (size_t)(end[b] - off[b]), /*ncclBfloat16*/ 9,
                                             /*ncclSum*/ 0, net.nccl_comm, cs));
                if (int rc = grads_from_bf16(net, off[b], end[b] - off[b], cs)) return rc;
            } else
            HP_NCCL_TRY(g_nccl.AllReduce(net.grads + off[b], net.grads + off[b], (size_t)(end[b] - off[b]), /*ncclFloat32*/ 7,
                                         /*ncclSum*/ 0, net.nccl_comm, cs));
            HP_CUDA_TRY(cudaEventRecord(net.ev_ar[b], cs));
            HP_CUDA_TRY(cudaStreamWaitEvent(us, net.ev_ar[b], 0));
        } else if (b == 2) {
            // one GPU: the conv bucket is the exposed end of the step and backward has just finished on the caller's
            // stream, so its update runs right there (no cross-stream hops); the side stream's FC work is joined after it
            HP_CUDA_TRY(cudaEventRecord(net.ev_tail, us));
            static const bool unfused_conv = getenv("HP_SGD_UNFUSED") != nullptr;
            if (precision == HP_PRECISION_TENSOR && !unfused_conv) {
                if (int rc = tc_sgd_conv_images(net, alpha, s)) return rc;
            } else {
                if (int rc = sgd_apply_range(net, alpha, off[b], end[b] - off[b], s)) return rc;
                if (precision == HP_PRECISION_TENSOR)
                    if (int rc = tc_refresh_bucket(net, b, s)) return rc;
            }
            HP_CUDA_TRY(cudaStreamWaitEvent(s, net.ev_tail, 0));
            net.tc_dirty = (precision != HP_PRECISION_TENSOR);
            return 0;
        } else {
            HP_CUDA_TRY(cudaStreamWaitEvent(us, net.ev_bucket[b], 0));
        }
        if (b < 2) HP_CUDA_TRY(cudaStreamWaitEvent(us, net.ev_dx[b], 0));
        static const bool unfused = getenv("HP_SGD_UNFUSED") != nullptr;   // A/B: separate SGD and shadow-refresh kernels
        if (net.world == 1 && precision == HP_PRECISION_TENSOR && b < 2 && !unfused) {
            if (int rc = tc_sgd_refresh_fc(net, b, alpha, us)) return rc;   // one pass over the bucket: update + both shadows
            continue;
        }
        if (int rc = sgd_apply_range(net, alpha, off[b], end[b] - off[b], us)) return rc;
        if (precision == HP_PRECISION_TENSOR)
            if (int rc = tc_refresh_bucket(net, b, us)) return rc;
    }
    HP_CUDA_TRY(cudaEventRecord(net.ev_tail, us));
    HP_CUDA_TRY(cudaStreamWaitEvent(s, net.ev_tail, 0));
    net.tc_dirty = (precision != HP_PRECISION_TENSOR);
    return 0;
}

// An exchange kernel of the peer-memory data-parallel path gave up waiting for a rank (hp_peer.cu): the ranks' weights
// can no longer be trusted to be in step, so training calls fail from here on instead of silently diverging.
static int check_peer(const Net &net)
{
    if (const int e = peer_failed(net)) {
        set_error("data-parallel exchange aborted: rank %d did not reach the barrier within the timeout (HP_PEER_TIMEOUT_S); "
                  "the weights were left at their last consistent state -- shut down and re-initialise the group", e - 1);
        return HP_ERR_PEER;
    }
    return 0;
}

static int check_precision(int precision)
{
    if (precision != HP_PRECISION_FP32 && precision != HP_PRECISION_TENSOR) {
        set_error("unknown precision %d", precision);
        return HP_ERR_INVALID;
    }
    return 0;
}

// forward over device buffers, chunked to the workspace
// call_n: crops of the whole API call this chunk belongs to (the host pipeline feeds 2,048-crop chunks)
static int eval_device(Net &net, const float *x, int64_t n, float *y, int precision, cudaStream_t s, int64_t call_n = -1)
{
    if (call_n < 0) call_n = n;
    if (precision == HP_PRECISION_TENSOR) {
        if (net.tc_dirty) {
            if (int rc = tc_refresh_weights(net, s)) return rc;
        }
        return tc_forward(net, x, n, y, s);
    }
    net.fp32_small_call = call_n <= 64;   // one FC summation order per call, not per chunk (hp_fp32.cu, fc_small)
    for (int64_t b = 0; b < n; b += FP32_CHUNK) {
        const int64_t m = std::min<int64_t>(FP32_CHUNK, n - b);
        if (int rc = ensure_workspace(net, m)) return rc;
        if (int rc = fp32_forward(net, x + b * N_IN, m, y + b * N_OUT, false, s)) return rc;
    }
    net.last_n = std::min<int64_t>(n, FP32_CHUNK);
    return 0;
}

// accumulate_first: add to the gradient sums already in the store (a batch that arrives in several host chunks)
static int grad_device(Net &net, const float *x, const float *t, int64_t n, float *mse, int precision, cudaStream_t s, bool accumulate_first = false)
{
    if (precision == HP_PRECISION_TENSOR) {
        for (int64_t b = 0; b < n; b += FP32_CHUNK) {
            const int64_t m = std::min<int64_t>(FP32_CHUNK, n - b);
            if (int rc = tc_train_grad(net, x + b * N_IN, t + b * N_OUT, m, mse ? mse + b : nullptr, accumulate_first || b > 0, s)) return rc;
        }
        net.last_n = std::min<int64_t>(n, FP32_CHUNK);
        return 0;
    }
    net.fp32_small_call = n <= 64 && !accumulate_first;
    for (int64_t b = 0; b < n; b += FP32_CHUNK) {
        const int64_t m = std::min<int64_t>(FP32_CHUNK, n - b);
        if (int rc = ensure_workspace(net, m)) return rc;
        if (int rc = fp32_forward(net, x + b * N_IN, m, nullptr, true, s)) return rc;
        if (int rc = fp32_backward(net, x + b * N_IN, t + b * N_OUT, m, mse ? mse + b : nullptr, accumulate_first || b > 0, s)) return rc;
    }
    net.last_n = std::min<int64_t>(n, FP32_CHUNK);
    return 0;
}

}  // namespace hp

using namespace hp;

struct hp_net {
    Net n;
};

static int check_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
        return HP_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d devices)", device, count); return HP_ERR_INVALID; }
    cudaDeviceProp prop;
    HP_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this build only contains sm_100a code", device, prop.major, prop.minor);
        return HP_ERR_NO_DEVICE;
    }
    return 0;
}

static bool is_handposedd(const hp_layer_desc *L, int n)
{
    if (n != 11 || !L) return false;
    auto conv = [](const hp_layer_desc &d, int ix, int iy, int iz, int kx, int ky, int ci, int co, int ox, int oy, int oz) {
        return d.kind == HP_LAYER_CONV && d.in_dims[0] == ix && d.in_dims[1] == iy && d.in_dims[2] == iz && d.w_dims[0] == kx &&
               d.w_dims[1] == ky && d.w_dims[2] == ci && d.w_dims[3] == co && d.out_dims[0] == ox && d.out_dims[1] == oy && d.out_dims[2] == oz;
    };
    auto pool = [](const hp_layer_desc &d, int x, int y, int z) {
        return d.kind == HP_LAYER_MAXPOOL && d.in_dims[0] == x && d.in_dims[1] == y && d.in_dims[2] == z;
    };
    auto full = [](const hp_layer_desc &d, int i, int o) { return d.kind == HP_LAYER_FULL && d.in_dims[0] == i && d.out_dims[0] == o; };
    auto act = [](const hp_layer_desc &d, int nn) { return d.kind == HP_LAYER_TANH && d.in_dims[0] == nn; };
    if (!conv(L[0], 64, 64, 1, 5, 5, 1, 16, 60, 60, 16)) return false;
    if (!act(L[1], 57600) || !pool(L[2], 60, 60, 16) || !pool(L[3], 30, 30, 16)) return false;
    if (!conv(L[4], 15, 15, 16, 4, 4, 16, 64, 12, 12, 64)) return false;
    if (!act(L[5], 9216) || !pool(L[6], 12, 12, 64)) return false;
    if (!full(L[7], 2304, 2048) || !act(L[8], 2048) || !full(L[9], 2048, 2304)) return false;
    if (L[10].kind != HP_LAYER_SOFTMAX_CHUNKED || L[10].n_spans != 24 || !L[10].spans) return false;
    for (int i = 0; i < 24; i++)
        if (L[10].spans[i] != (i < 8 ? 256 : 16)) return false;
    return true;
}

static int create_resources(Net &n)
{
    HP_CUDA_TRY(cudaStreamCreateWithFlags(&n.stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;   // NCCL's CTAs must win SM slots against the compute kernels as soon as any CTA retires
        HP_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HP_CUDA_TRY(cudaStreamCreateWithPriority(&n.comm_stream, cudaStreamNonBlocking, hi));
    }
    HP_CUDA_TRY(cudaStreamCreateWithFlags(&n.d2h_stream, cudaStreamNonBlocking));
    HP_CUDA_TRY(cudaStreamCreateWithFlags(&n.aux_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        HP_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HP_CUDA_TRY(cudaStreamCreateWithPriority(&n.hi_stream, cudaStreamNonBlocking, hi));
        for (int b = 0; b < 2; b++) HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_hi[b], cudaEventDisableTiming));
    }
    for (int b = 0; b < 3; b++) {
        HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_fork[b], cudaEventDisableTiming));
        HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_join[b], cudaEventDisableTiming));
    }
    for (int b = 0; b < 2; b++) {
        HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_in[b], cudaEventDisableTiming));
        HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_out[b], cudaEventDisableTiming));
        HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_comp[b], cudaEventDisableTiming));
    }
    // step-timeline events: timing costs ~20 us per step, so it is opt-in (hp_debug_step_times)
    n.step_timing = getenv("HP_STEP_TIMING") != nullptr;
    const unsigned evf = n.step_timing ? cudaEventDefault : cudaEventDisableTiming;
    for (int b = 0; b < 3; b++) HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_bucket[b], evf));
    HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_start, evf));
    HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_comm, cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_dx[b], evf));
    for (int b = 0; b < 3; b++) HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_ar[b], evf));
    HP_CUDA_TRY(cudaEventCreateWithFlags(&n.ev_tail, evf));
    HP_CUDA_TRY(cudaMalloc((void **)&n.params, (size_t)N_PARAMS * sizeof(float)));
    HP_CUDA_TRY(cudaMalloc((void **)&n.grads, (size_t)N_PARAMS * sizeof(float)));
    HP_CUDA_TRY(cudaMemset(n.params, 0, (size_t)N_PARAMS * sizeof(float)));
    HP_CUDA_TRY(cudaMemset(n.grads, 0, (size_t)N_PARAMS * sizeof(float)));
    if (int rc = fp32_init_attributes()) return rc;
    if (int rc = tc_init(n)) return rc;
    return HP_OK;
}


extern "C" {

int hp_create_handposedd(int device, hp_net **out)
{
    if (!out) { set_error("out is NULL"); return HP_ERR_INVALID; }
    *out = nullptr;
    if (int rc = check_device(device)) return rc;
    HP_CUDA_TRY(cudaSetDevice(device));
    hp_net *h = new hp_net;
    h->n.device = device;
    const int rc = create_resources(h->n);
    if (rc) {   // release whatever was created before the failure (hp_destroy tolerates a half-built net)
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        hp_destroy(h);
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    *out = h;
    return HP_OK;
}

int hp_create(const hp_layer_desc *layers, int n_layers, int device, hp_net **out)
{
    if (!is_handposedd(layers, n_layers)) {
        set_error("layer list is not handposedd (include/handtrack.h:108-118); no kernels for it");
        if (out) *out = nullptr;
        return HP_ERR_UNSUPPORTED;
    }
    return hp_create_handposedd(device, out);
}

int hp_retain(hp_net *net)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(g_ref_mutex);
    net->n.refcount++;
    return HP_OK;
}

int hp_destroy(hp_net *net)
{
    if (!net) return HP_OK;
    {
        std::lock_guard<std::mutex> lk(g_ref_mutex);
        if (--net->n.refcount > 0) return HP_OK;
    }
    Net &n = net->n;
    cudaSetDevice(n.device);
    cudaDeviceSynchronize();
    if (n.step_graph.exec) cudaGraphExecDestroy(n.step_graph.exec);
    if (n.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(n.nccl_comm);
    if (n.peer) peer_shutdown(n);   // normally done by hp_dp_shutdown after a host-side barrier
    tc_destroy(n);
    Workspace &w = n.ws;
    void *bufs[] = {n.params, n.grads, w.p1, w.idx1, w.col, w.c2, w.p2, w.idx2, w.h1, w.logits, w.y, w.dlog, w.da1, w.g2, w.colgrad,
                    w.g1, w.partial, w.w2p, w.p2_bf, w.h1_bf, n.dev_in[0], n.dev_in[1], n.dev_out[0], n.dev_out[1], n.dev_t, n.dev_mse, n.dev_norm[0], n.dev_norm[1], n.dev_dec[0], n.dev_dec[1]};
    for (void *p : bufs)
        if (p) cudaFree(p);
    for (int b = 0; b < 2; b++) {
        if (n.pin_in[b]) cudaFreeHost(n.pin_in[b]);
        if (n.pin_out[b]) cudaFreeHost(n.pin_out[b]);
        if (n.pin_dec[b]) cudaFreeHost(n.pin_dec[b]);
        if (n.ev_in[b]) cudaEventDestroy(n.ev_in[b]);
        if (n.ev_out[b]) cudaEventDestroy(n.ev_out[b]);
        if (n.ev_comp[b]) cudaEventDestroy(n.ev_comp[b]);
    }
    for (int b = 0; b < 3; b++)
        if (n.ev_bucket[b]) cudaEventDestroy(n.ev_bucket[b]);
    if (n.ev_comm) cudaEventDestroy(n.ev_comm);
    for (int b = 0; b < 2; b++)
        if (n.ev_dx[b]) cudaEventDestroy(n.ev_dx[b]);
    for (int b = 0; b < 3; b++)
        if (n.ev_ar[b]) cudaEventDestroy(n.ev_ar[b]);
    if (n.ev_tail) cudaEventDestroy(n.ev_tail);
    if (n.ev_start) cudaEventDestroy(n.ev_start);
    if (n.grads_bf) cudaFree(n.grads_bf);
    if (n.prof_ev) {
        for (int i = 0; i < Net::PROF_MAX; i++) cudaEventDestroy(n.prof_ev[i]);
        delete[] n.prof_ev;
    }
    if (n.stream) cudaStreamDestroy(n.stream);
    if (n.comm_stream) cudaStreamDestroy(n.comm_stream);
    if (n.d2h_stream) cudaStreamDestroy(n.d2h_stream);
    if (n.aux_stream) cudaStreamDestroy(n.aux_stream);
    if (n.hi_stream) cudaStreamDestroy(n.hi_stream);
    for (int b = 0; b < 2; b++)
        if (n.ev_hi[b]) cudaEventDestroy(n.ev_hi[b]);
    for (int b = 0; b < 3; b++) {
        if (n.ev_fork[b]) cudaEventDestroy(n.ev_fork[b]);
        if (n.ev_join[b]) cudaEventDestroy(n.ev_join[b]);
    }
    delete net;
    return HP_OK;
}

int hp_init_xavier(hp_net *net)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    std::vector<float> p(N_PARAMS);
    xavier_host(p);
    return hp_load_cnnb(net, p.data(), p.size() * sizeof(float));
}

int hp_load_cnnb(hp_net *net, const void *bytes, size_t n_bytes)
{
    if (!net || (!bytes && n_bytes)) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &n = net->n;
    HP_CUDA_TRY(cudaSetDevice(n.device));
    size_t take = std::min<size_t>(n_bytes, (size_t)HP_CNNB_BYTES);
    // like loadvb's istream::read (cnn.h:97) a short stream fills a prefix; a float torn by
    // the end of the stream is dropped here rather than half-written
    take -= take % sizeof(float);
    // the *_device entry points run on caller streams and side streams that a legacy-stream memcpy does not order against
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (take) HP_CUDA_TRY(cudaMemcpy(n.params, bytes, take, cudaMemcpyHostToDevice));
    n.tc_dirty = true;
    return HP_OK;
}

int hp_save_cnnb(const hp_net *net, void *bytes, size_t capacity, size_t *n_written)
{
    if (!net || !bytes) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (capacity < (size_t)HP_CNNB_BYTES) { set_error("buffer too small: need %d bytes", HP_CNNB_BYTES); return HP_ERR_IO; }
    const Net &n = net->n;
    HP_CUDA_TRY(cudaSetDevice(n.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (int rc = check_peer(n)) return rc;
    HP_CUDA_TRY(cudaMemcpy(bytes, n.params, (size_t)HP_CNNB_BYTES, cudaMemcpyDeviceToHost));
    if (n_written) *n_written = (size_t)HP_CNNB_BYTES;
    return HP_OK;
}

int hp_get_params_range(const hp_net *net, int64_t first, int64_t count, float *host)
{
    if (!net || !host || first < 0 || count < 0 || first + count > (int64_t)N_PARAMS) { set_error("bad argument"); return HP_ERR_INVALID; }
    const Net &n = net->n;
    HP_CUDA_TRY(cudaSetDevice(n.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (int rc = check_peer(n)) return rc;
    if (count) HP_CUDA_TRY(cudaMemcpy(host, n.params + first, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
    return HP_OK;
}

int hp_set_params_range(hp_net *net, int64_t first, int64_t count, const float *host)
{
    if (!net || !host || first < 0 || count < 0 || first + count > (int64_t)N_PARAMS) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &n = net->n;
    HP_CUDA_TRY(cudaSetDevice(n.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (count) HP_CUDA_TRY(cudaMemcpy(n.params + first, host, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
    n.tc_dirty = true;
    return HP_OK;
}

int hp_load_cnnb_file(hp_net *net, const char *path)
{
    if (!net || !path) { set_error("bad argument"); return HP_ERR_INVALID; }
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return HP_ERR_IO; }
    std::vector<char> buf((size_t)HP_CNNB_BYTES);
    size_t got = fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    return hp_load_cnnb(net, buf.data(), got);
}

int hp_save_cnnb_file(const hp_net *net, const char *path)
{
    if (!net || !path) { set_error("bad argument"); return HP_ERR_INVALID; }
    std::vector<char> buf((size_t)HP_CNNB_BYTES);
    size_t nw = 0;
    if (int rc = hp_save_cnnb(net, buf.data(), buf.size(), &nw)) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return HP_ERR_IO; }
    size_t put = fwrite(buf.data(), 1, nw, f);
    fclose(f);
    if (put != nw) { set_error("short write to %s", path); return HP_ERR_IO; }
    return HP_OK;
}

int hp_eval_batch_device(hp_net *net, const float *x_dev, int64_t n, float *y_dev, int precision, void *stream)
{
    if (!net || n < 0 || (n && (!x_dev || !y_dev))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return eval_device(net->n, x_dev, n, y_dev, precision, (cudaStream_t)stream);
}

// HOST buffers: chunks flow  host --H2D--> dev_in[b] --kernels--> dev_out[b] --D2H--> host  with
// two buffers per direction so that the copies of chunk c+1 / c-1 overlap the compute of chunk c.
// The upload is either fp32 crops (elem = 4) or 16-bit depth (elem = 2, normalised on the device,
// include/handtrack.h:700); the download is the 2304-float outputs and/or the 48-float decoded peaks.
struct DepthNorm { float scale, dmin, dmax; };
static int eval_host_pipeline(Net &N, const void *x, int elem, const DepthNorm *norm, int64_t n, float *y, float *dec, int precision)
{
    static const int64_t pipe_chunk = [] {   // HP_PIPE_CHUNK: tuning override (tools/dbg/e2e_chunks.py)
        const char *e = getenv("HP_PIPE_CHUNK");
        const int64_t v = e ? atoll(e) : 0;
        return (v >= 256 && v <= STAGE_CHUNK) ? v : PIPE_CHUNK;
    }();
    const int64_t chunk = std::min<int64_t>(n, pipe_chunk);
    const bool pin_x = is_pinned_host(x), pin_y = y ? is_pinned_host(y) : true, pin_d = dec ? is_pinned_host(dec) : true;
    if (int rc = ensure_staging(N, chunk, !pin_x, !pin_y)) return rc;
    cudaStream_t s = N.stream, h2d = N.comm_stream, d2h = N.d2h_stream;
    const int nc = (int)((n + chunk - 1) / chunk);
    const size_t in_row = (size_t)N_IN * elem;
    // wait until chunk c is back on the host (and bounce it into pageable memory)
    auto retire = [&](int c) -> int {
        const int b = c & 1;
        const int64_t m = std::min<int64_t>(chunk, n - (int64_t)c * chunk);
        HP_CUDA_TRY(cudaEventSynchronize(N.ev_out[b]));
        if (y && !pin_y) memcpy(y + (int64_t)c * chunk * N_OUT, N.pin_out[b], (size_t)m * N_OUT * sizeof(float));
        if (dec && !pin_d) memcpy(dec + (int64_t)c * chunk * 48, N.pin_dec[b], (size_t)m * 48 * sizeof(float));
        return 0;
    };
    for (int c = 0; c < nc; c++) {
        const int b = c & 1;
        const int64_t m = std::min<int64_t>(chunk, n - (int64_t)c * chunk);
        if (c >= 2)  // buffer pair b is reused: chunk c-2 must be fully out first
            if (int rc = retire(c - 2)) return rc;
        const char *src = (const char *)x + (size_t)c * chunk * in_row;
        if (!pin_x) {
            memcpy(N.pin_in[b], src, (size_t)m * in_row);
            src = (const char *)N.pin_in[b];
        }
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_in[b], src, (size_t)m * in_row, cudaMemcpyHostToDevice, h2d));
        HP_CUDA_TRY(cudaEventRecord(N.ev_in[b], h2d));
        HP_CUDA_TRY(cudaStreamWaitEvent(s, N.ev_in[b], 0));
        const float *xin = N.dev_in[b];
        bool decoded_in_epilogue = false;
        if (precision == HP_PRECISION_TENSOR && !N.tc->conv_v1 && (norm || dec)) {
            // tensor path: 16-bit depth goes straight into the conv kernel, which normalises in its loader, and the decode
            // runs in the fc2 epilogue (when y is not wanted it never exists in HBM)
            if (N.tc_dirty)
                if (int rc = tc_refresh_weights(N, s)) return rc;
            if (int rc = tc_forward_decode(N, norm ? nullptr : N.dev_in[b], norm ? (const uint16_t *)N.dev_in[b] : nullptr, norm ? norm->scale : 0.f,
                                           norm ? norm->dmin : 0.f, norm ? norm->dmax : 1.f, m, (y || !dec) ? N.dev_out[b] : nullptr,
                                           dec ? N.dev_dec[b] : nullptr, s))
                return rc;
            decoded_in_epilogue = dec != nullptr;
        } else {
            if (norm) {
                if (int rc = post_normalize_depth(N, (const uint16_t *)N.dev_in[b], m, norm->scale, norm->dmin, norm->dmax, N.dev_norm[b], s)) return rc;
                xin = N.dev_norm[b];
            }
            if (int rc = eval_device(N, xin, m, N.dev_out[b], precision, s, n)) return rc;
        }
        if (dec && !decoded_in_epilogue)
            if (int rc = post_decode(N, N.dev_out[b], m, N.dev_dec[b], s)) return rc;
        HP_CUDA_TRY(cudaEventRecord(N.ev_comp[b], s));
        HP_CUDA_TRY(cudaStreamWaitEvent(d2h, N.ev_comp[b], 0));
        if (y) {
            float *dst = pin_y ? y + (int64_t)c * chunk * N_OUT : N.pin_out[b];
            HP_CUDA_TRY(cudaMemcpyAsync(dst, N.dev_out[b], (size_t)m * N_OUT * sizeof(float), cudaMemcpyDeviceToHost, d2h));
        }
        if (dec) {
            float *dst = pin_d ? dec + (int64_t)c * chunk * 48 : N.pin_dec[b];
            HP_CUDA_TRY(cudaMemcpyAsync(dst, N.dev_dec[b], (size_t)m * 48 * sizeof(float), cudaMemcpyDeviceToHost, d2h));
        }
        HP_CUDA_TRY(cudaEventRecord(N.ev_out[b], d2h));
    }
    for (int c = std::max(0, nc - 2); c < nc; c++)
        if (int rc = retire(c)) return rc;
    return HP_OK;
}

int hp_eval_batch(hp_net *net, const float *x, int64_t n, float *y, int precision)
{
    if (!net || n < 0 || (n && (!x || !y))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return eval_host_pipeline(net->n, x, 4, nullptr, n, y, nullptr, precision);
}

int hp_eval_decode_batch(hp_net *net, const float *x, int64_t n, float *y, float *decoded, int precision)
{
    if (!net || n < 0 || (n && (!x || (!y && !decoded)))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return eval_host_pipeline(net->n, x, 4, nullptr, n, y, decoded, precision);
}

int hp_eval_depth_batch(hp_net *net, const uint16_t *depth, int64_t n, float depth_scale, float dmin, float dmax, float *y, float *decoded,
                        int precision)
{
    if (!net || n < 0 || (n && (!depth || (!y && !decoded))) || !(dmax > dmin)) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    DepthNorm nm{depth_scale, dmin, dmax};
    return eval_host_pipeline(net->n, depth, 2, &nm, n, y, decoded, precision);
}

int hp_eval_depth_batch_device(hp_net *net, const uint16_t *depth_dev, int64_t n, float depth_scale, float dmin, float dmax, float *y_dev,
                               float *decoded_dev, int precision, void *stream)
{
    if (!net || n < 0 || (n && (!depth_dev || (!y_dev && !decoded_dev))) || !(dmax > dmin)) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == HP_PRECISION_TENSOR && !N.tc->conv_v1) {
        if (N.tc_dirty)
            if (int rc = tc_refresh_weights(N, s)) return rc;
        return tc_forward_decode(N, nullptr, depth_dev, depth_scale, dmin, dmax, n, y_dev, decoded_dev, s);
    }
    // FP32 path (and the round-1 conv kernel): normalise chunk by chunk into the staging buffer (bit-identical values), then Eval
    for (int64_t b = 0; b < n; b += STAGE_CHUNK) {
        const int64_t m = std::min<int64_t>(STAGE_CHUNK, n - b);
        if (int rc = ensure_staging(N, m, false, false)) return rc;
        float *yb = y_dev ? y_dev + b * N_OUT : N.dev_out[0];
        if (int rc = post_normalize_depth(N, depth_dev + b * N_IN, m, depth_scale, dmin, dmax, N.dev_norm[0], s)) return rc;
        if (int rc = eval_device(N, N.dev_norm[0], m, yb, precision, s, n)) return rc;
        if (decoded_dev)
            if (int rc = post_decode(N, yb, m, decoded_dev + b * 48, s)) return rc;
    }
    return HP_OK;
}

int hp_resample_depth_device(hp_net *net, const uint16_t *frames_dev, int32_t width, int32_t height, const float src_intrinsics[4],
                             const int32_t *frame_of_crop_dev, const float *dst_cams_dev, int64_t n, uint16_t background, uint16_t *crops_dev,
                             void *stream)
{
    if (!net || n < 0 || width <= 0 || height <= 0 || !src_intrinsics || (n && (!frames_dev || !dst_cams_dev || !crops_dev))) {
        set_error("bad argument");
        return HP_ERR_INVALID;
    }
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return post_sample_d(net->n, frames_dev, width, height, src_intrinsics, frame_of_crop_dev, dst_cams_dev, n, background, crops_dev,
                         (cudaStream_t)stream);
}

int hp_eval_frames_device(hp_net *net, const uint16_t *frames_dev, int32_t width, int32_t height, const float src_intrinsics[4],
                          const int32_t *frame_of_crop_dev, const float *dst_cams_dev, int64_t n, uint16_t background, float depth_scale,
                          float dmin, float dmax, float *y_dev, float *decoded_dev, int precision, void *stream)
{
    if (!net || n < 0 || width <= 0 || height <= 0 || !src_intrinsics || (n && (!frames_dev || !dst_cams_dev || (!y_dev && !decoded_dev))) || !(dmax > dmin)) {
        set_error("bad argument");
        return HP_ERR_INVALID;
    }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    cudaStream_t s = (cudaStream_t)stream;
    // crops go through the (float-sized) staging buffer as 16-bit depth, STAGE_CHUNK crops at a time
    for (int64_t b = 0; b < n; b += STAGE_CHUNK) {
        const int64_t m = std::min<int64_t>(STAGE_CHUNK, n - b);
        if (int rc = ensure_staging(N, m, false, false)) return rc;
        uint16_t *crops = reinterpret_cast<uint16_t *>(N.dev_in[1]);
        // without an index, crop i samples frame i: the chunk's frames start b frames further on
        const uint16_t *fr = frame_of_crop_dev ? frames_dev : frames_dev + b * (int64_t)width * height;
        if (int rc = post_sample_d(N, fr, width, height, src_intrinsics, frame_of_crop_dev ? frame_of_crop_dev + b : nullptr,
                                   dst_cams_dev + b * HP_RESAMPLE_CAM_FLOATS, m, background, crops, s))
            return rc;
        if (int rc = hp_eval_depth_batch_device(net, crops, m, depth_scale, dmin, dmax, y_dev ? y_dev + b * N_OUT : nullptr,
                                                decoded_dev ? decoded_dev + b * 48 : nullptr, precision, stream))
            return rc;
    }
    return HP_OK;
}

int hp_normalize_depth_device(hp_net *net, const uint16_t *depth_dev, int64_t n, float depth_scale, float dmin, float dmax, float *x_dev, void *stream)
{
    if (!net || n < 0 || (n && (!depth_dev || !x_dev)) || !(dmax > dmin)) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return post_normalize_depth(net->n, depth_dev, n, depth_scale, dmin, dmax, x_dev, (cudaStream_t)stream);
}

int hp_decode_batch_device(hp_net *net, const float *y_dev, int64_t n, float *decoded_dev, void *stream)
{
    if (!net || n < 0 || (n && (!y_dev || !decoded_dev))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return post_decode(net->n, y_dev, n, decoded_dev, (cudaStream_t)stream);
}

int hp_render_labels_device(hp_net *net, const float *points_dev, const float *vals_dev, int64_t n, float *t_dev, void *stream)
{
    if (!net || n < 0 || (n && (!points_dev || !vals_dev || !t_dev))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (n == 0) return HP_OK;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return post_render_labels(net->n, points_dev, vals_dev, n, t_dev, (cudaStream_t)stream);
}

// HOST buffers: x[n][4096] crops, points[n][8][2] + vals[n][16] label parameters (128 B per sample instead of 9,216)
int hp_train_batch_points(hp_net *net, const float *x, const float *points, const float *vals, int64_t n, float alpha, float *mse_out, int precision)
{
    if (!net || n < 0 || (n && (!x || !points || !vals))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (int rc = check_peer(N)) return rc;
    const int64_t chunk = std::min<int64_t>(n, STAGE_CHUNK);
    if (int rc = ensure_staging(N, chunk, false, false)) return rc;
    cudaStream_t s = N.stream;
    float *dpts = N.dev_dec[0], *dvals = N.dev_dec[1];   // [chunk][48] scratch each: room for 16 floats per sample
    HP_CUDA_TRY(cudaEventRecord(N.ev_start, s));
    for (int64_t b = 0; b < n; b += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n - b);
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_in[0], x + b * N_IN, (size_t)m * N_IN * sizeof(float), cudaMemcpyHostToDevice, s));
        HP_CUDA_TRY(cudaMemcpyAsync(dpts, points + b * 16, (size_t)m * 16 * sizeof(float), cudaMemcpyHostToDevice, s));
        HP_CUDA_TRY(cudaMemcpyAsync(dvals, vals + b * 16, (size_t)m * 16 * sizeof(float), cudaMemcpyHostToDevice, s));
        if (int rc = post_render_labels(N, dpts, dvals, m, N.dev_t, s)) return rc;
        if (int rc = grad_device(N, N.dev_in[0], N.dev_t, m, N.dev_mse, precision, s, b > 0)) return rc;
        if (mse_out) HP_CUDA_TRY(cudaMemcpyAsync(mse_out + b, N.dev_mse, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (int rc = finish_step(N, alpha, precision, s)) return rc;
    HP_CUDA_TRY(cudaStreamSynchronize(s));
    return HP_OK;
}

int hp_render_labels(hp_net *net, const float *points, const float *vals, int64_t n, float *t)
{
    if (!net || n < 0 || (n && (!points || !vals || !t))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    for (int64_t b = 0; b < n; b += STAGE_CHUNK) {
        const int64_t m = std::min<int64_t>(STAGE_CHUNK, n - b);
        if (int rc = ensure_staging(N, m, false, false)) return rc;
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_dec[0], points + b * 16, (size_t)m * 16 * sizeof(float), cudaMemcpyHostToDevice, N.stream));
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_dec[1], vals + b * 16, (size_t)m * 16 * sizeof(float), cudaMemcpyHostToDevice, N.stream));
        if (int rc = post_render_labels(N, N.dev_dec[0], N.dev_dec[1], m, N.dev_t, N.stream)) return rc;
        HP_CUDA_TRY(cudaMemcpyAsync(t + b * N_OUT, N.dev_t, (size_t)m * N_OUT * sizeof(float), cudaMemcpyDeviceToHost, N.stream));
        HP_CUDA_TRY(cudaStreamSynchronize(N.stream));
    }
    return HP_OK;
}

int hp_decode_batch(hp_net *net, const float *y, int64_t n, float *decoded)
{
    if (!net || n < 0 || (n && (!y || !decoded))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    for (int64_t b = 0; b < n; b += STAGE_CHUNK) {
        const int64_t m = std::min<int64_t>(STAGE_CHUNK, n - b);
        if (int rc = ensure_staging(N, m, false, false)) return rc;
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_out[0], y + b * N_OUT, (size_t)m * N_OUT * sizeof(float), cudaMemcpyHostToDevice, N.stream));
        if (int rc = post_decode(N, N.dev_out[0], m, N.dev_dec[0], N.stream)) return rc;
        HP_CUDA_TRY(cudaMemcpyAsync(decoded + b * 48, N.dev_dec[0], (size_t)m * 48 * sizeof(float), cudaMemcpyDeviceToHost, N.stream));
        HP_CUDA_TRY(cudaStreamSynchronize(N.stream));
    }
    return HP_OK;
}

int hp_grad_batch_device(hp_net *net, const float *x_dev, const float *t_dev, int64_t n, float *mse_dev, int precision, void *stream)
{
    if (!net || n <= 0 || !x_dev || !t_dev) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return grad_device(net->n, x_dev, t_dev, n, mse_dev, precision, (cudaStream_t)stream);
}

int hp_apply_grads_device(hp_net *net, float alpha, void *stream)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    return sgd_apply(net->n, alpha, (cudaStream_t)stream);
}

// One step, eagerly: forward + backward on the caller's stream, the update tail pipelined on the side streams.
// One GPU, tensor path: the chain forward -> backward runs on an internal HIGH-priority stream forked from the caller's (and
// joined back at the end), so that its CTAs are dispatched ahead of the thousands of pending CTAs of the update tail that
// overlaps it (the 200 KB GEMM CTAs otherwise wait for the update's grid to drain: measured 11 us on one GEMM).  Stream
// priorities are recorded per kernel node when the step is captured, so the replayed graph keeps them.
static int train_step_eager(Net &N, const float *x_dev, const float *t_dev, int64_t n, float alpha, float *mse_dev, int precision, cudaStream_t s)
{
    static const bool no_hi = getenv("HP_NO_PRIORITY") != nullptr;   // A/B
    const bool hi = N.world == 1 && precision == HP_PRECISION_TENSOR && !N.step_timing && !no_hi;
    cudaStream_t m = s;
    if (hi) {
        HP_CUDA_TRY(cudaEventRecord(N.ev_hi[0], s));
        HP_CUDA_TRY(cudaStreamWaitEvent(N.hi_stream, N.ev_hi[0], 0));
        m = N.hi_stream;
    }
    HP_CUDA_TRY(cudaEventRecord(N.ev_start, m));
    if (int rc = grad_device(N, x_dev, t_dev, n, mse_dev, precision, m)) return rc;
    if (int rc = finish_step(N, alpha, precision, m)) return rc;
    if (hi) {
        HP_CUDA_TRY(cudaEventRecord(N.ev_hi[1], m));
        HP_CUDA_TRY(cudaStreamWaitEvent(s, N.ev_hi[1], 0));
    }
    return 0;
}

int hp_train_batch_device(hp_net *net, const float *x_dev, const float *t_dev, int64_t n, float alpha, float *mse_dev, int precision,
                          void *stream)
{
    if (!net || n < 0 || (n && (!x_dev || !t_dev))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (int rc = check_peer(N)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Net::StepGraph &G = N.step_graph;
    static const bool no_graph = getenv("HP_NO_GRAPH") != nullptr;
    // graph replay: one workspace chunk, nothing that needs host-visible events, shadows in step with the weights.
    // The legacy NULL stream cannot be captured.
    // Data parallel: only the peer-memory exchange (its kernels keep their barrier epochs in device memory, so a replayed
    // launch is a new exchange); NCCL steps stay eager.
    const bool dp_ok = N.world == 1 || (N.peer && N.peer->ready);
    const bool eligible = !no_graph && !G.disabled && dp_ok && !N.profiling && !N.step_timing && n <= FP32_CHUNK && s != nullptr &&
                          !(precision == HP_PRECISION_TENSOR && N.tc_dirty);
    // the key of a captured step: its arguments AND the generation of the device buffers it points into (an Eval with a
    // larger batch in between may reallocate the workspace or the activation buffers)
    const bool same = G.x == x_dev && G.t == t_dev && G.mse == mse_dev && G.n == n && G.alpha == alpha && G.precision == precision && G.stream == s &&
                      G.alloc_epoch == N.alloc_epoch;
    if (!eligible || !same) {
        if (!same) {
            if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
            G.x = x_dev; G.t = t_dev; G.mse = mse_dev; G.n = n; G.alpha = alpha; G.precision = precision; G.stream = s;
            G.seen = 0;
        }
        const int rc = train_step_eager(N, x_dev, t_dev, n, alpha, mse_dev, precision, s);
        G.alloc_epoch = N.alloc_epoch;   // what this step allocated is part of the key from here on
        if (eligible && rc == 0) G.seen = 1;
        return rc;
    }
    if (!G.exec) {
        if (G.seen < 1) { G.seen = 1; return train_step_eager(N, x_dev, t_dev, n, alpha, mse_dev, precision, s); }
        // second call with the same arguments: every buffer exists by now (the eager step allocated them), so capture
        const int64_t l0 = N.launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            G.disabled = true;
            return train_step_eager(N, x_dev, t_dev, n, alpha, mse_dev, precision, s);
        }
        const int rc = train_step_eager(N, x_dev, t_dev, n, alpha, mse_dev, precision, s);
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        G.launches_per_step = N.launches - l0;
        N.launches = l0;
        if (rc || ce != cudaSuccess || !graph || cudaGraphInstantiateWithFlags(&G.exec, graph, cudaGraphInstantiateFlagUseNodePriority) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            G.exec = nullptr;
            G.disabled = true;   // something in the step is not capturable here: stay eager
            N.tc_dirty = true;   // host-side flags advanced during the aborted capture; the shadows were not actually refreshed
            return train_step_eager(N, x_dev, t_dev, n, alpha, mse_dev, precision, s);
        }
        cudaGraphDestroy(graph);
    }
    HP_CUDA_TRY(cudaGraphLaunch(G.exec, s));
    N.launches += G.launches_per_step;
    N.last_n = std::min<int64_t>(n, FP32_CHUNK);
    N.tc_dirty = (precision != HP_PRECISION_TENSOR);
    return HP_OK;
}

int hp_train_batch(hp_net *net, const float *x, const float *t, int64_t n, float alpha, float *mse_out, int precision)
{
    if (!net || n < 0 || (n && (!x || !t))) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = check_precision(precision)) return rc;
    if (n == 0) return HP_OK;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (int rc = check_peer(N)) return rc;
    // the batch travels in chunks of at most STAGE_CHUNK samples (bounded staging); the gradient sums accumulate
    // across chunks at frozen weights and ONE update follows
    const int64_t chunk = std::min<int64_t>(n, STAGE_CHUNK);
    if (int rc = ensure_staging(N, chunk, false, false)) return rc;
    cudaStream_t s = N.stream;
    HP_CUDA_TRY(cudaEventRecord(N.ev_start, s));
    for (int64_t b = 0; b < n; b += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n - b);
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_in[0], x + b * N_IN, (size_t)m * N_IN * sizeof(float), cudaMemcpyHostToDevice, s));
        HP_CUDA_TRY(cudaMemcpyAsync(N.dev_t, t + b * N_OUT, (size_t)m * N_OUT * sizeof(float), cudaMemcpyHostToDevice, s));
        if (int rc = grad_device(N, N.dev_in[0], N.dev_t, m, N.dev_mse, precision, s, b > 0)) return rc;
        if (mse_out) HP_CUDA_TRY(cudaMemcpyAsync(mse_out + b, N.dev_mse, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (int rc = finish_step(N, alpha, precision, s)) return rc;
    HP_CUDA_TRY(cudaStreamSynchronize(s));
    return HP_OK;
}

int hp_get_grads(const hp_net *net, float *grads_host)
{
    if (!net || !grads_host) { set_error("bad argument"); return HP_ERR_INVALID; }
    HP_CUDA_TRY(cudaSetDevice(net->n.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    HP_CUDA_TRY(cudaMemcpy(grads_host, net->n.grads, (size_t)N_PARAMS * sizeof(float), cudaMemcpyDeviceToHost));
    return HP_OK;
}

int hp_device_ptrs(hp_net *net, float **params_dev, float **grads_dev)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    if (params_dev) *params_dev = net->n.params;
    if (grads_dev) *grads_dev = net->n.grads;
    net->n.tc_dirty = true;  // the caller may write the weights through this pointer
    return HP_OK;
}

// a captured training step bakes in the exchange mode, the grid sizes and the wire format: any change to those drops it
static void drop_step_graph(Net &N)
{
    Net::StepGraph &G = N.step_graph;
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    G.seen = 0;
    G.x = nullptr;
    G.disabled = false;
}

int hp_dp_unique_id(void *id128)
{
    if (!id128) { set_error("id128 is NULL"); return HP_ERR_INVALID; }
    if (int rc = nccl_bind()) return rc;
    HP_NCCL_TRY(g_nccl.GetUniqueId(id128));
    return HP_OK;
}

int hp_dp_init(hp_net *net, const void *id128, int rank, int world)
{
    if (!net || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (int rc = nccl_bind()) return rc;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    Id128 id;
    memcpy(id.b, id128, 128);
    HP_NCCL_TRY(g_nccl.CommInitRank(&N.nccl_comm, world, id, rank));
    drop_step_graph(N);
    N.rank = rank;
    N.world = world;
    // The persistent tensor-core kernels take one CTA with ~200 KB of shared memory on every SM, which leaves NCCL's
    // CTAs nowhere to run until a kernel drains.  In data-parallel mode a few SMs are therefore left out of the
    // persistent grids so that the gradient all-reduce really overlaps the backward pass.
    if (world > 1 && N.tc) {
        int reserve = 16;
        if (const char *e = getenv("HP_DP_RESERVE_SMS")) reserve = atoi(e);
        tc_set_reserved_sms(N, reserve);
    }
    return HP_OK;
}

int hp_dp_set_bf16_gradients(hp_net *net, int enable)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (enable && !N.grads_bf) HP_CUDA_TRY(cudaMalloc((void **)&N.grads_bf, (size_t)N_PARAMS * sizeof(__nv_bfloat16)));
    N.dp_bf16 = enable != 0;
    drop_step_graph(N);
    return HP_OK;
}

int hp_dp_peer_export(hp_net *net, void *handle_out)
{
    if (!net || !handle_out) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    return peer_export(N, handle_out);
}

int hp_dp_peer_init(hp_net *net, const void *all_handles, int rank, int world)
{
    if (!net || !all_handles || world < 2 || rank < 0 || rank >= world) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (N.peer && N.peer->ready) { set_error("peer exchange already initialised on this net; call hp_dp_shutdown first"); return HP_ERR_INVALID; }
    // the exchange kernels run one 1024-thread CTA on each SM that the persistent tensor-core grids leave free.  32: with
    // the step's chain at ~175 us (batch 256) the two FC exchanges are the critical path at 16 CTAs (70 us each, 2 GPUs:
    // 232 us/step); at 32 they are hidden again (193 us/step; 24: 202, 48: 196, 64: 209 -- profiles/r2_dp_exchange.md)
    int reserve = 32;
    if (const char *e = getenv("HP_DP_RESERVE_SMS")) reserve = atoi(e);
    if (int rc = peer_init(N, all_handles, rank, world, reserve)) return rc;
    drop_step_graph(N);
    N.rank = rank;
    N.world = world;
    if (N.tc) tc_set_reserved_sms(N, reserve);
    return HP_OK;
}

int hp_dp_peer_status(hp_net *net, int *timed_out_on_rank_plus_1)
{
    if (!net || !timed_out_on_rank_plus_1) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    return peer_status(N, timed_out_on_rank_plus_1);
}

/* diagnostics (tools/dbg/peer_bw.py): one exchange kernel of gradient bucket b (0 fc2, 1 fc1, 2 conv) on `stream`;
 * max_blocks > 0 overrides the CTA cap first.  All ranks must call it in the same order. */
HP_API int hp_debug_peer_exchange(hp_net *net, int bucket, float alpha, int max_blocks, void *stream)
{
    if (!net || bucket < 0 || bucket > 2 || !net->n.peer) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    const int off[3] = {OFF_F2W, OFF_F1W, 0};
    const int end[3] = {N_PARAMS, OFF_F2W, OFF_F1W};
    if (max_blocks > 0) {
        N.peer->max_blocks = max_blocks < PEER_MAX_BLOCKS ? max_blocks : PEER_MAX_BLOCKS;
        drop_step_graph(N);
    }
    return peer_sgd_bucket(N, alpha, off[bucket], end[bucket] - off[bucket], (cudaStream_t)stream);
}

int hp_dp_shutdown(hp_net *net)
{
    if (!net) return HP_OK;
    Net &N = net->n;
    cudaSetDevice(N.device);
    drop_step_graph(N);
    if (N.peer) {
        peer_shutdown(N);
    }
    if (N.nccl_comm) {
        cudaSetDevice(N.device);
        cudaDeviceSynchronize();
        g_nccl.CommDestroy(N.nccl_comm);
        N.nccl_comm = nullptr;
    }
    N.world = 1;
    N.rank = 0;
    return HP_OK;
}

int hp_profile(hp_net *net, int enable)
{
    if (!net) { set_error("net is NULL"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    if (enable && !N.prof_ev) {
        N.prof_ev = new cudaEvent_t[Net::PROF_MAX];
        for (int i = 0; i < Net::PROF_MAX; i++) HP_CUDA_TRY(cudaEventCreate(&N.prof_ev[i]));
    }
    if (enable) N.prof_used = 0;
    N.profiling = enable != 0;
    return HP_OK;
}

int hp_profile_read(hp_net *net, int n_stages, double *total_ms, int64_t *intervals)
{
    if (!net || n_stages < 0 || !total_ms || !intervals) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < n_stages; i++) { total_ms[i] = 0; intervals[i] = 0; }
    for (int i = 0; i + 1 < N.prof_used; i += 2) {
        const int st = N.prof_stage[i / 2];
        if (st < 0 || st >= n_stages) continue;
        float ms = 0;
        HP_CUDA_TRY(cudaEventElapsedTime(&ms, N.prof_ev[i], N.prof_ev[i + 1]));
        total_ms[st] += ms;
        intervals[st]++;
    }
    return HP_OK;
}

// Milliseconds from the start of the last hp_train_batch_device call to: bucket 0/1/2 gradients ready, dX 0/1 done,
// all-reduce 0/1/2 done (data parallel only), update tail done.  out[9].  Diagnostic for the overlap of the tail.
int hp_debug_step_times(hp_net *net, float *out)
{
    if (!net || !out) return HP_ERR_INVALID;
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    cudaEvent_t ev[9] = {N.ev_bucket[0], N.ev_bucket[1], N.ev_bucket[2], N.ev_dx[0], N.ev_dx[1], N.ev_ar[0], N.ev_ar[1], N.ev_ar[2], N.ev_tail};
    for (int i = 0; i < 9; i++) {
        out[i] = -1.f;
        if (!N.step_timing || (i >= 5 && i <= 7 && N.world <= 1)) continue;
        if (cudaEventElapsedTime(&out[i], N.ev_start, ev[i]) != cudaSuccess) { cudaGetLastError(); out[i] = -1.f; }
    }
    return HP_OK;
}

int64_t hp_launch_count(const hp_net *net) { return net ? net->n.launches : 0; }

int hp_peek(hp_net *net, int which, int64_t n, float *out_host)
{
    if (!net || !out_host || n <= 0) { set_error("bad argument"); return HP_ERR_INVALID; }
    Net &N = net->n;
    HP_CUDA_TRY(cudaSetDevice(N.device));
    HP_CUDA_TRY(cudaDeviceSynchronize());
    if (n > N.ws.cap) { set_error("peek beyond workspace (%lld > %lld)", (long long)n, (long long)N.ws.cap); return HP_ERR_INVALID; }
    const float *src = nullptr;
    size_t len = 0;
    switch (which) {
    case 3: src = N.ws.p1; len = P1_N; break;
    case 6: src = N.ws.p2; len = P2_N; break;
    case 8: src = N.ws.h1; len = FC1_OUT; break;
    case 9: src = N.ws.logits; len = N_OUT; break;
    case 109: src = N.ws.dlog; len = N_OUT; break;
    case 107: src = N.ws.da1; len = FC1_OUT; break;
    case 106: src = N.ws.g2; len = P2_N; break;
    case 103: src = N.ws.g1; len = P1_N; break;
    case 203:  // pool winners of the conv1 stage, uint8 [n][3600] (bytes, not floats)
        HP_CUDA_TRY(cudaMemcpy(out_host, N.ws.idx1, (size_t)n * P1_N, cudaMemcpyDeviceToHost));
        return HP_OK;
    case 206:  // pool winners of the conv2 stage, uint8 [n][2304]
        HP_CUDA_TRY(cudaMemcpy(out_host, N.ws.idx2, (size_t)n * P2_N, cudaMemcpyDeviceToHost));
        return HP_OK;
    default: set_error("unknown intermediate %d", which); return HP_ERR_INVALID;
    }
    HP_CUDA_TRY(cudaMemcpy(out_host, src, (size_t)n * len * sizeof(float), cudaMemcpyDeviceToHost));
    return HP_OK;
}

const char *hp_last_error(void) { return g_err; }
const char *hp_version(void) { return "handposedd-b200 0.1 sm_100a"; }

}  // extern "C"
