/*
 * handposedd.h -- C ABI of the B200-native replacement for the CNN hot path of
 * IntelRealSense/hand_tracking_samples (third_party/cnn.h, instantiated by
 * include/handtrack.h:103-130 as "handposedd").
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no
 * C++/torch types.  The cnn.h-compatible C++ class (include/handposedd/cnn.h)
 * and the Python mirror (hand_tracking_samples_b200/cnn.py) are thin callers of
 * these entry points.  Each entry point names the reference interface it
 * replaces.  There is no CPU fallback: every compute entry point fails with
 * HP_ERR_NO_DEVICE when no sm_100 device is usable.
 *
 * Threading: one hp_net is used by one host thread at a time (the reference has
 * at most one Eval in flight, include/handtrack.h:755); distinct nets are
 * independent.  hp_retain/hp_destroy are atomic.
 *
 * Layouts at the boundary are the reference's own:
 *   crops   x[n][4096]  float32, row-major 64x64, one channel (NHWC == NCHW),
 *           values as produced by include/handtrack.h:700 (nominally [0,1]);
 *   outputs y[n][2304]  float32: 8 heatmaps of 16x16 then 16 heatmaps of 16
 *           (spans of LSoftMaxChunked, include/handtrack.h:118);
 *   labels  t[n][2304]  float32 (GatherHandExpectedCNN, include/handtrack.h:160);
 *   weights the headerless little-endian float32 .cnnb stream of
 *           CNN::saveb (cnn.h:591,288-289,454-455): 9,458,400 floats.
 */
#ifndef HANDPOSEDD_H
#define HANDPOSEDD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HP_API
#else
#define HP_API __attribute__((visibility("default")))
#endif

typedef struct hp_net hp_net;

enum hp_status {
    HP_OK = 0,
    HP_ERR_INVALID = 1,     /* bad argument */
    HP_ERR_UNSUPPORTED = 2, /* layer list is not one this build has kernels for */
    HP_ERR_CUDA = 3,        /* a CUDA call failed; see hp_last_error() */
    HP_ERR_IO = 4,          /* file could not be opened / short buffer */
    HP_ERR_NCCL = 5,        /* NCCL unavailable or failed */
    HP_ERR_NO_DEVICE = 6,   /* no usable sm_100 device: there is no CPU fallback */
    HP_ERR_PEER = 7         /* a data-parallel exchange kernel gave up waiting for a rank; weights untouched since */
};

/* Arithmetic of the contraction layers (conv + FC). */
enum hp_precision {
    HP_PRECISION_FP32 = 0,  /* FFMA, reference summation order; parity bound 1e-5 */
    HP_PRECISION_TENSOR = 1 /* tcgen05 BF16 operands, FP32 accumulate in TMEM; parity bound 1e-2 */
};

/* Layer descriptors mirror the constructors client code calls on the
 * reference (cnn.h:139,203,403,459,496). */
enum hp_layer_kind {
    HP_LAYER_CONV = 1,            /* CNN::LConv(int3 indims, int4 dims, int3 outdims)   cnn.h:203 */
    HP_LAYER_TANH = 2,            /* CNN::LActivation<TanH>(int n)                       cnn.h:459 */
    HP_LAYER_MAXPOOL = 3,         /* CNN::LMaxPool(int3 indims)                          cnn.h:139 */
    HP_LAYER_FULL = 4,            /* CNN::LFull(int in, int out)                         cnn.h:403 */
    HP_LAYER_SOFTMAX_CHUNKED = 5  /* CNN::LSoftMaxChunked(std::vector<int> spans)        cnn.h:496 */
};
typedef struct hp_layer_desc {
    int kind;
    int in_dims[3];   /* x, y, z (z = channels); FULL: {in,0,0}; TANH: {n,0,0} */
    int w_dims[4];    /* CONV: kx, ky, cin, cout */
    int out_dims[3];  /* CONV: x, y, cout; FULL: {out,0,0} */
    int n_spans;      /* SOFTMAX_CHUNKED */
    const int *spans; /* SOFTMAX_CHUNKED */
} hp_layer_desc;

#define HP_N_IN 4096
#define HP_N_OUT 2304
#define HP_N_PARAMS 9458400
#define HP_CNNB_BYTES 37833600

/* ---- lifetime ------------------------------------------------------------ */

/* Replaces: the layer-list construction of PoseInitializerCNN
 * (include/handtrack.h:107-118).  Accepts exactly the handposedd list (the only
 * instantiation in the reference); anything else -> HP_ERR_UNSUPPORTED.
 * Weights start at zero (CNN ctor, cnn.h:203,403); call hp_init_xavier or
 * hp_load_cnnb next, as PoseInitializerCNN does. */
HP_API int hp_create(const hp_layer_desc *layers, int n_layers, int device, hp_net **out);
/* Convenience: the handposedd list without spelling it out. */
HP_API int hp_create_handposedd(int device, hp_net **out);
/* Replaces: CNN's by-value (shallow) copy semantics (include/handtrack.h:129,
 * train-hand-pose-cnn/train-cnn.cpp:116): copies share one device weight store. */
HP_API int hp_retain(hp_net *net);
HP_API int hp_destroy(hp_net *net);

/* ---- weights ------------------------------------------------------------- */

/* Replaces: CNN::Init (cnn.h:581-586): Xavier-uniform from one
 * std::default_random_engine (libstdc++: minstd_rand0, default seed) shared
 * across layers; biases 0.  Bit-identical to the reference under libstdc++. */
HP_API int hp_init_xavier(hp_net *net);
/* Replaces: CNN::loadb(std::istream&) (cnn.h:590).  Synchronises the whole device first (work in flight on caller
 * streams may still read the weights).  Like the reference's
 * loadvb (cnn.h:97), a short buffer fills a prefix and leaves the tail of the
 * weights unmodified; n_bytes beyond HP_CNNB_BYTES are ignored. */
HP_API int hp_load_cnnb(hp_net *net, const void *bytes, size_t n_bytes);
/* Replaces: CNN::saveb(std::ostream&) (cnn.h:591).  *n_written = HP_CNNB_BYTES. */
HP_API int hp_save_cnnb(const hp_net *net, void *bytes, size_t capacity, size_t *n_written);
/* Replaces: the per-layer streams LConv/LFull::loada/savea/loadb/saveb (cnn.h:286-289, 452-455) and their
 * operator>> / operator<< (cnn.h:606-609): one layer's W then B is a contiguous float range of the .cnnb-ordered
 * store (offsets: SURVEY.md 8c).  HOST buffers; first + count <= HP_N_PARAMS. */
HP_API int hp_get_params_range(const hp_net *net, int64_t first, int64_t count, float *host);
HP_API int hp_set_params_range(hp_net *net, int64_t first, int64_t count, const float *host);
/* Replaces: CNN::loadb(std::string) / CNN::saveb(std::string) (cnn.h:592-593).
 * hp_load_cnnb_file returns HP_ERR_IO if the file cannot be opened (the C++
 * wrapper keeps the reference's silent no-op by ignoring that status). */
HP_API int hp_load_cnnb_file(hp_net *net, const char *path);
HP_API int hp_save_cnnb_file(const hp_net *net, const char *path);

/* ---- inference ----------------------------------------------------------- */

/* Replaces: CNN::Eval (cnn.h:550-556), batched.  HOST buffers: x[n][4096] ->
 * y[n][2304]; uploads/downloads run through pinned staging buffers on an
 * internal stream, chunked and overlapped with compute; returns when y is
 * complete.  n == 0 is a no-op. */
HP_API int hp_eval_batch(hp_net *net, const float *x, int64_t n, float *y, int precision);
/* Same with DEVICE buffers on the caller's CUDA stream (cudaStream_t passed as
 * void*; NULL = the legacy default stream); asynchronous with respect to the
 * host. */
HP_API int hp_eval_batch_device(hp_net *net, const float *x_dev, int64_t n, float *y_dev,
                                int precision, void *stream);

/* ---- the steps on either side of Eval in the reference's tracker ---- */

#define HP_N_DECODED 48
/* Replaces: the numeric core of CNNOutputAnalysis::CNNOutputAnalysis (include/handtrack.h:218-241) built on
 * ImageFindMax / PeakSubPixel / PeakVolume / Peaks1D (include/misc_image.h:298-336, 340-350, 389-399).
 * decoded[n][48] = for each of the 8 landmark heatmaps (image_point.x, image_point.y, confidence, peak value
 * [crays.w]), then the 16 Peaks1D values (CNNOutputAnalysis::vals).  Bit-exact on identical y.  HOST buffers. */
HP_API int hp_decode_batch(hp_net *net, const float *y, int64_t n, float *decoded);
HP_API int hp_decode_batch_device(hp_net *net, const float *y_dev, int64_t n, float *decoded_dev, void *stream);
/* Eval followed by the decode in one pipelined call: y and decoded are each optional (not both NULL); with
 * y == NULL only 192 bytes per crop travel back over PCIe instead of 9,216. */
HP_API int hp_eval_decode_batch(hp_net *net, const float *x, int64_t n, float *y, float *decoded, int precision);
/* Replaces: the crop normalisation of HandTracker::update_cnn_model_threadsafe (include/handtrack.h:700) followed
 * by Eval (handtrack.h:701) and, optionally, the decode (handtrack.h:702): takes the segmented 64x64 crops as the
 * camera's 16-bit depth (8 KB per crop over PCIe instead of 16 KB) and applies
 *     x = clamp(1 - (d*depth_scale - dmin) / (dmax - dmin), 0, 1)
 * on the device, bit-exactly.  The reference uses drange = {0.1, 0.7} and depth_scale = 0.001. */
HP_API int hp_eval_depth_batch(hp_net *net, const uint16_t *depth, int64_t n, float depth_scale, float dmin, float dmax, float *y,
                               float *decoded, int precision);
/* Same chain with DEVICE buffers on the caller's stream: depth_dev[n][4096] uint16 -> y_dev[n][2304] and/or
 * decoded_dev[n][48] (each optional, not both NULL).  On the tensor path the normalisation runs inside the convolution
 * kernel's loader (no fp32 crop buffer exists in HBM: 8 KB read per crop instead of 8 + 16 + 16) and the decode runs inside
 * the fc2 + softmax kernel's epilogue; with y_dev == NULL the 9,216 bytes of y per crop are never written either.  The
 * decoded values equal hp_decode_batch of the same y bit for bit. */
HP_API int hp_eval_depth_batch_device(hp_net *net, const uint16_t *depth_dev, int64_t n, float depth_scale, float dmin, float dmax,
                                      float *y_dev, float *decoded_dev, int precision, void *stream);
HP_API int hp_normalize_depth_device(hp_net *net, const uint16_t *depth_dev, int64_t n, float depth_scale, float dmin, float dmax,
                                     float *x_dev, void *stream);

/* Replaces: SampleD<unsigned short> (include/misc_image.h:154-162) as HandSegmentVR calls it (include/handtrack.h:343):
 * the rotated / scaled point resample that turns a full depth frame into the 64x64 hand crop, for destination cameras
 * the host has already computed (the data-dependent search of HandSegmentVR, handtrack.h:280-341, stays on the host).
 *   frames[n_frames][height][width] uint16 depth, src_intrinsics = {focal.x, focal.y, principal.x, principal.y};
 *   dst_cams[n][11] = destination camera of each crop: focal xy, principal xy, pose position xyz, orientation xyzw
 *                     (dims are 64x64, HandSegmentVR's dstcam);
 *   frame_of_crop[n] (optional, NULL = crop i samples frame i);  background = HandSegmentVR's 4 m / depth_scale;
 *   crops[n][4096] uint16, bit-exact with the reference (including its x86 float -> int conversion of NaN / out-of-range
 *   values).  DEVICE buffers, caller's stream. */
#define HP_RESAMPLE_CAM_FLOATS 11
HP_API int hp_resample_depth_device(hp_net *net, const uint16_t *frames_dev, int32_t width, int32_t height, const float src_intrinsics[4],
                                    const int32_t *frame_of_crop_dev, const float *dst_cams_dev, int64_t n, uint16_t background,
                                    uint16_t *crops_dev, void *stream);
/* The tracker's chain from the full frame on (include/handtrack.h:698-702) without leaving the device: resample
 * (SampleD) -> normalise (handtrack.h:700, inside the convolution kernel's loader on the tensor path) -> Eval ->
 * optional decode.  y_dev[n][2304] and decoded_dev[n][48] are each optional (not both NULL). */
HP_API int hp_eval_frames_device(hp_net *net, const uint16_t *frames_dev, int32_t width, int32_t height, const float src_intrinsics[4],
                                 const int32_t *frame_of_crop_dev, const float *dst_cams_dev, int64_t n, uint16_t background,
                                 float depth_scale, float dmin, float dmax, float *y_dev, float *decoded_dev, int precision, void *stream);

/* Replaces: the label vector of GatherHandExpectedCNN (include/handtrack.h:160-173): RenderHeatMaps +
 * NormalizeHeatMap (include/misc_image.h:248-277) for the 8 image feature points, Render1DHeatMaps
 * (misc_image.h:279-295) for the 16 key values, u8 quantisation and c/255 (misc_image.h:169-171).
 * points[n][8][2] (heatmap pixel coordinates), vals[n][16] -> t[n][2304], bit-exact.  The pose-dependent inputs
 * (ImageFeaturePoints, HandPoseToKeyAngleSet) stay with the host solver. */
HP_API int hp_render_labels(hp_net *net, const float *points, const float *vals, int64_t n, float *t);
HP_API int hp_render_labels_device(hp_net *net, const float *points_dev, const float *vals_dev, int64_t n, float *t_dev, void *stream);

/* ---- recorded datasets (host side) ---------------------------------------- */

/* Replaces: load_dataset (include/dataset.h:109-163) for the parallel files DepthDataStreamOut writes
 * (dataset.h:62-105): <base>.json (DatasetInfo, dataset.h:21-37), <base>.rs (headerless 16-bit depth frames; with
 * "hasir" each frame is followed by its 8-bit IR image, dataset.h:135), optional <base>.ir and <base>.pose
 * (pose_array_size x 7 ASCII floats per frame: position xyz, orientation xyzw; 17 bones for the hand model,
 * train-cnn.cpp:75).  The binary files are memory-mapped; frames are copied on request straight into the caller's
 * batch buffers.  Same observable results as the reference: a trailing partial frame is dropped, a short .ir file
 * fills a prefix and leaves zeros, poses past the end of the .pose text are the default Pose, fields missing from
 * the .json read as 0 / false / "".  Errors instead of the reference's exceptions: HP_ERR_IO when .rs or .json
 * cannot be opened or dcamera.dims is not positive.  No GPU is needed for these calls. */
typedef struct hp_dataset hp_dataset;
typedef struct hp_dataset_info {
    int32_t width, height;           /* dcamera.dims */
    float focal[2], principal[2];    /* dcamera.focal, dcamera.principal */
    float depth_scale;               /* metres per depth unit */
    float mplane[4];
    int32_t hasir;                   /* .rs interleaves depth and IR (deprecated in the reference) */
    int32_t rgb_dim[2], feye_dim[2];
    float segment_scale;
    char camtype[32];
    int64_t n_frames;                /* complete frames in <base>.rs */
    int32_t pose_array_size;
    int32_t has_ir_file, has_pose_file;
} hp_dataset_info;
HP_API int hp_dataset_open(const char *basename, int pose_array_size, hp_dataset **out);
HP_API int hp_dataset_get_info(const hp_dataset *ds, hp_dataset_info *info);
/* Frames [first, first+count): depth[count][h][w], ir[count][h][w], poses[count][pose_array_size][7]; each may be NULL. */
HP_API int hp_dataset_read(hp_dataset *ds, int64_t first, int64_t count, uint16_t *depth, uint8_t *ir, float *poses);
/* Datasets already reduced to 64x64 hand crops (compress, train-cnn.cpp:31-50): frames go from the page cache through
 * hp_eval_depth_batch (depth_scale from the .json); other frame sizes -> HP_ERR_UNSUPPORTED (HandSegmentVR stays host). */
HP_API int hp_dataset_eval_depth(hp_net *net, hp_dataset *ds, int64_t first, int64_t count, float dmin, float dmax, float *y, float *decoded,
                                 int precision);
HP_API void hp_dataset_close(hp_dataset *ds);

/* ---- training ------------------------------------------------------------ */

/* Replaces: CNN::Train (cnn.h:558-580), batched.  One optimiser step on a
 * minibatch: all n samples see the same (pre-update) weights, and
 *     W <- W - alpha * sum_b g_b
 * where g_b is exactly what the reference's `update` (cnn.h:269-279, 438-445)
 * applies per unit alpha for sample b.  n == 1 is the reference's step.
 * mse_out (optional) receives the n per-sample values Train returns
 * (sum e^2 / 2304, cnn.h:566-569).  HOST buffers. */
HP_API int hp_train_batch(hp_net *net, const float *x, const float *t, int64_t n, float alpha,
                          float *mse_out, int precision);
/* Same step with the labels given as their 32 generating numbers per sample (hp_render_labels on the device):
 * 128 bytes of label upload per sample instead of 9,216.  HOST buffers. */
HP_API int hp_train_batch_points(hp_net *net, const float *x, const float *points, const float *vals, int64_t n, float alpha,
                                 float *mse_out, int precision);
/* DEVICE buffers, caller's stream: the step is ordered after the work already queued on `stream` and everything it
 * launches (also on the library's internal side streams: weight-gradient branches, the update of each gradient bucket,
 * the data-parallel exchange) has rejoined `stream` when the call returns -- the caller only ever synchronises `stream`.
 * A call that repeats the previous one exactly (same buffers, n, alpha, precision and a non-legacy stream) is replayed as
 * a CUDA graph captured on the second such call (HP_NO_GRAPH=1 disables); the result is bit-identical to the eager step.
 * With data parallelism enabled the gradient sums are exchanged behind the backward pass before the update:
 * hp_dp_peer_init -> one kernel per gradient bucket over NVLink peer memory (also graph-replayed), hp_dp_init -> NCCL
 * all-reduce.  The gradient store afterwards holds this rank's (peer path) or the all-reduced (NCCL path) sums. */
HP_API int hp_train_batch_device(hp_net *net, const float *x_dev, const float *t_dev, int64_t n,
                                 float alpha, float *mse_dev, int precision, void *stream);
/* Forward + backward WITHOUT the update: leaves sum_b g_b in the net's
 * gradient store (.cnnb order).  For gradient parity tests and custom
 * optimisers.  DEVICE buffers. */
HP_API int hp_grad_batch_device(hp_net *net, const float *x_dev, const float *t_dev, int64_t n,
                                float *mse_dev, int precision, void *stream);
/* Copy the gradient store (HP_N_PARAMS floats, .cnnb order) to a HOST buffer. */
HP_API int hp_get_grads(const hp_net *net, float *grads_host);
/* Device address of the weight / gradient stores (HP_N_PARAMS floats each). */
HP_API int hp_device_ptrs(hp_net *net, float **params_dev, float **grads_dev);
/* W <- W - alpha * grads (the SGD epilogue on its own) from THIS rank's gradient store: for single-GPU custom optimisers.
 * (hp_grad_batch_device does not exchange gradients; data-parallel steps go through hp_train_batch_device.)
 * DEVICE, caller's stream. */
HP_API int hp_apply_grads_device(hp_net *net, float alpha, void *stream);

/* ---- data parallelism (new; the reference is single-process) ------------ */

/* 128-byte NCCL unique id for rank 0 to broadcast by any host-side channel. */
HP_API int hp_dp_unique_id(void *id128);
/* One process per GPU: join a communicator of `world` ranks.  Afterwards
 * hp_train_batch_device all-reduces (sum) the 9,458,400-float gradient per
 * step, bucketed fc2 | fc1 | conv, launched as each bucket's weight gradient
 * finishes, and applies the identical update on every rank. */
HP_API int hp_dp_init(hp_net *net, const void *id128, int rank, int world);
/* Opt-in, NCCL exchange only (ignored once hp_dp_peer_init is in effect): steps run with HP_PRECISION_TENSOR send the two
 * FC gradient buckets (99.8 % of the bytes) over NVLink as bf16 and sum them in bf16 (18.9 MB instead of 37.8 MB per step).
 * FP32 steps always travel as fp32. */
HP_API int hp_dp_set_bf16_gradients(hp_net *net, int enable);
/* NVLink peer-memory exchange (one process per GPU of one NVSwitch box, 2..8 ranks): the reduce-scatter of the
 * gradient sums, the SGD update (cnn.h:438-445, 269-279) and the all-gather of the updated weights run as ONE kernel
 * per gradient bucket that reads the peers' gradient stores and writes the peers' weight stores directly
 * (csrc/hp_peer.cu); no NCCL on the data path.  Every rank calls hp_dp_peer_export (HP_PEER_HANDLE_BYTES of CUDA IPC
 * handles), the host program all-gathers them in rank order by any channel, every rank calls hp_dp_peer_init with
 * the world*HP_PEER_HANDLE_BYTES concatenation.  All ranks must then make the same sequence of training calls.
 * Takes precedence over hp_dp_init's NCCL all-reduce when both are set up.  Call hp_dp_shutdown on all ranks
 * (after a host-side barrier) before destroying the nets. */
#define HP_PEER_HANDLE_BYTES 256
HP_API int hp_dp_peer_export(hp_net *net, void *handle_out);
HP_API int hp_dp_peer_init(hp_net *net, const void *all_handles, int rank, int world);
/* 0, or 1 + the rank an exchange kernel gave up waiting for (HP_PEER_TIMEOUT_S seconds, default 30).  The failing
 * exchange and every later one skip their reduce / update / store phases, so the weights stay at the last state all
 * ranks agreed on, and hp_train_batch* / hp_save_cnnb return HP_ERR_PEER from then on (the latch is host-visible:
 * no polling of this call is needed).  Recover with hp_dp_shutdown + a fresh export/init on all ranks. */
HP_API int hp_dp_peer_status(hp_net *net, int *timed_out_on_rank_plus_1);
HP_API int hp_dp_shutdown(hp_net *net);

/* ---- diagnostics --------------------------------------------------------- */

/* Number of kernels this library has launched on behalf of `net` so far. */
HP_API int64_t hp_launch_count(const hp_net *net);
/* Per-stage device timing (CUDA events recorded on the caller's stream, read after a sync).
 * hp_profile(net, 1) starts recording, hp_profile(net, 0) stops; hp_profile_read returns, for
 * stage i < n_stages, the summed milliseconds and the number of intervals since recording
 * started.  Stages of the tensor-core Eval: 0 conv stages, 1 fc1 GEMM, 2 fc2 GEMM + softmax;
 * of the FP32 Eval: 0 conv stages, 1 fc1, 2 fc2, 3 softmax. */
HP_API int hp_profile(hp_net *net, int enable);
HP_API int hp_profile_read(hp_net *net, int n_stages, double *total_ms, int64_t *intervals);
/* Milliseconds from the start of the last hp_train_batch_device call to: gradient bucket 0/1/2 ready (fc2, fc1, conv),
 * dX GEMM 0/1 done, all-reduce 0/1/2 done (data parallel only, else -1), update tail done.  out[9]. */
HP_API int hp_debug_step_times(hp_net *net, float *out);
/* Peek at an intermediate of the last forward/backward pass (tests only):
 * which = 3 pooled conv1 stage [n][3600], 6 pooled conv2 stage [n][2304],
 * 8 fc1+tanh [n][2048]; copies n*len floats to HOST. */
HP_API int hp_peek(hp_net *net, int which, int64_t n, float *out_host);
/* Message for the last non-OK status returned on this thread. */
HP_API const char *hp_last_error(void);
/* "handposedd-b200 <version> sm_100a" */
HP_API const char *hp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HANDPOSEDD_H */
