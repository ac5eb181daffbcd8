// hp_post.cu -- the two steps that sit directly before and after the CNN in the reference's tracker
// (SURVEY.md 8f rows 3 and 1), as device kernels so that a caller can upload 16-bit depth crops (8 KB instead of
// 16 KB per crop) and download decoded peaks (192 B instead of 9 KB per crop):
//   normalize_depth_kernel  include/handtrack.h:700   depth -> [0,1] crop:  clamp(1 - (d*scale - dmin)/(dmax - dmin), 0, 1)
//   render_labels_kernel    include/handtrack.h:160-173 (label vector of GatherHandExpectedCNN from 8 image points + 16 key values)
//   sample_d_kernel         include/misc_image.h:154-162 (SampleD: the rotated / scaled point resample HandSegmentVR ends with,
//                           include/handtrack.h:343) for host-computed destination cameras
//   decode_kernel           include/handtrack.h:218-241 (numeric core of CNNOutputAnalysis): per 2-D heatmap ImageFindMax,
//                           PeakSubPixel, PeakVolume, peak value (include/misc_image.h:298-336); per 1-D heatmap
//                           max_element + PeakSubPixel1D (misc_image.h:340-350, 389-399)
// Both are bit-exact against the reference routines on identical inputs (explicitly un-fused multiply/add, the
// reference's sequential summation order, IEEE division).
#include "hp_common.cuh"

namespace hp {

#define LAUNCH_CHECK(net)                \
    do {                                 \
        (net).launches++;                \
        HP_CUDA_TRY(cudaGetLastError()); \
    } while (0)

__global__ void __launch_bounds__(256) normalize_depth_kernel(const uint16_t *__restrict__ d, float *__restrict__ x, int64_t count8, float depth_scale,
                                                              float dmin, float dmax)
{
    const float range = __fsub_rn(dmax, dmin);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 raw = reinterpret_cast<const uint4 *>(d)[i];   // 8 depth samples
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t v = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
            const float z = __fmul_rn((float)v, depth_scale);
            float a = __fsub_rn(1.0f, __fdiv_rn(__fsub_rn(z, dmin), range));
            a = (a < 0.0f) ? 0.0f : a;   // std::max(a, 0.0f), third_party/geometric.h:62
            a = (1.0f < a) ? 1.0f : a;   // std::min(.., 1.0f)
            o[k] = a;
        }
        reinterpret_cast<float4 *>(x)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4 *>(x)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// SampleD<unsigned short> (misc_image.h:154-162) for 64x64 destination cameras: one crop per CTA, 16 pixels per thread.
// cam[11] = destination focal xy, principal xy, pose position xyz, orientation xyzw (the data-dependent search that
// produces it, HandSegmentVR include/handtrack.h:280-341, stays on the host).  Every operation is separately rounded in
// the reference's order (linalg.h:284-288, misc_image.h:48-50); the float -> int conversions follow the pinned x86
// build (cvttss2si: INT_MIN for NaN / out of range; low 16 bits for the unsigned short).
__device__ __forceinline__ int cvtt_x86(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000u; }
struct SrcCam {
    int w, h;
    float fx, fy, px, py;
};
__global__ void __launch_bounds__(256) sample_d_kernel(const uint16_t *__restrict__ frames, SrcCam sc, const int32_t *__restrict__ frame_of_crop,
                                                       const float *__restrict__ cams, uint16_t background, uint16_t *__restrict__ out)
{
    __shared__ float c[11];
    const int64_t crop = blockIdx.x;
    if (threadIdx.x < 11) c[threadIdx.x] = cams[crop * 11 + threadIdx.x];
    __syncthreads();
    const float dfx = c[0], dfy = c[1], dpx = c[2], dpy = c[3];
    const float pos[3] = {c[4], c[5], c[6]};
    const float qx = c[7], qy = c[8], qz = c[9], qw = c[10];
    auto m = [](float a, float b) { return __fmul_rn(a, b); };
    auto ad = [](float a, float b) { return __fadd_rn(a, b); };
    auto sb = [](float a, float b) { return __fsub_rn(a, b); };
    const float ww = m(qw, qw), xx = m(qx, qx), yy = m(qy, qy), zz = m(qz, qz);
    const float xd[3] = {sb(sb(ad(ww, xx), yy), zz), m(ad(m(qx, qy), m(qz, qw)), 2.f), m(sb(m(qz, qx), m(qy, qw)), 2.f)};
    const float yd[3] = {m(sb(m(qx, qy), m(qz, qw)), 2.f), sb(ad(sb(ww, xx), yy), zz), m(ad(m(qy, qz), m(qx, qw)), 2.f)};
    const float zd[3] = {m(ad(m(qz, qx), m(qy, qw)), 2.f), m(sb(m(qy, qz), m(qx, qw)), 2.f), ad(sb(sb(ww, xx), yy), zz)};
    auto rot = [&](const float (&v)[3], float (&r)[3]) {   // position + (xd*v.x + yd*v.y) + zd*v.z
#pragma unroll
        for (int k = 0; k < 3; k++) r[k] = ad(pos[k], ad(ad(m(xd[k], v[0]), m(yd[k], v[1])), m(zd[k], v[2])));
    };
    float ppdir[3];
    {
        const float v[3] = {m(__fdiv_rn(sb(dpx, dpx), dfx), 1.0f), m(__fdiv_rn(sb(dpy, dpy), dfy), 1.0f), 1.0f};
        rot(v, ppdir);
    }
    const uint16_t *src = frames + (int64_t)(frame_of_crop ? frame_of_crop[crop] : crop) * sc.w * sc.h;
    for (int i = threadIdx.x; i < N_IN; i += 256) {
        const int x = i & 63, y = i >> 6;
        const float v[3] = {m(__fdiv_rn(sb((float)x, dpx), dfx), 1.0f), m(__fdiv_rn(sb((float)y, dpy), dfy), 1.0f), 1.0f};
        float p[3];
        rot(v, p);
        const int ix = cvtt_x86(ad(m(__fdiv_rn(p[0], p[2]), sc.fx), sc.px)), iy = cvtt_x86(ad(m(__fdiv_rn(p[1], p[2]), sc.fy), sc.py));
        uint16_t r = background;
        if (ix >= 0 && ix <= sc.w - 1 && iy >= 0 && iy <= sc.h - 1) {
            const float d = (float)src[(int64_t)iy * sc.w + ix];
            const float s0 = m(__fdiv_rn(sb((float)ix, sc.px), sc.fx), d), s1 = m(__fdiv_rn(sb((float)iy, sc.py), sc.fy), d), s2 = m(1.0f, d);
            r = (uint16_t)(cvtt_x86(ad(ad(m(ppdir[0], s0), m(ppdir[1], s1)), m(ppdir[2], s2))) & 0xffff);
        }
        out[crop * N_IN + i] = r;
    }
}

// one crop per CTA: warp w decodes 2-D heatmap w; threads 0..15 then decode the sixteen 1-D heatmaps
__global__ void __launch_bounds__(256) decode_kernel(const float *__restrict__ y, float *__restrict__ out)
{
    __shared__ float sy[N_OUT];
    const int64_t crop = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < N_OUT; i += 256) sy[i] = y[crop * N_OUT + i];
    __syncthreads();
    const float *m = sy + warp * 256;
    // ImageFindMax: first strict maximum in raster order, starting from pixel (0,0) (misc_image.h:300-304)
    float bv = m[lane];
    int bi = lane;
    if (lane != 0 && bv != bv) bv = -INFINITY;   // a NaN that is not the starting pixel never wins a `>` comparison
#pragma unroll
    for (int i = 1; i < 8; i++) {
        const float v = m[lane + 32 * i];
        if (v > bv) { bv = v; bi = lane + 32 * i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool keep_nan0 = (bi == 0 && bv != bv);     // pixel (0,0) NaN: nothing compares greater
        const bool other_nan0 = (oi == 0 && ov != ov);
        if (other_nan0 || (!keep_nan0 && (ov > bv || (ov == bv && oi < bi)))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        const int bx = bi & 15, by = bi >> 4;
        float wsum = 0.0f, vx = 0.0f, vy = 0.0f;
        for (int sy_ = max(0, by - 1); sy_ < min(16, by + 2); sy_++)
            for (int sx = max(0, bx - 1); sx < min(16, bx + 2); sx++) {   // PeakSubPixel, misc_image.h:316-322
                const float w = m[sy_ * 16 + sx];
                vx = __fadd_rn(vx, __fmul_rn((float)sx, w));
                vy = __fadd_rn(vy, __fmul_rn((float)sy_, w));
                wsum = __fadd_rn(wsum, w);
            }
        const float px = (wsum == 0) ? (float)bx : __fdiv_rn(vx, wsum);
        const float py = (wsum == 0) ? (float)by : __fdiv_rn(vy, wsum);
        // PeakVolume, misc_image.h:330: `int2 p(pf + float2(0.5f))`.  When pixel (0,0) is NaN the sub-pixel peak is NaN and
        // the reference's float->int conversion is undefined behaviour; the pinned x86 build (cvttss2si) yields INT_MIN
        // for NaN and out-of-range values, which empties both loops (volume 0), whereas CUDA's conversion yields 0 and
        // would sum a window.  Follow the pinned reference.
        const float fx = __fadd_rn(px, 0.5f), fy = __fadd_rn(py, 0.5f);
        const bool x_ok = fx >= -2147483648.0f && fx < 2147483648.0f, y_ok = fy >= -2147483648.0f && fy < 2147483648.0f;
        const int rx = x_ok ? (int)fx : 0, ry = y_ok ? (int)fy : 0;
        float vol = 0.0f;
        if (x_ok && y_ok)
            for (int sy_ = max(0, ry - 1); sy_ < min(16, ry + 2); sy_++)
                for (int sx = max(0, rx - 1); sx < min(16, rx + 2); sx++) vol = __fadd_rn(vol, m[sy_ * 16 + sx]);
        float *o = out + crop * 48 + 4 * warp;
        o[0] = px;
        o[1] = py;
        o[2] = vol;
        o[3] = m[bi];
    }
    if (tid < 16) {   // Peaks1D, misc_image.h:389-399
        const float *r = sy + 2048 + 16 * tid;
        int p = 0;
        for (int x = 1; x < 16; x++)
            if (r[p] < r[x]) p = x;
        float v = 0.0f, wsum = 0.0f;
        for (int i = max(0, p - 1); i < min(16, p + 2); i++) {
            const float w = r[i];
            v = __fadd_rn(v, __fmul_rn((float)i, w));
            wsum = __fadd_rn(wsum, w);
        }
        out[crop * 48 + 32 + tid] = __fdiv_rn((wsum == 0) ? (float)p : __fdiv_rn(v, wsum), 15.0f);
    }
}

// GatherHandExpectedCNN's label vector (include/handtrack.h:160-173): RenderHeatMap + NormalizeHeatMap
// (misc_image.h:248-270) for the 8 feature points, Render1DHeatMaps (misc_image.h:279-295) for the 16 key values,
// u8-quantised, then c/255 (misc_image.h:171).  One sample per CTA: warp w renders 2-D heatmap w, threads 0..15
// render the 1-D rows.  Integer sums are exact; exp is correctly rounded so the u8 truncation matches the reference.
__device__ __forceinline__ int to_gray(float x)
{
    float y = __fmul_rn(x, 255.0f);
    y = (y < 0.0f) ? 0.0f : y;
    y = (255.0f < y) ? 255.0f : y;
    return (int)(unsigned char)y;
}
__global__ void __launch_bounds__(256) render_labels_kernel(const float *__restrict__ points, const float *__restrict__ vals, float *__restrict__ t)
{
    __shared__ float st[N_OUT];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < N_OUT; i += 256) st[i] = 0.f;
    __syncthreads();
    {
        const float pkx = points[b * 16 + 2 * warp], pky = points[b * 16 + 2 * warp + 1];
        const int hx = (int)pkx, hy = (int)pky;
        const int px = hx - 2 + lane % 5, py = hy - 2 + lane / 5;
        const bool in = lane < 25 && px >= 0 && px < 16 && py >= 0 && py < 16;
        int c = 0;
        if (in) {
            const float dx = __fsub_rn(pkx, (float)px), dy = __fsub_rn(pky, (float)py);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            c = to_gray(exp_cr(__fdiv_rn(-d2, 2.0f * 0.33f)));
        }
        int sum = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (in) {
            const int q = sum ? (c * 255 / sum) : c;
            st[warp * 256 + py * 16 + px] = __fdiv_rn((float)(q & 255), 255.0f);
        }
    }
    if (tid < 16) {
        const float v = __fmul_rn(vals[b * 16 + tid], 15.0f);
        const int x0 = max(0, (int)v - 2), x1 = min(16, (int)v + 3);
        int r[5], sum = 0;
        for (int x = x0, k = 0; x < x1; x++, k++) {
            const float d = __fsub_rn((float)x, v);
            r[k] = to_gray(exp_cr(__fdiv_rn(-__fmul_rn(d, d), 2.0f * 0.5f)));
            sum += r[k];
        }
        for (int x = x0, k = 0; x < x1; x++, k++) {
            const int q = sum ? (r[k] * 255 / sum) : r[k];
            st[2048 + tid * 16 + x] = __fdiv_rn((float)(q & 255), 255.0f);
        }
    }
    __syncthreads();
    for (int i = tid; i < N_OUT; i += 256) t[b * N_OUT + i] = st[i];
}

int post_render_labels(Net &net, const float *points, const float *vals, int64_t n, float *t, cudaStream_t s)
{
    render_labels_kernel<<<(unsigned)n, 256, 0, s>>>(points, vals, t);
    LAUNCH_CHECK(net);
    return 0;
}

int post_normalize_depth(Net &net, const uint16_t *d, int64_t n, float depth_scale, float dmin, float dmax, float *x, cudaStream_t s)
{
    const int64_t count8 = n * (N_IN / 8);
    int blocks = (int)((count8 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    normalize_depth_kernel<<<blocks, 256, 0, s>>>(d, x, count8, depth_scale, dmin, dmax);
    LAUNCH_CHECK(net);
    return 0;
}

int post_sample_d(Net &net, const uint16_t *frames, int w, int h, const float *intr, const int32_t *frame_of_crop, const float *cams, int64_t n,
                  uint16_t background, uint16_t *out, cudaStream_t s)
{
    sample_d_kernel<<<(unsigned)n, 256, 0, s>>>(frames, SrcCam{w, h, intr[0], intr[1], intr[2], intr[3]}, frame_of_crop, cams, background, out);
    LAUNCH_CHECK(net);
    return 0;
}

int post_decode(Net &net, const float *y, int64_t n, float *out, cudaStream_t s)
{
    decode_kernel<<<(unsigned)n, 256, 0, s>>>(y, out);
    LAUNCH_CHECK(net);
    return 0;
}

}  // namespace hp
