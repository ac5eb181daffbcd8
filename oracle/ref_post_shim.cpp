// ref_post_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI access to the UNMODIFIED reference routines on either side of the CNN,
// compiled in place from /root/reference (oracle/Makefile, -I$(REFERENCE_ROOT)) into oracle/_ref/libpostref.so:
//   ref_decode            the numeric core of CNNOutputAnalysis::CNNOutputAnalysis (include/handtrack.h:218-241)
//                         built from the reference's own ImageFindMax / PeakSubPixel / PeakVolume / Peaks1D
//                         (include/misc_image.h:298-336, 389-399) in the constructor's call order
//   ref_normalize_depth   the depth -> [0,1] crop normalisation of include/handtrack.h:700
//   ref_sample_d          SampleD (include/misc_image.h:154-162): the rotated / scaled point resample HandSegmentVR ends with
//   ref_render_labels     the label vector of GatherHandExpectedCNN (include/handtrack.h:160-173) from feature points + key values
// handtrack.h itself is not included (it does not compile headless under g++), so the two call
// sequences are restated here; every arithmetic routine they call is the reference's.
#include <cfloat>
#include <cstring>
#include "third_party/linalg.h"
#include "third_party/geometric.h"
#include "include/misc_image.h"

extern "C" {

// out[48]: 8 x (image_point.x, image_point.y, confidence, peak value) then the 16 Peaks1D values
__attribute__((visibility("default"))) void ref_decode(const float *cnn_output, long n, float *out)
{
    const int2 hdim(16, 16);
    for (long b = 0; b < n; b++) {
        const float *y = cnn_output + b * 2304;
        float *o = out + b * 48;
        for (int i = 0; i < 8; i++) {                        // handtrack.h:221-234
            const float *base = y + product(hdim) * i;
            Image<float> fmap(hdim, std::vector<float>(base, base + product(hdim)));
            int2 mx = ImageFindMax(fmap);                    // handtrack.h:226
            float2 p = PeakSubPixel(fmap, mx);               // handtrack.h:230
            o[4 * i + 0] = p.x;
            o[4 * i + 1] = p.y;
            o[4 * i + 2] = PeakVolume(fmap, p);              // handtrack.h:232
            o[4 * i + 3] = base[hdim.x * mx.y + mx.x];       // handtrack.h:234 (crays.w)
        }
        const float *vptr = y + product(hdim) * 8;           // handtrack.h:236-239
        int2 vdim(16, 16);
        Image<float> vmap(vdim, std::vector<float>(vptr, vptr + product(vdim)));
        std::vector<float> vals = Peaks1D(vmap);
        for (int k = 0; k < 16; k++) o[32 + k] = vals[k];
    }
}

// handtrack.h:700: (float)clamp(1.0f - (d*depth_scale - drange.x) / (drange.y - drange.x), 0.0f, 1.0f)
__attribute__((visibility("default"))) void ref_normalize_depth(const unsigned short *d, long count, float depth_scale, float dmin, float dmax,
                                                                float *out)
{
    float2 drange = {dmin, dmax};
    for (long i = 0; i < count; i++)
        out[i] = (float)clamp(1.0f - (d[i] * depth_scale - drange.x) / (drange.y - drange.x), 0.0f, 1.0f);
}

// the label vector GatherHandExpectedCNN builds (include/handtrack.h:160-173) from 8 image
// feature points and 16 key values: RenderHeatMaps (misc_image.h:259-277) + Render1DHeatMaps (misc_image.h:279-295),
// u8-quantised, then GrayScaleToFloat (misc_image.h:171).  points[n][8][2], vals[n][16] -> t[n][2304].
__attribute__((visibility("default"))) void ref_render_labels(const float *points, const float *vals, long n, float *t)
{
    DCamera hcam(int2(16, 16));
    for (long b = 0; b < n; b++) {
        std::vector<float2> fp;
        for (int i = 0; i < 8; i++) fp.push_back(float2(points[b * 16 + 2 * i], points[b * 16 + 2 * i + 1]));
        auto hmaps = RenderHeatMaps(fp, hcam);
        std::vector<float> v(vals + b * 16, vals + b * 16 + 16);
        auto vmap = Render1DHeatMaps(v, 16);
        float *o = t + b * 2304;
        for (int i = 0; i < 8; i++)
            for (int k = 0; k < 256; k++) o[i * 256 + k] = GrayScaleToFloat(hmaps[i].raster[k]);
        for (int k = 0; k < 256; k++) o[2048 + k] = GrayScaleToFloat(vmap.raster[k]);
    }
}

// SURVEY.md 8f row 3, second half: SampleD<unsigned short> (misc_image.h:154-162) as HandSegmentVR calls it
// (include/handtrack.h:343).  src[h][w] depth with intrinsics (sfx, sfy, spx, spy); destination camera dw x dh with
// intrinsics (dfx, dfy, dpx, dpy) and pose7 = position xyz + orientation xyzw; out[dh][dw].
__attribute__((visibility("default"))) void ref_sample_d(const unsigned short *src, int w, int h, float sfx, float sfy, float spx, float spy,
                                                         int dw, int dh, float dfx, float dfy, float dpx, float dpy, const float *pose7,
                                                         unsigned short background, unsigned short *out)
{
    DCamera scam(int2(w, h), float2(sfx, sfy), float2(spx, spy), 0.001f);
    Image<unsigned short> simg(scam, std::vector<unsigned short>(src, src + (size_t)w * h));
    DCamera dcam(int2(dw, dh), float2(dfx, dfy), float2(dpx, dpy), 0.001f,
                 Pose(float3(pose7[0], pose7[1], pose7[2]), float4(pose7[3], pose7[4], pose7[5], pose7[6])));
    Image<unsigned short> r = SampleD(simg, dcam, background);
    memcpy(out, r.raster.data(), (size_t)dw * dh * sizeof(unsigned short));
}

}  // extern "C"
