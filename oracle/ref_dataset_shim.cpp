// ref_dataset_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI access to the UNMODIFIED reference dataset reader and writer, compiled in place from
// /root/reference (oracle/Makefile, -I$(REFERENCE_ROOT)) into oracle/_ref/libdatasetref.so:
//   ref_dataset_load   load_dataset (include/dataset.h:109-163) -> flat arrays
//   ref_dataset_save   DepthDataStreamOut(DatasetInfo) + SaveFrame (include/dataset.h:62-105): writes .json/.rs/.ir/.pose
// The forward declarations are the g++ shim of SURVEY.md Appendix A.2 (the reference relies on clang's delayed template
// parsing for from_json lookup); nothing in the reference tree is edited or copied.
#include <cfloat>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include "third_party/linalg.h"
namespace json { class value; }
struct Pose;
template <class T> void from_json(linalg::vec<T, 2> &, const json::value &);
template <class T> void from_json(linalg::vec<T, 3> &, const json::value &);
template <class T> void from_json(linalg::vec<T, 4> &, const json::value &);
template <class T, int M> void from_json(linalg::mat<T, M, 2> &, const json::value &);
template <class T, int M> void from_json(linalg::mat<T, M, 3> &, const json::value &);
template <class T, int M> void from_json(linalg::mat<T, M, 4> &, const json::value &);
void from_json(Pose &, const json::value &);
template <class T> json::value to_json(const linalg::vec<T, 2> &);
template <class T> json::value to_json(const linalg::vec<T, 3> &);
template <class T> json::value to_json(const linalg::vec<T, 4> &);
template <class T, int M> json::value to_json(const linalg::mat<T, M, 2> &);
template <class T, int M> json::value to_json(const linalg::mat<T, M, 3> &);
template <class T, int M> json::value to_json(const linalg::mat<T, M, 4> &);
json::value to_json(const Pose &);
#include "include/dataset.h"

static std::vector<Frame> g_frames;
static DatasetInfo g_dsi;

extern "C" {
#define EXPORT __attribute__((visibility("default")))

// returns the number of frames, or -1 when the reference throws; info[16] = w, h, fx, fy, px, py, depth_scale,
// mplane xyzw, hasir, rgb w h, feye w h   (segment_scale in info[16])
EXPORT long ref_dataset_load(const char *bname, int pose_array_size, float *info)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());          // load_dataset chats on stdout
    long n = -1;
    try {
        g_frames = load_dataset(bname, pose_array_size);
        from_json(g_dsi, json::parsefile(std::string(bname) + ".json"));
        n = (long)g_frames.size();
        const DCamera &c = g_dsi.dcamera;
        float v[17] = {(float)c.dim().x, (float)c.dim().y, c.focal().x, c.focal().y, c.principal().x, c.principal().y, c.depth_scale,
                       g_dsi.mplane.x, g_dsi.mplane.y, g_dsi.mplane.z, g_dsi.mplane.w, g_dsi.hasir ? 1.f : 0.f,
                       (float)g_dsi.rgb_dim.x, (float)g_dsi.rgb_dim.y, (float)g_dsi.feye_dim.x, (float)g_dsi.feye_dim.y, g_dsi.segment_scale};
        memcpy(info, v, sizeof(v));
    } catch (...) {
        n = -1;
    }
    std::cout.rdbuf(old);
    return n;
}

// copy frame i of the last load: depth[w*h] u16, ir[w*h] u8, poses[pose_array_size][7] (position xyz, orientation xyzw)
EXPORT void ref_dataset_frame(long i, unsigned short *depth, unsigned char *ir, float *poses)
{
    const Frame &f = g_frames[(size_t)i];
    memcpy(depth, f.depth.raster.data(), f.depth.raster.size() * 2);
    memcpy(ir, f.ir.raster.data(), f.ir.raster.size());
    for (size_t k = 0; k < f.pose.size(); k++) {
        const Pose &p = f.pose[k];
        const float v[7] = {p.position.x, p.position.y, p.position.z, p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w};
        memcpy(poses + 7 * k, v, sizeof(v));
    }
}

// write a dataset with the reference's own writer: n frames of w x h depth + ir, poses[n][np][7]
EXPORT int ref_dataset_save(const char *bname, int w, int h, float fx, float fy, float px, float py, float depth_scale, float segment_scale,
                            long n, int np, const unsigned short *depth, const unsigned char *ir, const float *poses)
{
    try {
        DCamera cam({w, h}, {fx, fy}, {px, py}, depth_scale);
        DatasetInfo dsi{cam, float4(0, 0, 0, FLT_MAX), bname, "synthetic", false, {0, 0}, {0, 0}, segment_scale};
        DepthDataStreamOut out(dsi);
        for (long i = 0; i < n; i++) {
            Image<unsigned short> d(cam, std::vector<unsigned short>(depth + i * w * h, depth + (i + 1) * w * h));
            Image<unsigned char> r(cam, std::vector<unsigned char>(ir + i * w * h, ir + (i + 1) * w * h));
            std::vector<Pose> ps(np);
            for (int k = 0; k < np; k++) {
                const float *v = poses + ((size_t)i * np + k) * 7;
                ps[k] = Pose(float3(v[0], v[1], v[2]), float4(v[3], v[4], v[5], v[6]));
            }
            out.SaveFrame(d, r, ps);
        }
        return 0;
    } catch (...) {
        return -1;
    }
}
}
