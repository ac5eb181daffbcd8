// The reference-side binding of the dataset reader ("Dataset reader binding"): load_dataset rebuilt on
// hp_dataset_*, returning the reference's own Frame objects, compared IN THE SAME BINARY with the reference's unmodified
// load_dataset (include/dataset.h:109-163).  Host-only: needs no GPU.  Built only where /root/reference exists
// (-I/root/reference); the forward declarations are the g++ shim of SURVEY.md Appendix A.2.
// usage: dataset_dropin <basename> <pose_array_size>
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

#include <handposedd.h>

// ---- the stub of INTEGRATION.md, verbatim except for its name --------------------------------------------------------
std::vector<Frame> load_dataset_hp(std::string bname, unsigned int pose_array_size, std::function<void(Frame&)> post_process = [](Frame&) {})
{
    hp_dataset *ds = nullptr;
    if (hp_dataset_open(bname.c_str(), (int)pose_array_size, &ds) != HP_OK) throw std::runtime_error(hp_last_error());
    hp_dataset_info di;  hp_dataset_get_info(ds, &di);
    DCamera cam({di.width, di.height}, {di.focal[0], di.focal[1]}, {di.principal[0], di.principal[1]}, di.depth_scale);
    std::vector<Frame> frames;
    std::vector<unsigned short> d(di.width * di.height);  std::vector<unsigned char> ir(d.size());  std::vector<float> p(pose_array_size * 7);
    for (int64_t k = 0; k < di.n_frames; k++) {
        hp_dataset_read(ds, k, 1, d.data(), ir.data(), p.data());
        std::vector<Pose> pose(pose_array_size);
        for (unsigned i = 0; i < pose_array_size; i++) pose[i] = Pose({p[7*i], p[7*i+1], p[7*i+2]}, {p[7*i+3], p[7*i+4], p[7*i+5], p[7*i+6]});
        auto f = MakeFrame(Image<unsigned short>(cam, d), pose, Image<unsigned char>{cam, ir}, Image<byte3>({di.rgb_dim[0], di.rgb_dim[1]}), Image<unsigned char>({di.feye_dim[0], di.feye_dim[1]}));
        f.fname = bname;  f.fid = (int)k;  post_process(f);  frames.push_back(f);
    }
    hp_dataset_close(ds);
    return frames;
}

int main(int argc, char **argv)
try {
    if (argc < 3) return 2;
    const unsigned np = (unsigned)atoi(argv[2]);
    int calls_ref = 0, calls_hp = 0;
    std::ostringstream sink;
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf());            // the reference's loader chats on stdout
    std::vector<Frame> want = load_dataset(argv[1], np, [&](Frame &) { calls_ref++; });
    std::cout.rdbuf(old);
    std::vector<Frame> got = load_dataset_hp(argv[1], np, [&](Frame &) { calls_hp++; });
    if (got.size() != want.size() || calls_ref != calls_hp) { fprintf(stderr, "frame count %zu vs %zu\n", got.size(), want.size()); return 1; }
    for (size_t k = 0; k < want.size(); k++) {
        const Frame &a = got[k], &b = want[k];
        bool ok = a.depth.raster == b.depth.raster && a.ir.raster == b.ir.raster && a.fid == b.fid && a.fname == b.fname &&
                  a.depth.cam.dim() == b.depth.cam.dim() && a.depth.cam.focal() == b.depth.cam.focal() &&
                  a.depth.cam.principal() == b.depth.cam.principal() && a.depth.cam.depth_scale == b.depth.cam.depth_scale &&
                  a.rgb.dim() == b.rgb.dim() && a.fisheye.dim() == b.fisheye.dim() && a.pose.size() == b.pose.size() &&
                  a.startpose.size() == b.startpose.size();
        for (size_t i = 0; ok && i < a.pose.size(); i++)
            ok = a.pose[i].position == b.pose[i].position && a.pose[i].orientation == b.pose[i].orientation &&
                 a.startpose[i].position == b.startpose[i].position;
        if (!ok) { fprintf(stderr, "frame %zu differs\n", k); return 1; }
    }
    printf("dataset dropin ok: %zu frames\n", want.size());
    return 0;
} catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
} catch (const char *e) {
    fprintf(stderr, "%s\n", e);
    return 1;
}
