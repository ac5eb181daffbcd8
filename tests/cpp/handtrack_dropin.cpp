// The reference's own include/handtrack.h, UNMODIFIED, compiled against the drop-in CNN class:
// handposedd/cnn.h defines the reference's include guard MINI_CNN_H, so handtrack.h:65's
// `#include "../third_party/cnn.h"` contributes nothing and PoseInitializerCNN (handtrack.h:103-130)
// builds a device-resident net.  Built only where /root/reference exists (-I/root/reference);
// the forward declarations below are the g++ shim of SURVEY.md Appendix A.2 (the reference needs
// clang's -fdelayed-template-parsing; nothing in the reference tree is edited or copied).
// usage: ht_dropin <crops.f32> <n> <out_eval.f32>
#include <cfloat>
#include <cstring>
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

#include <handposedd/cnn.h>   // must precede handtrack.h: claims MINI_CNN_H
#include "include/handtrack.h"

#include <cstdio>

int main(int argc, char **argv)
try {
    if (argc < 4) return 2;
    const int n = atoi(argv[2]);
    std::vector<float> crops((size_t)n * 4096), out;
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(crops.data(), 4, crops.size(), f) != crops.size()) return 2;
    fclose(f);
    CNN cnn = PoseInitializerCNN("");          // the reference's own factory, handtrack.h:103
    for (int i = 0; i < n; i++) {
        auto y = cnn.Eval(std::vector<float>(crops.begin() + i * 4096, crops.begin() + (i + 1) * 4096));  // handtrack.h:701
        out.insert(out.end(), y.begin(), y.end());
    }
    f = fopen(argv[3], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    printf("handtrack dropin ok\n");
    return 0;
} catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
}
