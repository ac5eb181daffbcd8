// Drop-in check: builds the handposedd net exactly the way include/handtrack.h:103-130 does,
// through the cnn.h-compatible header, and exercises Eval / Train / saveb / loadb / copy.
// usage: dropin_main <crops.f32> <labels.f32> <n> <out_eval.f32> <out_mse.f32> <out.cnnb> [tensor]
#include <handposedd/cnn.h>

#include <cstdio>
#include <cstring>
#include <iostream>

template <class T> std::vector<T> concat(std::vector<T> a, const std::vector<T> &b) { a.insert(a.end(), b.begin(), b.end()); return a; }
static const int key_angles_count = 16;

// verbatim call sequence of PoseInitializerCNN
inline CNN PoseInitializerCNN(std::string filename)
{
    CNN cnn({});
    cnn.layers.push_back(new CNN::LConv({64, 64, 1}, {5, 5, 1, 16}, {60, 60, 16}));
    cnn.layers.push_back(new CNN::LActivation<TanH>(60 * 60 * 16));
    cnn.layers.push_back(new CNN::LMaxPool({60, 60, 16}));
    cnn.layers.push_back(new CNN::LMaxPool({30, 30, 16}));
    cnn.layers.push_back(new CNN::LConv({15, 15, 16}, {4, 4, 16, 64}, {12, 12, 64}));
    cnn.layers.push_back(new CNN::LActivation<TanH>(12 * 12 * 64));
    cnn.layers.push_back(new CNN::LMaxPool({12, 12, 64}));
    cnn.layers.push_back(new CNN::LFull(6 * 6 * 64, 16 * 16 * 8));
    cnn.layers.push_back(new CNN::LActivation<TanH>(16 * 16 * 8));
    cnn.layers.push_back(new CNN::LFull(16 * 16 * 8, 16 * 16 * 8 + 16 * key_angles_count));
    cnn.layers.push_back(new CNN::LSoftMaxChunked(concat(std::vector<int>(8, 16 * 16), std::vector<int>(key_angles_count, 16))));
    cnn.Init();
    {
        std::ifstream is(filename, std::ios_base::in | std::ios_base::binary);
        if (is.is_open()) cnn.loadb(is);
    }
    return cnn;
}

static std::vector<float> read_f32(const char *path, size_t n)
{
    std::vector<float> v(n);
    FILE *f = fopen(path, "rb");
    if (!f || fread(v.data(), 4, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}
static void write_f32(const char *path, const std::vector<float> &v)
{
    FILE *f = fopen(path, "wb");
    fwrite(v.data(), 4, v.size(), f);
    fclose(f);
}

int main(int argc, char **argv)
try {
    if (argc < 7) { fprintf(stderr, "usage\n"); return 2; }
    const int n = atoi(argv[3]);
    auto crops = read_f32(argv[1], (size_t)n * 4096), labels = read_f32(argv[2], (size_t)n * 2304);
    CNN cnn = PoseInitializerCNN("");                 // returned by value, like handtrack.h:129
    if (argc > 7 && !strcmp(argv[7], "tensor")) cnn.precision = HP_PRECISION_TENSOR;
    std::vector<float> evals, mses;
    for (int i = 0; i < n; i++) {                     // handtrack.h:701
        std::vector<float> x(crops.begin() + i * 4096, crops.begin() + (i + 1) * 4096);
        auto y = cnn.Eval(x);
        evals.insert(evals.end(), y.begin(), y.end());
    }
    write_f32(argv[4], evals);
    CNN twin = cnn;                                   // shallow copy shares the weights (train-cnn.cpp:116 pattern)
    for (int i = 0; i < n; i++) {                     // train-cnn.cpp:160
        std::vector<float> x(crops.begin() + i * 4096, crops.begin() + (i + 1) * 4096);
        std::vector<float> t(labels.begin() + i * 2304, labels.begin() + (i + 1) * 2304);
        mses.push_back(twin.Train(x, t, 0.001f));
    }
    write_f32(argv[5], mses);
    cnn.saveb(std::string(argv[6]));                  // train-cnn.cpp:115; sees the twin's updates
    cnn.loadb(std::string("/nonexistent/file.cnnb")); // silent no-op
    cnn = PoseInitializerCNN("");                     // train-cnn.cpp:116 reset
    std::cout << "dropin ok\n";
    return 0;
} catch (const std::exception &e) {
    std::cerr << e.what() << "\n";
    return 1;
}
