// Drop-in check: builds the handposedd net exactly the way include/handtrack.h:103-130 does,
// through the cnn.h-compatible header, and exercises Eval / Train / saveb / loadb / copy.
// usage: dropin_main <crops.f32> <labels.f32> <n> <out_eval.f32> <out_mse.f32> <out.cnnb> [tensor]
#include <handposedd/cnn.h>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>

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
    {
        // per-layer streams (cnn.h:286-289, 606-609): conv1's W then B is the first 416 floats of the .cnnb order
        auto *c1 = dynamic_cast<CNN::LConv *>(cnn.layers[0]);
        std::ostringstream whole, part;
        cnn.saveb(whole);
        c1->saveb(part);
        if (part.str().size() != 416 * 4 || whole.str().compare(0, 416 * 4, part.str()) != 0) throw std::runtime_error("LConv::saveb differs from the net's prefix");
        std::stringstream txt;
        txt << *c1;                                   // operator<<, cnn.h:607
        std::vector<float> v;
        for (float w; txt >> w;) v.push_back(w);
        if (v.size() != 416) throw std::runtime_error("operator<< (LConv) wrote a wrong number of values");
        std::stringstream in;
        for (size_t i = 0; i < v.size(); i++) in << (float)i << ' ';
        in >> *c1;                                    // operator>>, cnn.h:606
        std::ostringstream again_s;
        cnn.saveb(again_s);
        const std::string again_bytes = again_s.str();
        const float *p = reinterpret_cast<const float *>(again_bytes.data());
        for (int i = 0; i < 416; i++)
            if (p[i] != (float)i) throw std::runtime_error("operator>> (LConv) did not reach the device store");
        if (again_bytes.compare(416 * 4, std::string::npos, whole.str(), 416 * 4, std::string::npos) != 0) throw std::runtime_error("operator>> (LConv) touched other layers");
        auto *f2 = dynamic_cast<CNN::LFull *>(cnn.layers[9]);
        std::ostringstream fpart;
        f2->saveb(fpart);                             // fc2: the last (2048 + 1) * 2304 floats
        if (fpart.str().size() != (size_t)(2048 + 1) * 2304 * 4 ||
            again_bytes.compare(again_bytes.size() - fpart.str().size(), std::string::npos, fpart.str()) != 0)
            throw std::runtime_error("LFull::saveb differs from the net's suffix");
    }
    {
        // copies taken BEFORE the device net exists share one weight store too (handtrack.h:129 returns by value)
        CNN a({});
        a.layers = cnn.layers;
        CNN b = a;
        a.Init();
        std::ostringstream sa, sb;
        a.saveb(sa);
        b.saveb(sb);
        if (sa.str() != sb.str()) throw std::runtime_error("copy made before first use has its own weight store");
    }
    {
        // layer types without kernels keep compiling and are rejected when the net is built (no CPU fallback)
        CNN m({});
        m.layers.push_back(new CNN::LConvS({8, 8}, 1, 2));
        m.layers.push_back(new CNN::LActivation<ReLU>(128));
        bool threw = false;
        try { m.Init(); } catch (const std::exception &) { threw = true; }
        if (!threw) throw std::runtime_error("an unsupported layer list was accepted");
    }
    std::cout << "dropin ok\n";
    return 0;
} catch (const std::exception &e) {
    std::cerr << e.what() << "\n";
    return 1;
}
