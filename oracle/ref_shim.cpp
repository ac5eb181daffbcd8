// ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI wrapper around the UNMODIFIED reference implementation
// (IntelRealSense/hand_tracking_samples third_party/cnn.h), compiled from the
// sources where they lie under /root/reference (oracle/Makefile passes
// -I$(REFERENCE_ROOT)); no reference source is copied into this repository.
// Output goes to oracle/_ref/ only.  The library is the strongest available
// oracle: it IS the reference's arithmetic.  It is used to
//   * pin the plain-C restatement (oracle/handposedd_oracle.c) bit-exactly,
//   * generate tests/golden/ fixtures (tests/golden/make_golden.py),
//   * serve as the CPU baseline / reference arm in bench.py (kind "reference").
//
// The only thing restated here is the architecture list of
// include/handtrack.h:108-118 (PoseInitializerCNN), because handtrack.h itself
// does not compile headless under g++.
#include "third_party/cnn.h"

#include <cstring>
#include <sstream>
#include <thread>

namespace {

struct Ref {
    CNN cnn{std::vector<int>{}};  // empty size list: no layers (cnn.h:595-604)
};

// handposedd: include/handtrack.h:108-118; key_angles_count = 16 -> 8x256 + 16x16 spans
void build_handposedd(CNN &cnn)
{
    std::vector<int> spans(8, 16 * 16);
    spans.insert(spans.end(), 16, 16);
    cnn.layers.push_back(new CNN::LConv({64, 64, 1}, {5, 5, 1, 16}, {60, 60, 16}));
    cnn.layers.push_back(new CNN::LActivation<TanH>(60 * 60 * 16));
    cnn.layers.push_back(new CNN::LMaxPool({60, 60, 16}));
    cnn.layers.push_back(new CNN::LMaxPool({30, 30, 16}));
    cnn.layers.push_back(new CNN::LConv({15, 15, 16}, {4, 4, 16, 64}, {12, 12, 64}));
    cnn.layers.push_back(new CNN::LActivation<TanH>(12 * 12 * 64));
    cnn.layers.push_back(new CNN::LMaxPool({12, 12, 64}));
    cnn.layers.push_back(new CNN::LFull(6 * 6 * 64, 16 * 16 * 8));
    cnn.layers.push_back(new CNN::LActivation<TanH>(16 * 16 * 8));
    cnn.layers.push_back(new CNN::LFull(16 * 16 * 8, 16 * 16 * 8 + 16 * 16));
    cnn.layers.push_back(new CNN::LSoftMaxChunked(spans));
}

// The reference never frees its layers and LBase has no virtual destructor
// (cnn.h:102-112), so delete through the concrete (final) types.
void free_layers(CNN &cnn)
{
    for (auto *l : cnn.layers) {
        if (auto *p = dynamic_cast<CNN::LConv *>(l)) delete p;
        else if (auto *p = dynamic_cast<CNN::LFull *>(l)) delete p;
        else if (auto *p = dynamic_cast<CNN::LMaxPool *>(l)) delete p;
        else if (auto *p = dynamic_cast<CNN::LActivation<TanH> *>(l)) delete p;
        else if (auto *p = dynamic_cast<CNN::LSoftMaxChunked *>(l)) delete p;
    }
    cnn.layers.clear();
}

struct membuf : std::streambuf {
    membuf(char *b, size_t n) { setg(b, b, b + n); setp(b, b + n); }
};

constexpr size_t kParams = 9458400;

void load_params(CNN &cnn, const float *p)
{
    membuf mb((char *)p, kParams * sizeof(float));
    std::istream is(&mb);
    cnn.loadb(is);
}
void save_params(const CNN &cnn, float *p)
{
    membuf mb((char *)p, kParams * sizeof(float));
    std::ostream os(&mb);
    cnn.saveb(os);
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) void *ref_create()
{
    auto *r = new Ref;
    build_handposedd(r->cnn);
    return r;
}
__attribute__((visibility("default"))) void ref_destroy(void *h)
{
    auto *r = (Ref *)h;
    free_layers(r->cnn);
    delete r;
}
// CNN::Init, cnn.h:581
__attribute__((visibility("default"))) void ref_init(void *h) { ((Ref *)h)->cnn.Init(); }
// CNN::loadb / saveb (stream overloads), cnn.h:590-591, over a memory buffer of 9,458,400 floats
__attribute__((visibility("default"))) void ref_load(void *h, const float *p) { load_params(((Ref *)h)->cnn, p); }
__attribute__((visibility("default"))) void ref_save(void *h, float *p) { save_params(((Ref *)h)->cnn, p); }
// CNN::saveb(std::string), cnn.h:593 -- the real file writer, for the .cnnb byte-layout test
__attribute__((visibility("default"))) void ref_saveb_file(void *h, const char *path) { ((Ref *)h)->cnn.saveb(std::string(path)); }
__attribute__((visibility("default"))) void ref_loadb_file(void *h, const char *path) { ((Ref *)h)->cnn.loadb(std::string(path)); }
__attribute__((visibility("default"))) void ref_set_simd(int on) { simd_enable = on != 0; }

// CNN::Eval, cnn.h:550, crop by crop
__attribute__((visibility("default"))) void ref_eval(void *h, const float *x, long n, float *y)
{
    auto &cnn = ((Ref *)h)->cnn;
    std::vector<float> in(4096);
    for (long b = 0; b < n; b++) {
        std::memcpy(in.data(), x + b * 4096, 4096 * sizeof(float));
        auto out = cnn.Eval(in);
        std::memcpy(y + b * 2304, out.data(), 2304 * sizeof(float));
    }
}

// CNN::Train, cnn.h:558, n sequential steps (train-cnn.cpp:160 semantics)
__attribute__((visibility("default"))) void ref_train_seq(void *h, const float *x, const float *t, long n, float alpha, float *mse)
{
    auto &cnn = ((Ref *)h)->cnn;
    std::vector<float> in(4096), tt(2304);
    for (long b = 0; b < n; b++) {
        std::memcpy(in.data(), x + b * 4096, 4096 * sizeof(float));
        std::memcpy(tt.data(), t + b * 2304, 2304 * sizeof(float));
        float m = cnn.Train(in, tt, alpha);
        if (mse) mse[b] = m;
    }
}

// Activations of every layer for one crop (forward chain of cnn.h:552-554),
// concatenated; returns total floats written. sizes[] gets the 11 lengths.
__attribute__((visibility("default"))) long ref_forward_trace(void *h, const float *x, float *out, int *sizes)
{
    auto &cnn = ((Ref *)h)->cnn;
    std::vector<float> cur(x, x + 4096);
    long off = 0;
    int li = 0;
    for (auto *l : cnn.layers) {
        cur = l->forward(cur);
        std::memcpy(out + off, cur.data(), cur.size() * sizeof(float));
        off += (long)cur.size();
        sizes[li++] = (int)cur.size();
    }
    return off;
}

// Per-sample gradient at frozen weights in .cnnb order, by the reference's own
// code: forward, loss (cnn.h:566-569), backward chain (cnn.h:571-572), then
// each layer's `update` (cnn.h:574-575) applied with alpha = -1 to a ZEROED
// twin of the layer, which leaves +sum(X*E) in the twin's W and B.
// errs (optional) receives errors[i] for all 11 layers, concatenated.
__attribute__((visibility("default"))) float ref_grad_sample(void *h, const float *x, const float *t, float *grad, float *errs)
{
    auto &cnn = ((Ref *)h)->cnn;
    const size_t L = cnn.layers.size();
    std::vector<std::vector<float>> outputs;
    std::vector<float> xin(x, x + 4096);
    for (auto *l : cnn.layers) outputs.push_back(l->forward(outputs.size() ? outputs.back() : xin));
    std::vector<std::vector<float>> errors(L);
    float mse = 0;
    errors.back().resize(outputs.back().size());
    for (size_t i = 0; i < outputs.back().size(); i++) {
        float e = outputs.back()[i] - t[i];
        mse += e * e;
        errors.back()[i] = e;
    }
    mse /= errors.back().size();
    for (auto i = L - 1; i > 0; i--) errors[i - 1] = cnn.layers[i]->backward(outputs[i - 1], outputs[i], errors[i]);

    Ref twin;
    build_handposedd(twin.cnn);  // W value-initialised to 0 by std::vector, B to 0.0f (cnn.h:203,403)
    for (size_t i = 0; i < L; i++)
        if (!errors[i].empty())
            twin.cnn.layers[i]->update(i ? outputs[i - 1] : xin, outputs[i], errors[i], -1.0f);
    save_params(twin.cnn, grad);
    free_layers(twin.cnn);
    if (errs) {
        size_t off = 0;
        for (size_t i = 0; i < L; i++) {
            if (errors[i].empty()) errors[i].assign(outputs[i].size(), 0.0f);  // errors[0] is never produced (cnn.h:571)
            std::memcpy(errs + off, errors[i].data(), errors[i].size() * sizeof(float));
            off += errors[i].size();
        }
    }
    return mse;
}

// "All host cores" figure: crops are independent, so run the
// reference's single-threaded Eval in nthreads threads, each on its own deep
// copy of the net (CNN copies are shallow pointer copies, handtrack.h:129).
__attribute__((visibility("default"))) void ref_eval_mt(void *h, const float *x, long n, float *y, int nthreads)
{
    if (nthreads <= 1) { ref_eval(h, x, n, y); return; }
    std::vector<float> params(kParams);
    save_params(((Ref *)h)->cnn, params.data());
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; k++) {
        long lo = n * k / nthreads, hi = n * (k + 1) / nthreads;
        th.emplace_back([=, &params]() {
            void *r = ref_create();
            ref_load(r, params.data());
            ref_eval(r, x + lo * 4096, hi - lo, y + lo * 2304);
            ref_destroy(r);
        });
    }
    for (auto &t : th) t.join();
}

// Same for Train: nthreads independent replicas each doing sequential Train
// on its slice (throughput figure only; the replicas' weights are discarded).
__attribute__((visibility("default"))) void ref_train_mt(void *h, const float *x, const float *t, long n, float alpha, int nthreads)
{
    std::vector<float> params(kParams);
    save_params(((Ref *)h)->cnn, params.data());
    std::vector<std::thread> th;
    if (nthreads < 1) nthreads = 1;
    for (int k = 0; k < nthreads; k++) {
        long lo = n * k / nthreads, hi = n * (k + 1) / nthreads;
        th.emplace_back([=, &params]() {
            void *r = ref_create();
            ref_load(r, params.data());
            ref_train_seq(r, x + lo * 4096, t + lo * 2304, hi - lo, alpha, nullptr);
            ref_destroy(r);
        });
    }
    for (auto &tt : th) tt.join();
}

}  // extern "C"
