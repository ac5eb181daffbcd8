// handposedd/cnn.h -- drop-in replacement for third_party/cnn.h of
// IntelRealSense/hand_tracking_samples: the same `struct CNN` surface (nested layer types whose
// constructors client code calls, public `layers`, Eval / Train / Init / loada / savea / loadb /
// saveb, the MLP convenience constructor and the stream operators), implemented as a thin caller
// of the C ABI in handposedd.h.  All arithmetic runs in hand-written sm_100a kernels on a B200;
// there is no CPU implementation behind this header.
//
// How it drops in: this header defines the reference's own include guard
// MINI_CNN_H, so pre-including it (`-include handposedd/cnn.h`, or replacing the body of
// third_party/cnn.h with `#include <handposedd/cnn.h>`) makes include/handtrack.h:65 and
// train-hand-pose-cnn/train-cnn.cpp compile against this class unchanged.
//
// Differences from the reference, all deliberate:
//  * layer objects are architecture DESCRIPTORS: weights live in the device-resident store owned
//    by the net handle, not in per-layer std::vectors, and LBase has no forward/backward;
//  * only the layer list of PoseInitializerCNN (include/handtrack.h:108-118) has kernels; any
//    other list throws std::runtime_error at first use instead of computing on the CPU;
//  * copies of a CNN share one device weight store (the reference's copies share raw layer
//    pointers, include/handtrack.h:129, train-cnn.cpp:116), reference-counted here;
//  * EvalBatch / TrainBatch are new: the batched entry points of the C ABI.
#ifndef MINI_CNN_H
#define MINI_CNN_H

#include <cstdint>
#include <fstream>
#include <istream>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../handposedd.h"

static bool simd_enable = true;  // cnn.h:22; kept for source compatibility, has no effect

// activation tags (cnn.h:24-43); only TanH has device kernels
struct Sigmoid {};
struct TanH {};
struct ReLU {};
struct LeakyReLU {};

namespace handposedd {
// accepts brace lists {x,y,z} as well as linalg::vec<int,3> / <int,4> (anything with .x .y .z [.w])
struct i3 {
    int x, y, z;
    i3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
    template <class V, class = decltype(V().x + V().y + V().z)> i3(const V &v) : x(v.x), y(v.y), z(v.z) {}
};
struct i4 {
    int x, y, z, w;
    i4(int x_, int y_, int z_, int w_) : x(x_), y(y_), z(z_), w(w_) {}
    template <class V, class = decltype(V().x + V().w)> i4(const V &v) : x(v.x), y(v.y), z(v.z), w(v.w) {}
};
inline void check(int status)
{
    if (status != HP_OK) throw std::runtime_error(std::string("handposedd: ") + hp_last_error());
}
}  // namespace handposedd

struct CNN {
    struct LBase {
        virtual ~LBase() {}
        virtual hp_layer_desc describe() const = 0;
    };
    struct LConv final : public LBase {  // cnn.h:194-290
        handposedd::i3 indims;
        handposedd::i4 dims;
        handposedd::i3 outdims;
        LConv(handposedd::i3 indims, handposedd::i4 dims, handposedd::i3 outdims) : indims(indims), dims(dims), outdims(outdims) {}
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = HP_LAYER_CONV;
            d.in_dims[0] = indims.x; d.in_dims[1] = indims.y; d.in_dims[2] = indims.z;
            d.w_dims[0] = dims.x; d.w_dims[1] = dims.y; d.w_dims[2] = dims.z; d.w_dims[3] = dims.w;
            d.out_dims[0] = outdims.x; d.out_dims[1] = outdims.y; d.out_dims[2] = outdims.z;
            return d;
        }
    };
    template <class F> struct LActivation final : public LBase {  // cnn.h:457-470
        int n;
        LActivation(int n) : n(n) {}
        hp_layer_desc describe() const override;
    };
    struct LMaxPool final : public LBase {  // cnn.h:136-165
        handposedd::i3 indims;
        LMaxPool(handposedd::i3 indims) : indims(indims) {}
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = HP_LAYER_MAXPOOL;
            d.in_dims[0] = indims.x; d.in_dims[1] = indims.y; d.in_dims[2] = indims.z;
            return d;
        }
    };
    struct LFull final : public LBase {  // cnn.h:398-456
        int M, N;
        LFull(int input_size, int output_size) : M(input_size), N(output_size) {}
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = HP_LAYER_FULL;
            d.in_dims[0] = M;
            d.out_dims[0] = N;
            return d;
        }
    };
    struct LSoftMaxChunked final : public LBase {  // cnn.h:493-528
        std::vector<int> spans;
        LSoftMaxChunked(std::vector<int> spans) : spans(spans) {}
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = HP_LAYER_SOFTMAX_CHUNKED;
            d.n_spans = (int)spans.size();
            d.spans = spans.data();
            return d;
        }
    };
    // Layer types of the reference that handposedd does not instantiate (cnn.h:113-135, 166-193,
    // 292-396, 471-492, 529-547).  They keep client code compiling; a net containing one is
    // rejected by hp_create with HP_ERR_UNSUPPORTED.
    struct LUnsupported : public LBase {
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = 0;
            return d;
        }
    };
    struct LAvgPool final : public LUnsupported { LAvgPool(handposedd::i3) {} };
    struct LSparsePool final : public LUnsupported { LSparsePool(handposedd::i3) {} };
    struct LSoftMax final : public LUnsupported { LSoftMax(int) {} };
    struct LCrossEntropy final : public LUnsupported { LCrossEntropy(int) {} };

    std::vector<LBase *> layers;  // cnn.h:548 (raw, never freed, exactly like the reference)
    int precision = HP_PRECISION_FP32;  // new: HP_PRECISION_TENSOR selects the tcgen05 path
    int device = 0;

    CNN(const std::vector<int> &s)  // cnn.h:595-604 ("quick test for simple NNs"): LFull + TanH pairs
    {
        for (unsigned int i = 1; i < s.size(); i++) {
            layers.push_back(new LFull(s[i - 1], s[i]));
            layers.push_back(new LActivation<TanH>(s[i]));
        }
        if (!layers.empty()) Init();
    }

    // cnn.h:550
    std::vector<float> Eval(const std::vector<float> &x)
    {
        std::vector<float> y(HP_N_OUT);
        if (x.size() < (size_t)HP_N_IN) throw std::runtime_error("handposedd: Eval input shorter than 4096 floats");
        handposedd::check(hp_eval_batch(net(), x.data(), 1, y.data(), precision));
        return y;
    }
    // cnn.h:558
    float Train(const std::vector<float> &x, const std::vector<float> &t, float alpha = 0.01f)
    {
        float mse = 0;
        if (x.size() < (size_t)HP_N_IN || t.size() < (size_t)HP_N_OUT) throw std::runtime_error("handposedd: Train input too short");
        handposedd::check(hp_train_batch(net(), x.data(), t.data(), 1, alpha, &mse, precision));
        return mse;
    }
    // new: n crops at once; x[n][4096] -> y[n][2304] (host pointers)
    void EvalBatch(const float *x, int64_t n, float *y) { handposedd::check(hp_eval_batch(net(), x, n, y, precision)); }
    // new: one optimiser step on n samples, W -= alpha * sum_b g_b; mse (optional) gets n values
    void TrainBatch(const float *x, const float *t, int64_t n, float alpha, float *mse = nullptr)
    {
        handposedd::check(hp_train_batch(net(), x, t, n, alpha, mse, precision));
    }
    // cnn.h:581
    void Init()
    {
        if (layers.empty()) return;  // CNN cnn({}) calls Init() on an empty list (handtrack.h:107)
        handposedd::check(hp_init_xavier(net()));
    }
    // cnn.h:590-593
    void loadb(std::istream &s)
    {
        std::vector<char> buf((size_t)HP_CNNB_BYTES);
        s.read(buf.data(), (std::streamsize)buf.size());
        handposedd::check(hp_load_cnnb(net(), buf.data(), (size_t)s.gcount()));
    }
    void saveb(std::ostream &s) const
    {
        std::vector<char> buf((size_t)HP_CNNB_BYTES);
        size_t n = 0;
        handposedd::check(hp_save_cnnb(net(), buf.data(), buf.size(), &n));
        s.write(buf.data(), (std::streamsize)n);
    }
    void loadb(std::string fname)
    {
        auto is = std::ifstream(fname, std::ios_base::binary | std::ios_base::in);
        if (is.is_open()) loadb(is);  // unopenable file: silent no-op, as cnn.h:592 behaves
    }
    void saveb(std::string fname) const
    {
        auto os = std::ofstream(fname, std::ios_base::binary);
        saveb(os);
    }
    // cnn.h:588-589: text form, every W then B value in layer order separated by ' '
    void loada(std::istream &s)
    {
        std::vector<float> p((size_t)HP_N_PARAMS);
        size_t i = 0;
        for (; i < p.size() && (s >> p[i]); i++) {}
        handposedd::check(hp_load_cnnb(net(), p.data(), i * sizeof(float)));
    }
    void savea(std::ostream &s) const
    {
        std::vector<float> p((size_t)HP_N_PARAMS);
        size_t n = 0;
        handposedd::check(hp_save_cnnb(net(), p.data(), p.size() * sizeof(float), &n));
        for (float w : p) s << w << ' ';
    }

    // the device net behind this object, created on first use from `layers`
    hp_net *net() const
    {
        if (!handle_ || built_for_ != layers.size()) {
            std::vector<hp_layer_desc> d;
            for (auto *l : layers) d.push_back(l->describe());
            hp_net *h = nullptr;
            handposedd::check(hp_create(d.data(), (int)d.size(), device, &h));
            handle_ = std::shared_ptr<hp_net>(h, [](hp_net *p) { hp_destroy(p); });
            built_for_ = layers.size();
        }
        return handle_.get();
    }

  private:
    mutable std::shared_ptr<hp_net> handle_;  // shared by copies: one weight store
    mutable size_t built_for_ = 0;
};

template <> inline hp_layer_desc CNN::LActivation<TanH>::describe() const
{
    hp_layer_desc d = {};
    d.kind = HP_LAYER_TANH;
    d.in_dims[0] = n;
    return d;
}
template <class F> inline hp_layer_desc CNN::LActivation<F>::describe() const
{
    hp_layer_desc d = {};
    d.kind = 0;  // Sigmoid / ReLU / LeakyReLU: no device kernels
    d.in_dims[0] = n;
    return d;
}

// cnn.h:606-611 (the per-layer operators of the reference stream one layer's own vectors, which
// do not exist here; the whole-net operators are kept)
inline std::istream &operator>>(std::istream &in, CNN &nn) { nn.loada(in); return in; }
inline std::ostream &operator<<(std::ostream &ot, const CNN &nn) { nn.savea(ot); return ot; }

#endif  // MINI_CNN_H
