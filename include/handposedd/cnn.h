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
struct i2 {
    int x, y;
    i2(int x_, int y_) : x(x_), y(y_) {}
    template <class V, class = decltype(V().x + V().y)> i2(const V &v) : x(v.x), y(v.y) {}
};
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
// which activation tags have device kernels (declared before CNN so that LActivation<F>::describe needs no
// specialisation after its first use)
template <class F> struct act_kind { static constexpr int value = 0; };
template <> struct act_kind<TanH> { static constexpr int value = HP_LAYER_TANH; };
// The device net behind a CNN and all of its by-value copies.  Allocated when the CNN is constructed -- not when the
// net is first used -- so that copies made before first use (include/handtrack.h:129 returns the CNN by value right
// after building it) share one weight store exactly like the reference's copies share their layer pointers.
struct NetHolder {
    hp_net *h = nullptr;
    size_t built_for = 0;
    ~NetHolder() { if (h) hp_destroy(h); }
};
}  // namespace handposedd

struct CNN {
    struct LBase {
        virtual ~LBase() {}
        virtual hp_layer_desc describe() const = 0;
        // cnn.h:107-110: per-layer streams; layers without weights stream nothing
        virtual void loada(std::istream &) {}
        virtual void savea(std::ostream &) const {}
        virtual void loadb(std::istream &) {}
        virtual void saveb(std::ostream &) const {}
    };
    // A weighted layer's W then B is one contiguous float range [first, first + count) of the device weight store
    // (.cnnb order).  The range is bound when the owning CNN builds its device net (CNN::net()); a layer that belongs
    // to no built net has nowhere to keep weights (there are no host-side vectors here) and its streams throw.
    struct LWeighted : public LBase {
        void loada(std::istream &s) override
        {
            std::vector<float> p = fetch();
            for (auto &w : p) s >> w;
            store(p);
        }
        void savea(std::ostream &s) const override
        {
            for (float w : fetch()) s << w << ' ';
        }
        void loadb(std::istream &s) override
        {
            std::vector<float> p = fetch();   // a short stream leaves the tail unmodified, like loadvb (cnn.h:97)
            s.read((char *)p.data(), (std::streamsize)(p.size() * sizeof(float)));
            store(p);
        }
        void saveb(std::ostream &s) const override
        {
            std::vector<float> p = fetch();
            s.write((const char *)p.data(), (std::streamsize)(p.size() * sizeof(float)));
        }
        void bind(std::shared_ptr<handposedd::NetHolder> h, int64_t first, int64_t count) { holder_ = h; first_ = first; count_ = count; }

      private:
        std::shared_ptr<handposedd::NetHolder> owner() const
        {
            auto h = holder_.lock();
            if (!h || !h->h) throw std::runtime_error("handposedd: layer is not part of a live, built CNN (its weights live in the net's device store)");
            return h;
        }
        std::vector<float> fetch() const
        {
            std::vector<float> p((size_t)count_);
            handposedd::check(hp_get_params_range(owner()->h, first_, count_, p.data()));
            return p;
        }
        void store(const std::vector<float> &p) { handposedd::check(hp_set_params_range(owner()->h, first_, count_, p.data())); }
        // weak: the layer objects are never freed (like the reference's), they must not keep a device net alive
        std::weak_ptr<handposedd::NetHolder> holder_;
        int64_t first_ = 0, count_ = 0;
    };
    struct LConv final : public LWeighted {  // cnn.h:194-290
        handposedd::i3 indims;
        handposedd::i4 dims;
        handposedd::i3 outdims;
        LConv(handposedd::i3 indims, handposedd::i4 dims, handposedd::i3 outdims) : indims(indims), dims(dims), outdims(outdims) {}
        int64_t weight_count() const { return (int64_t)dims.x * dims.y * dims.z * dims.w + dims.w; }
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
        hp_layer_desc describe() const override
        {
            hp_layer_desc d = {};
            d.kind = handposedd::act_kind<F>::value;  // Sigmoid / ReLU / LeakyReLU: 0, no device kernels
            d.in_dims[0] = n;
            return d;
        }
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
    struct LFull final : public LWeighted {  // cnn.h:398-456
        int M, N;
        LFull(int input_size, int output_size) : M(input_size), N(output_size) {}
        int64_t weight_count() const { return (int64_t)M * N + N; }
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
    struct LConvS final : public LUnsupported {  // cnn.h:292-396: same-size convolution with radius / stride
        handposedd::i2 rdims, radius, stride;
        int din, dout;
        LConvS(handposedd::i2 rdims, int din, int dout, handposedd::i2 radius = {1, 1}, handposedd::i2 stride = {1, 1})
            : rdims(rdims), radius(radius), stride(stride), din(din), dout(dout) {}
    };
    struct LSoftMax final : public LUnsupported { LSoftMax(int) {} };
    struct LCrossEntropy final : public LUnsupported { LCrossEntropy(int) {} };

    std::vector<LBase *> layers;  // cnn.h:548 (raw, never freed, exactly like the reference)
    int precision = HP_PRECISION_FP32;  // new: HP_PRECISION_TENSOR selects the tcgen05 path
    int device = 0;

    // cnn.h:595-604 ("quick test for simple NNs": LFull + TanH pairs).  The empty list -- the only use in the
    // reference, `CNN cnn({})` at include/handtrack.h:107 -- builds nothing; a non-empty list describes an MLP this
    // library has no kernels for, so Init() throws HP_ERR_UNSUPPORTED from inside the constructor instead of silently
    // building a CPU network ("What does not carry over").
    CNN(const std::vector<int> &s) : holder_(std::make_shared<handposedd::NetHolder>())
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

    // the device net behind this object and its copies, created on first use from `layers`; binds every weighted
    // layer to its float range of the store (.cnnb order: layers in order, W then B)
    hp_net *net() const
    {
        handposedd::NetHolder &H = *holder_;
        if (!H.h || H.built_for != layers.size()) {
            std::vector<hp_layer_desc> d;
            for (auto *l : layers) d.push_back(l->describe());
            hp_net *h = nullptr;
            handposedd::check(hp_create(d.data(), (int)d.size(), device, &h));
            if (H.h) hp_destroy(H.h);
            H.h = h;
            H.built_for = layers.size();
            int64_t off = 0;
            for (auto *l : layers) {
                int64_t cnt = 0;
                if (auto *c = dynamic_cast<LConv *>(l)) cnt = c->weight_count();
                else if (auto *f = dynamic_cast<LFull *>(l)) cnt = f->weight_count();
                if (auto *w = dynamic_cast<LWeighted *>(l)) w->bind(holder_, off, cnt);
                off += cnt;
            }
        }
        return H.h;
    }

  private:
    std::shared_ptr<handposedd::NetHolder> holder_;  // shared by copies from construction on: one weight store
};

// cnn.h:606-611.  The per-layer operators stream that layer's W then B range of the device store (the layer must
// belong to a CNN whose device net has been built: any Eval / Train / Init / load / save call does that).
inline std::istream &operator>>(std::istream &in, CNN::LConv &cl) { cl.loada(in); return in; }
inline std::ostream &operator<<(std::ostream &ot, const CNN::LConv &cl) { cl.savea(ot); return ot; }
inline std::istream &operator>>(std::istream &in, CNN::LFull &cl) { cl.loada(in); return in; }
inline std::ostream &operator<<(std::ostream &ot, const CNN::LFull &cl) { cl.savea(ot); return ot; }
inline std::istream &operator>>(std::istream &in, CNN &nn) { nn.loada(in); return in; }
inline std::ostream &operator<<(std::ostream &ot, const CNN &nn) { nn.savea(ot); return ot; }

#endif  // MINI_CNN_H
