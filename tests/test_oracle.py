"""CPU tests: the plain-C oracle against the committed golden vectors (generated from the
unmodified reference, tests/golden/make_golden.py) and, where oracle/_ref is present,
bit-exactly against the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.oracle import LAYOUT, N_PARAMS, Oracle, Ref, have_ref

META = json.load(open(os.path.join(GOLDEN, "meta.json")))
STRIDE = META["stride"]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def p0(orc):
    return orc.init_xavier()


def sample(params):
    return {k: (params[off:off + n][::STRIDE] if n > 100000 else params[off:off + n]) for k, (off, n) in LAYOUT.items()}


def test_init_known_answers(p0):
    # libstdc++ known answers for CNN::Init (cnn.h:581)
    assert abs(p0[0] - (-0.118816)) < 1e-6
    assert abs(p0[416] - (-0.0156417)) < 1e-7
    assert abs(p0[16864] - (-0.0280718)) < 1e-7
    assert abs(p0[4737504] - (-0.0141025)) < 1e-7
    for k, v in META["known"].items():
        name, idx = k.split("[")
        assert p0[LAYOUT[name][0] + int(idx[:-1])] == np.float32(v)
    assert hashlib.sha256(p0.tobytes()).hexdigest() == META["init_sha256"]
    for b in ("conv1.B", "conv2.B", "fc1.B", "fc2.B"):
        off, n = LAYOUT[b]
        assert not p0[off:off + n].any()


def test_cnnb_layout_offsets():
    # SURVEY.md 8c offsets table, in floats
    assert LAYOUT["conv1.B"][0] * 4 == 1600 and LAYOUT["conv2.W"][0] * 4 == 1664
    assert LAYOUT["conv2.B"][0] * 4 == 67200 and LAYOUT["fc1.W"][0] * 4 == 67456
    assert LAYOUT["fc1.B"][0] * 4 == 18941824 and LAYOUT["fc2.W"][0] * 4 == 18950016
    assert LAYOUT["fc2.B"][0] * 4 == 37824384 and N_PARAMS * 4 == 37833600


def test_eval_matches_golden_bit_exact(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    want = np.load(os.path.join(GOLDEN, "eval_init.npy"))
    got = orc.eval(p0, crops)
    assert np.array_equal(got, want)
    # every span is a softmax (cnn.h:497-511)
    sums = np.concatenate([got[:, :2048].reshape(-1, 8, 256).sum(-1), got[:, 2048:].reshape(-1, 16, 16).sum(-1)], 1)
    assert np.allclose(sums, 1.0, atol=2e-6)


def test_eval_peaky_matches_golden(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    pk = p0.copy()
    off, n = LAYOUT["fc2.W"]
    pk[off:off + n] *= 30.0
    assert np.array_equal(orc.eval(pk, crops), np.load(os.path.join(GOLDEN, "eval_peaky.npy")))


def test_trace_matches_golden(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    tr = np.load(os.path.join(GOLDEN, "trace_crop0.npz"))
    orc.eval(p0, crops[:1])
    assert np.array_equal(orc.peek(3), tr["pool1"])
    assert np.array_equal(orc.peek(6), tr["pool2"])
    assert np.array_equal(orc.peek(8), tr["fc1"])
    assert np.array_equal(orc.peek(9), tr["logits"])


def test_grads_match_golden(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    labels = np.load(os.path.join(GOLDEN, "labels.npy"))
    gold = np.load(os.path.join(GOLDEN, "grads_init.npz"))
    for i in (0, 2, 4):
        g, mse = orc.grad_sample(p0, crops[i], labels[i])
        assert np.float32(mse) == gold["mse"][i]
        for k, v in sample(g).items():
            assert np.array_equal(v, gold["%d/%s" % (i, k)]), (i, k)


def test_train24_matches_golden(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    labels = np.load(os.path.join(GOLDEN, "labels.npy"))
    gold = np.load(os.path.join(GOLDEN, "train24.npz"))
    p = p0.copy()
    mse = orc.train_seq(p, np.concatenate([crops] * 4), np.concatenate([labels] * 4), 0.001)
    assert np.array_equal(mse, gold["mse"])
    assert hashlib.sha256(p.tobytes()).hexdigest() == META["train24_sha256"]
    assert np.array_equal(orc.eval(p, crops), np.load(os.path.join(GOLDEN, "eval_train24.npy")))


def test_minibatch_is_sum_of_sample_grads(orc, p0):
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))[:3]
    labels = np.load(os.path.join(GOLDEN, "labels.npy"))[:3]
    p = p0.copy()
    gsum, mse = orc.train_minibatch(p, crops, labels, 0.001, apply=True)
    want = np.zeros(N_PARAMS, np.float64)
    for i in range(3):
        g, m = orc.grad_sample(p0, crops[i], labels[i])
        want += g
        assert np.float32(m) == mse[i]
    assert np.array_equal(gsum, want)
    assert np.array_equal(p, (p0 - (0.001 * want).astype(np.float32)).astype(np.float32)) or \
        np.allclose(p, p0 - 0.001 * want, atol=1e-9)


def test_tanh_nan_and_saturation(orc, p0):
    # SURVEY.md 8a note 2: (exp(2t)-1)/(exp(2t)+1) is NaN for t >~ 44.4; a crop of 1e3 drives conv1 there
    y = orc.eval(p0, np.full((1, 4096), 1e3, np.float32))
    assert np.isnan(y).any()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_is_bit_exact_against_reference(orc, p0):
    r = Ref()
    r.init()
    assert np.array_equal(r.save(), p0)
    rng = np.random.default_rng(7)
    x = rng.random((2, 4096), dtype=np.float32)
    x[1, rng.random(4096) < 0.6] = 0
    t = rng.random((2, 2304), dtype=np.float32) * 0.05
    assert np.array_equal(r.eval(x), orc.eval(p0, x))
    g_ref, m_ref, errs = r.grad_sample(x[1], t[1], want_errors=True)
    g_orc, m_orc = orc.grad_sample(p0, x[1], t[1])
    assert np.array_equal(g_ref, g_orc) and np.float32(m_ref) == np.float32(m_orc)
    assert np.array_equal(errs[9], orc.peek(109)) and np.array_equal(errs[4], orc.peek(104))
    assert np.array_equal(errs[0], orc.peek(100))
    p = p0.copy()
    m1 = orc.train_seq(p, x, t, 0.001)
    m2 = r.train_seq(x, t, 0.001)
    assert np.array_equal(m1, m2) and np.array_equal(r.save(), p)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_saveb_file_layout(tmp_path, p0):
    # the real file writer CNN::saveb(std::string) (cnn.h:593): headerless, 37,833,600 bytes, .cnnb order
    r = Ref()
    r.load(p0)
    path = str(tmp_path / "w.cnnb")
    r.saveb_file(path)
    raw = np.fromfile(path, np.float32)
    assert raw.size == N_PARAMS and np.array_equal(raw, p0)


# ---- decode and depth normalisation ---------------------------------
def _tricky_outputs():
    rng = np.random.default_rng(0)
    y = rng.random((40, 2304), dtype=np.float32)
    y[3, :256] = 0            # empty heatmap: wsum == 0 branch of PeakSubPixel
    y[4, 256:512] = 0.5       # constant heatmap: ImageFindMax keeps pixel (0,0)
    y[5, 2048:2064] = 0       # empty 1-D heatmap
    y[6, 512 + 255] = 9.0     # peak in the last corner (clamped 3x3 window)
    y[7, 2048 + 16 * 3 + 15] = 9.0
    return y


def test_decode_matches_golden(orc):
    for name in ("init", "peaky"):
        y = np.load(os.path.join(GOLDEN, "eval_%s.npy" % name))
        assert np.array_equal(orc.decode(y), np.load(os.path.join(GOLDEN, "decode_%s.npy" % name)))


def test_depth_normalisation_matches_golden(orc):
    g = np.load(os.path.join(GOLDEN, "depth_norm.npz"))
    x = orc.normalize_depth(g["depth"])
    assert np.array_equal(x, g["x"]) and x.min() == 0.0 and x.max() == 1.0


def test_decode_and_normalise_bit_exact_against_reference(orc):
    from oracle.oracle import PostRef, have_postref
    if not have_postref():
        pytest.skip("oracle/_ref/libpostref.so not built")
    r = PostRef()
    y = _tricky_outputs()
    assert np.array_equal(orc.decode(y), r.decode(y))
    d = np.random.default_rng(1).integers(0, 65536, (3, 4096)).astype(np.uint16)
    assert np.array_equal(orc.normalize_depth(d), r.normalize_depth(d))
    assert np.array_equal(orc.normalize_depth(d, 0.000125, 0.2, 0.9), r.normalize_depth(d, 0.000125, 0.2, 0.9))


def test_label_rendering_matches_golden_and_reference(orc):
    g = np.load(os.path.join(GOLDEN, "labels_render.npz"))
    want = g["t_u8"].astype(np.float32) / np.float32(255.0)
    got = orc.render_labels(g["points"], g["vals"])
    assert np.array_equal(got, want)
    # u8 truncation makes span sums fall short of 1 (SURVEY.md 8a note 5)
    assert 22.0 < got[5].sum() < 24.0
    from oracle.oracle import PostRef, have_postref
    if have_postref():
        rng = np.random.default_rng(4)
        p = rng.uniform(-1, 17, (500, 16)).astype(np.float32)
        v = rng.uniform(-0.1, 1.1, (500, 16)).astype(np.float32)
        assert np.array_equal(orc.render_labels(p, v), PostRef().render_labels(p, v))
