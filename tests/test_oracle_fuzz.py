"""Pins the plain-C oracle to the unmodified reference (oracle/_ref/libcnnref.so) where the arithmetic is least forgiving:
saturated tanh (exact +-1 ties in both max-pools and zero gradients), tanh arguments above 44.4 (NaN, cnn.h:31), softmax
overflow without max-subtraction (cnn.h:499), all-zero / all-one crops, small random weights.  Eval outputs and per-sample
gradients must agree bit for bit, NaN positions included.  CPU only."""
import numpy as np
import pytest

from hand_tracking_samples_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libcnnref.so not built (needs /root/reference)")


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


def test_oracle_equals_reference_under_saturation_nan_and_overflow():
    o, r = orc.Oracle(), orc.Ref()
    r.init()
    p0 = r.save()
    seen_nan = seen_tie = False
    for seed in range(24):
        rng = np.random.default_rng(seed)
        p, k = p0.copy(), seed % 6
        scale = {1: ("fc2.W", 30), 2: ("conv1.W", 40), 3: ("fc1.W", 200), 4: ("fc2.W", 3000)}.get(k)
        if scale:
            off, n = orc.LAYOUT[scale[0]]
            p[off:off + n] *= scale[1]
        elif k == 5:
            p = rng.normal(0, 0.05, p.shape).astype(np.float32)
        x = {0: synth.uniform_crops(2, seed), 1: synth.depthlike_crops(2, seed), 2: np.zeros((2, 4096), np.float32),
             3: np.ones((2, 4096), np.float32)}[seed % 4]
        t = synth.heatmap_labels(2, seed + 7)
        r.load(p)
        y_ref, y_orc = r.eval(x), o.eval(p, x)
        assert same_bits(y_ref, y_orc), ("eval", seed)
        g_ref, g_orc = r.grad_sample(x[0], t[0]), o.grad_sample(p, x[0], t[0])
        g_ref = g_ref[0] if isinstance(g_ref, tuple) else g_ref
        g_orc = g_orc[0] if isinstance(g_orc, tuple) else g_orc
        assert same_bits(g_ref, g_orc), ("grad", seed)
        seen_nan |= bool(np.isnan(y_ref).any())
        seen_tie |= k == 2
    assert seen_nan and seen_tie    # the generator really reached the NaN and saturation regimes
