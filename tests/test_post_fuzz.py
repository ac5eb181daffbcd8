"""Adversarial inputs for the rows either side of the CNN: heatmaps full of ties, NaN, inf and
negative values for the decode; label parameters off the map, on integer pixels, NaN; random depth ranges for the
normalisation.  The plain-C oracle against the reference's own routines (oracle/_ref/libpostref.so), and -- with the same
generators -- the device kernels against the oracle, bit for bit (the first device run of these generators found the
NaN-at-(0,0) PeakVolume case documented in csrc/hp_post.cu)."""
import numpy as np
import pytest

from oracle import oracle as orc


def same_bits(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))


def heatmaps(seed, n=4):
    rng = np.random.default_rng(seed)
    kind = seed % 8
    if kind == 0:
        return rng.random((n, 2304), dtype=np.float32)
    if kind == 1:
        return np.zeros((n, 2304), np.float32)                              # every value ties
    if kind == 2:
        y = np.zeros((n, 2304), np.float32)
        for i in range(n):
            y[i, rng.integers(0, 2304, 40)] = 1.0                           # many equal peaks, on the borders too
        return y
    if kind == 3:
        y = rng.random((n, 2304), dtype=np.float32)
        y[:, ::257] = np.nan
        return y
    if kind == 4:
        return rng.random((n, 2304), dtype=np.float32) ** 8                 # peaky
    if kind == 5:
        return np.float32(rng.integers(0, 4, (n, 2304))) / 3                # quantised: ties everywhere
    if kind == 6:
        return rng.normal(0, 1, (n, 2304)).astype(np.float32)               # negative values
    y = rng.random((n, 2304), dtype=np.float32)
    y[:, :256] = 0
    y[:, 0] = np.inf
    return y


def label_params(seed, n=6):
    rng = np.random.default_rng(1000 + seed)
    k = seed % 5
    pts = rng.uniform(-3, 19, (n, 8, 2)).astype(np.float32) if k < 3 else rng.integers(-2, 18, (n, 8, 2)).astype(np.float32)
    vals = rng.uniform(-0.3, 1.3, (n, 16)).astype(np.float32) if k != 2 else rng.choice([0.0, 1.0, 0.5, -0.0], (n, 16)).astype(np.float32)
    if k == 4:
        pts[0, 0] = [np.nan, 3]
        vals[0, 0] = np.nan
    return pts, vals


def depth_case(seed):
    rng = np.random.default_rng(5000 + seed)
    d = rng.integers(0, 65536, (3, 4096)).astype(np.uint16)
    dmin = float(rng.uniform(0.05, 0.3))
    return d, float(rng.choice([0.001, 0.000124987, 0.0001, 0.00025])), dmin, dmin + float(rng.uniform(0.1, 1.0))


@pytest.mark.skipif(not orc.have_postref(), reason="oracle/_ref/libpostref.so not built (needs /root/reference)")
def test_oracle_equals_reference_on_adversarial_inputs():
    o, r = orc.Oracle(), orc.PostRef()
    for seed in range(160):
        y = heatmaps(seed)
        assert same_bits(o.decode(y), r.decode(y)), ("decode", seed)
        pts, vals = label_params(seed)
        assert same_bits(o.render_labels(pts, vals), r.render_labels(pts, vals)), ("labels", seed)
    for seed in range(60):
        d, sc, dmin, dmax = depth_case(seed)
        assert same_bits(o.normalize_depth(d, sc, dmin, dmax), r.normalize_depth(d, sc, dmin, dmax)), ("normalize", seed)


@pytest.mark.gpu
def test_device_kernels_equal_oracle_on_adversarial_inputs():
    import torch
    from hand_tracking_samples_b200 import capi, cnn as hp
    net = hp.PoseInitializerCNN("")
    o = orc.Oracle()
    st = torch.cuda.current_stream().cuda_stream
    for seed in range(48):
        y = heatmaps(seed)
        assert same_bits(net.decode_batch(y), o.decode(y)), ("decode", seed)
        pts, vals = label_params(seed)
        assert same_bits(net.render_labels(pts, vals), o.render_labels(pts, vals)), ("labels", seed)
    for seed in range(16):
        d, sc, dmin, dmax = depth_case(seed)
        dd = torch.from_numpy(d.view(np.int16)).cuda()
        x = torch.empty((3, 4096), device="cuda")
        capi.check(net.L.hp_normalize_depth_device(net.h, dd.data_ptr(), 3, sc, dmin, dmax, x.data_ptr(), st))
        torch.cuda.synchronize()
        assert same_bits(x.cpu().numpy(), o.normalize_depth(d, sc, dmin, dmax)), ("normalize", seed)
