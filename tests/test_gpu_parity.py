"""GPU parity tests: the CUDA path, called through the C ABI (include/handposedd.h), against
the CPU oracle on the same seeded inputs, against the committed golden vectors produced by the
unmodified reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Metric: max-normalised error max|a-b|/max|b| per tensor.

Bounds (BASELINE.json north_star): FP32 path <= 1e-5 on outputs and gradients; tensor-core
path <= 1e-2 on outputs.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, maxnorm_err, record
from hand_tracking_samples_b200 import cnn as hp
from hand_tracking_samples_b200 import synth
from oracle import oracle as orc_mod
from oracle.oracle import LAYOUT, N_PARAMS, Oracle

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
TC_TOL = 1e-2
META = json.load(open(os.path.join(GOLDEN, "meta.json")))
STRIDE = META["stride"]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def p0(orc):
    return orc.init_xavier()


@pytest.fixture()
def net():
    n = hp.PoseInitializerCNN("")
    yield n
    del n


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def peaky(p0):
    pk = p0.copy()
    off, n = LAYOUT["fc2.W"]
    pk[off:off + n] *= 30.0
    return pk


# ---- weights: Init and .cnnb ---------------------------------------------------------------
def test_init_is_bit_identical_to_reference(net, p0):
    data = net.saveb()
    assert len(data) == 37833600
    assert hashlib.sha256(data).hexdigest() == META["init_sha256"]
    assert np.array_equal(np.frombuffer(data, np.float32), p0)


def test_cnnb_roundtrip_and_short_read(net, p0, tmp_path):
    rng = np.random.default_rng(3)
    w = rng.standard_normal(N_PARAMS).astype(np.float32)
    net.set_params(w)
    path = str(tmp_path / "a.cnnb")
    net.saveb(path)
    assert open(path, "rb").read() == w.tobytes()          # byte-identical file, reference layout
    net.Init()
    net.loadb(path)
    assert net.saveb() == w.tobytes()
    # short stream: prefix loaded, tail untouched (loadvb, cnn.h:97)
    net.Init()
    net.loadb(w.tobytes()[:1664 + 10])                      # conv1.W, conv1.B and 2.5 floats of conv2.W
    got = np.frombuffer(net.saveb(), np.float32)
    assert np.array_equal(got[:418], w[:418]) and np.array_equal(got[418:], p0[418:])
    # missing file: silent no-op like CNN::loadb(std::string) (cnn.h:592)
    before = net.saveb()
    net.loadb(str(tmp_path / "does_not_exist.cnnb"))
    assert net.saveb() == before


def test_copies_share_the_weight_store(net, p0):
    twin = net.copy()
    net.set_params(p0 * 2)
    assert np.array_equal(np.frombuffer(twin.saveb(), np.float32), p0 * 2)
    del twin
    assert np.array_equal(np.frombuffer(net.saveb(), np.float32), p0 * 2)   # still alive after one owner died


# ---- Eval ----------------------------------------------------------------------------------
def test_eval_fp32_matches_golden(net):
    crops = golden("crops.npy")
    got = net.eval_batch(crops)
    want = golden("eval_init.npy")
    for i in range(crops.shape[0]):
        assert maxnorm_err(got[i], want[i]) <= FP32_TOL, i
    assert maxnorm_err(net.Eval(crops[2]), want[2]) <= FP32_TOL     # the reference's own single-crop call


def test_eval_fp32_peaky_and_trained_weights(net, p0):
    crops = golden("crops.npy")
    net.set_params(peaky(p0))
    assert maxnorm_err(net.eval_batch(crops), golden("eval_peaky.npy")) <= FP32_TOL


def test_eval_fp32_stage_by_stage(net, orc, p0):
    crops = golden("crops.npy")
    tr = golden("trace_crop0.npz")
    net.eval_batch(crops[:1])
    assert maxnorm_err(net.peek(3, 1, 3600)[0], tr["pool1"]) <= FP32_TOL      # conv1+tanh+pool+pool
    assert maxnorm_err(net.peek(6, 1, 2304)[0], tr["pool2"]) <= FP32_TOL      # conv2+tanh+pool
    assert maxnorm_err(net.peek(8, 1, 2048)[0], tr["fc1"]) <= FP32_TOL        # fc1+tanh
    assert maxnorm_err(net.peek(9, 1, 2304)[0], tr["logits"]) <= FP32_TOL     # fc2


def test_eval_fp32_vs_oracle_seeded_and_ragged(net, orc, p0):
    x = np.concatenate([synth.uniform_crops(5, 11), synth.depthlike_crops(6, 12)])
    want = orc.eval(p0, x)
    got = net.eval_batch(x)
    assert maxnorm_err(got, want) <= FP32_TOL
    for n in (1, 2, 3, 7):                                   # ragged batch sizes give the same per-crop result
        assert np.array_equal(net.eval_batch(x[:n]), got[:n])
    assert net.eval_batch(np.zeros((0, 4096), np.float32)).shape == (0, 2304)    # empty batch


def test_eval_fp32_ragged_across_the_small_batch_boundary(net, orc, p0):
    # calls of <= 64 crops use split-K FC kernels (another summation order, hp_fp32.cu fc_small): the choice is per CALL,
    # so inside one call a crop's result does not depend on where it sits or on the workspace chunking
    x = synth.depthlike_crops(2048 + 3, 13)
    want = orc.eval(p0, x[:65])
    y_big = net.eval_batch(x)                                   # 2048-crop chunk + 3-crop tail chunk, one call
    assert np.array_equal(net.eval_batch(x[2048 - 70:])[-3:], y_big[-3:])          # tail chunk == same crops in another large call
    assert np.array_equal(net.eval_batch(x[:65]), y_big[:65])                      # 65 crops: large-call arithmetic
    for n in (63, 64):
        got = net.eval_batch(x[:n])
        assert maxnorm_err(got, want[:n]) <= FP32_TOL
        record("fp32 small-call vs large-call arithmetic, n=%d" % n, maxnorm_err(got, y_big[:n]))
        assert maxnorm_err(got, y_big[:n]) <= 2e-6
    assert maxnorm_err(y_big[:65], want) <= FP32_TOL


def test_eval_nan_propagation_like_reference(net, orc, p0):
    # TanH::f is NaN above ~44.4 (SURVEY.md 8a note 2); same crops must go NaN on both sides
    x = np.stack([np.full(4096, 1e3, np.float32), np.zeros(4096, np.float32), np.full(4096, -1e3, np.float32)])
    want = orc.eval(p0, x)
    got = net.eval_batch(x)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert maxnorm_err(got[ok], want[ok]) <= FP32_TOL


def test_eval_chunking_and_device_entry_agree(net):
    import torch
    n = 2048 + 517                                           # crosses the FP32 workspace chunk
    x = synth.uniform_crops(n, 5)
    y_host = net.eval_batch(x)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty((n, 2304), device="cuda")
    net.eval_batch_device(xd.data_ptr(), n, yd.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy(), y_host)
    xp = torch.from_numpy(x).pin_memory()                    # pinned host buffers take the zero-bounce route
    yp = torch.empty((n, 2304)).pin_memory()
    net.eval_batch(xp.numpy(), out=yp.numpy())
    assert np.array_equal(yp.numpy(), y_host)
    sums = y_host[:, :2048].reshape(n, 8, 256).sum(-1)
    assert np.allclose(sums, 1.0, atol=1e-5)


# ---- Train ---------------------------------------------------------------------------------
def sample(params):
    return {k: (params[off:off + n][::STRIDE] if n > 100000 else params[off:off + n]) for k, (off, n) in LAYOUT.items()}


def grads_of(net, x, t):
    import torch
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    td = torch.from_numpy(np.ascontiguousarray(t, np.float32)).cuda()
    mse = torch.empty(xd.shape[0], device="cuda")
    net.grad_batch_device(xd.data_ptr(), td.data_ptr(), xd.shape[0], mse.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return net.get_grads(), mse.cpu().numpy()


def test_single_sample_gradients_match_golden(net):
    crops, labels, gold = golden("crops.npy"), golden("labels.npy"), golden("grads_init.npz")
    for i in range(crops.shape[0]):
        g, mse = grads_of(net, crops[i:i + 1], labels[i:i + 1])
        assert abs(mse[0] - gold["mse"][i]) <= 1e-5 * gold["mse"][i]
        for k, v in sample(g).items():
            assert maxnorm_err(v, gold["%d/%s" % (i, k)]) <= FP32_TOL, (i, k)


def test_backward_stage_by_stage(net, orc, p0):
    x, t = synth.depthlike_crops(1, 21), synth.heatmap_labels(1, 22)
    g_want, _ = orc.grad_sample(p0, x[0], t[0])
    g, _ = grads_of(net, x, t)
    assert maxnorm_err(net.peek(109, 1, 2304)[0], orc.peek(109)) <= FP32_TOL   # softmax backward
    assert maxnorm_err(net.peek(107, 1, 2048)[0], orc.peek(107)) <= FP32_TOL   # fc2 dX * tanh'
    # winners-only forms: compare at the non-zero entries of the oracle's dense tensors
    e4 = orc.peek(104).reshape(64, 12, 12)
    g2 = net.peek(106, 1, 2304)[0].reshape(64, 6, 6)
    assert maxnorm_err(g2, e4.reshape(64, 6, 2, 6, 2).transpose(0, 1, 3, 2, 4).reshape(64, 6, 6, 4).sum(-1)) <= FP32_TOL
    e0 = orc.peek(100).reshape(16, 15, 4, 15, 4).transpose(0, 1, 3, 2, 4).reshape(16, 15, 15, 16).sum(-1)
    assert maxnorm_err(net.peek(103, 1, 3600)[0].reshape(16, 15, 15), e0) <= FP32_TOL
    for k, (off, n) in LAYOUT.items():
        assert maxnorm_err(g[off:off + n], g_want[off:off + n]) <= FP32_TOL, k


def test_minibatch_gradient_is_sum_of_reference_sample_gradients(net, orc, p0):
    n = 9
    x = np.concatenate([synth.depthlike_crops(5, 31), synth.uniform_crops(4, 32)])
    t = synth.heatmap_labels(n, 33)
    want, mse_want = orc.train_minibatch(p0.copy(), x, t, 0.001, apply=False)
    g, mse = grads_of(net, x, t)
    assert np.allclose(mse, mse_want, rtol=1e-5)
    for k, (off, cnt) in LAYOUT.items():
        assert maxnorm_err(g[off:off + cnt], want[off:off + cnt]) <= FP32_TOL, k


def test_train_batch1_sequence_matches_golden(net):
    # 24 sequential CNN::Train steps exactly as train-cnn.cpp:160 issues them (alpha = 0.001)
    crops, labels, gold = golden("crops.npy"), golden("labels.npy"), golden("train24.npz")
    xs, ts = np.concatenate([crops] * 4), np.concatenate([labels] * 4)
    mse = np.array([net.Train(xs[i], ts[i], 0.001) for i in range(24)], np.float32)
    assert np.allclose(mse, gold["mse"], rtol=2e-5)
    for k, v in sample(net.get_params()).items():
        assert maxnorm_err(v, gold[k]) <= FP32_TOL, k
    assert maxnorm_err(net.eval_batch(crops), golden("eval_train24.npy")) <= FP32_TOL


def test_minibatch_step_matches_oracle(net, orc, p0):
    n = 6
    x, t = synth.depthlike_crops(n, 41), synth.heatmap_labels(n, 42)
    p = p0.copy()
    _, mse_want = orc.train_minibatch(p, x, t, 0.001, apply=True)
    mse = net.train_batch(x, t, 0.001)
    assert np.allclose(mse, mse_want, rtol=1e-5)
    got = net.get_params()
    for k, (off, cnt) in LAYOUT.items():
        assert maxnorm_err(got[off:off + cnt], p[off:off + cnt]) <= FP32_TOL, k
    # update really is W - alpha * grads
    assert np.abs(got - p0).max() > 0


def test_train_chunked_batch_equals_unchunked_sum(net, p0):
    # n above the workspace chunk accumulates gradients across chunks at frozen weights
    n = 2048 + 64
    x, t = synth.uniform_crops(n, 51), synth.heatmap_labels(64, 52)
    t = np.concatenate([t] * 33)[:n]
    g_all, _ = grads_of(net, x, t)
    g_a, _ = grads_of(net, x[:2048], t[:2048])
    g_b, _ = grads_of(net, x[2048:], t[2048:])
    for k, (off, cnt) in LAYOUT.items():
        assert maxnorm_err(g_all[off:off + cnt], (g_a + g_b)[off:off + cnt]) <= 2e-6, k


# ---- tensor-core path ----------------------------------------------------------------------
def test_eval_tensor_path_within_bound(net, orc, p0):
    x = np.concatenate([golden("crops.npy"), synth.depthlike_crops(10, 61), synth.uniform_crops(10, 62)])
    want = orc.eval(p0, x)
    got = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    errs = [maxnorm_err(got[i], want[i]) for i in range(x.shape[0])]
    record("tensor eval, Init() weights, worst crop", max(errs))
    assert max(errs) <= TC_TOL
    # fc2.W x 30: peaky heatmaps (y_max ~ 0.999), the regime of a trained net.  fp16 forward operands hold the same bound.
    net.set_params(peaky(p0))
    want = orc.eval(peaky(p0), x)
    got = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    errs = [maxnorm_err(got[i], want[i]) for i in range(x.shape[0])]
    record("tensor eval, peaky (fc2.W x30) weights, worst crop", max(errs))
    assert max(errs) <= TC_TOL
    assert maxnorm_err(got[:golden("crops.npy").shape[0]], golden("eval_peaky.npy")) <= TC_TOL


def test_eval_tensor_path_on_trained_weights(net, orc, p0):
    crops, labels = golden("crops.npy"), golden("labels.npy")
    # (a) the golden 24-step reference run: weights reproduced by the FP32 path (bit-compatible within 1e-5, see
    #     test_train_batch1_sequence_matches_golden), evaluated on the tensor path against the reference's own outputs
    xs, ts = np.concatenate([crops] * 4), np.concatenate([labels] * 4)
    for i in range(24):
        net.Train(xs[i], ts[i], 0.001)
    e = maxnorm_err(net.eval_batch(crops, precision=hp.PRECISION_TENSOR), golden("eval_train24.npy"))
    record("tensor eval on the golden train24 weights", e)
    assert e <= TC_TOL
    # (b) weights after a real stretch of training (300 minibatch steps on 64 samples, enough for the heatmaps to sharpen)
    x, t = synth.depthlike_crops(64, 93), synth.heatmap_labels(64, 94)
    for _ in range(300):
        net.train_batch(x, t, 0.05 / 64)
    p = net.get_params()
    xe = np.concatenate([x[:12], synth.depthlike_crops(12, 95)])
    want = orc_mod.eval_mt(p, xe)
    got = net.eval_batch(xe, precision=hp.PRECISION_TENSOR)
    errs = [maxnorm_err(got[i], want[i]) for i in range(xe.shape[0])]
    record("tensor eval after 300 training steps, worst crop", max(errs))
    record("  peak softmax value of those outputs", float(want.max()))
    assert max(errs) <= TC_TOL
    assert maxnorm_err(net.eval_batch(xe), want) <= FP32_TOL


def test_eval_tensor_path_nan_quirk(net, orc, p0):
    # TanH::f = (e-1)/(e+1) with e = exp(2t) is NaN above t ~ 44.4 (SURVEY.md 8a note 2).  The tensor path evaluates the
    # same formula on the MUFU units (hp_tc.cuh, tanh_tc), so saturating crops go NaN on both sides.  Stated divergence:
    # the tensor path pools BEFORE tanh, so a window whose NaN is not at the first scan position is NaN here while
    # std::max(m, NaN) keeps m in the reference (cnn.h:146); uniform saturating crops have no such windows.
    x = np.stack([np.full(4096, 1e3, np.float32), np.zeros(4096, np.float32), np.full(4096, -1e3, np.float32)])
    want = orc.eval(p0, x)
    got = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.isnan(want[0]).any() and not np.isnan(want[1]).any()
    ok = ~np.isnan(want)
    assert maxnorm_err(got[ok], want[ok]) <= TC_TOL


def test_eval_tensor_path_ragged_and_against_fp32(net):
    x = synth.depthlike_crops(300, 71)
    y32 = net.eval_batch(x)
    ytc = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    assert maxnorm_err(ytc, y32) <= TC_TOL
    for n in (1, 3, 127, 129, 257):
        assert np.array_equal(net.eval_batch(x[:n], precision=hp.PRECISION_TENSOR), ytc[:n])


# ---- full-size properties (BASELINE.json configs[1]: 65,536 crops) --------------------------
def test_full_size_properties_and_oracle_slice():
    import torch
    n = 65536
    net = hp.PoseInitializerCNN("")
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand((n, 4096), device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    # a 1,024-crop slice of the batch, strided over the whole of it, against the CPU oracle
    pick = torch.arange(0, n, n // 1024, device="cuda")[:1024]
    want = orc_mod.eval_mt(net.get_params(), x[pick].cpu().numpy())
    for prec, tol in ((hp.PRECISION_TENSOR, TC_TOL), (hp.PRECISION_FP32, FP32_TOL)):
        y = torch.empty((n, 2304), device="cuda")
        net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
        got = y[pick].cpu().numpy()
        errs = [maxnorm_err(got[i], want[i]) for i in range(1024)]
        record("65,536-crop batch, 1,024-crop oracle slice, precision %d, worst crop" % prec, max(errs))
        assert max(errs) <= tol
        sums = torch.cat([y[:, :2048].reshape(n, 8, 256).sum(-1), y[:, 2048:].reshape(n, 16, 16).sum(-1)], 1)
        assert (sums - 1).abs().max().item() <= 1e-5        # 24 softmaxes per crop
        # shard consistency: evaluating a slice == slicing the evaluation (the multi-GPU inference contract)
        lo, hi = 3 * n // 8, 4 * n // 8
        ys = torch.empty((hi - lo, 2304), device="cuda")
        net.eval_batch_device(x[lo:hi].data_ptr(), hi - lo, ys.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(ys, y[lo:hi])
        # determinism
        y2 = torch.empty_like(y)
        net.eval_batch_device(x.data_ptr(), n, y2.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(y, y2)


# ---- tensor-core training ---------------------------------------------------------------------
TC_GRAD_TOL = 1e-2   # 16-bit operands in every contraction (measured on B200: 1e-4 ... 4e-3)
# The conv WEIGHT gradients get a looser bound: fp16 pre-activations flip ~0.03-0.09 % of the max-pool winners
# (near-ties), each flip re-routes one window's gradient to a neighbouring patch, and because these gradient sums
# cancel heavily, the error of the sum is a few times sqrt(flip fraction): measured 3.5e-2 (conv1.W) / 1.7e-2 (conv2.W)
# at 9 samples, 6.1e-3 / 3.7e-3 at 256 samples.  The acceptance criterion BASELINE.json names for this path is the
# 1k-step loss curve (test_loss_curves_over_1k_steps: measured 5e-6), not per-step gradients.
TC_CONV_W_GRAD_TOL = 5e-2


def test_tensor_path_minibatch_gradients_close_to_oracle(net, orc, p0):
    import torch
    n = 9
    x = np.concatenate([synth.depthlike_crops(5, 31), synth.uniform_crops(4, 32)])
    t = synth.heatmap_labels(n, 33)
    want, mse_want = orc.train_minibatch(p0.copy(), x, t, 0.001, apply=False)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    mse = torch.empty(n, device="cuda")
    net.grad_batch_device(xd.data_ptr(), td.data_ptr(), n, mse.data_ptr(), precision=hp.PRECISION_TENSOR,
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = net.get_grads()
    assert np.allclose(mse.cpu().numpy(), mse_want, rtol=2e-2)
    for k, (off, cnt) in LAYOUT.items():
        tol = TC_CONV_W_GRAD_TOL if k in ("conv1.W", "conv2.W") else TC_GRAD_TOL
        e = maxnorm_err(g[off:off + cnt], want[off:off + cnt])
        record("9-sample tensor-path gradient, %s" % k, e)
        assert e <= tol, k


def test_minibatch256_step_vs_oracle_both_paths(net, p0):
    # BASELINE.json configs[2] at its own size: one 256-sample optimiser step against the sum of the reference's
    # per-sample gradients at frozen weights (oracle, sample-parallel on the host cores)
    import torch
    n = 256
    x = np.concatenate([synth.depthlike_crops(160, 131), synth.uniform_crops(96, 132)])
    t = synth.heatmap_labels(n, 133)
    want, mse_want = orc_mod.grad_minibatch_mt(p0, x, t)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    mse = torch.empty(n, device="cuda")
    for prec, tols in ((hp.PRECISION_FP32, None), (hp.PRECISION_TENSOR, (TC_GRAD_TOL, 2e-2))):
        net.grad_batch_device(xd.data_ptr(), td.data_ptr(), n, mse.data_ptr(), precision=prec, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        g = net.get_grads()
        assert np.allclose(mse.cpu().numpy(), mse_want, rtol=1e-5 if prec == hp.PRECISION_FP32 else 2e-2)
        for k, (off, cnt) in LAYOUT.items():
            e = maxnorm_err(g[off:off + cnt], want[off:off + cnt])
            record("256-sample gradient, precision %d, %s" % (prec, k), e)
            tol = FP32_TOL if tols is None else (tols[1] if k in ("conv1.W", "conv2.W") else tols[0])
            assert e <= tol, (prec, k, e)
    # and the update itself: W - alpha * grads on the FP32 path
    net.train_batch(x, t, 0.001 / n)
    p = net.get_params()
    pw = (p0.astype(np.float64) - (0.001 / n) * want).astype(np.float32)
    for k, (off, cnt) in LAYOUT.items():
        assert maxnorm_err(p[off:off + cnt], pw[off:off + cnt]) <= FP32_TOL, k


def test_fused_updates_leave_the_16bit_shadows_equal_to_a_full_rebuild():
    # The single-GPU tensor-path step updates the 16-bit operand copies inside its update kernels (sgd_refresh_fc: fp16
    # W^T and bf16 W of an FC layer; sgd_conv_images: the conv weight images) instead of rebuilding them from the fp32
    # master weights.  After a few steps -- the later ones replayed as a CUDA graph -- outputs and gradients must be
    # bit-identical to what a full rebuild from the master weights gives (set_params marks every shadow stale).
    import torch
    n = 96
    x = np.concatenate([synth.depthlike_crops(64, 301), synth.uniform_crops(32, 302)])
    t = synth.heatmap_labels(n, 303)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    mse = torch.empty(n, device="cuda")
    side = torch.cuda.Stream()
    fresh = hp.PoseInitializerCNN("")
    for _ in range(4):
        fresh.train_batch_device(xd.data_ptr(), td.data_ptr(), n, 0.01 / n, mse.data_ptr(), precision=hp.PRECISION_TENSOR, stream=side.cuda_stream)
    torch.cuda.synchronize()

    def probe():
        y = fresh.eval_batch(x[:32], precision=hp.PRECISION_TENSOR)
        fresh.grad_batch_device(xd.data_ptr(), td.data_ptr(), n, mse.data_ptr(), precision=hp.PRECISION_TENSOR, stream=side.cuda_stream)
        torch.cuda.synchronize()
        return y, fresh.get_grads()

    y_a, g_a = probe()
    fresh.set_params(fresh.get_params())
    y_b, g_b = probe()
    assert np.array_equal(y_a, y_b)
    assert np.array_equal(g_a, g_b)
    assert not np.array_equal(fresh.get_params(), hp.PoseInitializerCNN("").get_params())   # the steps did move the weights


def test_step_graph_survives_buffer_reallocation_between_steps():
    # A replayed training step holds raw device pointers and tensor maps.  An Eval with a larger batch between two
    # steps reallocates the activation buffers (tensor path) and the workspace (FP32 path): the next step must notice
    # (the graph is keyed on the allocation generation) and give the same weights as a net that never ran those Evals.
    import torch
    n = 96
    x = np.concatenate([synth.depthlike_crops(64, 311), synth.uniform_crops(32, 312)])
    t = synth.heatmap_labels(n, 313)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    side = torch.cuda.Stream()
    big = synth.uniform_crops(700, 314)
    for prec in (hp.PRECISION_TENSOR, hp.PRECISION_FP32):
        a, b = hp.PoseInitializerCNN(""), hp.PoseInitializerCNN("")
        for step in range(6):
            for m in (a, b):
                m.train_batch_device(xd.data_ptr(), td.data_ptr(), n, 0.01 / n, None, precision=prec, stream=side.cuda_stream)
            torch.cuda.synchronize()
            if step == 2:      # a's step graph exists by now (captured on the second call)
                a.eval_batch(big, precision=hp.PRECISION_TENSOR)
                a.eval_batch(big, precision=hp.PRECISION_FP32)
        assert np.array_equal(a.get_params(), b.get_params()), prec


def test_tensor_path_pool_winners_agree_with_fp32_path(net):
    import torch
    n = 16
    x, t = synth.depthlike_crops(n, 5), synth.heatmap_labels(n, 9)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
    got = {}
    for prec in (hp.PRECISION_FP32, hp.PRECISION_TENSOR):
        net.grad_batch_device(xd.data_ptr(), td.data_ptr(), n, None, precision=prec, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        i1, i2 = np.empty((n, 3600), np.uint8), np.empty((n, 2304), np.uint8)
        from hand_tracking_samples_b200 import capi
        capi.check(net.L.hp_peek(net.h, 203, n, i1.ctypes.data))
        capi.check(net.L.hp_peek(net.h, 206, n, i2.ctypes.data))
        got[prec] = (i1, i2, net.peek(3, n, 3600))
    record("tensor vs FP32 pool winners agreeing, conv1 stage", float((got[0][0] == got[1][0]).mean()))
    record("tensor vs FP32 pool winners agreeing, conv2 stage", float((got[0][1] == got[1][1]).mean()))
    assert (got[0][0] == got[1][0]).mean() >= 0.99      # conv1-stage winners (hierarchical 4x4)
    assert (got[0][1] == got[1][1]).mean() >= 0.99      # conv2-stage winners
    assert maxnorm_err(got[1][2], got[0][2]) <= 1e-2     # pooled conv1 activations


def test_tensor_path_gradients_ragged_and_accumulating(net):
    # 70 samples (not a multiple of the 64-wide K block of the dW GEMMs) == 64 + 6 accumulated
    import torch
    x, t = synth.uniform_crops(70, 81), np.concatenate([synth.heatmap_labels(35, 82)] * 2)
    st = torch.cuda.current_stream().cuda_stream

    def grads(a, b):
        xd, td = torch.from_numpy(x[a:b].copy()).cuda(), torch.from_numpy(t[a:b].copy()).cuda()
        net.grad_batch_device(xd.data_ptr(), td.data_ptr(), b - a, None, precision=hp.PRECISION_TENSOR, stream=st)
        torch.cuda.synchronize()
        return net.get_grads()

    g_all, g_a, g_b = grads(0, 70), grads(0, 64), grads(64, 70)
    for k, (off, cnt) in LAYOUT.items():
        assert maxnorm_err(g_all[off:off + cnt], (g_a + g_b)[off:off + cnt]) <= 1e-5, k


def test_loss_curves_over_1k_steps(orc, p0):
    # BASELINE.json north_star: "training-loss curves matching over 1k steps".  1,000 batch-1 steps exactly as
    # train-cnn.cpp:160 issues them (alpha = 0.001), cycling 16 samples, against the CPU reference arithmetic.
    xs, ts = synth.depthlike_crops(16, 91), synth.heatmap_labels(16, 92)
    order = np.arange(1000) % 16
    p = p0.copy()
    want = orc.train_seq(p, xs[order], ts[order], 0.001)
    for prec, curve_tol, weight_tol in ((hp.PRECISION_FP32, 2e-4, 1e-4), (hp.PRECISION_TENSOR, 1e-3, None)):
        net = hp.PoseInitializerCNN("", precision=prec)
        got = np.array([net.Train(xs[i], ts[i], 0.001) for i in order], np.float32)
        rel = np.abs(got - want) / want
        record("1k-step loss curve, precision %d, worst relative deviation" % prec, float(rel.max()))
        assert rel.max() <= curve_tol, (prec, float(rel.max()), int(rel.argmax()))
        assert got[-16:].mean() < got[:16].mean()                   # the loss goes down (slowly at the reference's alpha)
        if weight_tol is not None:
            got_p = net.get_params()
            for k, (off, cnt) in LAYOUT.items():
                assert maxnorm_err(got_p[off:off + cnt], p[off:off + cnt]) <= weight_tol, k
        del net


# ---- output decode and 16-bit depth upload ----------------------------
def test_decode_is_bit_exact(net, orc):
    rng = np.random.default_rng(0)
    y = rng.random((40, 2304), dtype=np.float32)
    y[3, :256] = 0
    y[4, 256:512] = 0.5
    y[5, 2048:2064] = 0
    y[6, 512 + 255] = 9.0
    y[7, 2048 + 16 * 3 + 15] = 9.0
    assert np.array_equal(net.decode_batch(y), orc.decode(y))
    for name in ("init", "peaky"):
        assert np.array_equal(net.decode_batch(golden("eval_%s.npy" % name)), golden("decode_%s.npy" % name))
    assert net.decode_batch(np.zeros((0, 2304), np.float32)).shape == (0, 48)


def test_depth_upload_path(net, orc, p0):
    import torch
    g = golden("depth_norm.npz")
    d = np.concatenate([g["depth"], np.random.default_rng(5).integers(0, 900, (9, 4096)).astype(np.uint16)])
    n = d.shape[0]
    # device normalisation is bit-exact (handtrack.h:700)
    dd = torch.from_numpy(d.astype(np.int32)).to(torch.uint16).cuda() if hasattr(torch, "uint16") else None
    if dd is not None:
        xd = torch.empty((n, 4096), device="cuda")
        from hand_tracking_samples_b200 import capi
        capi.check(net.L.hp_normalize_depth_device(net.h, dd.data_ptr(), n, 0.001, 0.1, 0.7, xd.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(xd.cpu().numpy(), orc.normalize_depth(d))
        assert np.array_equal(xd.cpu().numpy()[:2], g["x"])
    # u16 upload + Eval + decode == fp32 upload of the normalised crops (same kernels downstream)
    x = orc.normalize_depth(d)
    y_ref = net.eval_batch(x)
    y, dec = net.eval_depth_batch(d)
    assert np.array_equal(y, y_ref)
    assert np.array_equal(dec, orc.decode(y_ref))
    assert maxnorm_err(y, orc.eval(p0, x)) <= FP32_TOL
    dec_only = net.eval_decode_batch(x, want_y=False)
    assert np.array_equal(dec_only, dec)
    ytc, dectc = net.eval_depth_batch(d, precision=hp.PRECISION_TENSOR)
    assert maxnorm_err(ytc, y_ref) <= TC_TOL
    # tensor path: the normalisation runs inside the conv kernel's loader; same values in, so the result must equal the
    # fp32-crop entry point bit for bit (the loader converts the identical fp32 value to fp16 either way)
    assert np.array_equal(ytc, net.eval_batch(x, precision=hp.PRECISION_TENSOR))
    if dd is not None:
        yd = torch.empty((n, 2304), device="cuda")
        decd = torch.empty((n, 48), device="cuda")
        for prec in (hp.PRECISION_TENSOR, hp.PRECISION_FP32):
            net.eval_depth_batch_device(dd.data_ptr(), n, yd.data_ptr(), decd.data_ptr(), precision=prec, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(yd.cpu().numpy(), ytc if prec == hp.PRECISION_TENSOR else y_ref)
            assert np.array_equal(decd.cpu().numpy(), orc.decode(yd.cpu().numpy()))


def test_decode_fused_into_fc2_epilogue_is_bit_exact(net, orc, p0):
    # tensor path: CNNOutputAnalysis' numeric core runs in the fc2 + softmax kernel's epilogue (hp_tc.cu, SOFTMAX_DECODE);
    # it must equal the stand-alone decode of the same y bit for bit, with and without y being written
    import torch
    x = np.concatenate([synth.depthlike_crops(200, 141), synth.uniform_crops(100, 142),
                        np.stack([np.full(4096, 1e3, np.float32), np.zeros(4096, np.float32)])])   # incl. a NaN row and a flat one
    for params in (p0, peaky(p0)):
        net.set_params(params)
        y, dec = net.eval_decode_batch(x, precision=hp.PRECISION_TENSOR)
        assert np.array_equal(y, net.eval_batch(x, precision=hp.PRECISION_TENSOR), equal_nan=True)
        want = orc.decode(y)
        assert np.array_equal(dec, want, equal_nan=True)
        dec_only = net.eval_decode_batch(x, precision=hp.PRECISION_TENSOR, want_y=False)
        assert np.array_equal(dec_only, want, equal_nan=True)
    # device entry point, 16-bit depth in, decoded only (no y anywhere in HBM)
    d = np.random.default_rng(7).integers(0, 900, (257, 4096)).astype(np.uint16)
    dd = torch.from_numpy(d.view(np.int16)).cuda()
    decd = torch.empty((257, 48), device="cuda")
    net.eval_depth_batch_device(dd.data_ptr(), 257, None, decd.data_ptr(), precision=hp.PRECISION_TENSOR, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    y16, _ = net.eval_depth_batch(d, precision=hp.PRECISION_TENSOR, want_decoded=False)
    assert np.array_equal(decd.cpu().numpy(), orc.decode(y16), equal_nan=True)


def test_label_rendering_is_bit_exact_and_trains(net, orc, p0):
    g = golden("labels_render.npz")
    want = g["t_u8"].astype(np.float32) / np.float32(255.0)
    assert np.array_equal(net.render_labels(g["points"], g["vals"]), want)
    rng = np.random.default_rng(4)
    p = rng.uniform(-1, 17, (3000, 16)).astype(np.float32)
    v = rng.uniform(-0.1, 1.1, (3000, 16)).astype(np.float32)
    assert np.array_equal(net.render_labels(p, v), orc.render_labels(p, v))
    # training from label parameters == training from the rendered labels
    x = synth.depthlike_crops(8, 77)
    twin = hp.PoseInitializerCNN("")
    m1 = net.train_batch_points(x, p[:8], v[:8], 0.001)
    m2 = twin.train_batch(x, orc.render_labels(p[:8], v[:8]), 0.001)
    assert np.array_equal(m1, m2) and np.array_equal(net.get_params(), twin.get_params())
