"""the dataset reader (csrc/hp_dataset.cu, hp_dataset_*) against the reference's own
load_dataset (include/dataset.h:109-163).  CPU tests: the committed golden dataset (written AND read back by the
unmodified reference, tests/golden/make_golden_dataset.py) and, where oracle/_ref exists, datasets written on the fly
by the reference writer including its edge cases (ragged tails, missing companions, interleaved IR).  The GPU test
feeds the golden 64x64-crop dataset through hp_dataset_eval_depth."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from hand_tracking_samples_b200 import capi
from hand_tracking_samples_b200.dataset import Dataset
from oracle import oracle as orc

GOLD = os.path.join(ROOT, "tests", "golden", "dataset")
need_ref = pytest.mark.skipif(not orc.have_datasetref(), reason="oracle/_ref/libdatasetref.so not built (needs /root/reference)")


def info_vector(ds):
    i = ds.info
    return np.array([i.width, i.height, i.focal[0], i.focal[1], i.principal[0], i.principal[1], i.depth_scale, *i.mplane, i.hasir,
                     i.rgb_dim[0], i.rgb_dim[1], i.feye_dim[0], i.feye_dim[1], i.segment_scale], np.float32)


def same_as_reference(base, np_):
    want = orc.DatasetRef().load(base, np_)
    assert want is not None
    info, depth, ir, poses = want
    ds = Dataset(base, np_)
    assert len(ds) == depth.shape[0]
    assert np.array_equal(info_vector(ds), info)
    d, r, p = ds.read()
    assert np.array_equal(d, depth) and np.array_equal(r, ir)
    assert np.array_equal(p.view(np.uint32), poses.view(np.uint32))     # bit-exact, including -0.0
    return ds


def test_golden_dataset_matches_reference_reader():
    want = np.load(os.path.join(GOLD, "crops64_expected.npz"))
    ds = Dataset(os.path.join(GOLD, "crops64"), 17)
    assert len(ds) == 6 and (ds.width, ds.height) == (64, 64)
    assert np.array_equal(info_vector(ds), want["info"])
    assert ds.info.camtype == b"synthetic" and ds.info.has_ir_file and ds.info.has_pose_file
    d, r, p = ds.read()
    assert np.array_equal(d, want["depth"]) and np.array_equal(r, want["ir"])
    assert np.array_equal(p.view(np.uint32), want["poses"].view(np.uint32))
    # ragged requests
    d2, r2, p2 = ds.read(2, 3)
    assert np.array_equal(d2, want["depth"][2:5]) and np.array_equal(r2, want["ir"][2:5]) and np.array_equal(p2, want["poses"][2:5])
    d0, _, _ = ds.read(6, 0)
    assert d0.shape[0] == 0
    with pytest.raises(capi.HpError):
        ds.read(4, 3)


def test_reference_example_header_fields():
    # the one dataset header the reference ships (datasets/example/hand_data_example.json), restated here
    ds_json = {"camtype": "ivycam", "dcamera": {"depth_scale": 0.000124987, "dims": [320, 240], "focal": [238.434, 238.433],
                                                 "principal": [157.717, 123.03]},
               "feyedim": [640, 480], "fname": "../hand_capture/hand_data_example", "hasir": False, "mplane": [0, 0, 0, 3.40282e+38],
               "rgb_dim": [640, 480], "segment_scale": 0.170}
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "ex")
        json.dump(ds_json, open(base + ".json", "w"))
        np.arange(320 * 240 * 2 + 7, dtype=np.uint16).tofile(base + ".rs")     # two frames and a ragged tail
        ds = Dataset(base, 17)
        i = ds.info
        assert (i.width, i.height, len(ds)) == (320, 240, 2) and not i.hasir and not i.has_ir_file and not i.has_pose_file
        assert np.float32(i.depth_scale) == np.float32(0.000124987) and np.float32(i.mplane[3]) == np.float32(3.40282e+38)
        assert tuple(i.rgb_dim) == (640, 480) and tuple(i.feye_dim) == (640, 480) and i.camtype == b"ivycam"
        d, r, p = ds.read()
        assert np.array_equal(d.ravel(), np.arange(320 * 240 * 2, dtype=np.uint16)) and not r.any()
        assert np.array_equal(p[..., :6], np.zeros_like(p[..., :6])) and np.all(p[..., 6] == 1)   # default Pose
        if orc.have_datasetref():
            same_as_reference(base, 17)


def test_missing_files_are_errors(tmp_path):
    base = str(tmp_path / "nothing")
    with pytest.raises(capi.HpError):
        Dataset(base, 17)                                 # no .rs  (dataset.h:114: runtime_error)
    np.zeros(16, np.uint16).tofile(base + ".rs")
    with pytest.raises(capi.HpError):
        Dataset(base, 17)                                 # no .json (dataset.h:117: throw)
    if orc.have_datasetref():
        assert orc.DatasetRef().load(base, 17) is None


@need_ref
@pytest.mark.parametrize("shape,np_", [((5, 10, 24), 3), ((3, 64, 64), 17), ((1, 7, 5), 1)])
def test_written_by_reference_read_by_both(tmp_path, shape, np_):
    from hand_tracking_samples_b200.synth import depth_frames as synth_frames
    depth, ir, poses = synth_frames(shape[0], shape[1], shape[2], 7 + shape[0], np_)
    base = str(tmp_path / "ds")
    orc.DatasetRef().save(base, (61.5, 60.25, shape[2] / 2, shape[1] / 2, 0.00025), depth, ir, poses, segment_scale=0.21)
    ds = same_as_reference(base, np_)
    assert np.array_equal(ds.read()[0], depth)


@need_ref
def test_ragged_and_missing_companions(tmp_path):
    from hand_tracking_samples_b200.synth import depth_frames as synth_frames
    depth, ir, poses = synth_frames(4, 6, 8, 99, 2)
    ref = orc.DatasetRef()
    base = str(tmp_path / "ds")
    ref.save(base, (10, 10, 4, 3, 0.001), depth, ir, poses)

    def truncate(ext, nbytes):
        data = open(base + ext, "rb").read()
        open(base + ext, "wb").write(data[:nbytes])

    truncate(".rs", 4 * 6 * 8 * 2 - 5)          # partial last depth frame: dropped (dataset.h:133)
    assert len(same_as_reference(base, 2)) == 3
    truncate(".ir", 6 * 8 + 11)                 # second IR frame partial, later ones missing: prefix + zeros
    same_as_reference(base, 2)
    text = open(base + ".pose").read()
    cut = text.index("\n") + 1 + len(text.split("\n")[1]) // 2
    open(base + ".pose", "w").write(text[:cut])          # pose text ends inside frame 1
    same_as_reference(base, 2)
    open(base + ".pose", "w").write(text[:cut] + " oops 1 2 3")   # a non-numeric token: failbit, the rest stays default
    same_as_reference(base, 2)
    same_as_reference(base, 5)                  # caller asks for more poses per frame than were recorded
    # end of file exactly where an orientation.w is due: the stream's sentry fails BEFORE num_get, so the element keeps its
    # default 1 (a non-numeric token there would store 0) -- found by tests/test_dataset_fuzz.py
    vals = text.split()
    open(base + ".pose", "w").write(" ".join(vals[:2 * 7 + 6]))
    same_as_reference(base, 2)
    open(base + ".pose", "w").write(" ".join(vals[:2 * 7 + 6]) + " zzz")
    same_as_reference(base, 2)
    os.remove(base + ".ir")
    os.remove(base + ".pose")
    ds = same_as_reference(base, 2)
    assert not ds.info.has_ir_file and not ds.info.has_pose_file


@need_ref
def test_interleaved_ir_and_sparse_header(tmp_path):
    # "hasir": depth and IR interleaved in the .rs file (dataset.h:135), and a header with fields missing
    rng = np.random.default_rng(5)
    w, h, n = 12, 9, 3
    base = str(tmp_path / "il")
    with open(base + ".rs", "wb") as f:
        for _ in range(n):
            f.write(rng.integers(0, 65536, w * h, dtype=np.uint16).tobytes())
            f.write(rng.integers(0, 256, w * h, dtype=np.uint8).tobytes())
        f.write(b"\x01\x02\x03")
    json.dump({"dcamera": {"dims": [w, h], "focal": [50, 51]}, "hasir": True, "camtype": "x"}, open(base + ".json", "w"))
    ds = same_as_reference(base, 0)
    assert len(ds) == n and ds.info.hasir
    # a separate .ir file overrides the interleaved bytes (dataset.h:137-138)
    rng.integers(0, 256, w * h * 2 + 5, dtype=np.uint8).tofile(base + ".ir")
    same_as_reference(base, 0)


@pytest.mark.gpu
def test_dataset_eval_depth_matches_batched_entry_point():
    from hand_tracking_samples_b200 import cnn as hp
    net = hp.PoseInitializerCNN("")
    ds = Dataset(os.path.join(GOLD, "crops64"), 17)
    depth, _, _ = ds.read()
    for prec in (hp.PRECISION_FP32, hp.PRECISION_TENSOR):
        y, dec = ds.eval_depth(net, precision=prec)
        y2, dec2 = net.eval_depth_batch(depth.reshape(len(ds), 4096), depth_scale=float(ds.info.depth_scale), precision=prec, want_y=True)
        assert np.array_equal(y, y2) and np.array_equal(dec, dec2)
    # and against the CPU oracle: handtrack.h:700 normalisation + Eval
    o = orc.Oracle()
    x = o.normalize_depth(depth.reshape(len(ds), 4096), float(ds.info.depth_scale), 0.1, 0.7)
    want = o.eval(o.init_xavier(), x)
    y, _ = ds.eval_depth(net, 1, 4, precision=hp.PRECISION_FP32)
    assert np.abs(y - want[1:5]).max() / np.abs(want).max() <= 1e-5
    with pytest.raises(capi.HpError):
        ds.eval_depth(net, 4, 5)
