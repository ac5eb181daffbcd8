"""The C++ side of the drop-in boundary: include/handposedd/cnn.h used exactly the way the
reference's callers use third_party/cnn.h (include/handtrack.h:103-130,701; train-cnn.cpp:115-116,160)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, maxnorm_err
from hand_tracking_samples_b200 import _build

BIN = os.path.join(ROOT, "tests", "_bin")


def test_dropin_sources_compile():
    outs = _build.build_dropin_test()
    assert os.path.exists(outs[0])
    if os.path.isdir("/root/reference"):
        # the reference's unmodified handtrack.h compiled against the drop-in CNN class
        assert os.path.exists(os.path.join(BIN, "ht_dropin"))
        assert os.path.exists(os.path.join(BIN, "dataset_dropin"))


def test_reference_side_load_dataset_binding_matches_reference_loader(tmp_path):
    """INTEGRATION.md's load_dataset stub (on hp_dataset_*) vs the reference's own load_dataset, same binary, same files:
    identical Frames (depth, ir, poses, camera, fid) for the golden dataset and for a ragged one.  Host-only."""
    exe = os.path.join(BIN, "dataset_dropin")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference"):
            pytest.skip("tests/_bin/dataset_dropin needs /root/reference at build time")
        _build.build_dropin_test()
    r = subprocess.run([exe, os.path.join(GOLDEN, "dataset", "crops64"), "17"], capture_output=True, text=True)
    assert r.returncode == 0 and "6 frames" in r.stdout, r.stdout + r.stderr
    # ragged copy: partial last frame, short .ir, pose text cut mid-frame, more poses requested than recorded
    import shutil
    for ext in (".json", ".rs", ".ir", ".pose"):
        shutil.copy(os.path.join(GOLDEN, "dataset", "crops64" + ext), tmp_path / ("r" + ext))
    for ext, keep in ((".rs", 8192 * 4 + 100), (".ir", 4096 * 2 + 7), (".pose", 700)):
        data = open(tmp_path / ("r" + ext), "rb").read()
        open(tmp_path / ("r" + ext), "wb").write(data[:keep])
    r = subprocess.run([exe, str(tmp_path / "r"), "19"], capture_output=True, text=True)
    assert r.returncode == 0 and "4 frames" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_dropin_cpp_matches_golden(tmp_path):
    exe = os.path.join(BIN, "dropin_main")
    if not os.path.exists(exe):
        _build.build_dropin_test()
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))
    labels = np.load(os.path.join(GOLDEN, "labels.npy"))
    n = crops.shape[0]
    c, l = tmp_path / "c.f32", tmp_path / "l.f32"
    crops.tofile(c)
    labels.tofile(l)
    ev, ms, wb = tmp_path / "e.f32", tmp_path / "m.f32", tmp_path / "w.cnnb"
    r = subprocess.run([exe, str(c), str(l), str(n), str(ev), str(ms), str(wb)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(ev, np.float32).reshape(n, 2304)
    assert maxnorm_err(got, np.load(os.path.join(GOLDEN, "eval_init.npy"))) <= 1e-5
    gold = np.load(os.path.join(GOLDEN, "train24.npz"))
    assert np.allclose(np.fromfile(ms, np.float32), gold["mse"][:n], rtol=2e-5)   # first 6 of the 24 reference steps
    assert os.path.getsize(wb) == 37833600
    # tensor-core precision through the same class
    r = subprocess.run([exe, str(c), str(l), str(n), str(ev), str(ms), str(wb), "tensor"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_reference_handtrack_h_runs_on_the_dropin(tmp_path):
    exe = os.path.join(BIN, "ht_dropin")
    if not os.path.exists(exe):
        pytest.skip("tests/_bin/ht_dropin was not prebuilt (needs /root/reference at build time)")
    crops = np.load(os.path.join(GOLDEN, "crops.npy"))[:3]
    c, ev = tmp_path / "c.f32", tmp_path / "e.f32"
    crops.tofile(c)
    # handtrack.h loads ../assets/model_hand.json lazily only when a HandTracker is built; the CNN factory needs no assets
    r = subprocess.run([exe, str(c), "3", str(ev)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(ev, np.float32).reshape(3, 2304)
    assert maxnorm_err(got, np.load(os.path.join(GOLDEN, "eval_init.npy"))[:3]) <= 1e-5
