"""Randomised parity of the dataset reader (hp_dataset_*) against the reference's load_dataset compiled in place
(oracle/_ref/libdatasetref.so): random frame shapes, frame counts, ragged .rs/.ir/.pose tails, interleaved IR, pose text
with mixed number formats and an occasional non-numeric token.  CPU only; skipped where the reference is not built."""
import json
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as orc
from test_dataset import same_as_reference

pytestmark = pytest.mark.skipif(not orc.have_datasetref(), reason="oracle/_ref/libdatasetref.so not built (needs /root/reference)")


def fmt(rng, v):
    k = rng.integers(0, 5)
    if k == 0:
        return "%g" % v
    if k == 1:
        return "%.9e" % v
    if k == 2:
        return "%+.4f" % v
    if k == 3:
        return "%d" % int(v * 100)
    return repr(float(np.float32(v)))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2**31 - 1))
def test_random_datasets_read_like_the_reference(tmp_path_factory, seed):
    rng = np.random.default_rng(seed)
    tmp = tmp_path_factory.mktemp("fz")
    w, h, n = int(rng.integers(1, 40)), int(rng.integers(1, 30)), int(rng.integers(0, 6))
    np_file, np_ask = int(rng.integers(0, 5)), int(rng.integers(0, 6))
    hasir = bool(rng.integers(0, 2))
    base = str(tmp / "d")
    hdr = {"dcamera": {"dims": [w, h], "focal": [float(rng.uniform(10, 500)), float(rng.uniform(10, 500))],
                       "principal": [w / 2.0, h / 2.0], "depth_scale": float(rng.choice([0.001, 0.000124987, 0.00025]))},
           "mplane": [0, 0, -1, 3.40282e+38], "fname": "x", "camtype": "fuzz", "hasir": hasir,
           "rgb_dim": [int(rng.integers(0, 3)), int(rng.integers(0, 3))], "feyedim": [0, 0], "segment_scale": float(rng.uniform(0.1, 0.3))}
    for k in list(hdr):                                  # drop a few top-level fields at random: they then read as zero
        if k != "dcamera" and rng.random() < 0.15:
            del hdr[k]
    json.dump(hdr, open(base + ".json", "w"), indent=int(rng.integers(0, 3)) or None)
    with open(base + ".rs", "wb") as f:
        for _ in range(n):
            f.write(rng.integers(0, 65536, w * h, dtype=np.uint16).tobytes())
            if hdr.get("hasir"):
                f.write(rng.integers(0, 256, w * h, dtype=np.uint8).tobytes())
        f.write(bytes(rng.integers(0, 256, int(rng.integers(0, 2 * w * h)), dtype=np.uint8)))     # ragged tail
    if rng.random() < 0.7:
        rng.integers(0, 256, int(rng.integers(0, (n + 1) * w * h + 1)), dtype=np.uint8).tofile(base + ".ir")
    if rng.random() < 0.8:
        vals = rng.normal(0, 1, n * np_file * 7 + int(rng.integers(0, 9)))
        toks = [fmt(rng, v) for v in vals]
        if toks and rng.random() < 0.3:
            toks[int(rng.integers(0, len(toks)))] = str(rng.choice(["oops", "nan", "x1", "--3", "inf"]))
        seps = [" ", "  ", "\n", "\t", " \n "]
        open(base + ".pose", "w").write("".join(t + seps[int(rng.integers(0, len(seps)))] for t in toks))
    ds = same_as_reference(base, np_ask)
    assert len(ds) >= n     # (the ragged tail may by chance hold one more whole frame)
