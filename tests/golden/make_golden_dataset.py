"""Generate tests/golden/dataset/* with the UNMODIFIED reference writer and reader (oracle/_ref/libdatasetref.so =
include/dataset.h DepthDataStreamOut + load_dataset behind oracle/ref_dataset_shim.cpp).

Run once in the build container:   python tests/golden/make_golden_dataset.py
Outputs (committed; the GPU box only reads them):
  crops64.{json,rs,ir,pose}   6 frames of 64x64 depth (the "compressed" training form, train-cnn.cpp:31-34), 17 poses/frame
  crops64_expected.npz        what the reference's load_dataset returns for it (info, depth, ir, poses)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import DatasetRef  # noqa: E402
from hand_tracking_samples_b200.synth import depth_frames as synth_frames  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataset")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = DatasetRef()
    depth, ir, poses = synth_frames(6, 64, 64, 20261018)
    base = os.path.join(OUT, "crops64")
    cwd = os.getcwd()
    os.chdir(OUT)   # DatasetInfo.fname is written into the .json: keep it relative
    ref.save("crops64", (60.0, 60.0, 32.0, 32.0, 0.001), depth, ir, poses)
    os.chdir(cwd)
    info, d, r, p = ref.load(base, 17)
    assert np.array_equal(d, depth) and np.array_equal(r, ir)
    np.savez_compressed(os.path.join(OUT, "crops64_expected.npz"), info=info, depth=d, ir=r, poses=p)
    print("wrote", OUT, "frames", d.shape, "pose text round trip max err", float(np.abs(p - poses).max()))


if __name__ == "__main__":
    main()
