"""SURVEY.md 8f-3, second half: SampleD (include/misc_image.h:154-162), the rotated / scaled point resample at the end of
HandSegmentVR (include/handtrack.h:343).  Three layers, all bit-exact: the plain-C oracle against the golden fixture the
reference's own template produced (tests/golden/sample_d.npz, make_golden.py --resample) and against the reference
routine itself where oracle/_ref is built; the device kernel against the oracle; and the fused device chain
frame -> crop -> normalise -> Eval against its separately-called stages."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc

import sys
sys.path.insert(0, GOLDEN)
from make_golden import resample_cases  # noqa: E402

INTR = (241.811768, 241.811768, 162.830505, 118.740089)


def fuzz_case(seed):
    rng = np.random.default_rng(9000 + seed)
    h, w = (240, 320) if seed % 3 else (120, 160)
    frame = rng.integers(0, 65536 if seed % 4 == 0 else 1500, (h, w)).astype(np.uint16)
    f = rng.uniform(20, 500)
    q = rng.normal(0, 1, 4)
    if seed % 2:
        q[3] += 4
    q /= np.linalg.norm(q)
    pos = rng.normal(0, 0.1, 3) if seed % 5 == 0 else np.zeros(3)
    cam = np.array([f, f * rng.uniform(0.9, 1.1), 32 + rng.normal(0, 2), 32, *pos, *q], np.float32)
    if seed % 13 == 0:
        cam[rng.integers(0, 11)] = np.nan
    if seed % 17 == 0:
        cam[0] = 0.0                      # division by zero focal
    intr = (INTR[0] * w / 320, INTR[1] * w / 320, INTR[2] * w / 320, INTR[3] * h / 240)
    return frame, intr, cam, int(rng.integers(0, 65536))


def test_oracle_matches_the_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "sample_d.npz"))
    frames, intr, cams = resample_cases()
    assert np.array_equal(cams, g["cams"], equal_nan=True)
    o = orc.Oracle()
    hit = 0
    for i in range(len(cams)):
        got = o.sample_d(frames[i], intr, cams[i], 4000)
        assert np.array_equal(got, g["crops"][i]), i
        hit += int((got != 4000).sum())
    assert hit > 20000        # the cases really sample the frames, not just the background


@pytest.mark.skipif(not orc.have_postref(), reason="oracle/_ref/libpostref.so not built (needs /root/reference)")
def test_oracle_equals_reference_sample_d_fuzz():
    o, r = orc.Oracle(), orc.PostRef()
    if not hasattr(r.L, "ref_sample_d"):
        pytest.skip("libpostref.so predates ref_sample_d")
    for seed in range(200):
        frame, intr, cam, bg = fuzz_case(seed)
        assert np.array_equal(o.sample_d(frame, intr, cam, bg), r.sample_d(frame, intr, cam, bg)), seed


@pytest.mark.gpu
def test_device_resample_is_bit_exact_and_chains_into_eval():
    import torch
    from hand_tracking_samples_b200 import cnn as hp
    net = hp.PoseInitializerCNN("")
    o = orc.Oracle()
    st = torch.cuda.current_stream().cuda_stream
    g = np.load(os.path.join(GOLDEN, "sample_d.npz"))
    frames, intr, cams = resample_cases()
    n = len(cams)
    fd = torch.from_numpy(frames.view(np.int16)).cuda()
    cd = torch.from_numpy(cams).cuda()
    out = torch.empty((n, 4096), dtype=torch.int16, device="cuda")
    net.resample_depth_device(fd.data_ptr(), 320, 240, intr, cd.data_ptr(), n, out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint16).reshape(n, 64, 64)
    assert np.array_equal(got, g["crops"])
    # explicit crop -> frame index (several crops of one frame) and other frame sizes / backgrounds
    for seed in range(60):
        frame, fintr, cam, bg = fuzz_case(seed)
        h, w = frame.shape
        f1 = torch.from_numpy(np.stack([frame, frame[::-1].copy()]).view(np.int16)).cuda()
        idx = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
        c3 = torch.from_numpy(np.stack([cam, cam, cam])).cuda()
        o3 = torch.empty((3, 4096), dtype=torch.int16, device="cuda")
        net.resample_depth_device(f1.data_ptr(), w, h, fintr, c3.data_ptr(), 3, o3.data_ptr(), frame_of_crop_ptr=idx.data_ptr(), background=bg, stream=st)
        torch.cuda.synchronize()
        r = o3.cpu().numpy().view(np.uint16).reshape(3, 64, 64)
        assert np.array_equal(r[1], o.sample_d(frame, fintr, cam, bg)), seed
        assert np.array_equal(r[0], o.sample_d(frame[::-1].copy(), fintr, cam, bg)) and np.array_equal(r[0], r[2]), seed
    # the fused chain equals its stages: resample -> (normalise in the conv loader) -> Eval -> decode
    for prec in (hp.PRECISION_TENSOR, hp.PRECISION_FP32):
        y = torch.empty((n, 2304), device="cuda")
        dec = torch.empty((n, 48), device="cuda")
        net.eval_frames_device(fd.data_ptr(), 320, 240, intr, cd.data_ptr(), n, y.data_ptr(), dec.data_ptr(), precision=prec, stream=st)
        y2 = torch.empty((n, 2304), device="cuda")
        dec2 = torch.empty((n, 48), device="cuda")
        net.eval_depth_batch_device(out.data_ptr(), n, y2.data_ptr(), dec2.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(y, y2) and torch.equal(dec, dec2)
