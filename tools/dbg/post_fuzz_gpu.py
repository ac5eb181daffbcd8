"""Device kernels either side of the CNN (decode, label rendering, depth normalisation) against the CPU oracle on the
adversarial inputs of tests/test_post_fuzz.py, bit for bit.  Run on a GPU box:  python tools/dbg/post_fuzz_gpu.py
Round-2 item: promote into tests/ as a @pytest.mark.gpu test once it has passed on hardware (its first run, before the
NaN fix in csrc/hp_post.cu, failed on seed 3: PeakVolume of a heatmap whose pixel (0,0) is NaN)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from test_post_fuzz import depth_case, heatmaps, label_params, same_bits  # noqa: E402


def main():
    import torch
    bad = 0
    from hand_tracking_samples_b200 import capi, cnn as hp
    net = hp.PoseInitializerCNN("")
    o = orc.Oracle()
    st = torch.cuda.current_stream().cuda_stream
    for seed in range(48):
        y = heatmaps(seed)
        ok = same_bits(net.decode_batch(y), o.decode(y)); bad += report(ok, "decode", seed)
        pts, vals = label_params(seed)
        ok = same_bits(net.render_labels(pts, vals), o.render_labels(pts, vals)); bad += report(ok, "labels", seed)
    for seed in range(16):
        d, sc, dmin, dmax = depth_case(seed)
        dd = torch.from_numpy(d.view(np.int16)).cuda()
        x = torch.empty((3, 4096), device="cuda")
        capi.check(net.L.hp_normalize_depth_device(net.h, dd.data_ptr(), 3, sc, dmin, dmax, x.data_ptr(), st))
        torch.cuda.synchronize()
        ok = same_bits(x.cpu().numpy(), o.normalize_depth(d, sc, dmin, dmax)); bad += report(ok, "normalize", seed)
    print("mismatches:", bad)
    return bad


def report(ok, what, seed):
    if not ok:
        print("MISMATCH", what, "seed", seed, flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
