"""Generate tests/golden/*.np[yz] by running the UNMODIFIED reference (oracle/_ref).

Run once in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py
The outputs are committed; the GPU box only ever reads the committed files.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import LAYOUT, Ref  # noqa: E402
from hand_tracking_samples_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STRIDE = 4099  # prime stride for sampling the two 4.7M-element FC weight tensors


def sample(params):
    d = {}
    for k, (off, n) in LAYOUT.items():
        v = params[off:off + n]
        d[k] = v[::STRIDE].copy() if n > 100000 else v.copy()
    return d


def main():
    r = Ref()
    r.init()
    p0 = r.save()
    meta = {"init_sha256": hashlib.sha256(p0.tobytes()).hexdigest(),
            "known": {"conv1.W[0]": float(p0[0]), "conv2.W[0]": float(p0[416]),
                      "fc1.W[0]": float(p0[16864]), "fc2.W[0]": float(p0[4737504])},
            "stride": STRIDE}

    crops = np.concatenate([synth.uniform_crops(2, 1234), synth.depthlike_crops(2, 1234),
                            np.zeros((1, 4096), np.float32), np.ones((1, 4096), np.float32)])
    labels = synth.heatmap_labels(crops.shape[0], 4321)
    np.save(os.path.join(OUT, "crops.npy"), crops)
    np.save(os.path.join(OUT, "labels.npy"), labels)

    # Eval with Init() weights
    np.save(os.path.join(OUT, "eval_init.npy"), r.eval(crops))
    tr = r.forward_trace(crops[0])
    np.savez_compressed(os.path.join(OUT, "trace_crop0.npz"), pool1=tr[3], pool2=tr[6], fc1=tr[8], logits=tr[9])

    # Eval with fc2.W x30 ("peaky" softmax, SURVEY.md section 4 fixture ii)
    pk = p0.copy()
    off, n = LAYOUT["fc2.W"]
    pk[off:off + n] *= 30.0
    r.load(pk)
    np.save(os.path.join(OUT, "eval_peaky.npy"), r.eval(crops))

    # per-sample gradients at Init() weights
    r.load(p0)
    gs = {}
    mses = []
    for i in range(crops.shape[0]):
        g, m = r.grad_sample(crops[i], labels[i])
        mses.append(m)
        for k, v in sample(g).items():
            gs["%d/%s" % (i, k)] = v
    gs["mse"] = np.array(mses, np.float32)
    np.savez_compressed(os.path.join(OUT, "grads_init.npz"), **gs)

    # 24 sequential reference Train steps (cycling the 6 crops), alpha as train-cnn.cpp:160
    r.load(p0)
    xs = np.concatenate([crops] * 4)
    ts = np.concatenate([labels] * 4)
    mse = r.train_seq(xs, ts, 0.001)
    p1 = r.save()
    meta["train24_sha256"] = hashlib.sha256(p1.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(OUT, "train24.npz"), mse=mse, **sample(p1))
    np.save(os.path.join(OUT, "eval_train24.npy"), r.eval(crops))

    json.dump(meta, open(os.path.join(OUT, "meta.json"), "w"), indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()


def main_post():
    """Fixtures for the decode / depth-normalisation rows, from oracle/_ref/libpostref.so."""
    from oracle.oracle import PostRef
    r = PostRef()
    np.save(os.path.join(OUT, "decode_init.npy"), r.decode(np.load(os.path.join(OUT, "eval_init.npy"))))
    np.save(os.path.join(OUT, "decode_peaky.npy"), r.decode(np.load(os.path.join(OUT, "eval_peaky.npy"))))
    rng = np.random.default_rng(77)
    d = rng.integers(0, 1200, (2, 4096)).astype(np.uint16)
    d[0, :64] = 0
    d[1, :64] = 65535
    np.savez_compressed(os.path.join(OUT, "depth_norm.npz"), depth=d, x=r.normalize_depth(d))


if __name__ == "__main__" and "--post" in sys.argv:
    main_post()


def main_labels():
    """Fixture for the label-rendering row, from oracle/_ref/libpostref.so; labels are k/255 so they are
    stored as the u8 numerators."""
    from oracle.oracle import PostRef
    r = PostRef()
    rng = np.random.default_rng(11)
    pts = rng.uniform(-1, 17, (64, 16)).astype(np.float32)
    vals = rng.uniform(-0.1, 1.1, (64, 16)).astype(np.float32)
    pts[0] = 0
    pts[1] = 15.999
    vals[0] = 0
    vals[1] = 1.0
    pts[2] = 7.5
    t = r.render_labels(pts, vals)
    np.savez_compressed(os.path.join(OUT, "labels_render.npz"), points=pts, vals=vals, t_u8=np.round(t * 255).astype(np.uint8))


if __name__ == "__main__" and "--labels" in sys.argv:
    main_labels()


def resample_cases(n=24, seed=314):
    """Frames + destination cameras for the SampleD row: HandSegmentVR-like cameras (64x64, focal
    avgdepth*64/diam, principal 32, rotation about the view axis and towards the blob), plus adversarial ones
    (looking away from the frame, NaN orientation, a shifted position)."""
    from hand_tracking_samples_b200 import synth
    rng = np.random.default_rng(seed)
    frames = synth.depth_frames(n, 240, 320, seed)[0]
    frames[3] = rng.integers(0, 65536, (240, 320)).astype(np.uint16)
    cams = np.zeros((n, 11), np.float32)
    for i in range(n):
        f = rng.uniform(0.2, 1.0) * 64.0 / 0.17
        ang = rng.uniform(-np.pi, np.pi)
        axis = rng.normal(0, 1, 3) * np.array([0.15, 0.15, 1.0])
        axis /= np.linalg.norm(axis)
        q = np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]])
        cams[i] = [f, f, 32, 32, 0, 0, 0, *q]
    cams[5, 4:7] = [0.02, -0.01, 0.05]
    cams[6, 7:] = [1, 0, 0, 0]           # looks backwards: negative z, negative distances
    cams[7, 9] = np.nan
    cams[8, 0:2] = 1e-3                  # tiny focal: everything projects far outside
    return frames, (241.811768, 241.811768, 162.830505, 118.740089), cams


def main_resample():
    """Fixture for the SampleD row (second half), from oracle/_ref/libpostref.so."""
    from oracle.oracle import PostRef
    r = PostRef()
    frames, intr, cams = resample_cases()
    out = np.stack([r.sample_d(frames[i], intr, cams[i], 4000) for i in range(len(cams))])
    np.savez_compressed(os.path.join(OUT, "sample_d.npz"), crops=out, frames_seed=314, cams=cams)


if __name__ == "__main__" and "--resample" in sys.argv:
    main_resample()
