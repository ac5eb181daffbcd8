"""Quick hardware check of the tensor-core path (run on a B200): parity of Eval and of the training gradients against
the CPU oracle, then the per-stage device times of a 16,384-crop chunk.  HP_CONV_V1=1 selects the round-1 conv kernel
(hp_tc_conv.cu) instead of the transposed / tap-paired one (hp_tc_conv2.cu), for A/B runs.

    python tools/conv_check.py [n_timing_crops]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())


def main():
    import torch
    from hand_tracking_samples_b200 import cnn as hp, synth
    from oracle import oracle as orc
    from oracle.oracle import LAYOUT
    out = {"conv_kernel": "v1" if os.environ.get("HP_CONV_V1") else "v2"}
    o = orc.Oracle()
    p0 = o.init_xavier()
    net = hp.PoseInitializerCNN("")
    x = np.concatenate([synth.depthlike_crops(10, 61), synth.uniform_crops(10, 62)])
    want = o.eval(p0, x)
    got = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    out["eval_init_worst"] = max(err(got[i], want[i]) for i in range(len(x)))
    pk = p0.copy()
    off, cnt = LAYOUT["fc2.W"]
    pk[off:off + cnt] *= 30.0
    net.set_params(pk)
    wantp = o.eval(pk, x)
    gotp = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    out["eval_peaky_worst"] = max(err(gotp[i], wantp[i]) for i in range(len(x)))
    net.set_params(p0)
    # 16-bit depth straight into the conv loader
    d = np.random.default_rng(5).integers(0, 900, (9, 4096)).astype(np.uint16)
    xn = o.normalize_depth(d)
    y16, _ = net.eval_depth_batch(d, precision=hp.PRECISION_TENSOR)
    out["u16_loader_equals_fp32_entry"] = bool(np.array_equal(y16, net.eval_batch(xn, precision=hp.PRECISION_TENSOR)))
    # gradients, 9 samples
    xt = np.concatenate([synth.depthlike_crops(5, 31), synth.uniform_crops(4, 32)])
    tt = synth.heatmap_labels(9, 33)
    gw, _ = o.train_minibatch(p0.copy(), xt, tt, 0.001, apply=False)
    xd, td = torch.from_numpy(xt).cuda(), torch.from_numpy(tt).cuda()
    net.grad_batch_device(xd.data_ptr(), td.data_ptr(), 9, None, precision=hp.PRECISION_TENSOR, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = net.get_grads()
    out["grad9"] = {k: err(g[a:a + c], gw[a:a + c]) for k, (a, c) in LAYOUT.items()}
    print(json.dumps(out), flush=True)
    # timing: one 16,384-crop chunk (the unit bench.py's roofline is quoted on), device-resident
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    xs = torch.rand((n, 4096), device="cuda")
    ys = torch.empty((n, 2304), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        net.eval_batch_device(xs.data_ptr(), n, ys.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
    torch.cuda.synchronize()
    net.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 20
    for _ in range(reps):
        net.eval_batch_device(xs.data_ptr(), n, ys.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = net.profile_read(3)
    net.profile(False)
    tot = e0.elapsed_time(e1) / reps
    out["timing"] = {"crops": n, "ms_per_pass": tot, "crops_per_s": n / tot * 1e3,
                     "stage_us": {k: 1e3 * ms[i] / max(cnt[i], 1) for i, k in enumerate(("conv", "fc1", "fc2"))}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
