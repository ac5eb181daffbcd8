"""Timeline of the fused conv kernel's warp roles (CTA 0, crops 6..9 of its share), from clock64 stamps compiled in with
HP_CONV_TRACE=1:   HP_CONV_TRACE=1 python -c "from hand_tracking_samples_b200 import _build; _build.build(force=True)"
                   python tools/dbg/conv2_trace.py
Events per role:  MMA1  0 loop top, 1 image ready, then per group g: 2+2g accumulator free, 3+2g issued+committed
                  MMA2  0 loop top, 1 pooled planes ready, per half h: 2+2h accumulator free, 3+2h issued+committed
                  EPI1x 0 loop top, 1 first half ready, 2 first half drained, 3 second half ready, 4 second half drained,
                        5 pooled buffer free, 6 pooled planes published
                  EPI2x 0 loop top, 1 accumulator ready, 2 accumulator released, 3 outputs stored
                  CONV  0 loop top, 1 staged crop landed, 2 image buffer free, 3 image published"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from hand_tracking_samples_b200 import capi, cnn as hp  # noqa: E402

net = hp.PoseInitializerCNN("")
n = 148 * 24
x = torch.rand((n, 4096), device="cuda")
y = torch.empty((n, 2304), device="cuda")
for _ in range(2):
    net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=hp.PRECISION_TENSOR)
    torch.cuda.synchronize()
L = capi.lib()
buf = np.zeros(8 * 24 * 16, np.int64)
L.hp_debug_conv2_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", L.hp_debug_conv2_trace(buf.ctypes.data, buf.size))
t = buf.reshape(8, 24, 16)
names = ["MMA1", "MMA2", "EPI1a", "EPI1b", "EPI2h0", "EPI2h1", "CONV", "-"]
for it in range(8, 12):
    t0 = t[0, it, 0]
    print("crop it=%d  (origin = MMA1 loop top)" % it)
    for r in range(7):
        row = t[r, it]
        print("   %-6s" % names[r], " ".join("%7d" % (v - t0) if v > 0 else "      -" for v in row[:10]))
print("per-crop period (MMA1 image ready):", np.diff(t[0, 4:22, 1]))
d = t[:, 6:22, :]
def mean(a):
    return float(np.mean(a))
print("MMA1: wait image %.0f | per group: wait acc %s issue %s" % (
    mean(d[0, :, 1] - d[0, :, 0]), [round(mean(d[0, :, 2 + 2 * g] - (d[0, :, 1 + 2 * g]))) for g in range(4)],
    [round(mean(d[0, :, 3 + 2 * g] - d[0, :, 2 + 2 * g])) for g in range(4)]))
print("MMA2: wait planes %.0f | wait acc h0 %.0f issue %.0f | wait acc h1 %.0f issue %.0f" % (
    mean(d[1, :, 1] - d[1, :, 0]), mean(d[1, :, 2] - d[1, :, 1]), mean(d[1, :, 3] - d[1, :, 2]), mean(d[1, :, 4] - d[1, :, 3]),
    mean(d[1, :, 5] - d[1, :, 4])))
for r in (2, 3):
    print("%s: wait half0 %.0f drain %.0f wait half1 %.0f drain %.0f wait pooled buf %.0f store+publish %.0f" % (
        names[r], mean(d[r, :, 1] - d[r, :, 0]), mean(d[r, :, 2] - d[r, :, 1]), mean(d[r, :, 3] - d[r, :, 2]), mean(d[r, :, 4] - d[r, :, 3]),
        mean(d[r, :, 5] - d[r, :, 4]), mean(d[r, :, 6] - d[r, :, 5])))
for r in (4, 5):
    print("%s: wait acc %.0f drain %.0f pool+store %.0f" % (names[r], mean(d[r, :, 1] - d[r, :, 0]), mean(d[r, :, 2] - d[r, :, 1]),
                                                          mean(d[r, :, 3] - d[r, :, 2])))
print("CONV: wait stage %.0f wait image buf %.0f convert %.0f" % (mean(d[6, :, 1] - d[6, :, 0]), mean(d[6, :, 2] - d[6, :, 1]),
                                                                mean(d[6, :, 3] - d[6, :, 2])))
