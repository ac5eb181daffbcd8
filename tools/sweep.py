"""BASELINE.json configs[4]: inference throughput sweep, batch 1 .. 1M synthetic crops on one GPU
(device-resident inputs), tensor-core and FP32 paths, with whole-step roofline fractions.
  python tools/sweep.py > profiles/sweep_r1.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hand_tracking_samples_b200 import cnn as hp

FLOP = 26472960
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
net = hp.PoseInitializerCNN("")
st = torch.cuda.current_stream().cuda_stream
rows = []
nmax = 1 << 20
x = torch.rand((nmax, 4096), device="cuda")
y = torch.empty((nmax, 2304), device="cuda")
for prec, name in ((hp.PRECISION_TENSOR, "tensor"), (hp.PRECISION_FP32, "fp32")):
    for k in range(0, 21):
        n = 1 << k
        if name == "fp32" and n > (1 << 17):
            break
        reps = max(3, min(200, (1 << 22) // max(n * (1 if name == "tensor" else 16), 1)))
        for _ in range(3):
            net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=prec, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cps = n / (ms * 1e-3)
        rows.append({"path": name, "crops": n, "ms": ms, "crops_per_s": cps, "tflops": cps * FLOP / 1e12,
                     "frac_of_sustained_bf16_peak": cps * FLOP / 1e12 / peak})
print(json.dumps({"gpu": torch.cuda.get_device_name(0), "peak_tflops": peak, "rows": rows}, indent=1))
