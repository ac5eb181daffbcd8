"""Why is a multi-chunk Eval in tools/sweep.py ~15 % slower per chunk than in bench.py?  A/B: legacy vs created stream,
stage events on/off, input allocation size."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand_tracking_samples_b200 import cnn as hp
net = hp.PoseInitializerCNN("")
N = 65536
def run(tag, x, y, stream_obj, profile, n=N, reps=20):
    with torch.cuda.stream(stream_obj) if stream_obj is not None else torch.cuda.stream(torch.cuda.default_stream()):
        st = torch.cuda.current_stream().cuda_stream
        net.profile(profile)
        for _ in range(3):
            net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
        e1.record(); torch.cuda.synchronize()
        net.profile(False)
        ms = e0.elapsed_time(e1) / reps
        print("%-46s stream %#x  %.3f ms  %.2f M crops/s" % (tag, st, ms, n / ms / 1e3), flush=True)
x = torch.rand((N, 4096), device="cuda"); y = torch.empty((N, 2304), device="cuda")
side = torch.cuda.Stream()
run("1 GiB input, legacy stream", x, y, None, False)
run("1 GiB input, created stream", x, y, side, False)
run("1 GiB input, created stream, stage events", x, y, side, True)
run("1 GiB input, legacy stream, stage events", x, y, None, True)
run("1 GiB input, legacy stream, 16384 crops", x, y, None, False, n=16384, reps=80)
run("1 GiB input, legacy stream, 32768 crops", x, y, None, False, n=32768, reps=40)
del x, y
xb = torch.rand((1 << 20, 4096), device="cuda"); yb = torch.empty((1 << 20, 2304), device="cuda")
run("16 GiB allocation, legacy stream", xb, yb, None, False)
run("16 GiB allocation, created stream", xb, yb, side, False)
run("16 GiB allocation, created, stage events", xb, yb, side, True)
