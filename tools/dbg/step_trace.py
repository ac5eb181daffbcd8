"""Device timeline of one training step (batch TB, tensor path) from CUPTI via torch.profiler: per kernel start offset,
duration and the gap to the previous kernel's end, eager (HP_NO_GRAPH=1) or graph-replayed."""
import os, sys, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand_tracking_samples_b200 import cnn as hp, synth
from torch.profiler import profile, ProfilerActivity
net = hp.PoseInitializerCNN("")
TB = int(os.environ.get("TB", "256"))
torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream().cuda_stream
tx = torch.rand((TB, 4096), device="cuda"); tt = torch.from_numpy(synth.heatmap_labels(TB, 1)).cuda(); mse = torch.empty(TB, device="cuda")
def step():
    net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-6, mse.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
for _ in range(20): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6): step()
    torch.cuda.synchronize()
path = "gpurun_out/step_trace_%s.json" % ("eager" if os.environ.get("HP_NO_GRAPH") else "graph")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
starts = [i for i, e in enumerate(ev) if "tc_conv" in e["name"] and "kernel" in e["name"]]
a, b = starts[-2], starts[-1]
t0 = ev[a]["ts"]; prev_end = t0
print("step = %.1f us, %d kernels" % (ev[b]["ts"] - t0, b - a))
for e in ev[a:b]:
    print("%7.1f  dur %6.1f  gap %5.1f  strm %3s  %s" % (e["ts"] - t0, e["dur"], e["ts"] - prev_end, e["args"].get("stream"), e["name"][:70]))
    prev_end = max(prev_end, e["ts"] + e["dur"])
