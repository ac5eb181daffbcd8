"""1-GPU training step timeline (HP_STEP_TIMING=1) for comparison with the data-parallel timelines."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand_tracking_samples_b200 import cnn as hp, synth, capi
net = hp.PoseInitializerCNN("")
TB = int(os.environ.get("TB", "256"))
tx = torch.rand((TB, 4096), device="cuda"); tt = torch.from_numpy(synth.heatmap_labels(TB, 1)).cuda(); mse = torch.empty(TB, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(10):
    net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-6, mse.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-6, mse.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
e1.record(); torch.cuda.synchronize()
tms = np.zeros(9, np.float32); capi.check(net.L.hp_debug_step_times(net.h, tms.ctypes.data))
print("1 GPU: %.1f us/step; bucket0/1/2 ready %s dx0/1 %s update0/1/2 %s tail %.0f" % (e0.elapsed_time(e1) * 10, (tms[:3]*1e3).round(), (tms[3:5]*1e3).round(), (tms[5:8]*1e3).round(), tms[8]*1e3))
