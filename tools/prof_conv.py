"""ncu target: nothing but a few 16,384-crop tensor-path Eval passes (conv kernel + the two FC GEMMs per pass).
    ncu --set full -k regex:tc_conv -s 2 -c 1 python tools/prof_conv.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hand_tracking_samples_b200 import cnn as hp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = hp.PoseInitializerCNN("")
x = torch.rand((n, 4096), device="cuda")
y = torch.empty((n, 2304), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(passes):
    net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=hp.PRECISION_TENSOR, stream=st)
torch.cuda.synchronize()
print("ok", float(y.sum().item()))
