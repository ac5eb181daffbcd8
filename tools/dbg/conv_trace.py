import numpy as np, sys, ctypes as C
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth, capi
import torch
net=hp.PoseInitializerCNN("")
n=148*24
x=torch.rand((n,4096),device='cuda'); y=torch.empty((n,2304),device='cuda')
for _ in range(2):
    net.eval_batch_device(x.data_ptr(), n, y.data_ptr(), precision=hp.PRECISION_TENSOR); torch.cuda.synchronize()
L=capi.lib()
buf=np.zeros(9*24*16,np.int64)
L.hp_debug_conv_trace.argtypes=[C.c_void_p,C.c_int]
print("rc",L.hp_debug_conv_trace(buf.ctypes.data, buf.size))
t=buf.reshape(9,24,16)
t0=t[t>0].min()
names=["MMA1","MMA2","EPI1a","EPI2","LOAD","EPI1b0","EPI1b1","EPI1b2","EPI1b3"]
for it in range(6,10):
    print("crop it=%d"%it)
    for r in range(9):
        row=t[r,it]; print("   %-6s"%names[r], " ".join("%7d"%(v-t0) if v>0 else "      -" for v in row[:11]))
print("per-crop period (MMA1 ev1):", np.diff(t[0,4:20,1]))
