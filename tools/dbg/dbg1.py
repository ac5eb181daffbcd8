import numpy as np, sys
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth
from oracle.oracle import Oracle, LAYOUT
import ctypes as C, torch
o=Oracle(); p0=o.init_xavier()
net=hp.PoseInitializerCNN("")
x=synth.uniform_crops(1,1234); t=synth.heatmap_labels(1,4321)
gw,_=o.grad_sample(p0,x[0],t[0])
xd=torch.from_numpy(x).cuda(); td=torch.from_numpy(t).cuda()
net.grad_batch_device(xd.data_ptr(),td.data_ptr(),1,None,stream=torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
g=net.get_grads()
for k,(off,n) in LAYOUT.items():
    print(k, np.abs(g[off:off+n]-gw[off:off+n]).max()/np.abs(gw[off:off+n]).max())
e0=o.peek(100).reshape(16,15,4,15,4).transpose(0,1,3,2,4).reshape(16,15,15,16)
nz=(e0!=0).sum(-1); print("nonzeros per window hist", np.bincount(nz.ravel()))
pos_want=np.abs(e0).argmax(-1)
buf=np.empty((1,3600),np.uint8); 
from hand_tracking_samples_b200 import capi
capi.check(net.L.hp_peek(net.h,203,1,buf.ctypes.data))
idx=buf.reshape(16,15,15)
g1=net.peek(103,1,3600)[0].reshape(16,15,15)
mism=(idx!=pos_want)&(nz>0)
print("mismatched winners", mism.sum(), "of", (nz>0).sum())
w=np.argwhere(mism)[:10]
a1=o.peek(1).reshape(16,60,60)
for c,py,px in w:
    win=a1[c,4*py:4*py+4,4*px:4*px+4]
    print(c,py,px,"mine",idx[c,py,px],"want",pos_want[c,py,px], "g1",g1[c,py,px], "e0sum", e0[c,py,px].sum()); print(win)
e4=o.peek(104).reshape(64,6,2,6,2).transpose(0,1,3,2,4).reshape(64,6,6,4)
nz=(e4!=0).sum(-1); print("stage2 nonzeros per window hist", np.bincount(nz.ravel()))
pos_want=np.abs(e4).argmax(-1)
buf=np.empty((1,2304),np.uint8)
capi.check(net.L.hp_peek(net.h,206,1,buf.ctypes.data))
idx=buf.reshape(64,6,6)
mism=(idx!=pos_want)&(nz>0)
print("stage2 mismatched winners", mism.sum(), "of", (nz>0).sum())
g2=net.peek(106,1,2304)[0].reshape(64,6,6)
print("g2 err", np.abs(g2-e4.sum(-1)).max()/np.abs(e4).max())
a5=o.peek(5).reshape(64,12,12)
for c,py,px in np.argwhere(mism)[:6]:
    print(c,py,px,"mine",idx[c,py,px],"want",pos_want[c,py,px]); print(a5[c,2*py:2*py+2,2*px:2*px+2])
