import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth, capi
from oracle.oracle import Oracle, LAYOUT
net=hp.PoseInitializerCNN("")
st=torch.cuda.current_stream().cuda_stream
def peek_u8(which, n, length):
    buf=np.empty((n,length),np.uint8); capi.check(net.L.hp_peek(net.h,which,n,buf.ctypes.data)); return buf
for name,x in (("depthlike",synth.depthlike_crops(16,5)),("uniform",synth.uniform_crops(16,6))):
    n=x.shape[0]; t=synth.heatmap_labels(n,9)
    xd,td=torch.from_numpy(x).cuda(),torch.from_numpy(t).cuda()
    res={}
    for prec in (hp.PRECISION_FP32, hp.PRECISION_TENSOR):
        net.grad_batch_device(xd.data_ptr(),td.data_ptr(),n,None,precision=prec,stream=st); torch.cuda.synchronize()
        res[prec]=(peek_u8(203,n,3600), peek_u8(206,n,2304), net.peek(3,n,3600), net.get_grads())
    a,b=res[0],res[1]
    print(name,"idx1 agree %.4f idx2 agree %.4f  p1 maxerr %.3e"%((a[0]==b[0]).mean(), (a[1]==b[1]).mean(), np.abs(a[2]-b[2]).max()))
    for k,(off,cnt) in LAYOUT.items():
        print("   %-8s grad err tc vs fp32 %.3e"%(k, np.abs(a[3][off:off+cnt]-b[3][off:off+cnt]).max()/np.abs(a[3][off:off+cnt]).max()))
