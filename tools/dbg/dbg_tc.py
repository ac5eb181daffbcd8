import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth
net=hp.PoseInitializerCNN("")
for n in (1, 5, 300, 4096):
    x=synth.depthlike_crops(min(n,300),71)
    x=np.concatenate([x]*((n+299)//300))[:n]
    y32=net.eval_batch(x)
    ytc=net.eval_batch(x, precision=hp.PRECISION_TENSOR)
    err=np.abs(ytc-y32).max()/np.abs(y32).max()
    print(n, "tc vs fp32 maxnorm err", err, "finite", np.isfinite(ytc).all(), "rowmax", np.abs(ytc-y32).max(1)[:5], flush=True)
x=synth.uniform_crops(64,5)
y32=net.eval_batch(x); ytc=net.eval_batch(x, precision=hp.PRECISION_TENSOR)
print("uniform err", np.abs(ytc-y32).max()/np.abs(y32).max())
