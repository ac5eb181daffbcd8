import sys, numpy as np
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth
net=hp.PoseInitializerCNN("")
x=synth.depthlike_crops(5,1); t=synth.heatmap_labels(5,2)
for prec in (hp.PRECISION_FP32, hp.PRECISION_TENSOR):
    y=net.eval_batch(x, precision=prec); print(prec, y.shape, float(y.sum()))
    m=net.train_batch(x, t, 0.001, precision=prec); print(prec, m)
x=synth.uniform_crops(300,3)
print(net.eval_batch(x, precision=hp.PRECISION_TENSOR).sum())
