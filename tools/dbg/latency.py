import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth
net=hp.PoseInitializerCNN("")
x=synth.depthlike_crops(1,3)
for prec,name in ((hp.PRECISION_FP32,"fp32"),(hp.PRECISION_TENSOR,"tensor")):
    for _ in range(20): net.Eval(x[0]) if prec==0 else net.eval_batch(x,precision=prec)
    t0=time.perf_counter()
    for _ in range(200): net.eval_batch(x,precision=prec)
    dt=(time.perf_counter()-t0)/200
    print(name,"host-call latency of Eval (n=1, pageable host buffers): %.1f us"%(dt*1e6))
    t=synth.heatmap_labels(1,4)
    for _ in range(20): net.train_batch(x,t,1e-6,precision=prec)
    t0=time.perf_counter()
    for _ in range(100): net.train_batch(x,t,1e-6,precision=prec)
    print(name,"host-call latency of Train (n=1): %.1f us"%((time.perf_counter()-t0)/100*1e6))
