import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
from hand_tracking_samples_b200 import cnn as hp, synth
prec = hp.PRECISION_TENSOR if (len(sys.argv)<2 or sys.argv[1]=="tensor") else hp.PRECISION_FP32
net=hp.PoseInitializerCNN("")
TB=256
tx=torch.rand((TB,4096),device='cuda'); tt=torch.from_numpy(synth.heatmap_labels(TB,1)).cuda(); mse=torch.empty(TB,device='cuda')
st=torch.cuda.current_stream().cuda_stream
for _ in range(6):
    net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-5, mse.data_ptr(), precision=prec, stream=st)
torch.cuda.synchronize()
