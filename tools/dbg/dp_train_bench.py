"""torchrun helper: TC training step time under data parallelism, for HP_DP_RESERVE_SMS experiments."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand_tracking_samples_b200 import cnn as hp, synth, dp
rank, world, local = dp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
net = hp.PoseInitializerCNN("", device=local)
dp.init_data_parallel(net, mode=os.environ.get('HP_DP_MODE', 'peer'))
if os.environ.get('HP_BF16_WIRE'): net.dp_set_bf16_gradients(True)
TB = 256
tx = torch.rand((TB, 4096), device="cuda"); tt = torch.from_numpy(synth.heatmap_labels(TB, 1)).cuda(); mse = torch.empty(TB, device="cuda")
torch.cuda.set_stream(torch.cuda.Stream(device=local))   # the legacy NULL stream cannot be captured into the step graph
st = torch.cuda.current_stream().cuda_stream
for prec, name in ((hp.PRECISION_TENSOR, "tensor"),):
    for _ in range(10):
        net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-6, mse.data_ptr(), precision=prec, stream=st)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 1e-6, mse.data_ptr(), precision=prec, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 100], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    import numpy as np
    from hand_tracking_samples_b200 import capi
    tms = np.zeros(9, np.float32); capi.check(net.L.hp_debug_step_times(net.h, tms.ctypes.data))
    if rank == 0:
        print("step timeline us: bucket0/1/2 ready %s  dx0/1 %s  allreduce0/1/2 done %s  tail %.0f" % ((tms[:3]*1e3).round(), (tms[3:5]*1e3).round(), (tms[5:8]*1e3).round(), tms[8]*1e3), flush=True)
        print("mode", os.environ.get("HP_DP_MODE", "peer"), "world %d reserve %s %s: %.1f us/step  %.0f samples/s" % (world, os.environ.get("HP_DP_RESERVE_SMS", "default"), name, ms.item() * 1e3, world * TB / (ms.item() * 1e-3)), flush=True)
dp.shutdown_data_parallel(net); dist.destroy_process_group()
