"""torchrun helper: the peer-memory exchange kernel alone (no concurrent compute), time per gradient bucket vs CTA count."""
import os, sys, ctypes as C, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hand_tracking_samples_b200 import cnn as hp, dp, capi
rank, world, local = dp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
net = hp.PoseInitializerCNN("", device=local)
dp.init_data_parallel(net, mode="peer")
L = net.L
L.hp_debug_peer_exchange.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
sizes = {0: (9458400 - 4737504) * 4, 1: (4737504 - 16864) * 4, 2: 16864 * 4}
for blocks in (4, 8, 16, 32, 64, 128):
    for b in (0, 2):
        for _ in range(5):
            capi.check(L.hp_debug_peer_exchange(net.h, b, 0.0, blocks, st))
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            capi.check(L.hp_debug_peer_exchange(net.h, b, 0.0, blocks, st))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        t = torch.tensor([us], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wire = sizes[b] * (world - 1) / world
        if rank == 0:
            print("world %d bucket %d blocks %3d: %7.1f us/launch  remote read %6.1f GB/s + remote write %6.1f GB/s per GPU" % (world, b, blocks, t.item(), wire / t.item() / 1e3, wire / t.item() / 1e3), flush=True)
assert net.dp_peer_status() == 0
dp.shutdown_data_parallel(net); dist.destroy_process_group()
