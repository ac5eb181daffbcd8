import os, sys, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
for n in (4720896, 4720640, 16864, 9458400):
    t=torch.ones(n,device="cuda")
    for _ in range(10): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): dist.all_reduce(t)
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)/50*1e3
    if rank==0: print("world %d allreduce %d floats (%.1f MB): %.1f us  busbw %.0f GB/s"%(world,n,n*4/1e6,us,2*(world-1)/world*n*4/us/1e3),flush=True)
dist.destroy_process_group()
