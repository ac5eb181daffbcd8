"""N-GPU data-parallel check (launched by tests/test_multi_gpu.py under torch.distributed.run):
the NCCL-all-reduced step on sharded data must equal the single-GPU step on the concatenated batch, and every
rank must end with bit-identical weights.  Also checks that sharded inference equals unsharded inference."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_tracking_samples_b200 import cnn as hp, synth, dp
rank, world, local = dp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
net = hp.PoseInitializerCNN("", device=local)
dp.init_data_parallel(net)
n = 64
x, t = synth.depthlike_crops(n, 7), synth.heatmap_labels(n, 8)
lo, hi = dp.shard_range(n, rank, world)
st = torch.cuda.current_stream().cuda_stream
for prec, tol in ((hp.PRECISION_FP32, 2e-5), (hp.PRECISION_TENSOR, 2e-2)):
    net.Init()
    xd, td = torch.from_numpy(x[lo:hi].copy()).cuda(), torch.from_numpy(t[lo:hi].copy()).cuda()
    mse = torch.empty(hi - lo, device="cuda")
    net.train_batch_device(xd.data_ptr(), td.data_ptr(), hi - lo, 0.001, mse.data_ptr(), precision=prec, stream=st)
    torch.cuda.synchronize()
    p_dp = net.get_params()
    # single-GPU reference of the concatenated batch on this rank
    ref = hp.PoseInitializerCNN("", device=local)
    ref.train_batch(x, t, 0.001, precision=prec)
    p_1 = ref.get_params()
    p0 = hp.PoseInitializerCNN("", device=local).get_params()
    rel = np.abs((p_dp - p0) - (p_1 - p0)).max() / np.abs(p_1 - p0).max()
    g = torch.tensor([rel], device="cuda"); dist.all_reduce(g, op=dist.ReduceOp.MAX)
    same = torch.from_numpy(p_dp).cuda(); ref0 = same.clone(); dist.broadcast(ref0, 0)
    identical = bool(torch.equal(same, ref0))
    if rank == 0:
        print("precision", prec, "update rel err DP(%d ranks) vs 1 GPU: %.3e (tol %g)  ranks identical: %s" % (world, g.item(), tol, identical), flush=True)
    assert g.item() <= tol and identical
# opt-in bf16 gradient transport (tensor-precision steps only): same step within the tensor-path bound
net.Init()
net.dp_set_bf16_gradients(True)
xd, td = torch.from_numpy(x[lo:hi].copy()).cuda(), torch.from_numpy(t[lo:hi].copy()).cuda()
net.train_batch_device(xd.data_ptr(), td.data_ptr(), hi - lo, 0.001, None, precision=hp.PRECISION_TENSOR, stream=st)
torch.cuda.synchronize()
p_bf = net.get_params()
ref = hp.PoseInitializerCNN("", device=local)
ref.train_batch(x, t, 0.001, precision=hp.PRECISION_TENSOR)
p_1 = ref.get_params()
p0 = hp.PoseInitializerCNN("", device=local).get_params()
rel = np.abs((p_bf - p0) - (p_1 - p0)).max() / np.abs(p_1 - p0).max()
same = torch.from_numpy(p_bf).cuda(); ref0 = same.clone(); dist.broadcast(ref0, 0)
if rank == 0:
    print("bf16 wire: update rel err %.3e, ranks identical %s" % (rel, bool(torch.equal(same, ref0))), flush=True)
assert rel <= 2e-2 and torch.equal(same, ref0)
net.dp_set_bf16_gradients(False)
# inference shards: contiguous slices, replicated weights, no collective
net.Init()
y_shard = net.eval_batch(x[lo:hi], precision=hp.PRECISION_TENSOR)
y_full = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
assert np.array_equal(y_shard, y_full[lo:hi])
net.dp_shutdown()
dist.destroy_process_group()
if rank == 0: print("dp ok")
