"""N-GPU data-parallel check (launched by tests/test_multi_gpu.py under torch.distributed.run), for both exchange
paths -- the NVLink peer-memory kernel (reduce-scatter + SGD + all-gather of weights, csrc/hp_peer.cu) and the NCCL
all-reduce baseline: the step on sharded data must equal the single-GPU step on the concatenated batch, every rank
must end with bit-identical weights, and several consecutive steps must keep doing so (epoch-counted flag barriers).
Also checks that sharded inference equals unsharded inference."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_tracking_samples_b200 import cnn as hp, synth, dp
rank, world, local = dp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
n = 64
x, t = synth.depthlike_crops(n, 7), synth.heatmap_labels(n, 8)
lo, hi = dp.shard_range(n, rank, world)
st = torch.cuda.current_stream().cuda_stream
xd, td = torch.from_numpy(x[lo:hi].copy()).cuda(), torch.from_numpy(t[lo:hi].copy()).cuda()
p0 = hp.PoseInitializerCNN("", device=local).get_params()


def ranks_identical(p):
    same = torch.from_numpy(p).cuda(); ref0 = same.clone(); dist.broadcast(ref0, 0)
    ok = torch.tensor([1.0 if torch.equal(same, ref0) else 0.0], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item() == 1.0)


for mode in ("peer", "nccl"):
    net = hp.PoseInitializerCNN("", device=local)
    dp.init_data_parallel(net, mode=mode)
    for prec, tol, steps in ((hp.PRECISION_FP32, 2e-5, 1), (hp.PRECISION_TENSOR, 2e-2, 1), (hp.PRECISION_FP32, 5e-5, 4)):
        net.Init()
        mse = torch.empty(hi - lo, device="cuda")
        for _ in range(steps):
            net.train_batch_device(xd.data_ptr(), td.data_ptr(), hi - lo, 0.001, mse.data_ptr(), precision=prec, stream=st)
        torch.cuda.synchronize()
        p_dp = net.get_params()
        # single-GPU reference of the concatenated batch on this rank
        ref = hp.PoseInitializerCNN("", device=local)
        for _ in range(steps):
            ref.train_batch(x, t, 0.001, precision=prec)
        p_1 = ref.get_params()
        rel = np.abs((p_dp - p0) - (p_1 - p0)).max() / np.abs(p_1 - p0).max()
        g = torch.tensor([rel], device="cuda"); dist.all_reduce(g, op=dist.ReduceOp.MAX)
        identical = ranks_identical(p_dp)
        if rank == 0:
            print("%s precision %d steps %d: update rel err DP(%d ranks) vs 1 GPU: %.3e (tol %g)  ranks identical: %s"
                  % (mode, prec, steps, world, g.item(), tol, identical), flush=True)
        assert g.item() <= tol and identical
    if mode == "peer":
        assert net.dp_peer_status() == 0
        # the same steps on a capturable stream: the second identical call captures the whole step (exchange kernels
        # included -- their barrier epochs live in device memory) and later calls replay the graph
        side = torch.cuda.Stream()
        for prec, tol in ((hp.PRECISION_FP32, 5e-5), (hp.PRECISION_TENSOR, 2e-2)):
            res = []
            for stream in (st, side.cuda_stream):
                net.Init()
                torch.cuda.synchronize()
                for _ in range(5):
                    net.train_batch_device(xd.data_ptr(), td.data_ptr(), hi - lo, 0.001, None, precision=prec, stream=stream)
                torch.cuda.synchronize()
                res.append(net.get_params())
            ref = hp.PoseInitializerCNN("", device=local)
            for _ in range(5):
                ref.train_batch(x, t, 0.001, precision=prec)
            p_1 = ref.get_params()
            rel = np.abs((res[1] - p0) - (p_1 - p0)).max() / np.abs(p_1 - p0).max()
            g = torch.tensor([rel], device="cuda"); dist.all_reduce(g, op=dist.ReduceOp.MAX)
            identical = ranks_identical(res[1])
            if rank == 0:
                print("peer, graph replay, precision %d, 5 steps: rel err vs 1 GPU %.3e (tol %g), ranks identical %s, equal to the eager DP steps: %s"
                      % (prec, g.item(), tol, identical, np.array_equal(res[0], res[1])), flush=True)
            assert g.item() <= tol and identical and np.array_equal(res[0], res[1])
        assert net.dp_peer_status() == 0
    else:
        # opt-in bf16 gradient transport (NCCL path, tensor-precision steps only): same step within the tensor-path bound
        net.Init()
        net.dp_set_bf16_gradients(True)
        net.train_batch_device(xd.data_ptr(), td.data_ptr(), hi - lo, 0.001, None, precision=hp.PRECISION_TENSOR, stream=st)
        torch.cuda.synchronize()
        p_bf = net.get_params()
        ref = hp.PoseInitializerCNN("", device=local)
        ref.train_batch(x, t, 0.001, precision=hp.PRECISION_TENSOR)
        p_1 = ref.get_params()
        rel = np.abs((p_bf - p0) - (p_1 - p0)).max() / np.abs(p_1 - p0).max()
        identical = ranks_identical(p_bf)
        if rank == 0:
            print("bf16 wire: update rel err %.3e, ranks identical %s" % (rel, identical), flush=True)
        assert rel <= 2e-2 and identical
        net.dp_set_bf16_gradients(False)
    dp.shutdown_data_parallel(net)
# inference shards: contiguous slices, replicated weights, no collective
net = hp.PoseInitializerCNN("", device=local)
y_shard = net.eval_batch(x[lo:hi], precision=hp.PRECISION_TENSOR)
y_full = net.eval_batch(x, precision=hp.PRECISION_TENSOR)
assert np.array_equal(y_shard, y_full[lo:hi])
dist.destroy_process_group()
if rank == 0: print("dp ok")
