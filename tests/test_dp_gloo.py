"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard partitioning, the
unique-id side channel, and the data-parallel identity the training path relies on --
sum over ranks of per-shard gradient sums == gradient sum of the concatenated batch."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from hand_tracking_samples_b200 import dp, synth
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        p0 = o.init_xavier()
        n = 3
        x = synth.depthlike_crops(n, 99)
        t = synth.heatmap_labels(n, 98)
        lo, hi = dp.shard_range(n, rank, world)
        g_local, mse_local = o.train_minibatch(p0.copy(), x[lo:hi], t[lo:hi], 0.001, apply=False)
        g = torch.from_numpy(g_local.copy())
        dist.all_reduce(g)  # what ncclAllReduce(sum) does inside libhandposedd
        uid = dp.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        # the peer-memory exchange (csrc/hp_peer.cu) restated on the host: IPC handles all-gathered in rank order;
        # rank r sums ITS slice of every rank's gradient store in rank order, updates its slice of the weights and
        # publishes the new weights to everybody
        handles = dp.all_gather_bytes(bytes([rank + 1]) * 256)
        out["handles_ok_%d" % rank] = handles == b"".join(bytes([r + 1]) * 256 for r in range(world))
        stores = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(stores, torch.from_numpy(g_local.copy()))     # "peer-mapped gradient stores"
        w_new = torch.zeros(len(p0), dtype=torch.float64)
        for b_lo, b_hi in dp.BUCKETS:
            lo4, hi4 = dp.peer_slice((b_hi - b_lo) // 4, rank, world)
            sl = slice(b_lo + 4 * lo4, b_lo + 4 * hi4)
            acc = stores[0][sl].clone()
            for r in range(1, world):
                acc += stores[r][sl]
            w_new[sl] = torch.from_numpy(p0.astype(np.float64))[sl] - 0.001 * acc
        dist.all_reduce(w_new)    # slices are disjoint and cover every bucket: the sum is the all-gather
        if rank == 0:
            g_full, mse_full = o.train_minibatch(p0.copy(), x, t, 0.001, apply=False)
            out["max_abs_diff"] = float(np.abs(g.numpy() - g_full).max())
            out["max_abs"] = float(np.abs(g_full).max())
            out["uid_ok"] = uid == bytes(range(128))
            out["peer_w_diff"] = float(np.abs(w_new.numpy() - (p0.astype(np.float64) - 0.001 * g_full)).max())
        else:
            out["uid_ok_1"] = uid == bytes(range(128))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from hand_tracking_samples_b200 import dp
    for n in (0, 1, 7, 256, 65536, 1000003):
        for world in (1, 2, 4, 8):
            edges = [dp.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_peer_slices_partition_every_bucket():
    from hand_tracking_samples_b200 import dp
    assert sorted(dp.BUCKETS) == [(0, 16864), (16864, 4737504), (4737504, 9458400)]
    for lo, hi in dp.BUCKETS:
        assert (hi - lo) % 4 == 0 and lo % 4 == 0
        for world in range(2, 9):
            edges = [dp.peer_slice((hi - lo) // 4, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == (hi - lo) // 4
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_data_parallel_gradient_identity_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["uid_ok"] and out["uid_ok_1"]
    assert out["handles_ok_0"] and out["handles_ok_1"]
    assert out["peer_w_diff"] <= 1e-12
    # double accumulation on both sides: the shard sums add up to the full-batch sum to rounding
    assert out["max_abs_diff"] <= 1e-12 * max(1.0, out["max_abs"]) + 1e-15
