"""Host-side plumbing for the multi-GPU paths: one process per GPU.

* inference: contiguous batch slices per rank, weights replicated, NO collective;
* training: data parallel -- each rank computes the gradient sum of its slice, the
  C ABI all-reduces it over NCCL (hp_dp_init) and every rank applies the same update.

torch.distributed is only the rendezvous/side channel here (unique-id broadcast,
barriers, max-over-ranks timing); the gradient all-reduce itself runs inside
libhandposedd on its own CUDA stream.
"""
import os


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n units owned by `rank` ([g*N/G, (g+1)*N/G))."""
    return n * rank // world, n * (rank + 1) // world


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(data, src=0):
    """Broadcast a small bytes object from `src` over the default torch.distributed group
    (works on gloo and nccl)."""
    import torch
    import torch.distributed as dist
    obj = [data if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def init_data_parallel(cnn):
    """Join the net to an NCCL communicator spanning the default process group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = cnn.dp_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0)
    cnn.dp_init(uid, rank, world)
    return rank, world
