"""Host-side plumbing for the multi-GPU paths: one process per GPU.

* inference: contiguous batch slices per rank, weights replicated, NO collective;
* training: data parallel -- each rank computes the gradient sum of its slice; the exchange
  step is either ONE kernel per gradient bucket over NVLink peer memory (reduce-scatter + SGD +
  all-gather of the updated weights, hp_dp_peer_init; the default) or an NCCL all-reduce
  followed by the local SGD kernel (hp_dp_init; the comparison baseline).

torch.distributed is only the rendezvous/side channel here (unique-id broadcast,
barriers, max-over-ranks timing); the gradient all-reduce itself runs inside
libhandposedd on its own CUDA stream.
"""
import os


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n units owned by `rank` ([g*N/G, (g+1)*N/G))."""
    return n * rank // world, n * (rank + 1) // world


# gradient buckets of the flat .cnnb-ordered stores, in the order backward produces them: fc2 | fc1 | conv1+conv2
BUCKETS = ((4737504, 9458400), (16864, 4737504), (0, 16864))


def peer_slice(count4, rank, world):
    """float4 range [lo, hi) of a bucket of count4 float4s that `rank` reduces, updates and publishes in the
    peer-memory exchange kernel (csrc/hp_peer.cu, peer_sgd_kernel: same integer arithmetic)."""
    return count4 * rank // world, count4 * (rank + 1) // world


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


_ORIG_AFFINITY = None


def unbind_host():
    """Undo bind_host_to_gpu (e.g. before a CPU-side job that should use every core)."""
    if _ORIG_AFFINITY is not None:
        os.sched_setaffinity(0, _ORIG_AFFINITY)


def bind_host_to_gpu(local_rank):
    """Pin this process's host threads to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging
    buffers it allocates afterwards (first touch) and the threads that fill them are local to the GPU's PCIe root
    (matters for the host-buffer entry points when several ranks share one box).  Best effort: returns a small dict
    describing what was done, never raises."""
    global _ORIG_AFFINITY
    info = {"bound": False}
    try:
        import torch
        if _ORIG_AFFINITY is None:
            _ORIG_AFFINITY = os.sched_getaffinity(0)
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        node_path = "/sys/bus/pci/devices/%s/numa_node" % bus
        node = int(open(node_path).read().strip())
        info.update(pci=bus, numa_node=node)
        if node < 0:
            return info
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:  # containers without sysfs access, CPU-only hosts, ...
        info["note"] = str(e)[:80]
    return info


def broadcast_bytes(data, src=0):
    """Broadcast a small bytes object from `src` over the default torch.distributed group
    (works on gloo and nccl)."""
    import torch
    import torch.distributed as dist
    obj = [data if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def all_gather_bytes(data):
    """All-gather equal-length bytes objects over the default group, concatenated in rank order."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, data)
    return b"".join(out)


def init_data_parallel(cnn, mode="peer"):
    """Join the net to the data-parallel group spanning the default process group.
    mode "peer": NVLink peer-memory exchange kernel (CUDA IPC handles all-gathered here);
    mode "nccl": NCCL all-reduce + local SGD."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if mode == "peer":
        handles = all_gather_bytes(cnn.dp_peer_export())
        cnn.dp_peer_init(handles, rank, world)
        dist.barrier()
    elif mode == "nccl":
        uid = cnn.dp_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 0)
        cnn.dp_init(uid, rank, world)
    else:
        raise ValueError(mode)
    return rank, world


def shutdown_data_parallel(cnn):
    """Host-side barrier (no rank may unmap a store a peer's kernel still writes), then hp_dp_shutdown."""
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    cnn.dp_shutdown()
    dist.barrier()
