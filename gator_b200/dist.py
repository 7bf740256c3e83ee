"""Batch sharding for multi-GPU runs (one process per GPU).  Samples are independent in every stage of
the forward (eval-mode BatchNorm uses running stats), so ranks take contiguous batch slices, weights are
replicated, and there is no collective on the data path; gather_meshes() is the optional final
all_gather of the (B/G, 6890, 3) output shards over NVLink."""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `total` samples for `rank` (first total % world ranks get +1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError('bad world/rank')
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_meshes(local_mesh, total: int, group=None):
    """all_gather of per-rank mesh shards (possibly ragged by one sample) into a (total, 6890, 3) tensor on
    every rank.  Works with nccl (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local_mesh.new_zeros((mx,) + tuple(local_mesh.shape[1:]))
    pad[:local_mesh.shape[0]] = local_mesh
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards are local to the GPU's PCIe root (with 8 ranks per node, device-to-host copies otherwise cross sockets).
    Best effort: returns the node id, or None when the topology cannot be read."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
