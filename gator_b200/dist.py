"""Batch sharding for multi-GPU runs (one process per GPU).  Samples are independent in every stage of
the forward (eval-mode BatchNorm uses running stats), so ranks take contiguous batch slices, weights are
replicated, and there is no collective on the data path.  The optional final gather of the (B/G, 6890, 3) output shards
over NVLink is forward_gathered(): block-cyclic deal + chunked all_gather_into_tensor on a side
stream, overlapped with the next chunk's compute; gather_meshes() is the plain (non-overlapped) form for contiguous shards."""
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
    if total % world == 0:                      # equal shards: gather straight into the output, no pad / cat copies
        full = local_mesh.new_empty((total,) + tuple(local_mesh.shape[1:]))
        dist.all_gather_into_tensor(full, local_mesh.contiguous(), group=group)
        return full
    pad = local_mesh.new_zeros((mx,) + tuple(local_mesh.shape[1:]))
    pad[:local_mesh.shape[0]] = local_mesh
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def round_plan(total: int, world: int, block: int):
    """Block-cyclic deal of `total` samples for the overlapped gather: round k covers the global samples
    [start_k, start_k + world * n_k) and rank r owns [start_k + r * n_k, start_k + (r + 1) * n_k) of it, so the gathered
    round is ONE contiguous slice of the full output - all_gather_into_tensor writes it in place, no pad / cat copies.
    n_k = block except in the last round (n = ceil(rest / world); ranks past the end own fewer or no samples).
    Returns [(start, n)]."""
    if world <= 0 or block <= 0:
        raise ValueError('bad world/block')
    plan, start = [], 0
    while start < total:
        rest = total - start
        n = block if rest >= world * block else -(-rest // world)
        plan.append((start, n))
        start += world * n
    return plan


def rank_span(total: int, start: int, n: int, rank: int) -> Tuple[int, int]:
    """Global [lo, hi) that `rank` owns in the round (start, n), clipped to the batch."""
    lo = min(total, start + rank * n)
    return lo, min(total, lo + n)


def forward_gathered(fn, total: int, feat_shape, block: int, *, dtype=None, device=None, group=None, out=None):
    """Every rank computes its share of a `total`-sample batch with fn(lo, hi) -> (hi - lo, *feat_shape) and ends up with
    the FULL (total, *feat_shape) result (BASELINE configs[4]: batch-sharded forward + NCCL output gather).
    The batch is dealt block-cyclically (round_plan); the all_gather of round k runs on a side stream while fn computes
    round k + 1, so only the last round's gather is exposed.  Works with nccl (CUDA) and gloo (CPU tensors, tests)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan = round_plan(total, world, block)
    cuda = device is not None and torch.device(device).type == 'cuda'
    if out is None:
        out = torch.empty((total,) + tuple(feat_shape), dtype=dtype or torch.float32, device=device)
    comm = torch.cuda.Stream(device=device) if cuda else None
    works, keep = [], []
    for start, n in plan:
        lo, hi = rank_span(total, start, n, rank)
        y = fn(lo, hi)
        full_round = start + world * n <= total
        if full_round:
            dst, src = out[start:start + world * n], y
        else:                                   # ragged last round: equal-size padded blocks, valid rows copied out below
            dst = out.new_empty((world * n,) + tuple(feat_shape))
            src = out.new_zeros((n,) + tuple(feat_shape))
            src[:hi - lo] = y
        if cuda:
            ev = torch.cuda.Event()
            ev.record()
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                works.append(dist.all_gather_into_tensor(dst, src, group=group, async_op=True))
            src.record_stream(comm)
        else:
            dist.all_gather_into_tensor(dst, src, group=group)
        keep.append((full_round, dst, src, start, n))
    for w in works:
        w.wait()                                # the current stream waits for the collective
    if cuda:
        torch.cuda.current_stream(device).wait_stream(comm)
    for full_round, dst, _, start, n in keep:
        if not full_round:
            for r in range(world):
                lo, hi = rank_span(total, start, n, r)
                if hi > lo:
                    out[lo:hi] = dst[r * n:r * n + (hi - lo)]
    return out


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
