"""Batch sharding for multi-GPU runs (one process per GPU).  Samples are independent in every stage of
the forward (eval-mode BatchNorm uses running stats), so ranks take contiguous batch slices, weights are
replicated, and there is no collective on the data path.  The optional final gather of the (B/G, 6890, 3) output shards
over NVLink is forward_gathered(): block-cyclic deal + chunked all_gather_into_tensor on a side
stream, overlapped with the next chunk's compute; gather_meshes() is the plain (non-overlapped) form for contiguous shards.
P2PGather / forward_gathered_p2p() is the same deal without a collective kernel: the output buffer is symmetric memory
(mapped into every rank over NVLink / NVSwitch), the decoder writes its slice in place and the copy engines push it to
the peers - no SM is taken from the compute kernels, which is what NCCL's all-gather kernels cost at N = 8."""
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


def round_plan(total: int, world: int, block: int, min_block: int = 0):
    """Block-cyclic deal of `total` samples for the overlapped gather: round k covers the global samples
    [start_k, start_k + world * n_k) and rank r owns [start_k + r * n_k, start_k + (r + 1) * n_k) of it, so the gathered
    round is ONE contiguous slice of the full output - all_gather_into_tensor writes it in place, no pad / cat copies.
    n_k = block except in the last round (n = ceil(rest / world); ranks past the end own fewer or no samples).
    With 0 < min_block < block the plan TAPERS: once fewer than two blocks per rank remain the block is halved, down to
    min_block, so that the only transfer left exposed at the end of the step is a small one (per rank 8192, block 2048,
    min_block 512: 2048, 2048, 2048, 1024, 512, 512).  Returns [(start, n)]."""
    if world <= 0 or block <= 0:
        raise ValueError('bad world/block')
    plan, start = [], 0
    while start < total:
        rest = total - start
        n = block
        if 0 < min_block < block:
            per_rank = -(-rest // world)
            while n > min_block and per_rank < 2 * n:
                n //= 2
            n = max(n, min_block)
        if rest < world * n:
            n = -(-rest // world)
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


class P2PGather:
    """The full (total, *feat_shape) fp32 output of a batch-sharded forward as SYMMETRIC MEMORY
    (torch.distributed._symmetric_memory: cuMem allocations mapped into every rank of the node over NVLink / NVSwitch):
    `out` is this rank's copy, `peers[r]` a view of rank r's copy.  push(lo, hi) sends rows [lo, hi) of `out` - just
    written by this rank - to every peer with device-to-device copies on side streams (copy engines, no SMs): one copy to
    the NVSwitch multicast address when the platform offers one, else one copy per peer;
    finish() orders the current stream after this rank's pushes and after a cross-rank barrier, so that every rank's
    buffer is complete.  NCCL group required (the rendezvous exchanges the handles through its store)."""

    def __init__(self, total: int, feat_shape, device, group=None, streams: int = 7, multicast: bool = False):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.out = symm_mem.empty((total,) + tuple(feat_shape), dtype=torch.float32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.out, group)
        self.peers = [self.out if r == self.rank else self.hdl.get_buffer(r, tuple(self.out.shape), self.out.dtype)
                      for r in range(self.world)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, min(streams, self.world - 1)))]
        # Optional NVSwitch multicast (NVLS): ONE copy-engine write to the multicast address of the buffer lands in every
        # rank's copy, so a rank sends each round once instead of world - 1 times (tools/mc_probe.py: it works, 330-470 GB/s
        # of source data per copy on 2 GPUs, every destination verified).  Off by default: on 8 GPUs the multicast copies
        # were slower end to end than one unicast copy per peer on 7 streams (21.4 vs 19.8 ms per 65 536-sample step,
        # tools/gather_sweep.py, profiles/r02_gather_sweep.txt)
        self._mc, self._cudart = 0, None
        if multicast and self.world > 1:
            try:
                from cuda import cudart
                mc = int(self.hdl.multicast_ptr)
                if mc:
                    self._mc, self._cudart = mc, cudart
            except Exception:
                self._mc = 0
        self.row_bytes = 4
        for d in feat_shape:
            self.row_bytes *= int(d)

    def push(self, lo: int, hi: int):
        import torch
        if hi <= lo:
            return
        ev = torch.cuda.Event()
        ev.record()                                         # rows [lo, hi) of self.out are complete on the current stream
        src = self.out[lo:hi]
        if self._mc:
            st = self.streams[0]
            st.wait_event(ev)
            rt = self._cudart
            err, = rt.cudaMemcpyAsync(self._mc + lo * self.row_bytes, src.data_ptr(), (hi - lo) * self.row_bytes,
                                      rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, st.cuda_stream)
            if int(err) != 0:
                raise RuntimeError(f'P2PGather: multicast copy failed: {err}')
            return
        for k in range(1, self.world):                      # peer order rotated by rank: the pushes of a round spread over
            r = (self.rank + k) % self.world                # all links instead of converging on one receiver
            st = self.streams[(k - 1) % len(self.streams)]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                self.peers[r][lo:hi].copy_(src, non_blocking=True)

    def finish(self):
        import torch
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)
        self.hdl.barrier(channel=0)                         # stream-ordered: every rank's pushes have landed everywhere


def forward_gathered_p2p(fn, total: int, block: int, gather: 'P2PGather', min_block: int = 0):
    """forward_gathered() over symmetric memory: fn(lo, hi, out) must write its (hi - lo, *feat) result INTO `out` (a
    slice of the symmetric buffer); round k's rows are pushed to the peers by the copy engines while round k + 1
    computes.  Ragged last rounds need no padding (every push is an independent copy); `min_block` tapers the last
    rounds (round_plan) so that only a small push is left exposed.  Returns gather.out."""
    for start, n in round_plan(total, gather.world, block, min_block):
        lo, hi = rank_span(total, start, n, gather.rank)
        if hi > lo:
            fn(lo, hi, gather.out[lo:hi])
            gather.push(lo, hi)
    gather.finish()
    return gather.out


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
