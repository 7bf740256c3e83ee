"""World-size-2 gloo test of the multi-GPU host logic (sharding + optional gather) on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gator_b200.dist import forward_gathered, gather_meshes, rank_span, round_plan, shard_range


def test_shard_range_partitions():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    full = torch.arange(total * 5 * 3, dtype=torch.float32).reshape(total, 5, 3)
    lo, hi = shard_range(total, world, rank)
    got = gather_meshes(full[lo:hi].clone(), total)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


def test_round_plan_covers_batch_once():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            for block in (1, 5, 1024):
                seen = 0
                for start, n in round_plan(total, world, block):
                    assert start == seen and n > 0
                    spans = [rank_span(total, start, n, r) for r in range(world)]
                    assert all(0 <= hi - lo <= n for lo, hi in spans)
                    assert spans[0][0] == start and all(a[1] == b[0] for a, b in zip(spans, spans[1:]) if b[1] > b[0])
                    seen = max(hi for _, hi in spans)
                assert seen == total


def _worker_rounds(rank, world, port, total, block, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    full = torch.arange(total * 5 * 3, dtype=torch.float32).reshape(total, 5, 3)
    calls = []

    def fn(lo, hi):
        calls.append((lo, hi))
        return full[lo:hi].clone()
    got = forward_gathered(fn, total, (5, 3), block)
    mine = sum(hi - lo for lo, hi in calls)
    q.put((rank, bool(torch.equal(got, full)), mine))
    dist.destroy_process_group()


@pytest.mark.parametrize('total,block', [(16, 2), (13, 3), (5, 4)])
def test_forward_gathered_world2_gloo(total, block):
    """Block-cyclic deal + chunked all_gather_into_tensor: every rank ends with the full batch, each sample computed once."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_rounds, args=(r, 2, port, total, block, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)]
    assert sum(r[2] for r in res) == total


class _GlooGather:
    """Stand-in for gator_b200.dist.P2PGather on CPU (same interface: world, rank, out, push, finish): push() records
    the rows this rank wrote in place, finish() ships them with gloo broadcasts - the transport is a copy engine on the
    GPU; what the test exercises is forward_gathered_p2p's round logic (tapered plans, ragged rounds, in-place writes)."""

    def __init__(self, total, feat):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.out = torch.full((total,) + tuple(feat), float('nan'))
        self.sent = []

    def push(self, lo, hi):
        self.sent.append((lo, hi))

    def finish(self):
        spans = [None] * self.world
        dist.all_gather_object(spans, self.sent)
        for r, lst in enumerate(spans):
            for lo, hi in lst:
                buf = self.out[lo:hi].clone()
                dist.broadcast(buf, src=r)
                self.out[lo:hi] = buf
        self.sent = []


def _worker_p2p(rank, world, port, total, block, min_block, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gator_b200.dist import forward_gathered_p2p
    full = torch.arange(total * 5 * 3, dtype=torch.float32).reshape(total, 5, 3)
    g = _GlooGather(total, (5, 3))
    calls = []

    def fn(lo, hi, out):
        calls.append((lo, hi))
        out.copy_(full[lo:hi])                      # written IN PLACE into the gathered buffer
    got = forward_gathered_p2p(fn, total, block, g, min_block)
    q.put((rank, bool(torch.equal(got, full)), sum(hi - lo for lo, hi in calls)))
    dist.destroy_process_group()


@pytest.mark.parametrize('total,block,min_block', [(64, 16, 4), (13, 4, 1), (5, 4, 2), (16, 2, 0)])
def test_forward_gathered_p2p_round_logic_world2_gloo(total, block, min_block):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_p2p, args=(r, 2, port, total, block, min_block, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)]
    assert sum(r[2] for r in res) == total


@pytest.mark.parametrize('total', [8, 7])
def test_gather_world2_gloo(total):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_host_pipeline_slice_plan():
    """gator_b200.pipeline.plan_slices: geometric slices that cover the batch exactly, never below two CTAs per SM."""
    from gator_b200.pipeline import HostPipeline, plan_slices
    for B in (1, 37, 296, 297, 700, 4096, 8192, 65536):
        b = plan_slices(B, HostPipeline.RATIO)
        sizes = [hi - lo for lo, hi in zip(b[:-1], b[1:])]
        assert b[0] == 0 and b[-1] == B and all(s > 0 for s in sizes)
        assert all(x >= y for x, y in zip(sizes[:-2], sizes[1:-1]))           # shrinking (the last one absorbs the remainder)
        if B > 592:
            assert min(sizes) >= 296
        if B >= 1500:
            assert len(sizes) >= 2 and sizes[-1] <= 0.4 * B                    # only a short copy is left exposed
    assert plan_slices(4096, 0.35, 148) == [0, 2662, 3593, 3919, 4096]
    assert plan_slices(4096, 0.8)[:3] == [0, 819, 1474] and len(plan_slices(4096, 0.8)) == 10


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa_node never raises: without a GPU / topology it reports None and leaves the affinity alone."""
    import os
    from gator_b200.dist import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)
    assert os.sched_getaffinity(0) <= before


def test_round_plan_taper():
    """round_plan(min_block): the rounds tile the batch exactly, every rank's span is contiguous inside its round, and the
    last rounds shrink by halves down to min_block (only a small transfer is left exposed at the end of the step)."""
    from gator_b200.dist import rank_span, round_plan
    assert [n for _, n in round_plan(65536, 8, 2048, 512)] == [2048, 2048, 2048, 1024, 512, 512]
    assert round_plan(65536, 8, 2048) == round_plan(65536, 8, 2048, 0) == [(i * 8 * 2048, 2048) for i in range(4)]
    for total, world, block, mb in ((65536, 4, 2048, 512), (1000, 3, 128, 32), (37, 8, 16, 4), (5, 2, 8, 2), (301, 1, 128, 64)):
        plan = round_plan(total, world, block, mb)
        covered = []
        for start, n in plan:
            assert n >= 1
            for r in range(world):
                lo, hi = rank_span(total, start, n, r)
                covered += list(range(lo, hi))
        assert covered == list(range(total)), (total, world, block, mb)
        sizes = [n for _, n in plan]
        assert all(a >= b for a, b in zip(sizes[:-1], sizes[1:-1] + sizes[-1:])) or sizes[-1] <= sizes[0]
