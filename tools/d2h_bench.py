"""Device-to-host copy bandwidth per rank with N ranks copying at once (the limiter of the host-to-host `e2e` figure at
N > 1): every rank copies a (4096, 6890, 3) fp32 mesh block to pinned host memory, all ranks together and rank 0 alone.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_bench.py
Prints one JSON line on rank 0."""
import json, os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gator_b200.dist import bind_to_gpu_numa_node
node = bind_to_gpu_numa_node(local)
src = torch.randn(4096, 6890, 3, device=dev)
dst = torch.empty(src.shape, dtype=src.dtype).pin_memory()
nbytes = src.numel() * 4


def run(active: bool, iters=10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9 if active else 0.0
    t = torch.tensor([gbs], device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(o.item(), 1) for o in out]
    return [round(gbs, 1)]


run(True, 3)
together = run(True)
alone = run(rank == 0)
if rank == 0:
    print(json.dumps({'ranks': world, 'mb_per_copy': nbytes / 1e6, 'numa_node_rank0': node,
                      'd2h_gbs_per_rank_all_ranks_copying': together, 'aggregate_gbs': round(sum(together), 1),
                      'd2h_gbs_rank0_alone': alone[0]}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
