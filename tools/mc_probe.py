"""Probe: does a copy-engine cudaMemcpyAsync to the NVLS multicast address of a symmetric-memory buffer broadcast the data
to every rank?  torchrun --nproc-per-node 2 tools/mc_probe.py"""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from cuda import cudart
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 64 << 20
buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
buf.zero_(); torch.cuda.synchronize(); dist.barrier()
mc = hdl.multicast_ptr
print(rank, 'multicast_ptr', hex(mc) if mc else mc, flush=True)
src = torch.full((n // world,), float(rank + 1), device=dev)
off = rank * (n // world) * 4
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    err, = cudart.cudaMemcpyAsync(mc + off, src.data_ptr(), src.numel() * 4, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, st)
    e1.record()
    torch.cuda.synchronize()
    print(rank, it, 'multicast cudaMemcpyAsync ->', err, 'ms', round(e0.elapsed_time(e1), 3), 'GB/s', round(src.numel() * 4 / e0.elapsed_time(e1) / 1e6, 1), flush=True)
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    peer[rank * (n // world):(rank + 1) * (n // world)].copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(rank, it, 'unicast peer copy ms', round(e0.elapsed_time(e1), 3), 'GB/s', round(src.numel() * 4 / e0.elapsed_time(e1) / 1e6, 1), flush=True)
dist.barrier(); torch.cuda.synchronize()
ok = all(bool((buf[r * (n // world):(r + 1) * (n // world)] == r + 1).all()) for r in range(world))
print(rank, 'broadcast ok' if ok else 'broadcast FAILED', [float(buf[r * (n // world)]) for r in range(world)], flush=True)
dist.barrier(); dist.destroy_process_group()
