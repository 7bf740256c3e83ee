"""P2P (symmetric-memory, copy-engine) output gather at batch 65 536 over N GPUs: ms per step for several
(copy streams, block, min_block) settings, against the sharded step without any gather.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gather_sweep.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
from builders import build_b200_gator, golden, synthetic
from gator_b200.dist import P2PGather, forward_gathered_p2p, rank_span, round_plan

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
total, J = 65536, 19
model = build_b200_gator('coco', dev).set_precision('bf16x3')
xg = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], total)).to(dev)


def timed(step, iters=5):
    for _ in range(2):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


res = {}
with torch.no_grad():
    lo, hi = rank * (total // world), (rank + 1) * (total // world)
    res['sharded_no_gather'] = timed(lambda: model(xg[lo:hi]))
    for streams, block, mb, mc in ((7, 2048, 1024, False), (4, 2048, 0, True), (4, 2048, 1024, True), (4, 1024, 0, True), (4, 4096, 1024, True)):
        pg = P2PGather(total, (6890, 3), dev, streams=streams, multicast=mc)
        plan = round_plan(total, world, block, mb)
        spans = [rank_span(total, st, n, rank) for st, n in plan]
        idx = torch.cat([torch.arange(a, b) for a, b in spans]).to(dev)
        offs, acc = {}, 0
        for a, b in spans:
            offs[a] = acc; acc += b - a
        xm = xg[idx].contiguous()
        st_ = {}

        def fn(a, b, out):
            o = offs[a]
            model.pose2mesh.forward_parts(xm[o:o + b - a], st_['p3'][o:o + b - a], st_['feat'][o:o + b - a], out=out)

        def step():
            p3, feat = model.pose_lifter(xm.reshape(xm.shape[0], -1))
            st_['p3'], st_['feat'] = p3.reshape(-1, J, 3), feat
            forward_gathered_p2p(fn, total, block, pg, mb)
        res[f'multicast={bool(pg._mc)} streams={streams} block={block} min_block={mb} rounds={[n for _, n in plan]}'] = timed(step)
        if rank == 0:
            probe = [total - 1, total // 2 + 3, 2049]
            res['verified ' + str(len(res))] = all(bool(torch.equal(model(xg[i:i + 1])[0][0], pg.out[i])) for i in probe)
        del pg
        torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(res, indent=1), flush=True)
dist.barrier()
dist.destroy_process_group()
