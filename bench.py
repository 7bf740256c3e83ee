#!/usr/bin/env python
"""Headline benchmark: meshes/sec of the GATOR pose->mesh forward (J=19, alpha=True, synthetic COCO poses).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--precision fp32|bf16|bf16x3]

N = 1: BASELINE configs[3] - batch 4096 on one B200.  N > 1 (torchrun, one rank per GPU): BASELINE configs[4] - a
batch of 65 536 sharded over the N GPUs (65 536 / N samples per rank), reported without and with the NCCL output gather.

One JSON line on stdout (rank 0):
  value           whole-job meshes/s, inputs resident in HBM, CUDA events per step, L2 flushed before every step, max over ranks
  with_gather     (N > 1) the same job ending with the FULL (65 536, 6890, 3) result on every rank: block-cyclic deal,
                  chunked all_gather_into_tensor (NCCL) on a side stream overlapped with the next chunk's kernels;
                  with_gather.p2p = the same over symmetric memory: decoder writes in place, copy engines push each round
                  to the peers over NVLink (no SMs); value_with_gather = the better of the two
  e2e             host -> host through the public API (gator_b200.pipeline.HostPipeline.submit/result): pinned inputs, H2D,
                  forward, full mesh + pose3d copied back to pinned host memory, all inside the timed region; the copy of
                  step i overlaps the kernels of step i+1 (double-buffered pinned outputs, the last copy is exposed and counted)
  e2e_sync        the same through the synchronous HostPipeline.forward (a batch's own copy overlaps its own sliced decoder)
  e2e_eval        host -> device -> host with the evaluation epilogue (row f1) on the device: only per-sample errors return
  roofline        the dominant kernel timed alone through its C-ABI entry (plus the other hot kernels)
  other_workloads BASELINE configs[2]: SMPL_Layer alone at batch 16 384, fp32 and tensor-core path, HBM roofline; the fused
                  two-level Mesh.upsample at batch 4096; the MANO layer at batch 16 384
  cpu_baseline    (N = 1) the CPU oracle (port of the reference forward, same ATen ops) on the box's host cores
  gpu_eager_baseline (N = 1, informational) the same oracle ops run eagerly on this GPU (fp32, TF32 off, chunks of 256)
--impl reference times the CPU oracle as the reference arm (the reference is plain PyTorch and is not installable offline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'meshes_per_sec'
UNIT = 'meshes/s'
TAG = 'coco'                      # J=19, alpha=True (3dpw configuration of demo/run.py)
FLOP_PER_MESH = 418.1e6           # GATOR.forward J=19, dense, 2*MAC
SA_FLOP_PER_SAMPLE = 2 * 2 * 2 * 431 * 431 * 32   # self-attention core: 2 heads x (QK^T + PV) x 2*MAC
MESH_BYTES = 6890 * 3 * 4
GLOBAL_BATCH_MULTI = 65536        # BASELINE configs[4]
GATHER_BLOCK = 2048               # samples per rank and gather round
GATHER_MIN_BLOCK = 1024           # P2P gather: the last rounds taper down to this (gator_b200.dist.round_plan)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p['bf16_tflops']),
                'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), 'src': 'measured'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'src': 'fallback'}


def ncu_traffic():
    """DRAM bytes per launch of the hot kernels from the committed `ncu` capture (profiles/r02_traffic.json, written by
    tools/launch_summary.py --json from the launch list of a 4096-sample forward); None when absent."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None                        # timed region (perf_counter); samples outside it are dropped

    def mark(self, begin: bool):
        if begin:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.12)]
        if not rows:                                    # region shorter than one polling period: closest samples
            rows = [r for _, r in self.rows[-2:]]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, col in (('hw_slowdown', 5), ('hw_thermal_slowdown', 6), ('sw_thermal_slowdown', 7), ('sw_power_cap', 8)):
                    if r[col].lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def make_inputs(batch, seed=2):
    from builders import golden, synthetic
    base = golden('fixtures')['demo_pose19']
    return synthetic.coco_poses2d(base, batch, seed=seed)


def cpu_forward_rate(sample_batch, min_seconds, threads):
    """Reference arm / cpu_baseline: the CPU oracle (oracle/gator_oracle.py, same ATen ops as the reference's
    forward) on `threads` host threads, batches of `sample_batch` until `min_seconds` elapsed."""
    import torch
    from helpers import oracle_setup, orc
    torch.set_num_threads(threads)
    sd, gc, mc, alpha = oracle_setup(TAG)
    x = torch.from_numpy(make_inputs(sample_batch))
    with torch.no_grad():
        orc.gator_forward(sd, gc, mc, x, alpha)            # warm-up
        t0, n = time.perf_counter(), 0
        times = []
        while time.perf_counter() - t0 < min_seconds or n < 2:
            t1 = time.perf_counter()
            orc.gator_forward(sd, gc, mc, x, alpha)
            times.append(time.perf_counter() - t1)
            n += 1
    times.sort()
    med = times[len(times) // 2]
    return sample_batch / med, med, n


def gpu_eager_rate(dev, batch=4096, chunk=256):
    """Informational same-GPU baseline: the oracle's ATen op sequence - the reference forward - executed
    eagerly on this GPU in fp32 (TF32 off), `batch` samples in chunks of `chunk` (the reference materialises the
    (B,2,431,431) score tensors: 1.5 MB per sample and layer).  Checker code: only bench.py's baseline legs run it."""
    import torch
    from helpers import oracle_setup, orc
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd, gc, mc, alpha = oracle_setup(TAG)
    mv = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}
    sd, gc, mc = mv(sd), mv(gc), mv(mc)
    x = torch.from_numpy(make_inputs(batch)).to(dev)

    def run():
        for lo in range(0, batch, chunk):
            orc.gator_forward(sd, gc, mc, x[lo:lo + chunk], alpha)
    with torch.no_grad():
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {'value': batch / (ms * 1e-3), 'unit': UNIT, 'ms_per_4096': ms * 4096 / batch, 'batch': batch, 'chunk': chunk,
            'what': 'oracle port of the reference forward, PyTorch eager (cuBLAS / ATen, fp32, allow_tf32=False) on this GPU; informational'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sample = 64
    # one "step" = one forward over a 64-sample slice of the workload (bounded sample)
    from helpers import oracle_setup, orc
    torch.set_num_threads(threads)
    sd, gc, mc, alpha = oracle_setup(TAG)
    x = torch.from_numpy(make_inputs(sample))
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            orc.gator_forward(sd, gc, mc, x, alpha)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.gator_forward(sd, gc, mc, x, alpha)
        dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    world = args.gpus
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong' if world > 1 else 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': workload_name(world, per_gpu_batch(args, world)),
                       'note': 'CPU reference forward (oracle port, same ATen ops) on rank 0; each step = a 64-sample slice of the workload'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': f'{args.steps} forwards of batch {sample}'},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def per_gpu_batch(args, world):
    if args.batch:
        return args.batch
    return 4096 if world == 1 else GLOBAL_BATCH_MULTI // world


def workload_name(world, B):
    if world == 1:
        return (f'GATOR.forward J=19 alpha=True, batch {B} on one GPU (BASELINE configs[3]); random-init weights, synthetic '
                'SMPL-shaped template/bases')
    return (f'GATOR.forward J=19 alpha=True, batch {B * world} sharded over {world} GPUs = {B} per GPU (BASELINE configs[4]); '
            'random-init weights, synthetic SMPL-shaped template/bases')


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from gator_b200 import _lib
    from gator_b200.dist import bind_to_gpu_numa_node, forward_gathered, rank_span, round_plan, shard_range
    numa_node = bind_to_gpu_numa_node(local)              # before any pinned buffer is allocated
    from builders import build_b200_gator
    L = _lib.lib()
    model = build_b200_gator(TAG, dev).set_precision(args.precision)
    J = model.num_joint
    B = per_gpu_batch(args, world)                         # per GPU
    total = B * world
    lo, hi = shard_range(total, world, rank)
    x_all = make_inputs(total)
    x_host = torch.from_numpy(x_all[lo:hi]).pin_memory()
    x = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident throughput ----
    # nvidia-smi is started well before the timed region: its start-up (NVML initialisation) was seen to stall the GPU for
    # tens of milliseconds when it coincided with the first timed steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        t_ramp = time.perf_counter()                       # untimed: first call packs the weights, then ~0.8 s of
        while True:                                        # forwards so that the SM clock has ramped before the W
            model(x)                                       # warm-up steps the contract asks for
            torch.cuda.synchronize()
            if time.perf_counter() - t_ramp > 0.8:
                break
        mesh = p3 = None
        for _ in range(args.warmup):
            mesh, p3 = model(x)                            # same binding pattern as the timed loop: the previous step's
        barrier()                                          # outputs are still alive while the next ones are allocated
        L.gator_launch_count(1)
        sampler.mark(True)
        evs = []
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mesh, p3 = model(x)
            e1.record()
            evs.append((e0, e1))
        barrier()
        t_wall = time.perf_counter() - t_wall
        sampler.mark(False)
        launches = L.gator_launch_count(1)
        clocks = sampler.stop() if rank == 0 else None
    per_step = [a.elapsed_time(b) for a, b in evs]
    step_ms = max_over_ranks(sum(per_step) / args.steps)
    value = total / (step_ms * 1e-3)
    del mesh, p3

    # ---- N > 1: the same job ending with the full result on every rank (NCCL gather over NVLink) ----
    gather = None
    if world > 1:
        xg = torch.from_numpy(x_all).to(dev)               # every rank holds the (tiny) inputs: 152 B per sample
        full = torch.empty((total, 6890, 3), dtype=torch.float32, device=dev)

        # the rank's samples in round order (block-cyclic deal): the lifter runs once over all of them, the decoder per round
        plan = round_plan(total, world, GATHER_BLOCK)
        spans = [rank_span(total, st, n, rank) for st, n in plan]
        idx = torch.cat([torch.arange(a, b) for a, b in spans]).to(dev)
        offs, acc = {}, 0
        for a, b in spans:
            offs[a] = acc
            acc += b - a
        x_mine = xg[idx].contiguous()
        state = {}

        def lift():
            p3_, feat_ = model.pose_lifter(x_mine.reshape(x_mine.shape[0], -1))
            state['p3'], state['feat'] = p3_.reshape(-1, J, 3), feat_

        def fn(a, b):
            o = offs[a]
            return model.pose2mesh.forward_parts(x_mine[o:o + b - a], state['p3'][o:o + b - a], state['feat'][o:o + b - a])

        def gathered_step():
            lift()
            forward_gathered(fn, total, (6890, 3), GATHER_BLOCK, device=dev, out=full)
        with torch.no_grad():
            for _ in range(2):
                gathered_step()
            barrier()
            gevs = []
            for _ in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gathered_step()
                e1.record()
                gevs.append((e0, e1))
            barrier()
        g_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in gevs) / args.steps)
        # check on rank 0: the gathered batch equals a local forward of samples the other ranks computed
        ok = True
        probe = [total - 1, total // 2 + 3, GATHER_BLOCK + 1]
        if rank == 0:
            with torch.no_grad():
                ok = all(torch.equal(model(xg[i:i + 1])[0][0], full[i]) for i in probe)
        recv = (world - 1) / world * total * MESH_BYTES
        # same slicing without the collective, to separate the cost of the gather from the cost of running in 1024-sample calls
        def sliced_step():
            lift()
            for a, b in spans:
                if b > a:
                    fn(a, b)
        with torch.no_grad():
            for _ in range(2):
                sliced_step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                sliced_step()
            e1.record()
            barrier()
        s_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        gather = {'value': total / (g_ms * 1e-3), 'unit': UNIT, 'ms_per_step': g_ms, 'efficiency_vs_no_gather': step_ms / g_ms,
                  'ms_per_step_same_slicing_no_gather': s_ms,
                  'bytes_received_per_gpu': int(recv), 'nvlink_gbs_in_per_gpu_over_step': recv / (g_ms * 1e-3) / 1e9,
                  'block_samples': GATHER_BLOCK, 'rounds': -(-total // (world * GATHER_BLOCK)), 'verified': bool(ok),
                  'how': 'block-cyclic deal; the lifter runs once over the rank\'s samples, the decoder per round; all_gather_into_tensor of round k on a side stream while round k+1 computes; '
                         'output (65536, 6890, 3) fp32 written in place, no pad / cat copies'}
        # the same deal without a collective kernel: symmetric-memory output, decoder writes in place, copy engines push each
        # round to the peers over NVLink (gator_b200.dist.P2PGather) - no SMs are taken from the compute kernels
        try:
            from gator_b200.dist import P2PGather, forward_gathered_p2p
            pg = P2PGather(total, (6890, 3), dev)
            plan_p = round_plan(total, world, GATHER_BLOCK, GATHER_MIN_BLOCK)       # tapered: 2048 ... 2048, 1024, 1024
            spans_p = [rank_span(total, st, n, rank) for st, n in plan_p]
            idx_p = torch.cat([torch.arange(a, b) for a, b in spans_p]).to(dev)
            offs_p, acc = {}, 0
            for a, b in spans_p:
                offs_p[a] = acc
                acc += b - a
            x_mine_p = xg[idx_p].contiguous()
            state_p = {}

            def fn_out(a, b, out):
                o = offs_p[a]
                model.pose2mesh.forward_parts(x_mine_p[o:o + b - a], state_p['p3'][o:o + b - a], state_p['feat'][o:o + b - a], out=out)

            def p2p_step():
                p3_, feat_ = model.pose_lifter(x_mine_p.reshape(x_mine_p.shape[0], -1))
                state_p['p3'], state_p['feat'] = p3_.reshape(-1, J, 3), feat_
                forward_gathered_p2p(fn_out, total, GATHER_BLOCK, pg, GATHER_MIN_BLOCK)
            with torch.no_grad():
                for _ in range(2):
                    p2p_step()
                barrier()
                pevs = []
                for _ in range(args.steps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    p2p_step()
                    e1.record()
                    pevs.append((e0, e1))
                barrier()
            p_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in pevs) / args.steps)
            okp = True
            if rank == 0:
                with torch.no_grad():
                    okp = all(torch.equal(model(xg[i:i + 1])[0][0], pg.out[i]) for i in probe)
            try:
                _mc = bool(pg.hdl.multicast_ptr)
            except Exception:
                _mc = None
            gather['p2p'] = {'value': total / (p_ms * 1e-3), 'unit': UNIT, 'ms_per_step': p_ms, 'efficiency_vs_no_gather': step_ms / p_ms,
                             'nvlink_gbs_in_per_gpu_over_step': recv / (p_ms * 1e-3) / 1e9, 'verified': bool(okp),
                             'rounds_per_rank': [n for _, n in plan_p],
                             'multicast_ptr': _mc,
                             'how': 'same block-cyclic deal; the output buffer is symmetric memory (torch.distributed._symmetric_memory), the decoder '
                                    'writes its rows in place and device-to-device copies on side streams (copy engines) push them to every peer '
                                    'while the next round computes (one copy per peer on up to 7 streams); one cross-rank barrier at the end of the step'}
            del pg, x_mine_p, state_p
        except Exception as e:      # symmetric memory unavailable: the NCCL figure above stands alone
            gather['p2p'] = {'error': f'{type(e).__name__}: {e}'[:300]}
        del full, xg, x_mine, state

    # ---- end to end: pinned host input -> H2D -> forward -> mesh + pose3d D2H ----
    # (a) throughput mode of the public host-to-host API: submit() / result(), the D2H of step i (340 MB) runs on the copy
    #     stream while the kernels of step i+1 execute; every step's H2D and D2H - the last one's included - are inside
    #     the timed region.  (b) the synchronous call (one batch in, its mesh out, sliced so that its own copy overlaps
    #     its own kernels) - the latency-oriented figure, reported as e2e_sync.
    from gator_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, B)
    mesh_host, p3_host = pipe.mesh_host, pipe.pose3d_host
    with torch.no_grad():
        def e2e_steps(n):
            prev = None
            for _ in range(n):
                t = pipe.submit(x_host)
                if prev is not None:
                    pipe.result(prev)                # the host "consumes" step i-1 while step i computes
                prev = t
            pipe.result(prev)                        # the last copy is exposed, and counted
        e2e_steps(max(args.warmup, 2))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_steps(args.steps)
        torch.cuda.current_stream().wait_stream(pipe.copy_stream)
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        for _ in range(max(args.warmup, 1)):
            pipe.forward(x_host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            pipe.forward(x_host)
        e1.record()
        barrier()
    e2e_sync_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # ---- end to end with the evaluation epilogue on the device: only per-sample errors come back ----
    from gator_b200.evaluate import EvalEpilogue
    from builders import regressor as _reg
    ep = EvalEpilogue(_reg('h36m'), device=dev)
    gt_mesh = torch.randn(B, 6890, 3, device=dev) * 0.3
    gt_pose = torch.randn(B, 17, 3, device=dev) * 300
    err_host = torch.empty((2, B), dtype=torch.float32).pin_memory()
    with torch.no_grad():
        def eval_step():
            xd = x_host.to(dev, non_blocking=True)
            m_, _ = model(xd)
            r = ep(m_, gt_mesh, gt_pose)
            err_host[0].copy_(r.joint_err, non_blocking=True)
            err_host[1].copy_(r.surface_err, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(max(args.warmup, 1)):
            eval_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            eval_step()
        e1.record()
        barrier()
    ev_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    del gt_mesh, gt_pose

    # ---- batch-1 latency (BASELINE metric: p50 batch-1 latency), eager and under a CUDA graph ----
    latency = None
    if rank == 0 and world == 1:
        x1 = x[:1].clone()
        with torch.no_grad():
            for _ in range(5):
                model(x1)
            torch.cuda.synchronize()

            def p50(fn, iters=300):
                ts = []
                for _ in range(iters):
                    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b_.record(); b_.synchronize()
                    ts.append(a.elapsed_time(b_))
                ts.sort()
                return ts[len(ts) // 2]
            eager = p50(lambda: model(x1))
            graph_ms = None
            try:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    model(x1)
                torch.cuda.current_stream().wait_stream(side)
                with torch.cuda.graph(g):
                    gm, gp = model(x1)
                g.replay(); torch.cuda.synchronize()
                ref_m, _ = model(x1)
                assert torch.equal(gm, ref_m)
                graph_ms = p50(g.replay)
            except Exception as e:   # report, do not hide
                graph_ms = f'capture failed: {type(e).__name__}: {e}'
        latency = {'batch': 1, 'p50_ms_eager': eager, 'p50_ms_cuda_graph': graph_ms, 'iters': 300,
                   'timing': 'CUDA events around each forward, inputs resident'}

    line = None
    if rank == 0:
        peaks = measured_peaks()
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'strong' if world > 1 else 'weak', 'vs_baseline': None,
                'dtype': args.precision, 'data': 'synthetic',
                'config': {'workload': workload_name(world, B), 'global_batch': total, 'per_gpu_batch': B,
                           'l2': 'flushed (256 MiB memset) before every timed step',
                           'precision': 'bf16x3 = 3-term bf16 split GEMMs (fp32-level accuracy on tcgen05), fp16 self-attention core',
                           'parallelism': f'batch-sharded x{world}, no collective on the data path' + (' (strong scaling: the 65 536-sample job of configs[4] at every N > 1; N = 1 runs configs[3])' if world > 1 else '')},
                'clocks': clocks,
                'e2e': {'value': total / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                        'h2d_bytes_per_step': int(x_host.numel() * 4), 'd2h_bytes_per_step': int((mesh_host.numel() + p3_host.numel()) * 4),
                        'mode': 'HostPipeline.submit()/result(): pinned poses -> H2D -> whole-batch forward -> mesh + pose3d D2H on a copy '
                                'stream that overlaps the NEXT step\'s kernels (two pinned output sets); all K steps\' copies, the '
                                'last one included, are inside the timed region'},
                'e2e_sync': {'value': total / (e2e_sync_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_sync_ms,
                             'mode': 'HostPipeline.forward(): one batch in, its mesh out before the call returns; the decoder is '
                                     'sliced so that a slice\'s D2H overlaps the next slice\'s kernels'},
                'e2e_eval': {'value': total / (ev_ms * 1e-3), 'unit': UNIT, 'ms_per_step': ev_ms, 'h2d_bytes_per_step': int(x_host.numel() * 4),
                             'd2h_bytes_per_step': int(err_host.numel() * 4),
                             'what': 'pinned host poses -> forward -> device-side evaluation epilogue (sparse J-regression, root alignment, '
                                     'MPJPE / MPVPE per sample, csrc/eval.cu) -> per-sample errors to pinned host memory; the mesh never leaves the GPU'},
                'gpu_launches': int(launches),
                'tflops_effective': FLOP_PER_MESH * value / 1e12,
                'numa_node': numa_node, 'wall_s_timed_region': t_wall, 'step_ms': [round(t_, 3) for t_ in per_step],
                'whole_step_fraction_of_tensor_peak': FLOP_PER_MESH * value / world / 1e12 / peaks['bf16_tflops_sustained']}
        if gather is not None:
            line['with_gather'] = gather
            line['value_with_gather'] = max(gather['value'], gather.get('p2p', {}).get('value', 0.0))
        if latency is not None:
            line['latency_b1'] = latency
        if world == 1:
            line.update(single_gpu_extras(args, model, x, dev, peaks))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def single_gpu_extras(args, model, x, dev, peaks):
    """Rank 0, N = 1: kernel rooflines, configs[2] (SMPL), parity against the oracle, CPU and eager-GPU baselines."""
    import torch
    from gator_b200 import _lib
    L = _lib.lib()
    J = model.num_joint
    out = {}
    traffic = ncu_traffic()
    nb = 4096                                               # launches of the step's own size
    pcode = _lib.PRECISIONS[args.precision]
    s = _lib.stream_ptr()

    def time_launch(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def tensor_entry(name, ms, flop, key, note=None):
        tf = flop / (ms * 1e-3) / 1e12
        d = {'kernel': name, 'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
             'frac': tf / peaks['bf16_tflops'], 'launch_ms': ms, 'flops_per_launch': flop, 'traffic': traffic.get(key)}
        if note:
            d['note'] = note
        return d

    kernels = []
    qkv = torch.randn(nb * 431, 192, device=dev)
    att = torch.empty(nb * 431, 64, device=dev)
    if pcode == 0:
        ms = time_launch(lambda: _lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), att.data_ptr(), nb, 0, s), 'self_attention'))
        kernels.append(tensor_entry('mdr_self_attn_kernel (fp32 FFMA)', ms, SA_FLOP_PER_SAMPLE * nb, 'mdr_self_attn_kernel'))
    else:
        model(x[:2])                                            # make sure the MDR weights are packed
        (_, _, _, _), table, _ = model.pose2mesh._packed
        img = torch.empty(L.gator_mdr_self_attention_image_bytes(nb), dtype=torch.uint8, device=dev)
        xin = torch.randn(nb * 431, 64, device=dev)
        kvb = torch.randn(nb * J, 128, device=dev)
        x3o = torch.empty(nb * 431, 64, device=dev)
        ch_ms = time_launch(lambda: _lib.check(L.gator_mdr_layer_chain(table, 1, J, pcode, xin.data_ptr(), att.data_ptr(), kvb.data_ptr(),
                                                                        x3o.data_ptr(), None, img.data_ptr(), nb, s), 'layer_chain'))
        ch_flop = nb * 431 * (14 * 64 * 64 * 2 + 2 * J * 32 * 2 * 2)
        kernels.append(tensor_entry('mdr_chain2_kernel<J> (tcgen05 fused layer chain, operands in tensor memory)', ch_ms, ch_flop, 'mdr_chain2_kernel',
                                    'algorithmic 2*MAC of the layer\'s 14 64x64 products + cross-attention; the 3-term bf16 split issues 3x the MMAs'))
        _lib.check(L.gator_mdr_self_attention_f16(qkv.data_ptr(), img.data_ptr(), att.data_ptr(), nb, s), 'qkv image')
        sa_ms = time_launch(lambda: _lib.check(L.gator_mdr_self_attention_core(img.data_ptr(), att.data_ptr(), nb, s), 'self_attention_core'))
        kernels.append(tensor_entry('mdr_self_attn2_kernel (tcgen05 fp16, P in tensor memory, TMA-fed)', sa_ms, SA_FLOP_PER_SAMPLE * nb, 'mdr_self_attn2_kernel',
                                    'bound by the exponentials: 2 x 431 x 431 ex2 per sample = 0.76 G per launch against 16 MUFU results per clock and SM'))
        del img, xin, kvb, x3o
    del qkv, att
    if pcode == 2:
        # upsample_conv's product on the wide tcgen05 + TMA kernel (bias-only epilogue instead of the conv3 scatter)
        from gator_b200.packing import pack_umma_wide
        Mw, Nw, Kw = 3 * 4096, 6890, 1296
        Aw = torch.randn(Mw, Kw, device=dev)
        Ww = pack_umma_wide(torch.randn(Nw, Kw, device=dev) / Kw ** 0.5)
        Cw = torch.empty(Mw, 6892, device=dev)
        wsw = torch.empty(L.gator_umma_wide_a_bytes(Mw, Kw), dtype=torch.uint8, device=dev)
        ga = _lib.GemmArgs(M=Mw, N=Nw, K=Kw, lda=Kw, ldw=Kw, ldc=6892, precision=2, A=_lib.ptr(Aw), W_wide=_lib.ptr(Ww),
                           a_image=_lib.ptr(wsw), a_image_bytes=wsw.numel(), C=_lib.ptr(Cw))
        wg_ms = time_launch(lambda: _lib.check(L.gator_gemm(ga, s), 'gator_gemm wide'))
        wg_flop = 2.0 * Mw * Nw * Kw
        e = tensor_entry('umma_gemm_wide_kernel (upsample_conv shape, 4096 samples; incl. the A-image pre-pass)', wg_ms, wg_flop, 'umma_gemm_wide_kernel')
        e.update({'tensor_flops_issued': 3 * wg_flop, 'frac_issued': 3 * wg_flop / (wg_ms * 1e-3) / 1e12 / peaks['bf16_tflops']})
        kernels.append(e)
        del Aw, Ww, Cw, wsw
        # evaluation epilogue (row f1): HBM-bound, 2 x 82 680 B read per sample
        from gator_b200.evaluate import EvalEpilogue
        from builders import regressor as _reg
        ep = EvalEpilogue(_reg('h36m'), device=dev)
        pm = torch.randn(4096, 6890, 3, device=dev) * 0.3
        gm = pm + 0.02 * torch.randn_like(pm)
        gj = torch.randn(4096, 17, 3, device=dev) * 300
        ev_ms = time_launch(lambda: ep(pm, gm, gj))
        ev_bytes = 4096 * 2 * MESH_BYTES
        kernels.append({'kernel': 'eval_sample_kernel (+ eval_mean_kernel)', 'bound': 'hbm', 'achieved': ev_bytes / (ev_ms * 1e-3) / 1e9,
                        'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': ev_bytes / (ev_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                        'launch_ms': ev_ms, 'bytes_per_launch': ev_bytes, 'traffic': traffic.get('eval_sample_kernel')})
        del pm, gm, gj
    roofline = dict(kernels[0])
    roofline.update({'peak_source': peaks['src'] + ' (cuBLAS bf16 burst)',
                     'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu launch list of a '
                                       '4096-sample forward (profiles/r02_traffic.json); null if not captured',
                     'other_kernels': kernels[1:]})
    out['roofline'] = roofline

    # ---- BASELINE configs[2]: SMPL LBS layer alone, batch 16384 ----
    from builders import build_b200_smpl, synthetic
    Bs = 16384
    pose, betas, trans = [torch.from_numpy(a).to(dev) for a in synthetic.smpl_inputs(Bs)]
    smpl = {}
    layer = build_b200_smpl(device=dev)
    for prec in ('fp32', 'bf16x3'):
        layer.set_precision(prec)
        with torch.no_grad():
            ms = time_launch(lambda: layer(pose, betas, trans), reps=10)
        by = Bs * (MESH_BYTES + 24 * 12 + 72 * 4 + 10 * 4 + 12)
        smpl[prec] = {'value': Bs / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'bound': 'hbm', 'bytes_per_mesh': by // Bs,
                      'achieved': by / (ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'frac': by / (ms * 1e-3) / 1e9 / peaks['hbm_gbs']}
    out['other_workloads'] = {'smpl_layer_b16384': {'workload': 'SMPL_Layer.forward(pose, betas, trans), batch 16384 (BASELINE configs[2]), synthetic '
                                                                'SMPL-shaped buffers; device-resident inputs, CUDA events over 10 calls',
                                                    'roofline': 'HBM: 83.3 KB of mandatory traffic per mesh (SURVEY 8(d))', **smpl}}
    del pose, betas, trans
    # Mesh.upsample 431 -> 1723 -> 6890 (row a21) in one launch, batch 4096; MANO layer (row f4) at batch 16384, fp32 kernels
    try:
        from builders import base_data_root
        from gator_b200.mesh import Mesh
        mesh_op = Mesh(os.path.join(base_data_root(), 'data', 'base_data', 'mesh_downsampling.npz'), device=dev)
        xc = torch.randn(4096, 431, 3, device=dev)
        ms = time_launch(lambda: mesh_op.upsample(xc, n1=2, n2=0), reps=10)
        by = 4096 * (431 + 6890) * 12
        out['other_workloads']['mesh_upsample_b4096'] = {
            'workload': 'Mesh.upsample(x, n1=2, n2=0): 431 -> 1723 -> 6890 vertices, batch 4096, one launch (gator_mesh_upsample2)',
            'ms_per_step': ms, 'value': 4096 / (ms * 1e-3), 'unit': UNIT, 'bound': 'hbm', 'bytes_per_mesh': by // 4096,
            'achieved': by / (ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'frac': by / (ms * 1e-3) / 1e9 / peaks['hbm_gbs']}
        del xc
        from gator_b200.mano_layer import ManoLayer
        hand = ManoLayer(mano_data=synthetic.mano_data(), ncomps=6).to(dev)
        hp, hb, ht = [torch.from_numpy(a).to(dev) for a in synthetic.mano_inputs(Bs, ncomps=6)]
        with torch.no_grad():
            ms = time_launch(lambda: hand(hp, hb, ht), reps=10)
        out['other_workloads']['mano_layer_b16384'] = {
            'workload': 'ManoLayer.forward(pose_coeffs (B,9), betas, trans), batch 16384, synthetic MANO-shaped buffers (778 vertices, '
                        '16 joints), fp32 kernels', 'ms_per_step': ms, 'value': Bs / (ms * 1e-3), 'unit': 'hands/s'}
        del hp, hb, ht
    except Exception as e:      # secondary workloads: report, do not fail the bench
        out['other_workloads']['error'] = f'{type(e).__name__}: {e}'

    # ---- parity of this very configuration against the CPU oracle (64 samples) ----
    from helpers import oracle_setup, orc, regressor
    sd, gc, mc, alpha = oracle_setup(TAG)
    xs = torch.from_numpy(make_inputs(64, seed=7))
    with torch.no_grad():
        ref_mesh, _ = orc.gator_forward(sd, gc, mc, xs, alpha)
        got, _ = model(xs.to(dev))
    mp, pa = orc.mpjpe_pa(got.cpu().numpy(), ref_mesh.numpy(), regressor('h36m'))
    out['parity'] = {'max_abs_vertex_err_m': (got.cpu() - ref_mesh).abs().max().item(), 'mpjpe_drift_mm': mp,
                     'pa_mpjpe_drift_mm': pa, 'samples': 64, 'tolerance_m': 1e-4}
    try:
        out['gpu_eager_baseline'] = gpu_eager_rate(dev)
    except Exception as e:      # informational leg: report, do not fail the bench
        out['gpu_eager_baseline'] = {'error': f'{type(e).__name__}: {e}'}
    threads = os.cpu_count() or 1
    cpu_v, cpu_med, cpu_n = cpu_forward_rate(64, 12.0, threads)
    out['cpu_baseline'] = {'value': cpu_v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                           'sample': f'{cpu_n} forwards of batch 64 (median {cpu_med * 1e3:.0f} ms), oracle port of the reference forward'}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=0, help='samples per GPU (default: 4096 on one GPU, 65536 / N on N)')
    ap.add_argument('--precision', default='bf16x3', choices=['fp32', 'bf16', 'bf16x3'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
