#!/usr/bin/env python
"""Headline benchmark: meshes/sec of the GATOR pose->mesh forward (BASELINE.json configs[3]: full
GAT+MDR+upsample forward, synthetic COCO poses, J=19, alpha=True, batch 4096 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--precision fp32|bf16]

One JSON line on stdout (rank 0).  `value` = whole-job meshes/s with inputs resident in HBM;
`e2e` = the same through the nn.Module API with pinned-host inputs and the mesh copied back to the host
inside the timed region; `roofline` = the dominant kernel (MDR 431x431 self-attention) timed on its own
with CUDA events; `cpu_baseline` = the CPU oracle (port of the reference forward, same torch ops) timed
on the box's host cores on a bounded sample.  --impl reference times that CPU path as the reference arm.
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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p['bf16_tflops']),
                'bf16_tflops_sustained': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), 'src': 'measured'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'src': 'fallback'}


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
    import numpy as np
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


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sample = 64
    # one "step" = one forward over a 64-sample slice of the 4096-sample workload (bounded sample)
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
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': f'GATOR.forward J=19 alpha=True, batch {args.batch}/GPU (BASELINE configs[3])',
                       'note': 'CPU reference forward (oracle port, same ATen ops); each step = a 64-sample slice'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': f'{args.steps} forwards of batch {sample}'},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


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
    from gator_b200.dist import bind_to_gpu_numa_node, shard_range
    numa_node = bind_to_gpu_numa_node(local)              # before any pinned buffer is allocated
    from builders import build_b200_gator
    L = _lib.lib()
    model = build_b200_gator(TAG, dev).set_precision(args.precision)
    J = model.num_joint
    B = args.batch                                         # per GPU (weak scaling)
    lo, hi = shard_range(B * world, world, rank)
    x_host = torch.from_numpy(make_inputs(B * world)[lo:hi]).pin_memory()
    x = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    # nvidia-smi is started well before the timed region: its start-up (NVML initialisation) was seen to stall the GPU for
    # tens of milliseconds when it coincided with the first timed steps (one run measured 41 ms per step instead of 20)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        t_ramp = time.perf_counter()                       # untimed: first call packs the weights, then ~0.8 s of
        while True:                                        # forwards so that the SM clock has ramped before the W
            mesh_w = model(x)                              # warm-up steps the contract asks for
            torch.cuda.synchronize()
            if time.perf_counter() - t_ramp > 0.8:
                break
        mesh = p3 = None
        for _ in range(args.warmup):
            mesh, p3 = model(x)                            # same binding pattern as the timed loop: the previous step's
        barrier()                                          # outputs are still alive while the next ones are allocated, so
        L.gator_launch_count(1)                            # the caching allocator's second 340 MB block exists before timing
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
    step_ms = sum(per_step) / args.steps
    t = torch.tensor([step_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = t.item()
    value = B * world / (step_ms * 1e-3)

    # ---- end to end: pinned host input -> H2D -> forward -> mesh + pose3d D2H ----
    from gator_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, B)                    # public host-to-host API: sliced decoder, D2H overlapped
    mesh_host, p3_host = pipe.mesh_host, pipe.pose3d_host
    with torch.no_grad():
        def e2e_step():
            pipe.forward(x_host)
        for _ in range(max(args.warmup, 1)):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()

    # ---- batch-1 latency (BASELINE metric: p50 batch-1 latency), eager and under a CUDA graph ----
    latency = None
    if rank == 0:
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
        # ---- dominant kernels alone (C-ABI entry points, CUDA events on the launching stream) ----
        nb = 1184                                               # 8 x 148 samples: 3987 chain tiles, 2368 attention CTAs
        pcode = _lib.PRECISIONS[args.precision]
        s = _lib.stream_ptr()

        def time_launch(fn, reps=20):
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

        qkv = torch.randn(nb * 431, 192, device=dev)
        out = torch.empty(nb * 431, 64, device=dev)
        sa_ms = time_launch(lambda: _lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), out.data_ptr(), nb, pcode, s), 'self_attention'))
        sa_tf = SA_FLOP_PER_SAMPLE * nb / (sa_ms * 1e-3) / 1e12
        sa_name = 'mdr_self_attn_kernel (fp32 FFMA)' if pcode == 0 else f'mdr_self_attn_umma_kernel<{"true" if pcode == 2 else "false"}> (tcgen05)'
        kernels = [{'kernel': sa_name, 'bound': 'tensor', 'achieved': sa_tf, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                    'frac': sa_tf / peaks['bf16_tflops'], 'launch_ms': sa_ms,
                    'flops_per_launch': SA_FLOP_PER_SAMPLE * nb}]
        if pcode != 0:
            model(x[:2])                                            # make sure the MDR weights are packed
            (_, _, _, _), table, _ = model.pose2mesh._packed
            xin = torch.randn(nb * 431, 64, device=dev)
            att = torch.randn(nb * 431, 64, device=dev)
            kvb = torch.randn(nb * J, 128, device=dev)
            x3o = torch.empty(nb * 431, 64, device=dev)
            qko = torch.empty(nb * 431, 192, device=dev)
            ch_ms = time_launch(lambda: _lib.check(L.gator_mdr_layer_chain(table, 1, J, pcode, xin.data_ptr(), att.data_ptr(), kvb.data_ptr(),
                                                                            x3o.data_ptr(), qko.data_ptr(), nb, s), 'layer_chain'))
            ch_flop = nb * 431 * (14 * 64 * 64 * 2 + 2 * J * 32 * 2 * 2)
            ch_tf = ch_flop / (ch_ms * 1e-3) / 1e12
            kernels.insert(0, {'kernel': 'mdr_chain_kernel<1,J> (tcgen05 fused layer chain)', 'bound': 'tensor', 'achieved': ch_tf,
                               'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': ch_tf / peaks['bf16_tflops'],
                               'launch_ms': ch_ms, 'flops_per_launch': ch_flop})
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
            kernels.append({'kernel': 'umma_gemm_wide_kernel (upsample_conv shape, 4096 samples; incl. the A-image pre-pass)',
                            'bound': 'tensor', 'achieved': wg_flop / (wg_ms * 1e-3) / 1e12, 'peak': peaks['bf16_tflops'],
                            'unit': 'TFLOP/s', 'frac': wg_flop / (wg_ms * 1e-3) / 1e12 / peaks['bf16_tflops'], 'launch_ms': wg_ms,
                            'flops_per_launch': wg_flop, 'tensor_flops_issued': 3 * wg_flop,
                            'frac_issued': 3 * wg_flop / (wg_ms * 1e-3) / 1e12 / peaks['bf16_tflops']})
            del Aw, Ww, Cw, wsw
            # evaluation epilogue (row f1): HBM-bound, 2 x 82 680 B read per sample
            from gator_b200.evaluate import EvalEpilogue
            from builders import regressor as _reg
            ep = EvalEpilogue(_reg('h36m'), device=dev)
            pm = torch.randn(4096, 6890, 3, device=dev) * 0.3
            gm = pm + 0.02 * torch.randn_like(pm)
            gj = torch.randn(4096, 17, 3, device=dev) * 300
            ev_ms = time_launch(lambda: ep(pm, gm, gj))
            ev_bytes = 4096 * 2 * 82680
            kernels.append({'kernel': 'eval_sample_kernel (+ eval_mean_kernel)', 'bound': 'hbm', 'achieved': ev_bytes / (ev_ms * 1e-3) / 1e9,
                            'peak': peaks.get('hbm_gbs'), 'unit': 'GB/s',
                            'frac': (ev_bytes / (ev_ms * 1e-3) / 1e9 / peaks['hbm_gbs']) if peaks.get('hbm_gbs') else None,
                            'launch_ms': ev_ms, 'bytes_per_launch': ev_bytes, 'traffic': 686.8e6})
            del pm, gm, gj
        roofline = dict(kernels[0])
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu launch list of a 4096-sample forward
        # (profiles/r01_launches_bf16x3_b4096_flat.csv, layers 1-2: 944.5 MB read + 1750 MB written per 4096 samples), scaled to
        # this launch's 1184 samples; ncu flushes caches between kernels, so this is an upper bound of a warm launch
        roofline.update({'traffic': 779e6 if pcode != 0 else None, 'peak_source': peaks['src'] + ' (cuBLAS bf16 burst)',
                         'note': 'algorithmic flops (2*MAC of the layer\'s 14 64x64 products + cross-attention, or of QK^T + PV) per launch of '
                                 '1184 samples; the 3-term split issues 3x that many tensor-core MACs, which are not counted; both kernels are '
                                 'bound by CUDA-core softmax/GELU/LayerNorm/operand-conversion work and MMA round-trip latency, not by the tensor pipe',
                         'other_kernels': kernels[1:]})
        # parity of this very configuration against the CPU oracle (64 samples)
        from helpers import oracle_setup, orc, regressor
        sd, gc, mc, alpha = oracle_setup(TAG)
        xs = torch.from_numpy(make_inputs(64, seed=7))
        with torch.no_grad():
            ref_mesh, _ = orc.gator_forward(sd, gc, mc, xs, alpha)
            got, _ = model(xs.to(dev))
        mp, pa = orc.mpjpe_pa(got.cpu().numpy(), ref_mesh.numpy(), regressor('h36m'))
        parity = {'max_abs_vertex_err_m': (got.cpu() - ref_mesh).abs().max().item(), 'mpjpe_drift_mm': mp,
                  'pa_mpjpe_drift_mm': pa, 'samples': 64, 'tolerance_m': 1e-4}
        threads = os.cpu_count() or 1
        cpu_v, cpu_med, cpu_n = cpu_forward_rate(64, 12.0, threads)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': args.precision, 'data': 'synthetic',
                'config': {'workload': f'GATOR.forward J=19 alpha=True, batch {B}/GPU (BASELINE configs[3]); '
                                       'random-init weights, synthetic SMPL-shaped template/bases',
                           'global_batch': B * world, 'l2': 'flushed (256 MiB memset) before every timed step',
                           'parallelism': f'batch-sharded x{world}, no collective on the data path'},
                'clocks': clocks,
                'e2e': {'value': B * world / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                        'h2d_bytes_per_step': int(x_host.numel() * 4), 'd2h_bytes_per_step': int((mesh_host.numel() + p3_host.numel()) * 4)},
                'gpu_launches': int(launches),
                'tflops_effective': FLOP_PER_MESH * value / 1e12,
                'numa_node': numa_node, 'wall_s_timed_region': t_wall, 'step_ms': [round(t_, 3) for t_ in per_step],
                'latency_b1': latency,
                'roofline': roofline, 'parity': parity,
                'cpu_baseline': {'value': cpu_v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                 'sample': f'{cpu_n} forwards of batch 64 (median {cpu_med * 1e3:.0f} ms), oracle port of the reference forward'}}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=4096, help='samples per GPU')
    ap.add_argument('--precision', default='bf16x3', choices=['fp32', 'bf16', 'bf16x3'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
