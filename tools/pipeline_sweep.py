"""HostPipeline slice-plan sweep at B = 4096 (J = 19): end-to-end ms per step for geometric ratios / floors / equal slices,
the D2H copy alone, and the decoder alone per slice size.  python tools/pipeline_sweep.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_gator, synthetic
from gator_b200.pipeline import HostPipeline, plan_slices
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
model = build_b200_gator('coco', dev).set_precision('bf16x3')
J = model.num_joint
base19 = synthetic.poses2d(1, J)[0]
x_host = torch.from_numpy(synthetic.coco_poses2d(base19, B)).pin_memory() if J == 19 else torch.from_numpy(synthetic.poses2d(B, J)).pin_memory()


def timed(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    x = x_host.to(dev)
    print(f'device-resident forward: {timed(lambda: model(x)):.2f} ms', flush=True)
    pipe = HostPipeline(model, B)
    mesh = torch.empty((B, 6890, 3), device=dev)
    print(f'D2H of the mesh alone ({mesh.numel() * 4 / 1e6:.0f} MB): {timed(lambda: pipe.mesh_host.copy_(mesh, non_blocking=True)):.2f} ms', flush=True)
    p3, feat = model.pose_lifter(x.reshape(B, J * 2))
    p3 = p3.reshape(B, J, 3)
    print(f'lifter alone: {timed(lambda: model.pose_lifter(x.reshape(B, J * 2))):.2f} ms', flush=True)
    for n in (148, 296, 512, 1024, 2048, 4096):
        t = timed(lambda: model.pose2mesh.forward_parts(x[:n], p3[:n], feat[:n]))
        print(f'decoder alone, slice of {n}: {t:.3f} ms = {t / n * 1e3:.2f} us per mesh', flush=True)
    for ratio in (0.35, 0.5, 0.6, 0.7, 0.8, 0.9):
        for floor in (296, 592):
            pipe._bounds = plan_slices(B, ratio, floor)
            print(f'geometric ratio {ratio} floor {floor} ({len(pipe._bounds) - 1} slices): {timed(lambda: pipe.forward(x_host)):.2f} ms', flush=True)
    for n in (512, 1024, 2048):
        pipe.slice_samples = n
        print(f'equal slices of {n}: {timed(lambda: pipe.forward(x_host)):.2f} ms', flush=True)
