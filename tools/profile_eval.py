"""Evaluation epilogue and 2D-pose pre-processing timed alone (CUDA events, L2 flushed between iterations).
Prints one JSON line per kernel with the achieved HBM rate: epilogue algorithmic bytes = 2 x 82 680 B per sample."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from builders import golden, regressor
from gator_b200.evaluate import EvalEpilogue
from gator_b200.preprocess import COCO_MID_PAIRS, Pose2DPreprocessor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
peaks = {}
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peaks = json.load(open(p))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20, reps=1):
    """median over `iters` of (time of `reps` back-to-back calls) / reps; reps > 1 hides the host-side launch
    path of a sub-millisecond call (inputs are far larger than L2, so every call still streams from HBM)."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return float(np.median(ts))


ep = EvalEpilogue(regressor('h36m'), device=dev)
pred = torch.randn(B, 6890, 3, device=dev) * 0.3
gt = pred + 0.02 * torch.randn_like(pred)
gtj = torch.randn(B, 17, 3, device=dev) * 300
for pa in (False, True):
    ms = timed(lambda: ep(pred, gt, gtj, pa=pa), reps=8)
    by = B * 2 * 82680
    print(json.dumps({'kernel': 'eval_sample_kernel', 'pa': pa, 'batch': B, 'ms': ms, 'meshes_per_s': B / ms * 1e3,
                      'achieved_GBps': by / ms / 1e6, 'algorithmic_bytes': by, 'peak_GBps': peaks.get('hbm_gbs')}))
# what the reference does instead on the device side (dense regression) + the D2H it needs for numpy
regd = torch.from_numpy(regressor('h36m')).to(dev)


def ref_like():
    pm, gm = pred * 1000, gt * 1000
    pp = torch.matmul(regd[None], pm)
    return pm.cpu(), gm.cpu(), pp.cpu()


ms = timed(ref_like, iters=3)
print(json.dumps({'kernel': 'reference-style (x1000, dense matmul, D2H of both meshes; numpy metrics not included)',
                  'batch': B, 'ms': ms}))
base = golden('fixtures')['coco_joint_input'].reshape(17, -1).astype(np.float32)
x = torch.from_numpy(np.repeat(base[None], B, 0)).to(dev) + torch.randn(B, 17, 3, device=dev)
pre = Pose2DPreprocessor((384, 288), COCO_MID_PAIRS)
ms = timed(lambda: pre(x))
print(json.dumps({'kernel': 'pose2d_preprocess_kernel', 'batch': B, 'ms': ms, 'poses_per_s': B / ms * 1e3}))
x1 = x[:1].contiguous()
ms = timed(lambda: pre(x1), iters=50)
print(json.dumps({'kernel': 'pose2d_preprocess_kernel', 'batch': 1, 'ms': ms}))

# ground-truth mesh generation (row f2): batched get_smpl_coord, camera fix-up + SMPL forward in mm
from builders import build_b200_smpl, synthetic
from gator_b200.gt_mesh import GtMeshGenerator
Bg = 16384
gen = GtMeshGenerator(build_b200_smpl(device=dev).set_precision('bf16x3'))
ann = [torch.from_numpy(a).to(dev) for a in synthetic.camera_annotations(Bg)]
ms = timed(lambda: gen.h36m(*ann), iters=10)
print(json.dumps({'kernel': 'GtMeshGenerator.h36m (smpl_cam_fixup + SMPL forward bf16x3, mm out)', 'batch': Bg, 'ms': ms,
                  'meshes_per_s': Bg / ms * 1e3, 'achieved_GBps_out': Bg * 82680 / ms / 1e6}))
ms = timed(lambda: gen.layer.forward(ann[0], ann[1], ann[2]), iters=10)
print(json.dumps({'kernel': 'SMPL_Layer.forward alone (same batch, metres)', 'batch': Bg, 'ms': ms}))
