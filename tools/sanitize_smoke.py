"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from builders import build_b200_gator, build_b200_smpl, golden, regressor, synthetic
from gator_b200.evaluate import EvalEpilogue
from gator_b200.gt_mesh import GtMeshGenerator
from gator_b200.preprocess import COCO_MID_PAIRS, Pose2DPreprocessor

dev = 'cuda:0'
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'gator'):
    m = build_b200_gator('coco', dev)
    x = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], 13)).to(dev)
    for prec in ('fp32', 'bf16x3'):
        m.set_precision(prec)
        mesh, p3 = m(x)
        torch.cuda.synchronize()
        print(prec, float(mesh.abs().max()))
if which in ('all', 'smpl'):
    layer = build_b200_smpl(device=dev).set_precision('bf16x3')
    ann = [torch.from_numpy(a).to(dev) for a in synthetic.camera_annotations(9)]
    v, j = GtMeshGenerator(layer).h36m(*ann)
    torch.cuda.synchronize()
    print('gtmesh', float(v.abs().max()))
if which in ('all', 'eval'):
    reg = regressor('h36m')
    pred, gt, gtj = [torch.from_numpy(a).to(dev) for a in synthetic.eval_inputs(7, reg)]
    r = EvalEpilogue(reg, device=dev)(pred, gt, gtj, pa=True)
    pose = Pose2DPreprocessor((384, 288), COCO_MID_PAIRS)(torch.from_numpy(golden('preproc')['input'].astype(np.float32)).to(dev))
    torch.cuda.synchronize()
    print('eval', float(r.joint_error), float(r.pa_joint_error), float(pose.abs().max()))
if which in ('all', 'mesh'):
    from builders import base_data_root
    from gator_b200.mesh import Mesh
    mesh_op = Mesh(os.path.join(base_data_root(), 'data', 'base_data', 'mesh_downsampling.npz'), device=torch.device(dev))
    xc = torch.randn(7, 431, 3, device=dev)
    up = mesh_op.upsample(xc, n1=2, n2=0)          # fused two-level kernel, ragged last group
    torch.cuda.synchronize()
    print('upsample2', float(up.abs().max()))
if which in ('all', 'mano'):
    from gator_b200.mano_layer import ManoLayer
    hand = ManoLayer(mano_data=synthetic.mano_data(), ncomps=6, center_idx=8).to(dev)
    hp, hb, ht = [torch.from_numpy(a).to(dev) for a in synthetic.mano_inputs(21)]
    v, j = hand(hp, hb, torch.zeros_like(ht))
    torch.cuda.synchronize()
    print('mano', float(v.abs().max()), float(j.abs().max()))
