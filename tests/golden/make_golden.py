"""Generate the committed golden vectors by running the UNMODIFIED reference (CPU, fp32).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Inputs are the deterministic synthetic base_data / weights of ``gator_b200.synthetic`` plus the three
fixtures the reference does ship (demo/coco_joint_input.npy and the two 17x6890 J-regressors, stored
sparsely).  Outputs of the reference modules - final and inter-stage - are stored so that both the
CPU oracle (``oracle/gator_oracle.py``) and the CUDA path can be checked on a box without the reference.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from gator_b200 import synthetic  # noqa: E402
from oracle import refshim        # noqa: E402

REF = refshim.REF


def sparse_rows(a):
    r, c = np.nonzero(a)
    return r.astype(np.int32), c.astype(np.int32), a[r, c].astype(np.float64)


def demo_pose19():
    """demo/run.py:103-134,193-198: coco_joint_input -> +pelvis,+neck -> bbox/affine/normalise/standardise."""
    cfg = refshim.install_shims()
    sys.path.insert(0, os.path.join(REF, 'demo'))
    from coord_utils import get_bbox, process_bbox
    from aug_utils import j2d_processing
    names = synthetic.COCO_JOINTS_NAME
    j = np.load(os.path.join(REF, 'demo', 'coco_joint_input.npy')).reshape(17, -1)

    def add_mid(jc, a, b):
        ia, ib = names.index(a), names.index(b)
        m = (jc[ia, :] + jc[ib, :]) * 0.5
        m[2] = jc[ia, 2] * jc[ib, 2]
        return np.concatenate((jc, m.reshape(1, 3)))
    j = add_mid(j, 'L_Hip', 'R_Hip')
    j = add_mid(j, 'L_Shoulder', 'R_Shoulder')
    j = j[:, :2]
    bbox = get_bbox(j)
    bbox2 = process_bbox(bbox.copy())
    ji, _ = j2d_processing(j.copy(), (cfg.MODEL.input_shape[1], cfg.MODEL.input_shape[0]), bbox2, 0, 0, None)
    ji = ji[:, :2]
    ji /= np.array([[cfg.MODEL.input_shape[1], cfg.MODEL.input_shape[0]]])
    mean, std = np.mean(ji, axis=0), np.std(ji, axis=0)
    return ((ji.copy() - mean) / std).astype(np.float32)


def run_gator(root, category, alpha, regressor, x, tag, out):
    model = refshim.build_gator(root, category, alpha, regressor)
    synthetic.load_synth_weights(model)
    trace = {}
    hooks = []
    pl, pm = model.pose_lifter, model.pose2mesh
    for i, blk in enumerate(pl.blocks):
        hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: trace.__setitem__(f'gat_block{i}', o[0].detach().clone())))
    hooks.append(pl.get_hop_path_encoding.register_forward_hook(lambda m, a, o: trace.__setitem__('hop_path_bias', o.detach().clone())))
    for li, sfx in enumerate(('', '_1', '_2')):
        hooks.append(getattr(pm, 'encoder' + sfx).register_forward_hook(
            lambda m, a, o, li=li: trace.__setitem__(f'mdr_cross{li}', o.detach().clone())))
    orig_conv = pm.upsample_conv.forward
    hooks.append(pm.upsample_conv.register_forward_hook(lambda m, a, o: trace.__setitem__('mdr_coarse', a[0].detach().clone())))
    hooks.append(pm.motion_linear.register_forward_hook(lambda m, a, o: trace.__setitem__('mdr_layer2', a[0].detach().clone())))
    with torch.no_grad():
        mesh, pose3d = model(torch.from_numpy(x))
        feat = pl(torch.from_numpy(x).view(len(x), -1))[1]
    for h in hooks:
        h.remove()
    sd = model.state_dict()
    out[f'{tag}/keys'] = np.array([f'{k}|{",".join(map(str, v.shape))}|{str(v.dtype).replace("torch.", "")}' for k, v in sd.items()])
    out[f'{tag}/pose2d'] = x
    out[f'{tag}/mesh'] = mesh.numpy()
    out[f'{tag}/pose3d'] = pose3d.numpy()
    out[f'{tag}/feat'] = feat.numpy()
    out[f'{tag}/vj_relation'] = np.asarray(pm.vj_relation).astype(np.int64)
    out[f'{tag}/graph_adj'] = sd['pose_lifter.graph_adj'].numpy()
    out[f'{tag}/init_vertices_431'] = sd['pose2mesh.init_vertices'].numpy()
    out[f'{tag}/edge_input'] = pl.get_hop_path_encoding.edg_adj.numpy()
    for k, v in trace.items():
        v = v.numpy()
        if k.startswith('mdr_cross') or k.startswith('mdr_layer'):
            v = v[:1]                        # sample 0 only (size)
        out[f'{tag}/trace/{k}'] = v
    # the post-step every caller runs (base.py:221, run.py:142)
    out[f'{tag}/joints'] = torch.matmul(torch.from_numpy(regressor)[None], mesh).numpy()
    return model


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    root = tempfile.mkdtemp(prefix='gator_golden_')
    reg_h36m = np.load(os.path.join(REF, 'data', 'Human36M', 'J_regressor_h36m_correct.npy'))
    reg_coco = np.load(os.path.join(REF, 'data', 'COCO', 'J_regressor_coco.npy'))
    fx = {}
    for name, a in (('h36m', reg_h36m), ('coco', reg_coco)):
        r, c, v = sparse_rows(a)
        fx[f'J_regressor_{name}/row'], fx[f'J_regressor_{name}/col'], fx[f'J_regressor_{name}/val'] = r, c, v
    fx['coco_joint_input'] = np.load(os.path.join(REF, 'demo', 'coco_joint_input.npy'))
    fx['demo_pose19'] = demo_pose19()
    np.savez_compressed(os.path.join(HERE, 'fixtures.npz'), **fx)

    reg_h36m32, reg_coco32 = reg_h36m.astype(np.float32), reg_coco.astype(np.float32)
    synthetic.write_base_data(root, reg_h36m32)

    out = {}
    run_gator(root, 'human36', False, reg_h36m32, synthetic.poses2d(3, 17), 'h36m', out)
    x19 = np.concatenate([fx['demo_pose19'][None], synthetic.coco_poses2d(fx['demo_pose19'], 2)], 0)
    run_gator(root, 'coco', True, reg_coco32, x19, 'coco', out)
    np.savez_compressed(os.path.join(HERE, 'gator.npz'), **out)

    # ---- SMPL_Layer (smpl_layer.py:65-158) ----
    buf = synthetic.smpl_buffers()
    pose, betas, trans = synthetic.smpl_inputs(4)
    s = {}
    with torch.no_grad():
        layer = refshim.build_smpl_layer(buf)
        v, j = layer(torch.from_numpy(pose), torch.from_numpy(betas), torch.from_numpy(trans))
        s['full/verts'], s['full/jtr'] = v.numpy(), j.numpy()
        v, j = layer(torch.from_numpy(pose))                       # default betas/trans = zeros(1)
        s['nobetas/verts'], s['nobetas/jtr'] = v.numpy()[:2], j.numpy()
        layer_c = refshim.build_smpl_layer(buf, center_idx=0)
        v, j = layer_c(torch.from_numpy(pose), torch.from_numpy(betas))
        s['center/verts'], s['center/jtr'] = v.numpy()[:2], j.numpy()
    np.savez_compressed(os.path.join(HERE, 'smpl.npz'), **s)

    # ---- Mesh.downsample / upsample (mesh.py:93-123) ----
    m = refshim.ref_mesh(root)
    r = np.random.Generator(np.random.PCG64(5))
    x = r.standard_normal((2, 6890, 3)).astype(np.float32)
    g = {'x': x}
    d1 = m.downsample(torch.from_numpy(x))
    d2 = m.downsample(d1, n1=1, n2=2)
    g['down1'], g['down2'] = d1.numpy(), d2.numpy()
    u1 = m.upsample(d2, n1=2, n2=1)
    u0 = m.upsample(u1, n1=1, n2=0)
    g['up1'], g['up0'] = u1.numpy(), u0.numpy()
    g['down2d'] = m.downsample(torch.from_numpy(x[0]), n1=0, n2=2).numpy()
    np.savez_compressed(os.path.join(HERE, 'mesh.npz'), **g)
    for f in ('fixtures', 'gator', 'smpl', 'mesh'):
        print(f, os.path.getsize(os.path.join(HERE, f + '.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
