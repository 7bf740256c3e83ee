"""Golden vectors of the evaluation epilogue, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_eval.py

The inputs are regenerated from a seed (``gator_b200.synthetic.eval_inputs``), so only the reference's outputs are
stored: ``J_regressor[None] @ (pred_mesh*1000)`` as base.py:219-221 computes it, ``compute_both_err`` of both
datasets called unbound on a stand-in ``self``, the totals ``evaluate_joint`` prints, and per-sample PA-MPJPE through
the reference's own ``rigid_align``.
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from gator_b200 import synthetic  # noqa: E402
from oracle import refshim        # noqa: E402

EVAL_JOINTS = (1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16)       # data/Human36M/dataset.py:62
B = 6


def main():
    reg = np.load(os.path.join(refshim.REF, 'data', 'Human36M', 'J_regressor_h36m_correct.npy')).astype(np.float32)
    pred, gt, gt_pose = synthetic.eval_inputs(B, reg)
    H36M = refshim.dataset_class('Human36M')
    PW3D = refshim.dataset_class('PW3D')
    from coord_utils import rigid_align
    me = types.SimpleNamespace(human36_eval_joint=EVAL_JOINTS)
    out = {}
    with torch.no_grad():
        # lib/core/base.py:219-223 verbatim (J_regressor as torch.Tensor(...), base.py:194)
        J_regressor = torch.Tensor(reg)
        pred_mesh, gt_mesh = torch.from_numpy(pred) * 1000, torch.from_numpy(gt) * 1000
        gt_pose3d = torch.from_numpy(gt_pose)
        pred_pose = torch.matmul(J_regressor[None, :, :], pred_mesh)
        j_error, s_error = H36M.compute_both_err(me, pred_mesh, gt_mesh, pred_pose, gt_pose3d)
        j2, s2 = PW3D.compute_both_err(me, pred_mesh, gt_mesh, pred_pose, gt_pose3d)
        assert j2 == j_error and s2 == s_error
        per_j, per_s = [], []
        for b in range(B):                                   # per-sample values: the same method at batch 1
            jb, sb = H36M.compute_both_err(me, pred_mesh[b:b + 1], gt_mesh[b:b + 1], pred_pose[b:b + 1], gt_pose3d[b:b + 1])
            per_j.append(jb)
            per_s.append(sb)
    out['pred_pose'] = pred_pose.numpy()
    out['joint_error'], out['surface_error'] = np.float64(j_error), np.float64(s_error)
    out['joint_err'], out['surface_err'] = np.asarray(per_j, np.float64), np.asarray(per_s, np.float64)

    # evaluate_joint (data/Human36M/dataset.py:480-504) only prints its totals: capture them, and take the
    # per-sample PA values through the reference's rigid_align with the same pre-processing.
    me.datalist = [{'joint_cam': gt_pose[b]} for b in range(B)]
    outs = [{'joint_coord': out['pred_pose'][b]} for b in range(B)]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        H36M.evaluate_joint(me, outs)
    tot = [float(x) for x in re.findall(r'tot: ([0-9.]+)', buf.getvalue())]
    out['evaluate_joint_printed'] = np.asarray(tot, np.float64)             # [MPJPE, PA-MPJPE] to 2 decimals
    pa = []
    ev = list(EVAL_JOINTS)
    for b in range(B):
        o, g = out['pred_pose'][b] - out['pred_pose'][b][:1], gt_pose[b] - gt_pose[b][:1]
        o, g = o[ev, :], g[ev, :]
        pa.append(np.sqrt(np.sum((rigid_align(o, g) - g) ** 2, 1)).mean())
    out['pa_joint_err'] = np.asarray(pa, np.float64)
    # a reflected configuration: rigid_align's det(R) < 0 branch (lib/coord_utils.py:135-138)
    o = out['pred_pose'][0][ev, :].astype(np.float64)
    g = o * np.array([1.0, 1.0, -1.0]) + 3.0
    out['reflect_A'], out['reflect_B'] = o, g
    out['reflect_aligned'] = rigid_align(o, g)
    np.savez_compressed(os.path.join(HERE, 'eval.npz'), **out)
    print({k: (v.shape if v.ndim else float(v)) for k, v in out.items()})
    print('printed totals', tot, 'means', out['joint_err'].mean(), out['pa_joint_err'].mean())


if __name__ == '__main__':
    main()
