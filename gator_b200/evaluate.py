"""Device-side evaluation epilogue.

Mirrors what every caller of the forward does with its output:

  * ``Tester.test`` (lib/core/base.py:219-223): ``pred_mesh*1000``, ``gt_mesh*1000``,
    ``pred_pose = J_regressor[None] @ pred_mesh``, ``val_dataset.compute_both_err(...)``;
  * ``compute_both_err`` (data/Human36M/dataset.py:466-478, data/PW3D/dataset.py:273-286): root-align on joint
    0, mean L2 over the vertices (MPVPE) and over the H36M evaluation joints (MPJPE), after copying both meshes
    to the host;
  * ``evaluate_joint`` (data/Human36M/dataset.py:480-504): per-sample MPJPE and PA-MPJPE with ``rigid_align``
    (lib/coord_utils.py:127-149).

Here one kernel launch (csrc/eval.cu) reads the two meshes once and leaves only (B,)-sized results, so neither the
dense 17x6890 regression nor a device->host copy of a mesh remains on the evaluation path.  CUDA only.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .mesh import _Csr

H36M_EVAL_JOINTS = (1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16)   # data/Human36M/dataset.py:75


class EvalResult(NamedTuple):
    joint_error: torch.Tensor       # () batch mean of joint_err            (MPJPE, mm)
    surface_error: torch.Tensor     # () batch mean of surface_err          (MPVPE, mm); 0 without gt_mesh
    pa_joint_error: torch.Tensor    # () batch mean of pa_joint_err         (PA-MPJPE, mm); 0 unless requested
    pred_pose: torch.Tensor         # (B, Jt, 3)  J_regressor @ (pred_mesh * scale)
    joint_err: torch.Tensor         # (B,)
    surface_err: Optional[torch.Tensor]
    pa_joint_err: Optional[torch.Tensor]


class EvalEpilogue:
    """``epilogue(pred_mesh_m, gt_mesh_m, gt_pose3d_mm)`` -> :class:`EvalResult`, everything left on the device."""

    def __init__(self, J_regressor, eval_joints: Sequence[int] = H36M_EVAL_JOINTS, root: int = 0,
                 scale: float = 1000.0, device='cuda'):
        a = J_regressor.detach().cpu().numpy() if torch.is_tensor(J_regressor) else np.asarray(J_regressor)
        import scipy.sparse
        self.device = torch.device(device)
        self.num_joints = int(a.shape[0])
        self.num_verts = int(a.shape[1])
        if self.num_joints > 32:
            raise ValueError('EvalEpilogue: at most 32 regressed joints')
        self._csr = _Csr(scipy.sparse.csr_matrix(a.astype(np.float32)), self.device)
        ev = [int(j) for j in eval_joints]
        if not ev or min(ev) < 0 or max(ev) >= self.num_joints:
            raise ValueError('EvalEpilogue: evaluation joints out of range')
        self.eval_joints = torch.tensor(ev, dtype=torch.int32, device=self.device)
        self.root = int(root)
        self.scale = float(scale)

    def _check(self, t, name, shape):
        if t is None:
            return None
        if not t.is_cuda:
            raise RuntimeError(f'EvalEpilogue: {name} must be a CUDA tensor (no CPU fallback)')
        if t.dtype != torch.float32:
            raise TypeError(f'EvalEpilogue: {name} must be float32')
        if tuple(t.shape[1:]) != shape:
            raise ValueError(f'EvalEpilogue: {name} must be (B, {shape[0]}, {shape[1]}), got {tuple(t.shape)}')
        return t.contiguous()

    def __call__(self, pred_mesh, gt_mesh, gt_pose3d, pa: bool = False, pred_pose=None, scale=None) -> EvalResult:
        """pred_mesh / gt_mesh (B,6890,3) in metres (gt_mesh may be None), gt_pose3d (B,Jt,3) in mm.
        `pred_pose` (B,Jt,3) in mm, when given, replaces the regression (compute_both_err's own signature)."""
        pred_mesh = self._check(pred_mesh, 'pred_mesh', (self.num_verts, 3))
        B = pred_mesh.shape[0]
        gt_mesh = self._check(gt_mesh, 'gt_mesh', (self.num_verts, 3))
        gt_pose3d = self._check(gt_pose3d, 'gt_pose3d', (self.num_joints, 3))
        pred_pose = self._check(pred_pose, 'pred_pose', (self.num_joints, 3))
        for t, name in ((gt_mesh, 'gt_mesh'), (gt_pose3d, 'gt_pose3d'), (pred_pose, 'pred_pose')):
            if t is not None and t.shape[0] != B:
                raise ValueError(f'EvalEpilogue: {name} has batch {t.shape[0]}, pred_mesh has {B}')
        dev = pred_mesh.device
        out_pose = torch.empty((B, self.num_joints, 3), dtype=torch.float32, device=dev)
        j_err = torch.empty((B,), dtype=torch.float32, device=dev)
        s_err = torch.empty((B,), dtype=torch.float32, device=dev) if gt_mesh is not None else None
        pa_err = torch.empty((B,), dtype=torch.float32, device=dev) if pa else None
        means = torch.zeros((3,), dtype=torch.float32, device=dev)
        a = _lib.EvalArgs(batch=B, verts=self.num_verts, joints=self.num_joints, n_eval=self.eval_joints.numel(),
                          root=self.root, scale=self.scale if scale is None else float(scale),
                          jreg_rowptr=_lib.ptr(self._csr.rowptr), jreg_colidx=_lib.ptr(self._csr.colidx),
                          jreg_values=_lib.ptr(self._csr.values), eval_joints=_lib.ptr(self.eval_joints),
                          pred_mesh=_lib.ptr(pred_mesh), gt_mesh=_lib.ptr(gt_mesh), gt_joints=_lib.ptr(gt_pose3d),
                          pred_joints_in=_lib.ptr(pred_pose), pred_joints=_lib.ptr(out_pose),
                          joint_err=_lib.ptr(j_err), surface_err=_lib.ptr(s_err), pa_joint_err=_lib.ptr(pa_err),
                          batch_mean=_lib.ptr(means))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gator_eval_epilogue(a, _lib.stream_ptr()), 'gator_eval_epilogue')
        return EvalResult(means[0], means[1], means[2], out_pose, j_err, s_err, pa_err)

    def compute_both_err(self, pred_mesh, target_mesh, pred_joint, target_joint):
        """Same signature, units (everything already in mm) and return value as the datasets'
        ``compute_both_err`` (data/Human36M/dataset.py:466-478): ``(joint_mean_error, mesh_mean_error)``."""
        r = self(pred_mesh, target_mesh, target_joint, pred_pose=pred_joint, scale=1.0)
        jm, sm = torch.stack([r.joint_error, r.surface_error]).tolist()      # the one host read-back
        return jm, sm
