"""Ground-truth mesh generation on the device.

The reference's datasets run SMPL once per item, at batch 1, on the CPU inside every DataLoader worker
(``get_smpl_coord``: data/Human36M/dataset.py:254-298, data/PW3D/dataset.py:84-102) and then regress joints from
the mesh with dense matmuls (``get_coco_from_mesh`` :311-320, ``get_h36mJ_from_mesh``).  Here a whole batch of
annotations becomes camera-frame meshes in millimetres with three launches: the camera fix-up kernel
(csrc/smpl_cam.cu), the SMPL forward (csrc/smpl.cu) with the translation and the ``*1000`` folded into the skinning
kernel, and sparse J-regressions (csrc/sparse.cu).  CUDA only - no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from .ops import JointRegressor
from .smpl_layer import SMPL_Layer


class GtMeshGenerator:
    """``gen.h36m(pose, shape, trans, R, t)`` / ``gen.pw3d(pose, shape, trans)`` -> ``(mesh_cam, joint_cam_smpl)``
    in millimetres, the batched ``get_smpl_coord`` of the two datasets.  ``layer`` is a gator_b200 ``SMPL_Layer``
    (one per gender, as in lib/smpl.py:12-13)."""

    def __init__(self, layer: SMPL_Layer, root_joint_idx: int = 0):
        if not isinstance(layer, SMPL_Layer):
            raise TypeError('GtMeshGenerator needs a gator_b200.smpl_layer.SMPL_Layer')
        if layer.center_idx is not None:
            raise ValueError('GtMeshGenerator: the datasets use SMPL layers without a centre joint')
        if root_joint_idx != 0:
            raise NotImplementedError('the SMPL root joint (Pelvis) is joint 0 (lib/smpl.py:48)')
        self.layer = layer

    @staticmethod
    def _f32(t, name, shape, dev):
        if not torch.is_tensor(t) or not t.is_cuda:
            raise RuntimeError(f'GtMeshGenerator: {name} must be a CUDA tensor (no CPU fallback)')
        t = t.detach().to(dev).float().reshape(shape).contiguous()
        return t

    def h36m(self, pose, shape, trans, cam_R, cam_t, scale: float = 1000.0):
        """pose (B,72) world-frame axis-angle, shape (B,10), trans (B,3) m, cam_R (B,3,3), cam_t (B,3) mm."""
        L = self.layer
        if L._packed is None:
            L.pack()
        p = L._packed
        dev = p['dev']
        B = pose.shape[0]
        pose = self._f32(pose, 'pose', (B, 72), dev)
        shape = self._f32(shape, 'shape', (B, 10), dev)
        trans = self._f32(trans, 'trans', (B, 3), dev)
        cam_R = self._f32(cam_R, 'cam_R', (B, 9), dev)
        cam_t = self._f32(cam_t, 'cam_t', (B, 3), dev)
        pose_o, shape_o, trans_o = torch.empty_like(pose), torch.empty_like(shape), torch.empty_like(trans)
        a = _lib.SmplCamArgs(batch=B, reserved=0, j_template=_lib.ptr(p['j_template']),
                             j_shapedirs=_lib.ptr(p['j_shapedirs']), default_betas=_lib.ptr(p['default_betas']),
                             pose=_lib.ptr(pose), betas=_lib.ptr(shape), trans=_lib.ptr(trans), cam_R=_lib.ptr(cam_R),
                             cam_t=_lib.ptr(cam_t), pose_out=_lib.ptr(pose_o), betas_out=_lib.ptr(shape_o),
                             trans_out=_lib.ptr(trans_o))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gator_smpl_cam_fixup(a, _lib.stream_ptr()), 'gator_smpl_cam_fixup')
        return L.forward_scaled(pose_o, shape_o, trans_o, scale)

    def pw3d(self, pose, shape, trans, scale: float = 1000.0):
        """pose (B,72), shape (B,10), trans (B,3) m -> SMPL forward with the world translation, then mm."""
        return self.layer.forward_scaled(pose, shape, trans, scale)


class MeshJoints:
    """``get_coco_from_mesh`` / ``get_h36mJ_from_mesh`` (data/Human36M/dataset.py:311-320, data/PW3D/dataset.py):
    sparse regression of a joint set from the generated meshes, optional pelvis / neck midpoints and the
    pin-hole projection ``cam2pixel`` (lib/coord_utils.py:104-109)."""

    def __init__(self, J_regressor, mid_pairs=(), device='cuda'):
        self.reg = JointRegressor(J_regressor, device=device)
        self.mid_pairs = tuple(mid_pairs)

    def __call__(self, mesh_cam: torch.Tensor, focal=None, princpt=None):
        joints = self.reg(mesh_cam)                                           # csr_spmm kernel
        for a, b in self.mid_pairs:                                           # add_pelvis_and_neck (:322-334)
            joints = torch.cat([joints, ((joints[:, a] + joints[:, b]) * 0.5)[:, None]], 1)
        if focal is None:
            return joints
        f = torch.as_tensor(focal, dtype=torch.float32, device=joints.device).reshape(-1, 2)
        c = torch.as_tensor(princpt, dtype=torch.float32, device=joints.device).reshape(-1, 2)
        img = torch.empty_like(joints)
        img[..., 0] = joints[..., 0] / joints[..., 2] * f[:, None, 0] + c[:, None, 0]
        img[..., 1] = joints[..., 1] / joints[..., 2] * f[:, None, 1] + c[:, None, 1]
        img[..., 2] = 1.0                                                     # dataset.py:319
        return joints, img
