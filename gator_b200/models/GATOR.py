"""Drop-in for ``models.GATOR`` (lib/models/GATOR.py:8-27): GAT lifter -> MDR decoder, same signature,
same ``state_dict`` keys (``pose_lifter.*``, ``pose2mesh.*``), same return value.

The module owns no arithmetic: it wires the two kernel-backed stages together, keeps their precision mode in
sync and skips the (B, J, 133) ``torch.cat`` of the reference by handing the three pieces to the decoder separately.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, config
from . import GAT, MDR

_PRECISION_NAMES = {v: k for k, v in _lib.PRECISIONS.items()}


class GATOR(nn.Module):
    def __init__(self, num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor):
        super().__init__()
        self.num_joint = int(num_joint)
        lifter_kwargs = dict(pretrained=config.get_cfg().MODEL.posenet_pretrained)       # GATOR.py:13
        self.pose_lifter = GAT.get_model(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor, **lifter_kwargs)
        self.pose2mesh = MDR.get_model(num_joint, embed_dim)

    # -- precision of the matrix products (accumulation is always fp32) ---------------------------------------------
    @property
    def precision(self) -> str:
        return _PRECISION_NAMES[self.pose_lifter.precision]

    def set_precision(self, precision: str):
        """'fp32' (FFMA parity path), 'bf16x3' (tcgen05: 3-term bf16 split GEMMs + bf16 attention cores) or
        'bf16' (tcgen05, single bf16 products)."""
        if precision not in _lib.PRECISIONS:
            raise ValueError(f'precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}')
        code = _lib.PRECISIONS[precision]
        for stage in (self.pose_lifter, self.pose2mesh):
            stage.precision = code
        return self

    def extra_repr(self) -> str:
        return f'num_joint={self.num_joint}, precision={self.precision}'

    # -- forward -------------------------------------------------------------------------------------------------------
    def forward(self, pose2d):
        """pose2d (B,J,2) -> (cam_mesh (B,6890,3) metres, pose3d (B,J,3) millimetres)  (GATOR.py:16-22).
        The (B,J,133) concat is never materialised; the /1000 happens inside the MDR embedding kernel."""
        J = self.num_joint
        if pose2d.dim() != 3 or pose2d.shape[1] != J or pose2d.shape[2] != 2:
            raise ValueError(f'GATOR.forward expects pose2d of shape (B, {J}, 2), got {tuple(pose2d.shape)}')
        batch = pose2d.shape[0]
        lifted, joint_feat = self.pose_lifter(pose2d.reshape(batch, J * 2))               # (B, 3J) mm, (B, J, 128)
        pose3d = lifted.reshape(batch, J, 3)
        return self.pose2mesh.forward_parts(pose2d, pose3d, joint_feat), pose3d


def get_model(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor):
    return GATOR(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor)
