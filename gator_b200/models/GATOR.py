"""Drop-in for ``models.GATOR`` (lib/models/GATOR.py:8-27): GAT lifter -> MDR decoder, same signature,
same ``state_dict`` keys (``pose_lifter.*``, ``pose2mesh.*``), same return value."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import config
from . import GAT, MDR


class GATOR(nn.Module):
    def __init__(self, num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor):
        super().__init__()
        cfg = config.get_cfg()
        self.num_joint = num_joint
        self.pose_lifter = GAT.get_model(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor,
                                         pretrained=cfg.MODEL.posenet_pretrained)
        self.pose2mesh = MDR.get_model(num_joint, embed_dim)

    def set_precision(self, precision: str):
        """'fp32' (FFMA parity path), 'bf16x3' (tcgen05: 3-term bf16 split GEMMs + bf16 attention cores) or
        'bf16' (tcgen05, single bf16 products)."""
        from .. import _lib
        p = _lib.PRECISIONS[precision]
        self.pose_lifter.precision = p
        self.pose2mesh.precision = p
        return self

    def forward(self, pose2d):
        """pose2d (B,J,2) -> (cam_mesh (B,6890,3) metres, pose3d (B,J,3) millimetres)  (GATOR.py:16-22).
        The (B,J,133) concat is never materialised; the /1000 happens inside the MDR embedding kernel."""
        pose3d, pose3d_feat = self.pose_lifter(pose2d.reshape(len(pose2d), self.num_joint * 2))
        pose3d = pose3d.reshape(-1, self.num_joint, 3)
        cam_mesh = self.pose2mesh.forward_parts(pose2d, pose3d, pose3d_feat)
        return cam_mesh, pose3d


def get_model(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor):
    return GATOR(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor)
