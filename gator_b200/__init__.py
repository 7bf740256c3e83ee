"""gator_b200 - B200-native (sm_100a) drop-in for the GATOR pose->mesh forward path.

Public surface (mirrors the reference's module layout):
  gator_b200.models.GATOR / GAT / MDR     nn.Module replacements (same signatures + state_dict keys)
  gator_b200.smpl_layer.SMPL_Layer        SMPL linear blend skinning layer
  gator_b200.mesh.Mesh                    sparse mesh up/down-sampling
  gator_b200.install()                    rebind the reference's classes in place
The compute path is the C-ABI library include/gator_b200.h (csrc/*.cu); there is no CPU fallback.
"""
from .install import install  # noqa: F401

__all__ = ['install']
