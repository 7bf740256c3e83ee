"""Device-side 2D-pose pre-processing.

Mirrors the per-sample host chain in front of every forward - ``add_pelvis`` / ``add_neck``
(demo/run.py:103-121), ``get_bbox`` + ``process_bbox`` (lib/coord_utils.py:21-66), ``j2d_processing`` with
rot = 0 / flip = 0 (lib/aug_utils.py:51-64,140-184), ``/[W, H]`` and the per-axis standardisation
(demo/run.py:130-133, data/Human36M/dataset.py:383-389) - as one kernel over the batch (csrc/preprocess.cu),
so that detector pixels go in and the forward's ``pose2d`` comes out without host arithmetic.  CUDA only.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib

# indices in the COCO joint order of demo/run.py:190: L_Hip, R_Hip -> Pelvis; L_Shoulder, R_Shoulder -> Neck
COCO_MID_PAIRS = ((11, 12), (5, 6))


class Pose2DPreprocessor:
    """``pre(joint_input)`` -> standardised ``pose2d`` (B, J', 2) for ``GATOR.forward``.

    joint_input: (B, J, C>=2) or (J, C) float32 CUDA tensor of pixel coordinates (x, y[, confidence]).
    input_shape: cfg.MODEL.input_shape = (height, width) of the virtual crop.
    mid_pairs:   joints to synthesise as midpoints and append (the demo's pelvis and neck); () for H36M."""

    def __init__(self, input_shape: Tuple[int, int] = (384, 288), mid_pairs: Sequence[Tuple[int, int]] = (),
                 bbox_scale: float = 1.0):
        if len(mid_pairs) > 2:
            raise ValueError('Pose2DPreprocessor: at most two synthesised joints')
        self.input_shape = (int(input_shape[0]), int(input_shape[1]))
        self.mid_pairs = tuple((int(a), int(b)) for a, b in mid_pairs)
        self.bbox_scale = float(bbox_scale)

    def __call__(self, joint_input: torch.Tensor, return_aux: bool = False):
        if not joint_input.is_cuda:
            raise RuntimeError('Pose2DPreprocessor: CUDA tensors only (no CPU fallback)')
        if joint_input.dtype != torch.float32:
            raise TypeError('Pose2DPreprocessor: float32 only')
        squeeze = joint_input.dim() == 2
        x = (joint_input.unsqueeze(0) if squeeze else joint_input).contiguous()
        if x.dim() != 3 or x.shape[2] < 2:
            raise ValueError(f'Pose2DPreprocessor: expected (B, J, >=2), got {tuple(joint_input.shape)}')
        B, Jin, Cin = x.shape
        J = Jin + len(self.mid_pairs)
        out = torch.empty((B, J, 2), dtype=torch.float32, device=x.device)
        aux = None
        if return_aux:
            aux = (torch.empty((B, J, 2), dtype=torch.float32, device=x.device),
                   torch.empty((B, 4), dtype=torch.float32, device=x.device),
                   torch.empty((B,), dtype=torch.int32, device=x.device))
        H, W = self.input_shape
        pairs = list(self.mid_pairs) + [(0, 0)] * (2 - len(self.mid_pairs))
        a = _lib.Pose2dArgs(batch=B, joints_in=Jin, in_stride=Cin, n_mid=len(self.mid_pairs),
                            mid_a=(_lib.C.c_int32 * 2)(pairs[0][0], pairs[1][0]),
                            mid_b=(_lib.C.c_int32 * 2)(pairs[0][1], pairs[1][1]),
                            out_w=float(W), out_h=float(H), aspect=W / H, bbox_scale=self.bbox_scale,
                            joints=_lib.ptr(x), pose2d=_lib.ptr(out),
                            joint_img=_lib.ptr(aux[0] if aux else None), bbox=_lib.ptr(aux[1] if aux else None),
                            valid=_lib.ptr(aux[2] if aux else None))
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().gator_pose2d_preprocess(a, _lib.stream_ptr()), 'gator_pose2d_preprocess')
        if squeeze:
            out = out[0]
            aux = tuple(t[0] for t in aux) if aux else None
        return (out,) + aux if return_aux else out
