"""Golden vectors of the 2D-pose pre-processing from the UNMODIFIED reference
functions get_bbox / process_bbox (lib/coord_utils.py) and j2d_processing (lib/aug_utils.py), chained as
demo/run.py:124-133 chains them.  Container only (needs /root/reference):

    python tests/golden/make_golden_preproc.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from gator_b200 import synthetic  # noqa: E402
from oracle import refshim        # noqa: E402


def reference_chain(j, cfg):
    from coord_utils import get_bbox, process_bbox
    from aug_utils import j2d_processing
    bbox = get_bbox(j)
    bbox2 = process_bbox(bbox.copy())
    ji, _ = j2d_processing(j.copy(), (cfg.MODEL.input_shape[1], cfg.MODEL.input_shape[0]), bbox2, 0, 0, None)
    joint_img = ji[:, :2].copy()
    ji = ji[:, :2]
    ji /= np.array([[cfg.MODEL.input_shape[1], cfg.MODEL.input_shape[0]]])
    mean, std = np.mean(ji, axis=0), np.std(ji, axis=0)
    return ((ji.copy() - mean) / std), joint_img, bbox2


def main():
    cfg = refshim.install_shims()
    sys.path.insert(0, os.path.join(refshim.REF, 'demo'))
    names = synthetic.COCO_JOINTS_NAME
    base = np.load(os.path.join(refshim.REF, 'demo', 'coco_joint_input.npy')).reshape(17, -1)
    inputs = synthetic.pixel_poses(base, 8)                     # (8,17,3) float64: tall, wide and tiny boxes
    out = {'input': inputs}
    p19, i19, b19, p17, i17, b17 = [], [], [], [], [], []
    for j in inputs:
        # demo/run.py:103-121,193-198 (add_pelvis, add_neck; the run.py module itself needs a renderer)
        jc = j.copy()
        for a, b in (('L_Hip', 'R_Hip'), ('L_Shoulder', 'R_Shoulder')):
            ia, ib = names.index(a), names.index(b)
            m = (jc[ia, :] + jc[ib, :]) * 0.5
            m[2] = jc[ia, 2] * jc[ib, 2]
            jc = np.concatenate((jc, m.reshape(1, 3)))
        p, i, b = reference_chain(jc[:, :2], cfg)
        p19.append(p), i19.append(i), b19.append(b)
        p, i, b = reference_chain(j[:, :2].copy(), cfg)         # 17-joint sets: no synthesised joints
        p17.append(p), i17.append(i), b17.append(b)
    out['pose19'], out['joint_img19'], out['bbox19'] = np.asarray(p19), np.asarray(i19), np.asarray(b19)
    out['pose17'], out['joint_img17'], out['bbox17'] = np.asarray(p17), np.asarray(i17), np.asarray(b17)
    np.savez_compressed(os.path.join(HERE, 'preproc.npz'), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == '__main__':
    main()
