"""Golden vectors of the ground-truth mesh generation: the UNMODIFIED
``Human36M.get_smpl_coord`` / ``PW3D.get_smpl_coord`` (data/Human36M/dataset.py:254-298, data/PW3D/dataset.py:84-102)
called unbound on a stand-in ``self`` whose ``mesh_model.layer`` is the reference SMPL_Layer with synthetic buffers,
and ``get_coco_from_mesh`` (:311-334).  transforms3d is not installed here; the shim supplies the oracle's
restatement of its two functions.  Container only:

    python tests/golden/make_golden_gtmesh.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from gator_b200 import synthetic  # noqa: E402
from oracle import refshim        # noqa: E402

B = 5
KEEP = 512           # vertices stored per mesh (a strided subset) - the joints cover the rest


def main():
    H36M = refshim.dataset_class('Human36M')
    PW3D = refshim.dataset_class('PW3D')
    layer = refshim.build_smpl_layer(synthetic.smpl_buffers())
    reg_coco = np.load(os.path.join(refshim.REF, 'data', 'COCO', 'J_regressor_coco.npy')).astype(np.float32)
    me = types.SimpleNamespace(mesh_model=types.SimpleNamespace(layer={'neutral': layer}), smpl_root_joint_idx=0,
                               joint_regressor_coco=reg_coco, coco_joints_name=synthetic.COCO_JOINTS_NAME)
    me.add_pelvis_and_neck = lambda jc: H36M.add_pelvis_and_neck(me, jc)
    pose, shape, trans, R, t = synthetic.camera_annotations(B)
    sel = np.arange(0, 6890, 6890 // KEEP)[:KEEP]
    out = {'vertex_subset': sel}
    hm, hj, pm, pj, cc, ci = [], [], [], [], [], []
    for b in range(B):
        smpl_param = {'pose': pose[b].tolist(), 'shape': shape[b].tolist(), 'trans': trans[b].tolist(), 'gender': 'neutral'}
        cam_param = {'R': R[b].tolist(), 't': t[b].tolist(), 'focal': [1145.0, 1143.0], 'princpt': [512.5, 515.4]}
        mesh, joints = H36M.get_smpl_coord(me, smpl_param, cam_param)
        jc, ji = H36M.get_coco_from_mesh(me, mesh, cam_param)
        hm.append(mesh[sel]), hj.append(joints), cc.append(jc), ci.append(ji)
        mesh, joints = PW3D.get_smpl_coord(me, smpl_param)
        pm.append(mesh[sel]), pj.append(joints)
    out['h36m/mesh'], out['h36m/joints'] = np.asarray(hm), np.asarray(hj)
    out['h36m/coco_cam'], out['h36m/coco_img'] = np.asarray(cc), np.asarray(ci)
    out['pw3d/mesh'], out['pw3d/joints'] = np.asarray(pm), np.asarray(pj)
    np.savez_compressed(os.path.join(HERE, 'gtmesh.npz'), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, 'gtmesh.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
