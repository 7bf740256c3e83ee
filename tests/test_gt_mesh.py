"""Ground-truth mesh generation: oracle vs goldens from the reference's own
``get_smpl_coord`` on CPU; batched CUDA path vs both on the GPU.

Tolerance: outputs are camera-frame millimetres of magnitude ~6000 (4.5 m camera distance), i.e. one float32 ulp is
5e-4 mm, and the reference itself adds a ~5 m translation to metre-scale vertices in float32 before the *1000.
Bound: 1e-2 mm (1.7e-6 relative); pixel coordinates (~500 px): 1e-2 px.
"""
import numpy as np
import pytest
import torch

from helpers import build_b200_smpl, golden, orc, regressor, synthetic

TOL_MM = 1e-2
B = 5


def _buf():
    return {k: torch.from_numpy(v) for k, v in synthetic.smpl_buffers().items()}


def test_oracle_matches_reference_get_smpl_coord():
    g = golden('gtmesh')
    sel = g['vertex_subset']
    pose, shape, trans, R, t = synthetic.camera_annotations(B)
    buf = _buf()
    for b in range(B):
        mesh, joints = orc.h36m_smpl_coord(buf, synthetic.SMPL_PARENTS, pose[b], shape[b], trans[b], R[b], t[b])
        assert np.abs(mesh[sel] - g['h36m/mesh'][b]).max() <= 1e-3 and np.abs(joints - g['h36m/joints'][b]).max() <= 1e-3
        mesh, joints = orc.pw3d_smpl_coord(buf, synthetic.SMPL_PARENTS, pose[b], shape[b], trans[b])
        assert np.abs(mesh[sel] - g['pw3d/mesh'][b]).max() <= 1e-3 and np.abs(joints - g['pw3d/joints'][b]).max() <= 1e-3


def test_restated_axangle_round_trip():
    """mat2axangle(axangle2mat(a)) = a for rotation vectors with |a| < pi, incl. tiny angles."""
    r = np.random.Generator(np.random.PCG64(3))
    for scale in (1e-6, 0.3, 2.0, 3.1):
        a = r.standard_normal(3)
        a = a / np.linalg.norm(a) * scale
        axis, ang = orc.mat2axangle(orc.axangle2mat(a / np.linalg.norm(a), np.linalg.norm(a)))
        assert np.abs(axis * ang - a).max() < 1e-9


@pytest.fixture(scope='module')
def gen():
    from gator_b200 import build
    from gator_b200.gt_mesh import GtMeshGenerator
    build.build()
    return GtMeshGenerator(build_b200_smpl(device='cuda:0'))


def _cuda(*arrays):
    return [torch.from_numpy(a).cuda() for a in arrays]


@pytest.mark.gpu
def test_gt_mesh_matches_reference(gen):
    from gator_b200.gt_mesh import MeshJoints
    from gator_b200.preprocess import COCO_MID_PAIRS
    g = golden('gtmesh')
    sel = torch.from_numpy(g['vertex_subset']).cuda()
    pose, shape, trans, R, t = _cuda(*synthetic.camera_annotations(B))
    mesh, joints = gen.h36m(pose, shape, trans, R, t)
    assert mesh.shape == (B, 6890, 3) and joints.shape == (B, 24, 3)
    assert np.abs(mesh[:, sel].cpu().numpy() - g['h36m/mesh']).max() < TOL_MM
    assert np.abs(joints.cpu().numpy() - g['h36m/joints']).max() < TOL_MM
    # get_coco_from_mesh: sparse regression + pelvis / neck + projection
    cam, img = MeshJoints(regressor('coco'), COCO_MID_PAIRS)(mesh, focal=[1145.0, 1143.0], princpt=[512.5, 515.4])
    assert np.abs(cam.cpu().numpy() - g['h36m/coco_cam']).max() < TOL_MM
    assert np.abs(img.cpu().numpy() - g['h36m/coco_img']).max() < 1e-2
    mesh, joints = gen.pw3d(pose, shape, trans)
    assert np.abs(mesh[:, sel].cpu().numpy() - g['pw3d/mesh']).max() < TOL_MM
    assert np.abs(joints.cpu().numpy() - g['pw3d/joints']).max() < TOL_MM


@pytest.mark.gpu
def test_gt_mesh_matches_oracle_and_properties(gen):
    n = 1100
    pose, shape, trans, R, t = synthetic.camera_annotations(n, seed=17)
    pose[3, :3] = 0.0                                           # identity root: angle 0 would divide 0/0 in the reference
    pose[3, 0] = 1e-4
    pose[4, :3] = np.array([3.0, 0.5, -0.2], np.float32)        # |a| close to pi
    dp, ds, dt, dR, dtt = _cuda(pose, shape, trans, R, t)
    mesh, joints = gen.h36m(dp, ds, dt, dR, dtt)
    buf = _buf()
    for b in (0, 1, 3, 4, 1030, 1099):
        om, oj = orc.h36m_smpl_coord(buf, synthetic.SMPL_PARENTS, pose[b], shape[b], trans[b], R[b], t[b])
        assert np.abs(mesh[b].cpu().numpy() - om).max() < TOL_MM, b
        assert np.abs(joints[b].cpu().numpy() - oj).max() < TOL_MM, b
    # identity camera: the fix-up must be the identity, so the result equals the plain forward + trans, in mm
    eye = torch.eye(3, device='cuda').repeat(n, 1, 1)
    zero_t = torch.zeros(n, 3, device='cuda')
    ds_ok = ds.clamp(-2.9, 2.9)
    m1, j1 = gen.h36m(dp, ds_ok, dt, eye, zero_t)
    m2, j2 = gen.pw3d(dp, ds_ok, dt)
    assert (m1 - m2).abs().max() < TOL_MM and (j1 - j2).abs().max() < TOL_MM
    # rigidity: the camera rotation changes no pairwise distance of the SMPL joints
    _, j0 = gen.h36m(dp, ds_ok, dt, dR, dtt)
    d1 = (j1[:, :, None] - j1[:, None]).norm(dim=-1)
    d0 = (j0[:, :, None] - j0[:, None]).norm(dim=-1)
    assert (d1 - d0).abs().max() < TOL_MM
    # metres output when no scale is requested is exactly the mm output / 1000 up to rounding
    m3, _ = gen.h36m(dp[:4], ds[:4], dt[:4], dR[:4], dtt[:4], scale=1.0)
    assert (m3 * 1000 - mesh[:4]).abs().max() < TOL_MM
    assert gen.h36m(dp[:0], ds[:0], dt[:0], dR[:0], dtt[:0])[0].shape == (0, 6890, 3)
    with pytest.raises(RuntimeError):
        gen.h36m(dp.cpu(), ds, dt, dR, dtt)
