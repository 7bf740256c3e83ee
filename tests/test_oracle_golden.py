"""Pins the CPU oracle (oracle/gator_oracle.py) to golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Same torch ops in the same order => bit-exact."""
import numpy as np
import pytest
import torch

from helpers import golden, oracle_setup, orc, synthetic, to_dtype


@pytest.mark.parametrize('tag', ['h36m', 'coco'])
def test_gator_forward_matches_reference(tag):
    g = golden('gator')
    sd, gc, mc, alpha = oracle_setup(tag)
    tr = {}
    with torch.no_grad():
        mesh, p3 = orc.gator_forward(sd, gc, mc, torch.from_numpy(g[f'{tag}/pose2d']), alpha, trace=tr)
    assert np.abs(mesh.numpy() - g[f'{tag}/mesh']).max() <= 1e-6
    assert np.abs(p3.numpy() - g[f'{tag}/pose3d']).max() <= 1e-3          # millimetres, |x| ~ 300
    n = 0
    for k, v in tr.items():
        gk = f'{tag}/trace/{k}'
        if gk in g:
            n += 1
            assert np.abs(v.numpy()[:len(g[gk])] - g[gk]).max() <= 1e-5, k
    assert n >= 12


@pytest.mark.parametrize('tag', ['h36m', 'coco'])
def test_constructor_constants_match_reference(tag):
    g = golden('gator')
    sd, gc, mc, _ = oracle_setup(tag)
    assert (np.asarray(mc['vj_relation']) == g[f'{tag}/vj_relation']).all()
    assert (sd['pose_lifter.graph_adj'].numpy() == g[f'{tag}/graph_adj']).all()
    assert np.abs(gc['edge_input'].numpy() - g[f'{tag}/edge_input']).max() == 0
    assert np.abs(sd['pose2mesh.init_vertices'].numpy() - g[f'{tag}/init_vertices_431']).max() == 0


def test_fp64_oracle_bounds_reference_rounding():
    """fp64 run of the same restatement: the reference's own fp32 rounding is ~5e-7 m."""
    g = golden('gator')
    sd, gc, mc, alpha = oracle_setup('h36m')
    with torch.no_grad():
        mesh, _ = orc.gator_forward(to_dtype(sd, torch.float64), gc, mc,
                                    torch.from_numpy(g['h36m/pose2d']).double(), alpha)
    assert np.abs(mesh.numpy() - g['h36m/mesh']).max() < 5e-6


def test_joint_regression_post_step():
    from helpers import regressor
    g = golden('gator')
    j = orc.joint_regress(torch.from_numpy(regressor('h36m')), torch.from_numpy(g['h36m/mesh']))
    assert np.abs(j.numpy() - g['h36m/joints']).max() <= 1e-6


def test_smpl_layer_matches_reference():
    s = golden('smpl')
    buf = {k: torch.from_numpy(v) for k, v in synthetic.smpl_buffers().items()}
    pose, betas, trans = [torch.from_numpy(a) for a in synthetic.smpl_inputs(4)]
    v, j, _ = orc.smpl_forward(buf, synthetic.SMPL_PARENTS, pose, betas, trans)
    assert np.abs(v.numpy() - s['full/verts']).max() <= 1e-6 and np.abs(j.numpy() - s['full/jtr']).max() <= 1e-6
    v, j, _ = orc.smpl_forward(buf, synthetic.SMPL_PARENTS, pose)
    assert np.abs(v.numpy()[:2] - s['nobetas/verts']).max() <= 1e-6 and np.abs(j.numpy() - s['nobetas/jtr']).max() <= 1e-6
    v, j, _ = orc.smpl_forward(buf, synthetic.SMPL_PARENTS, pose, betas, None, center_idx=0)
    assert np.abs(v.numpy()[:2] - s['center/verts']).max() <= 1e-6 and np.abs(j.numpy() - s['center/jtr']).max() <= 1e-6


def test_mesh_resampling_matches_reference():
    m = golden('mesh')
    A, D, U = synthetic.mesh_sampling_matrices()
    x = torch.from_numpy(m['x'])
    d1 = orc.mesh_downsample(D, x, 0, 1)
    d2 = orc.mesh_downsample(D, d1, 1, 2)
    u1 = orc.mesh_upsample(U, d2, 2, 1)
    u0 = orc.mesh_upsample(U, u1, 1, 0)
    for a, k in ((d1, 'down1'), (d2, 'down2'), (u1, 'up1'), (u0, 'up0')):
        assert np.abs(a.numpy() - m[k]).max() <= 1e-6, k
    assert np.abs(orc.mesh_downsample(D, x[0], 0, 2).numpy() - m['down2d']).max() <= 1e-6
