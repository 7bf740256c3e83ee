"""Evaluation epilogue: oracle vs the reference's golden values on CPU, CUDA kernel
vs oracle and goldens on the GPU.

Tolerance: values are millimetres of magnitude 10..1000 accumulated in fp32 - the reference's own result moves by
~1e-4 mm with the summation order (dense 6890-term matmul vs the 6-term sparse row), so the bound is 2e-3 mm.
"""
import numpy as np
import pytest
import torch

from helpers import golden, orc, regressor, synthetic

TOL_MM = 2e-3
B = 6


def _inputs():
    return synthetic.eval_inputs(B, regressor('h36m'))


def test_oracle_matches_reference_eval():
    g = golden('eval')
    pred, gt, gt_pose = _inputs()
    reg = torch.from_numpy(regressor('h36m'))
    pred_pose = orc.eval_pred_pose(reg, torch.from_numpy(pred))
    assert np.array_equal(pred_pose.numpy(), g['pred_pose'])                      # same torch ops: bit-exact
    j, s = orc.compute_both_err(torch.from_numpy(pred) * 1000, torch.from_numpy(gt) * 1000, pred_pose,
                                torch.from_numpy(gt_pose))
    assert float(j) == float(g['joint_error']) and float(s) == float(g['surface_error'])
    mpjpe, mpvpe, pa = orc.per_sample_errors(pred * np.float32(1000), gt * np.float32(1000), pred_pose.numpy(), gt_pose)
    assert np.abs(mpjpe - g['joint_err']).max() < 1e-4
    assert np.abs(mpvpe - g['surface_err']).max() < 1e-4
    assert np.abs(pa - g['pa_joint_err']).max() < 1e-4
    # totals printed by the reference's evaluate_joint ('%.2f')
    assert abs(round(float(mpjpe.mean()), 2) - g['evaluate_joint_printed'][0]) < 0.011
    assert abs(round(float(pa.mean()), 2) - g['evaluate_joint_printed'][1]) < 0.011
    # reflection branch of rigid_align
    assert np.abs(orc.rigid_align(g['reflect_A'], g['reflect_B']) - g['reflect_aligned']).max() < 1e-8


@pytest.fixture(scope='module')
def epilogue():
    from gator_b200 import build
    from gator_b200.evaluate import EvalEpilogue
    build.build()
    return EvalEpilogue(regressor('h36m'), device='cuda:0')


@pytest.mark.gpu
def test_eval_epilogue_matches_reference(epilogue):
    g = golden('eval')
    pred, gt, gt_pose = (torch.from_numpy(a).cuda() for a in _inputs())
    r = epilogue(pred, gt, gt_pose, pa=True)
    assert np.abs(r.pred_pose.cpu().numpy() - g['pred_pose']).max() < TOL_MM
    assert np.abs(r.joint_err.cpu().numpy() - g['joint_err']).max() < TOL_MM
    assert np.abs(r.surface_err.cpu().numpy() - g['surface_err']).max() < TOL_MM
    assert np.abs(r.pa_joint_err.cpu().numpy() - g['pa_joint_err']).max() < TOL_MM
    assert abs(float(r.joint_error) - float(g['joint_error'])) < TOL_MM
    assert abs(float(r.surface_error) - float(g['surface_error'])) < TOL_MM
    assert abs(float(r.pa_joint_error) - g['pa_joint_err'].mean()) < TOL_MM
    # the datasets' own signature: everything already in mm, joints supplied by the caller
    j, s = epilogue.compute_both_err(pred * 1000, gt * 1000, torch.from_numpy(g['pred_pose']).cuda(), gt_pose)
    assert abs(j - float(g['joint_error'])) < TOL_MM and abs(s - float(g['surface_error'])) < TOL_MM


@pytest.mark.gpu
@pytest.mark.parametrize('batch', [1, 257])
def test_eval_epilogue_matches_oracle(epilogue, batch):
    reg = regressor('h36m')
    pred, gt, gt_pose = synthetic.eval_inputs(batch, reg, seed=11)
    pp = orc.eval_pred_pose(torch.from_numpy(reg), torch.from_numpy(pred)).numpy()
    mpjpe, mpvpe, pa = orc.per_sample_errors(pred * np.float32(1000), gt * np.float32(1000), pp, gt_pose)
    r = epilogue(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(gt_pose).cuda(), pa=True)
    assert np.abs(r.pred_pose.cpu().numpy() - pp).max() < TOL_MM
    assert np.abs(r.joint_err.cpu().numpy() - mpjpe).max() < TOL_MM
    assert np.abs(r.surface_err.cpu().numpy() - mpvpe).max() < TOL_MM
    assert np.abs(r.pa_joint_err.cpu().numpy() - pa).max() < TOL_MM
    assert abs(float(r.joint_error) - mpjpe.mean()) < TOL_MM
    # meshes that are only 4-byte aligned take the narrow-copy path and must give the same numbers
    def shifted(a):
        buf = torch.empty(a.size + 1, device='cuda')
        buf[1:] = torch.from_numpy(a).cuda().reshape(-1)
        return buf[1:].view(a.shape)
    r4 = epilogue(shifted(pred), shifted(gt), torch.from_numpy(gt_pose).cuda())
    assert r4.pred_pose.data_ptr() % 8 == 0 and shifted(pred).data_ptr() % 8 == 4
    assert torch.equal(r4.surface_err, r.surface_err) and torch.equal(r4.joint_err, r.joint_err)
    # without a ground-truth mesh and without PA: optional outputs are absent, means are zero
    r2 = epilogue(torch.from_numpy(pred).cuda(), None, torch.from_numpy(gt_pose).cuda())
    assert r2.surface_err is None and r2.pa_joint_err is None and float(r2.surface_error) == 0.0
    assert torch.equal(r2.joint_err, r.joint_err)


@pytest.mark.gpu
def test_eval_epilogue_edge_cases(epilogue):
    g = golden('eval')
    reg = regressor('h36m')
    # reflected joints: the det(R) < 0 repair of rigid_align; perfect alignment => PA error ~ 0 while MPJPE is large
    A, Bm = g['reflect_A'], g['reflect_B']
    full_p = np.zeros((1, 17, 3), np.float32)
    full_g = np.zeros((1, 17, 3), np.float32)
    ev = list(orc.H36M_EVAL_JOINTS)
    full_p[0, ev], full_g[0, ev] = A, Bm
    mesh = torch.zeros(1, synthetic.V_FULL, 3, device='cuda')
    r = epilogue(mesh, None, torch.from_numpy(full_g).cuda(), pa=True, pred_pose=torch.from_numpy(full_p).cuda(), scale=1.0)
    o, gg = full_p[0] - full_p[0][:1], full_g[0] - full_g[0][:1]
    want = np.sqrt(((orc.rigid_align(o[ev].astype(np.float64), gg[ev].astype(np.float64)) - gg[ev]) ** 2).sum(1)).mean()
    assert abs(float(r.pa_joint_err[0]) - want) < TOL_MM and float(r.joint_err[0]) > 50.0
    # empty batch
    e = epilogue(torch.zeros(0, synthetic.V_FULL, 3, device='cuda'), None, torch.zeros(0, 17, 3, device='cuda'))
    assert e.joint_err.shape == (0,) and e.pred_pose.shape == (0, 17, 3)
    # argument validation: CPU tensors and wrong shapes are refused, never routed to a fallback
    with pytest.raises(RuntimeError):
        epilogue(torch.zeros(1, synthetic.V_FULL, 3), None, torch.zeros(1, 17, 3))
    with pytest.raises(ValueError):
        epilogue(mesh, None, torch.zeros(1, 19, 3, device='cuda'))
