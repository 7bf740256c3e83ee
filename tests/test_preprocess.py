"""2D-pose pre-processing: oracle vs goldens from the reference's own functions on
CPU; CUDA kernel vs both on the GPU.

Tolerance: the outputs are standardised coordinates (unit scale) in float32.  The kernel takes float32 pixels where
the reference reads float64 and solves the 3-point affine in closed form where cv2 runs an LU, so crop pixels may
differ by an ulp (3e-5 px at 300 px) and the standardised value by ~1e-6: bound 1e-5.
"""
import numpy as np
import pytest
import torch

from helpers import golden, orc, synthetic

TOL = 1e-5
MID = ((11, 12), (5, 6))


def test_oracle_matches_reference_preprocessing():
    g = golden('preproc')
    for b, j in enumerate(g['input']):
        p, ji, bb = orc.preprocess_pose2d(j, mid_pairs=MID)
        assert np.array_equal(p, g['pose19'][b]) and np.array_equal(bb, g['bbox19'][b])
        assert np.abs(ji - g['joint_img19'][b]).max() < 1e-9
        p, ji, bb = orc.preprocess_pose2d(j)
        assert np.array_equal(p, g['pose17'][b]) and np.array_equal(bb, g['bbox17'][b])
    # the demo fixture shipped by the reference, pre-processed when the model goldens were made
    assert np.array_equal(g['pose19'][0], golden('fixtures')['demo_pose19'])
    # a degenerate box (all joints on one point) is rejected like process_bbox's `return None`
    assert orc.preprocess_pose2d(np.ones((17, 2))) is None


@pytest.fixture(scope='module')
def built():
    from gator_b200 import build
    build.build()


@pytest.mark.gpu
@pytest.mark.parametrize('n_mid', [2, 0])
def test_preprocess_matches_reference(built, n_mid):
    from gator_b200.preprocess import COCO_MID_PAIRS, Pose2DPreprocessor
    g = golden('preproc')
    tag = '19' if n_mid else '17'
    pre = Pose2DPreprocessor((384, 288), COCO_MID_PAIRS if n_mid else ())
    x = torch.from_numpy(g['input'].astype(np.float32)).cuda()                 # (8,17,3) with confidence column
    pose, joint_img, bbox, valid = pre(x, return_aux=True)
    assert valid.cpu().numpy().all()
    assert np.abs(pose.cpu().numpy() - g['pose' + tag]).max() < TOL
    assert np.abs(bbox.cpu().numpy() - g['bbox' + tag]).max() < 2e-3           # pixels, |x| up to 3000
    assert np.abs(joint_img.cpu().numpy() - g['joint_img' + tag]).max() < 2e-3
    # (J, C) input, 2 columns
    p1 = pre(x[3, :, :2].contiguous())
    assert p1.shape == (17 + n_mid, 2) and np.abs(p1.cpu().numpy() - g['pose' + tag][3]).max() < TOL


@pytest.mark.gpu
def test_preprocess_matches_oracle_large_batch(built):
    from gator_b200.preprocess import COCO_MID_PAIRS, Pose2DPreprocessor
    base = golden('fixtures')['coco_joint_input'].reshape(17, -1)
    x64 = synthetic.pixel_poses(base, 301, seed=21)
    x32 = x64.astype(np.float32)
    want = np.stack([orc.preprocess_pose2d(j.astype(np.float64), mid_pairs=MID)[0] for j in x32])
    got = Pose2DPreprocessor((384, 288), COCO_MID_PAIRS)(torch.from_numpy(x32).cuda())
    assert np.abs(got.cpu().numpy() - want).max() < TOL


@pytest.mark.gpu
def test_preprocess_edge_cases(built):
    from gator_b200.preprocess import Pose2DPreprocessor
    pre = Pose2DPreprocessor()
    # degenerate box -> NaN + valid = 0 (the reference's process_bbox returns None and the caller crashes)
    x = torch.ones(2, 17, 2, device='cuda')
    x[1] = torch.from_numpy(golden('preproc')['input'][0, :, :2].astype(np.float32))
    pose, _, _, valid = pre(x, return_aux=True)
    assert valid.tolist() == [0, 1] and torch.isnan(pose[0]).all() and torch.isfinite(pose[1]).all()
    assert pre(torch.zeros(0, 17, 2, device='cuda')).shape == (0, 17, 2)
    with pytest.raises(RuntimeError):
        pre(torch.zeros(1, 17, 2))
    with pytest.raises(RuntimeError):
        Pose2DPreprocessor(mid_pairs=((40, 1),))(torch.zeros(1, 17, 2, device='cuda'))
    with pytest.raises(RuntimeError):
        pre(torch.zeros(1, 33, 2, device='cuda'))


@pytest.mark.gpu
def test_pixels_to_mesh_on_device(built):
    """demo/run.py end to end on the device: detector pixels -> pre-processing -> GATOR.forward, against the
    reference's mesh for its shipped demo input."""
    from helpers import build_b200_gator
    from gator_b200.preprocess import COCO_MID_PAIRS, Pose2DPreprocessor
    g = golden('gator')
    model = build_b200_gator('coco', 'cuda:0')
    px = torch.from_numpy(golden('fixtures')['coco_joint_input'].reshape(1, 17, -1).astype(np.float32)).cuda()
    pose = Pose2DPreprocessor((384, 288), COCO_MID_PAIRS)(px)
    with torch.no_grad():
        mesh, _ = model(pose)
    assert np.abs(mesh.cpu().numpy()[0] - g['coco/mesh'][0]).max() < 1e-4
