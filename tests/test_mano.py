"""MANO hand layer (manopth/manopth/manolayer.py:109-273).
CPU: the oracle restatement against goldens of the unmodified reference forward (tests/golden/make_golden_mano.py).
GPU: gator_b200.mano_layer.ManoLayer (generic LBS core + MANO pre/post kernels, through the C ABI) against the goldens
and, at a larger batch, against the oracle; size-independent properties."""
import numpy as np
import pytest
import torch

from helpers import golden, orc, synthetic

DEV = 'cuda:0'


def _oracle_buffers(data, ncomps, flat=True):
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    mean = np.zeros(45, np.float32) if flat else data['hands_mean']
    return {'th_betas': t(data['betas'])[None], 'th_shapedirs': t(data['shapedirs']), 'th_posedirs': t(data['posedirs']),
            'th_v_template': t(data['v_template'])[None], 'th_J_regressor': t(data['J_regressor']), 'th_weights': t(data['weights']),
            'th_hands_mean': t(mean)[None], 'th_selected_comps': t(data['hands_components'][:ncomps])}


CASES = [   # golden prefix, input prefix, layer kwargs, forward kwargs
    ('pca', 'pca', dict(ncomps=6), dict(use=('betas', 'trans'))),
    ('pca_plain', 'pca', dict(ncomps=6), dict(use=())),
    ('centre', 'centre', dict(ncomps=12, center_idx=9, flat_hand_mean=False, side='left'), dict(use=('betas',), root_palm=True, share_betas=True)),
    ('tip', 'centre', dict(ncomps=12, center_idx=8, flat_hand_mean=False, side='left'), dict(use=('betas',), zero_trans=True)),
    ('full', 'full', dict(ncomps=45, use_pca=False), dict(use=('betas', 'trans'))),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(case):
    name, inp, lk, fk = case
    g = golden('mano')
    data = synthetic.mano_data()
    buf = _oracle_buffers(data, lk['ncomps'], lk.get('flat_hand_mean', True))
    pose = torch.from_numpy(g[f'{inp}/pose'])
    betas = torch.from_numpy(g[f'{inp}/betas']) if 'betas' in fk['use'] else None
    trans = torch.from_numpy(g[f'{inp}/trans']) if 'trans' in fk['use'] else (torch.zeros(pose.shape[0], 3) if fk.get('zero_trans') else None)
    v, j = orc.mano_forward(buf, pose, betas, trans, ncomps=lk['ncomps'], use_pca=lk.get('use_pca', True), side=lk.get('side', 'right'),
                            center_idx=lk.get('center_idx'), root_palm=fk.get('root_palm', False), share_betas=fk.get('share_betas', False))
    assert v.shape == (pose.shape[0], 778, 3) and j.shape == (pose.shape[0], 21, 3)
    # millimetres, |x| ~ 100; bit-exact in the build container (same ops, same order)
    assert np.abs(v.numpy() - g[f'{name}/verts']).max() <= 1e-4, name
    assert np.abs(j.numpy() - g[f'{name}/jtr']).max() <= 1e-4, name


def test_state_dict_and_constructor_contract():
    from gator_b200.mano_layer import ManoLayer
    m = ManoLayer(mano_data=synthetic.mano_data(), ncomps=6)
    assert list(m.state_dict().keys()) == ['th_betas', 'th_shapedirs', 'th_posedirs', 'th_v_template', 'th_J_regressor', 'th_weights',
                                           'th_faces', 'th_hands_mean', 'th_selected_comps']
    assert m.th_selected_comps.shape == (6, 45) and m.kintree_parents[1:] == synthetic.MANO_PARENTS[1:]
    with pytest.raises(NotImplementedError):
        ManoLayer(mano_data=synthetic.mano_data(), root_rot_mode='rot6d')
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 9))                       # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_mano_layer_matches_reference_golden(case):
    from gator_b200.mano_layer import ManoLayer
    name, inp, lk, fk = case
    g = golden('mano')
    layer = ManoLayer(mano_data=synthetic.mano_data(), **lk).to(DEV)
    pose = torch.from_numpy(g[f'{inp}/pose']).to(DEV)
    kw = {}
    if 'betas' in fk['use']:
        kw['th_betas'] = torch.from_numpy(g[f'{inp}/betas']).to(DEV)
    if 'trans' in fk['use']:
        kw['th_trans'] = torch.from_numpy(g[f'{inp}/trans']).to(DEV)
    elif fk.get('zero_trans'):
        kw['th_trans'] = torch.zeros(pose.shape[0], 3, device=DEV)
    if fk.get('root_palm'):
        kw['root_palm'] = torch.Tensor([1])
    if fk.get('share_betas'):
        kw['share_betas'] = torch.Tensor([1])
    v, j = layer(pose, **kw)
    # 1e-3 mm = 1e-6 m on values of ~100 mm
    assert np.abs(v.cpu().numpy() - g[f'{name}/verts']).max() <= 1e-3, name
    assert np.abs(j.cpu().numpy() - g[f'{name}/jtr']).max() <= 1e-3, name


@pytest.mark.gpu
def test_mano_layer_batch_vs_oracle_and_properties():
    from gator_b200.mano_layer import ManoLayer
    data = synthetic.mano_data()
    B = 8192 + 37                                  # crosses the 8192-sample workspace chunk
    pose, betas, trans = synthetic.mano_inputs(B, ncomps=6, seed=11)
    layer = ManoLayer(mano_data=data, ncomps=6, flat_hand_mean=False).to(DEV)
    P, Bt, T = [torch.from_numpy(a).to(DEV) for a in (pose, betas, trans)]
    v, j = layer(P, Bt, T)
    idx = np.r_[0:16, 8184:8200, B - 8:B]
    buf = _oracle_buffers(data, 6, flat=False)
    vo, jo = orc.mano_forward(buf, torch.from_numpy(pose[idx]), torch.from_numpy(betas[idx]), torch.from_numpy(trans[idx]), ncomps=6)
    assert np.abs(v[idx].cpu().numpy() - vo.numpy()).max() <= 1e-3
    assert np.abs(j[idx].cpu().numpy() - jo.numpy()).max() <= 1e-3
    # sample independence (bit exact), translation additivity, empty batch
    for i in (0, 8191, 8192, B - 1):
        vi, ji = layer(P[i:i + 1], Bt[i:i + 1], T[i:i + 1])
        assert torch.equal(vi[0], v[i]) and torch.equal(ji[0], j[i]), i
    v2, j2 = layer(P[:64], Bt[:64], T[:64] + 0.25)
    assert torch.allclose(v2, v[:64] + 250.0, atol=2e-3) and torch.allclose(j2, j[:64] + 250.0, atol=2e-3)
    ve, je = layer(P[:0], Bt[:0], T[:0])
    assert ve.shape == (0, 778, 3) and je.shape == (0, 21, 3)
    # zero pose coefficients with a flat hand mean and default betas: the template, up to the regressed joints
    flat = ManoLayer(mano_data=data, ncomps=6).to(DEV)
    vz, jz = flat(torch.zeros(3, 9, device=DEV))
    assert torch.allclose(vz[0], flat.th_v_template[0] * 1000, atol=1e-3)
    assert torch.allclose(jz[0, 0], (flat.th_J_regressor @ flat.th_v_template[0])[0] * 1000, atol=1e-3)
