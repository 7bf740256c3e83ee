"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes mirrors match the compiled structs, the replacement modules keep the reference's
state_dict contract, and nothing silently falls back to the CPU."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT, CONFIGS, build_b200_gator, build_b200_smpl, golden, key_spec, synthetic


@pytest.fixture(scope='module')
def lib():
    from gator_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'gator_b200.h')).read()
    declared = set(re.findall(r'\b(gator_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), name
    from gator_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_struct_mirrors_and_slot_names(lib):
    from gator_b200 import _lib
    g, b = _lib.slot_names('gat')
    assert g[0] == "EMB_W1" and g[-1] == "CHAIN_PRM" and len(g) == 16 and len(b) == 21
    g, l = _lib.slot_names('mdr')
    assert g[-2:] == ["CHAIN_FINAL", "UP_W_WIDE"] and len(g) == 16 and len(l) == 19
    assert lib.gator_gat_workspace_bytes(64, 17, 0) > 0
    assert lib.gator_mdr_workspace_bytes(64, 17, 0) > 0
    assert lib.gator_smpl_workspace_bytes(64) > 0


def test_error_reporting_without_gpu(lib):
    from gator_b200 import _lib
    a = _lib.GemmArgs(M=4, N=4, K=3, lda=3, ldw=3, ldc=4)       # K not a multiple of 4, null pointers
    assert lib.gator_gemm(a, None) == -1
    assert b'null' in lib.gator_last_error() or b'multiple' in lib.gator_last_error()
    assert lib.gator_gat_forward(None, None) == -1


@pytest.mark.parametrize('tag', ['h36m', 'coco'])
def test_state_dict_contract(tag):
    """Same keys, order, shapes and dtypes as the reference model (golden key list) -> strict load works."""
    m = build_b200_gator(tag)
    keys = [(k, tuple(v.shape), str(v.dtype).replace('torch.', '')) for k, v in m.state_dict().items()]
    assert keys == key_spec(tag)
    g = golden('gator')
    assert (m.pose2mesh.vj_relation == g[f'{tag}/vj_relation']).all()
    assert (m.pose_lifter.graph_adj.numpy() == g[f'{tag}/graph_adj']).all()
    assert (m.pose_lifter.get_hop_path_encoding.edg_adj.numpy() == g[f'{tag}/edge_input']).all()
    assert (m.pose2mesh.init_vertices.numpy() == g[f'{tag}/init_vertices_431']).all()
    # a checkpoint round trip through torch.save, as funcs_utils.save/load_checkpoint do
    import io
    buf = io.BytesIO()
    torch.save({'model_state_dict': m.state_dict()}, buf)
    buf.seek(0)
    m.load_state_dict(torch.load(buf)['model_state_dict'], strict=True)
    assert m.pose_lifter._packed is None and m.pose2mesh._packed is None


def test_no_cpu_fallback():
    m = build_b200_gator('h36m')
    x = torch.from_numpy(synthetic.poses2d(2, 17))
    with pytest.raises(RuntimeError):
        m(x)
    with pytest.raises(NotImplementedError):
        m.train()(x)
    s = build_b200_smpl()
    with pytest.raises(RuntimeError):
        s(torch.zeros(1, 72))


def test_product_and_tools_never_touch_the_oracle():
    """The oracle is test infrastructure: nothing under gator_b200/ or tools/ may import it, and bench.py only inside
    its checker legs (cpu_baseline / parity / --impl reference)."""
    import glob
    import os
    import re
    from builders import ROOT
    pat = re.compile(r'^\s*(from|import)\s+(oracle|helpers)\b|gator_oracle|refshim', re.M)
    files = glob.glob(os.path.join(ROOT, 'gator_b200', '**', '*.py'), recursive=True) + glob.glob(os.path.join(ROOT, 'tools', '*.py'))
    assert files
    for f in files:
        assert not pat.search(open(f).read()), f
    bench = open(os.path.join(ROOT, 'bench.py')).read()
    top_level = [l for l in bench.splitlines() if re.match(r'(from|import)\s+(oracle|helpers)\b', l)]
    assert not top_level, top_level
    for f in glob.glob(os.path.join(ROOT, 'gator_b200', 'csrc', '*')):
        assert 'oracle' not in open(f, errors='ignore').read().lower() or f.endswith('.md'), f


def test_unsupported_configurations_fail_loudly():
    from gator_b200.models import GAT
    with pytest.raises(NotImplementedError):
        GAT.GAT(num_joint=17, embed_dim=256, depth=4, graph_adj=[np.eye(17)], J_regressor=torch.zeros(17, 6890))


@pytest.mark.needs_reference
def test_install_rebinds_reference_modules():
    import tempfile
    from oracle import refshim
    import gator_b200
    refshim.install_shims()
    root = tempfile.mkdtemp()
    synthetic.write_base_data(root)
    with refshim.chdir(root):
        import models.GATOR as ref_GATOR          # the reference module
        orig = ref_GATOR.GATOR
        try:
            done = gator_b200.install()
            from gator_b200.models import GATOR as b
            assert ref_GATOR.GATOR is b.GATOR and 'models.GATOR.get_model' in done
        finally:
            ref_GATOR.GATOR = orig
            import importlib
            for name in ('models.GATOR', 'models.GAT', 'models.MDR', 'models.backbones.mesh'):
                importlib.reload(importlib.import_module(name))


@pytest.mark.needs_reference
def test_adjmat_equals_reference():
    """gator_b200.mesh.adjmat_sparse (CSR pattern + row counts) against the reference's lil/coo construction
    (lib/models/backbones/mesh.py:29-48) on every synthetic mesh level, nsize 1 and 2: same pattern, same values."""
    import tempfile
    from oracle import refshim
    from gator_b200.mesh import adjmat_sparse
    refshim.install_shims()
    root = tempfile.mkdtemp()
    synthetic.write_base_data(root)
    with refshim.chdir(root):
        from models.backbones import mesh as ref_mesh
        A, _, _ = synthetic.mesh_sampling_matrices()
        for a in A:
            for nsize in (1, 2):
                mine, ref = adjmat_sparse(a, nsize).coalesce(), ref_mesh.adjmat_sparse(a, nsize).coalesce()
                assert torch.equal(mine.indices(), ref.indices()) and torch.equal(mine.values(), ref.values())


def test_ell_records_of_the_upsampling_operators():
    """graph.csr_to_ell4: the ELL form gator_mesh_upsample2 reads reproduces the CSR operator exactly (a dense product
    through the records equals the scipy product), padding has weight 0 on a column of the same row, and operators it
    cannot hold are refused."""
    import scipy.sparse
    from gator_b200 import graph
    _, D, U = synthetic.mesh_sampling_matrices()
    rng = np.random.default_rng(0)
    for m in list(U) + list(D):
        rp, ci, va, shape = graph.to_csr(m)
        w, col, val = graph.csr_to_ell4(rp, ci, va)
        assert 1 <= w <= 4 and col.shape == val.shape == (shape[0], 4) and col.dtype == np.int32 and val.dtype == np.float32
        x = rng.standard_normal((shape[1], 3)).astype(np.float32)
        assert np.abs((val[:, :, None] * x[col]).sum(1) - scipy.sparse.csr_matrix(m).astype(np.float32) @ x).max() <= 1e-6
        nnz = np.diff(rp)
        pad = np.arange(4)[None, :] >= nnz[:, None]
        assert (val[pad] == 0).all() and (col[pad] == np.broadcast_to(col[:, :1], col.shape)[pad]).all()
    dense5 = scipy.sparse.csr_matrix(np.ones((3, 5), np.float32))
    assert graph.csr_to_ell4(*graph.to_csr(dense5)[:3]) is None              # five non-zeros per row
    empty_row = scipy.sparse.csr_matrix(np.array([[1.0, 0.0], [0.0, 0.0]], np.float32))
    assert graph.csr_to_ell4(*graph.to_csr(empty_row)[:3]) is None


def test_new_entry_points_validate_before_touching_the_gpu(lib):
    """gator_lbs_forward / gator_mesh_upsample2 / gator_mano_post / gator_mano_pose reject bad arguments with a status
    code and a message (no CUDA call is made before validation, so this runs without a GPU)."""
    from gator_b200 import _lib
    a = _lib.LbsArgs(batch=4, n_verts=777, n_joints=16, k_blend=148)                 # odd vertex count
    assert lib.gator_lbs_forward(a, None) == -1 and b'even vertex count' in lib.gator_last_error()
    a = _lib.LbsArgs(batch=4, n_verts=778, n_joints=16, k_blend=145)                 # k_blend not padded to a multiple of 4
    assert lib.gator_lbs_forward(a, None) == -1
    a = _lib.LbsArgs(batch=0, n_verts=778, n_joints=16, k_blend=148)                 # empty batch is fine
    assert lib.gator_lbs_forward(a, None) == 0
    assert lib.gator_lbs_workspace_bytes(8, 778, 16) > 0 and lib.gator_lbs_workspace_bytes(8, 777, 16) == 0
    u = _lib.Upsample2Args(batch=2, cols=431, rows1=1723, rows2=6890, width1=5, width2=3, scale=1.0)
    assert lib.gator_mesh_upsample2(u, None) == -1 and b'ELL width' in lib.gator_last_error()
    u = _lib.Upsample2Args(batch=2, cols=4000, rows1=9000, rows2=20000, width1=3, width2=3, scale=1.0)
    assert lib.gator_mesh_upsample2(u, None) == -1                                   # null buffers / levels too large
    u = _lib.Upsample2Args(batch=0, cols=431, rows1=1723, rows2=6890, width1=3, width2=3, scale=1.0)
    assert lib.gator_mesh_upsample2(u, None) == 0
    q = _lib.ManoPostArgs(batch=1, n_verts=778, center_idx=21)                       # 21 joints: valid indices are 0..20
    assert lib.gator_mano_post(q, None) == -1
    assert lib.gator_mano_pose(None, 9, 46, None, None, None, 1, None) == -1        # more than 45 PCA components
