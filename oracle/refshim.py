"""TEST INFRASTRUCTURE ONLY - imports the *unmodified* reference from /root/reference on CPU.

Used by ``tests/golden/make_golden.py`` (to produce committed golden vectors) and by the
``needs_reference`` tests, in the build container only: /root/reference does not exist on the
GPU box.  Nothing under ``gator_b200/`` may import this file.

Shims; none of them touches arithmetic on the hot path:
  * sys.path += reference {lib, data, smplpytorch}          (main/__init_path.py:16-25)
  * stub modules: easydict, timm (DropPath = identity in eval, Mlp = fc1/act/drop/fc2/drop),
    matplotlib; ``core.config`` is pre-seeded because importing the real one mkdir's inside the
    read-only tree (lib/core/config.py:21-39)
  * ``.cuda()`` -> identity and ``Mesh(device='cpu')`` because this container has no GPU
  * ``MDR.vj_relation`` is a float64 numpy array (graph_utils.py:78) which current torch refuses
    as an index; it is cast to int64 after construction (value-preserving)
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REF = os.environ.get('GATOR_REFERENCE', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'lib', 'models'))


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


_cfg = None


def install_shims(alpha: bool = False):
    """Idempotent; returns the stub cfg."""
    global _cfg
    import torch
    import torch.nn as nn
    if _cfg is not None:
        _cfg.MODEL.alpha = alpha
        return _cfg
    for p in ('lib', 'data', 'smplpytorch'):
        sys.path.insert(0, os.path.join(REF, p))

    ed = types.ModuleType('easydict')
    ed.EasyDict = _AttrDict
    sys.modules['easydict'] = ed

    class DropPath(nn.Module):
        def __init__(self, p=0.):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training
            return x

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)
            self.drop = nn.Dropout(drop)

        def forward(self, x):
            return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))

    timm = types.ModuleType('timm')
    tm = types.ModuleType('timm.models')
    tl = types.ModuleType('timm.models.layers')
    tv = types.ModuleType('timm.models.vision_transformer')
    tl.DropPath = DropPath
    tv.Mlp = Mlp
    timm.models = tm
    tm.layers = tl
    tm.vision_transformer = tv
    sys.modules.update({'timm': timm, 'timm.models': tm, 'timm.models.layers': tl,
                        'timm.models.vision_transformer': tv})
    for name in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))

    cfg = _AttrDict()
    cfg.DATASET = _AttrDict(BASE_DATA_DIR='data/base_data')
    cfg.MODEL = _AttrDict(alpha=alpha, posenet_pretrained=False, posenet_path='', input_shape=(384, 288))
    cfg.data_dir = os.path.join(REF, 'data')
    cfg.smpl_dir = os.path.join(REF, 'smplpytorch')
    core = types.ModuleType('core')
    core.__path__ = [os.path.join(REF, 'lib', 'core')]
    cc = types.ModuleType('core.config')
    cc.cfg = cfg
    core.config = cc
    sys.modules['core'] = core
    sys.modules['core.config'] = cc

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    _cfg = cfg
    return cfg


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def build_gator(root: str, joint_category: str, alpha: bool, regressor):
    """Construct the reference GATOR (GATOR.py:24-27) with CWD=root (holding data/base_data).
    `regressor` is the (J_t, 6890) float32 J-regressor the caller would pass (base.py:53)."""
    import numpy as np
    import scipy.sparse
    import torch
    cfg = install_shims(alpha)
    cfg.MODEL.alpha = alpha
    from gator_b200 import synthetic
    J, skel, flip, _ = synthetic.joint_set(joint_category)
    with chdir(root):
        import graph_utils
        from models.backbones import mesh as ref_mesh
        if not torch.cuda.is_available():
            ref_mesh.Mesh.__init__.__defaults__ = ('data/base_data/mesh_downsampling.npz', 1, 1, torch.device('cpu'))
        from models import GATOR as ref_GATOR
        graph_adj = [scipy.sparse.csr_matrix(graph_utils.build_adj(J, skel, flip))]
        model = ref_GATOR.get_model(J, embed_dim=128, depth=6, graph_adj=graph_adj, GCN_depth=1,
                                    J_regressor=torch.as_tensor(np.asarray(regressor, dtype=np.float32)))
    model.pose2mesh.vj_relation = np.asarray(model.pose2mesh.vj_relation).astype(np.int64)
    model.eval()
    return model


def build_smpl_layer(buffers, center_idx=None):
    """Reference SMPL_Layer with __init__ bypassed (it needs chumpy + the licensed pkl,
    smpl_layer.py:15-63); the reference ``forward`` then runs unmodified."""
    import torch
    from gator_b200 import synthetic
    install_shims()
    ser = types.ModuleType('smplpytorch.native.webuser.serialization')
    ser.ready_arguments = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('pkl loader unavailable'))
    sys.modules.setdefault('smplpytorch.native.webuser.serialization', ser)
    from smplpytorch.pytorch.smpl_layer import SMPL_Layer
    layer = SMPL_Layer.__new__(SMPL_Layer)
    torch.nn.Module.__init__(layer)
    layer.center_idx = center_idx
    layer.gender = 'neutral'
    for k, v in buffers.items():
        layer.register_buffer(k, torch.as_tensor(v))
    layer.kintree_parents = [4294967295] + synthetic.SMPL_PARENTS[1:]
    layer.num_joints = 24
    return layer.eval()


def build_mano_layer(data, center_idx=None, ncomps=6, use_pca=True, side='right', flat_hand_mean=True):
    """Reference ManoLayer with __init__ bypassed (it needs chumpy + the licensed MANO pkl, manolayer.py:62-103): the
    buffers and attributes __init__ would set are filled from `data` (gator_b200.synthetic.mano_data), then the
    reference ``forward`` runs unmodified.  `mano.webuser.smpl_handpca_wrapper_HAND_only` (imports chumpy at module
    level) is stubbed - only its ``ready_arguments`` name is needed for the import of manolayer.py to succeed."""
    import numpy as np
    import torch
    install_shims()
    if os.path.join(REF, 'manopth') not in sys.path:
        sys.path.insert(0, os.path.join(REF, 'manopth'))
    for name in ('mano', 'mano.webuser', 'mano.webuser.smpl_handpca_wrapper_HAND_only'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            m.ready_arguments = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('pkl loader unavailable'))
            sys.modules[name] = m
    from manopth.manolayer import ManoLayer
    layer = ManoLayer.__new__(ManoLayer)
    torch.nn.Module.__init__(layer)
    layer.center_idx, layer.robust_rot, layer.rot = center_idx, False, 3
    layer.flat_hand_mean, layer.side, layer.use_pca = flat_hand_mean, side, use_pca
    layer.joint_rot_mode = layer.root_rot_mode = 'axisang'
    layer.ncomps = ncomps if use_pca else 45
    T = lambda a: torch.Tensor(np.asarray(a, dtype=np.float32))
    layer.register_buffer('th_betas', T(data['betas']).unsqueeze(0))
    layer.register_buffer('th_shapedirs', T(data['shapedirs']))
    layer.register_buffer('th_posedirs', T(data['posedirs']))
    layer.register_buffer('th_v_template', T(data['v_template']).unsqueeze(0))
    layer.register_buffer('th_J_regressor', T(data['J_regressor']))
    layer.register_buffer('th_weights', T(data['weights']))
    layer.register_buffer('th_faces', torch.as_tensor(np.asarray(data['f']).astype(np.int32)).long())
    mean = np.zeros(45, np.float32) if flat_hand_mean else np.asarray(data['hands_mean'], np.float32)
    layer.register_buffer('th_hands_mean', T(mean).unsqueeze(0))
    layer.register_buffer('th_selected_comps', T(np.asarray(data['hands_components'])[:ncomps]))
    layer.kintree_table = np.asarray(data['kintree_table'])
    layer.kintree_parents = list(layer.kintree_table[0].tolist())
    return layer.eval()


def ref_mesh(root: str):
    import torch
    install_shims()
    with chdir(root):
        from models.backbones import mesh as m
        return m.Mesh(device=torch.device('cpu'))


def dataset_class(name: str = 'Human36M'):
    """The reference dataset class (data/<name>/dataset.py) without constructing it: its evaluation methods
    (`compute_both_err`, `evaluate_joint`) only read `self.human36_eval_joint` / `self.datalist`, so they can be
    called unbound on a stand-in object.  Import-time dependencies that are absent here and unused by those
    methods are stubbed: pycocotools, vis (needs mpl_toolkits).  transforms3d (absent, unpinned in requirements.sh)
    is provided by the oracle's restatement of axangle2mat / mat2axangle."""
    install_shims()
    if 'transforms3d' not in sys.modules:
        # third-party, absent here: its two functions the datasets call are restated in the oracle
        from oracle import gator_oracle as orc
        t3 = types.ModuleType('transforms3d')
        t3.axangles = types.SimpleNamespace(axangle2mat=orc.axangle2mat, mat2axangle=orc.mat2axangle)
        sys.modules['transforms3d'] = t3
    for mod in ('pycocotools', 'pycocotools.coco', 'vis', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if mod not in sys.modules:
            m = types.ModuleType(mod)
            m.COCO = object
            m.vis_2d_pose = m.vis_3d_pose = m.Axes3D = None
            sys.modules[mod] = m
    import importlib
    return getattr(importlib.import_module(f'{name}.dataset'), name)
