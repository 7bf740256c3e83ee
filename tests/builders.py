"""Golden fixtures and product-side builders shared by tests, tools and bench.py.

Nothing here imports ``oracle/``: tools and the GPU arm of bench.py build models through this module, the checker
(``tests/helpers.py``) adds the oracle on top."""
from __future__ import annotations

import functools
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gator_b200 import synthetic          # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


@functools.lru_cache(None)
def golden(name: str):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


@functools.lru_cache(None)
def regressor(name: str) -> np.ndarray:
    """The reference's shipped 17x6890 J-regressors (stored sparsely in fixtures.npz), as float32."""
    fx = golden('fixtures')
    a = np.zeros((17, synthetic.V_FULL), np.float64)
    a[fx[f'J_regressor_{name}/row'], fx[f'J_regressor_{name}/col']] = fx[f'J_regressor_{name}/val']
    return a.astype(np.float32)


CONFIGS = {
    # tag: (joint category, alpha, regressor name)
    'h36m': ('human36', False, 'h36m'),
    'coco': ('coco', True, 'coco'),
}


def key_spec(tag: str):
    spec = []
    for s in golden('gator')[f'{tag}/keys']:
        k, shape, dt = str(s).split('|')
        spec.append((k, tuple(int(x) for x in shape.split(',')) if shape else (), dt))
    return spec


# ---------------------------------------------------------------------------------------------
# product-side builders (gator_b200 modules); the CUDA library is only touched at forward time
# ---------------------------------------------------------------------------------------------
_BASE_ROOT = None


def base_data_root() -> str:
    """Synthetic data/base_data tree (written once per process) + gator_b200.config pointed at it."""
    global _BASE_ROOT
    import tempfile
    from gator_b200 import config
    if _BASE_ROOT is None:
        _BASE_ROOT = tempfile.mkdtemp(prefix='gator_base_')
        synthetic.write_base_data(_BASE_ROOT, regressor('h36m'))
    config.configure(root=_BASE_ROOT)
    return _BASE_ROOT


def build_b200_gator(tag: str, device=None):
    import scipy.sparse
    from gator_b200 import config, graph
    from gator_b200.models import GATOR
    base_data_root()
    category, alpha, regname = CONFIGS[tag]
    config.configure(alpha=alpha)
    J, skel, flip, _ = synthetic.joint_set(category)
    adj = [scipy.sparse.csr_matrix(graph.build_adj(J, skel, flip))]
    model = GATOR.get_model(J, 128, 6, adj, 1, torch.from_numpy(regressor(regname)))
    synthetic.load_synth_weights(model)
    model.eval()
    if device is not None:
        model = model.to(device)
    return model


def build_b200_smpl(center_idx=None, device=None):
    from gator_b200.smpl_layer import SMPL_Layer
    layer = SMPL_Layer.from_buffers(synthetic.smpl_buffers(), synthetic.SMPL_PARENTS, center_idx=center_idx).eval()
    return layer.to(device) if device is not None else layer
