"""The handful of settings the reference's model code reads from its global ``core.config.cfg``
(lib/core/config.py): ``cfg.DATASET.BASE_DATA_DIR`` (GAT.py:13-14, MDR.py:16), ``cfg.MODEL.alpha``
(MDR.py:115,162), ``cfg.MODEL.posenet_pretrained/posenet_path`` (GATOR.py:13, GAT.py:130).

When the replacement runs inside the reference tree (``core.config`` already imported by the caller)
that object is used, so yaml overrides keep working; otherwise the local defaults below apply and can
be changed with :func:`configure`.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

_local = SimpleNamespace(
    DATASET=SimpleNamespace(BASE_DATA_DIR='data/base_data'),
    MODEL=SimpleNamespace(alpha=False, posenet_pretrained=False, posenet_path=''),
    root='.',   # directory that './data/base_data/*.npy' (GAT.py:89-93, mesh.py:62) is relative to
)


def get_cfg():
    mod = sys.modules.get('core.config')
    if mod is not None and hasattr(mod, 'cfg'):
        return mod.cfg
    return _local


def configure(base_data_dir: str | None = None, alpha: bool | None = None, root: str | None = None):
    """Set the defaults (also written through to the reference's cfg when that is loaded)."""
    ext = get_cfg()
    if base_data_dir is not None:
        _local.DATASET.BASE_DATA_DIR = base_data_dir
        if ext is not _local:
            ext.DATASET.BASE_DATA_DIR = base_data_dir
    if alpha is not None:
        _local.MODEL.alpha = bool(alpha)
        if ext is not _local:
            ext.MODEL.alpha = bool(alpha)
    if root is not None:
        _local.root = root


def base_data_path(name: str) -> str:
    """Resolve a base_data file the way the reference does: relative to the CWD, or to `root`."""
    cfg = get_cfg()
    rel = os.path.join(cfg.DATASET.BASE_DATA_DIR, name)
    if os.path.exists(rel):
        return rel
    return os.path.join(getattr(cfg, 'root', _local.root) if cfg is _local else _local.root, rel)
