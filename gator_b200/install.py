"""Rebind the reference's hot-path classes to the B200 replacements without touching its call sites.

    import gator_b200; gator_b200.install()      # after main/__init_path.py put lib/ etc. on sys.path

After this ``models.GATOR.get_model`` (lib/core/base.py:57, demo/run.py:96), ``models.GAT.get_model``
(base.py:59), ``models.MDR.get_model``, ``smplpytorch.pytorch.smpl_layer.SMPL_Layer`` (lib/smpl.py:8) and
``models.backbones.mesh.Mesh`` and ``manopth.manolayer.ManoLayer`` resolve to gator_b200's modules; constructor signatures, state_dict keys
and return values are identical, so ``model.load_state_dict(checkpoint['model_state_dict'])`` works.
"""
from __future__ import annotations

import importlib
import sys


def install(verbose: bool = False):
    from . import mano_layer as b_mano, mesh as b_mesh, smpl_layer as b_smpl
    from .models import GAT as b_GAT, GATOR as b_GATOR, MDR as b_MDR
    done = []

    def rebind(mod_name, attrs):
        try:
            mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        except Exception:
            return
        for name, obj in attrs.items():
            setattr(mod, name, obj)
            done.append(f'{mod_name}.{name}')

    rebind('models.GATOR', {'GATOR': b_GATOR.GATOR, 'get_model': b_GATOR.get_model})
    rebind('models.GAT', {'GAT': b_GAT.GAT, 'get_model': b_GAT.get_model})
    rebind('models.MDR', {'MDR': b_MDR.MDR, 'get_model': b_MDR.get_model})
    rebind('models.backbones.mesh', {'Mesh': b_mesh.Mesh})
    rebind('smplpytorch.pytorch.smpl_layer', {'SMPL_Layer': b_smpl.SMPL_Layer})
    rebind('smpl', {'SMPL_Layer': b_smpl.SMPL_Layer})
    rebind('manopth.manolayer', {'ManoLayer': b_mano.ManoLayer})
    if verbose:
        print('gator_b200.install: rebound', ', '.join(done))
    return done
