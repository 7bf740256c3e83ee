"""Shared test plumbing: everything in ``builders`` (golden fixtures, synthetic checkpoints, product-side builders)
plus the CPU oracle and its constants.  Only tests, ``__graft_entry__.smoke()`` and the checker legs of bench.py
(`cpu_baseline`, `parity`, `--impl reference`) import this module."""
from __future__ import annotations

import functools

import torch

from builders import *                      # noqa: F401,F403
from builders import CONFIGS, GOLDEN, ROOT, golden, key_spec, regressor, synthetic   # noqa: F401
from oracle import gator_oracle as orc      # noqa: E402


@functools.lru_cache(None)
def oracle_setup(tag: str):
    """(state_dict of torch fp32 tensors, gat constants, mdr constants, alpha) for the oracle."""
    category, alpha, regname = CONFIGS[tag]
    J, skel, flip, _ = synthetic.joint_set(category)
    mv = synthetic.mean_vertices()
    shortest, path = synthetic.floyd_warshall(J, skel)
    A, D, U = synthetic.mesh_sampling_matrices()
    gadj = orc.gat_graph_adj(J, skel, flip)
    gat_c = orc.gat_constants(J, gadj, torch.from_numpy(regressor(regname)), mv, shortest, path)
    mdr_c = orc.mdr_constants(mv, regressor('h36m'), D)
    sd = {k: torch.from_numpy(v) for k, v in synthetic.synth_state_dict(key_spec(tag)).items()}
    sd['pose_lifter.graph_adj'] = gadj
    sd['pose_lifter.init_vertices'] = gat_c['init_vertices']
    sd['pose2mesh.init_vertices'] = mdr_c['init_vertices']
    sd['pose2mesh.init_vertices_6890'] = mdr_c['init_vertices_6890']
    return sd, gat_c, mdr_c, alpha


def to_dtype(sd, dt):
    return {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}


