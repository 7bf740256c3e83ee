"""Init-time graph constants of the reference constructors (host side, numpy; run once per model).

file:line references are into the reference tree.
"""
from __future__ import annotations

import math

import numpy as np
import torch

NO_VIA = 510   # modules.py:6-11,22


def build_adj(joint_num, skeleton, flip_pairs) -> np.ndarray:
    """graph_utils.py:60-69."""
    adj = np.zeros((joint_num, joint_num))
    for a, b in skeleton:
        adj[a, b] = adj[b, a] = 1
    for a, b in flip_pairs:
        adj[a, b] = adj[b, a] = 1
    return adj + np.eye(joint_num)


def dense_graph_adj(graph_adj) -> torch.Tensor:
    """GAT.py:57-65: last level of the coarsening list -> dense fp32, minus the hard-coded H36M flip
    edges (applied to every joint set, as the reference does)."""
    m = graph_adj[-1]
    m = m.toarray() if hasattr(m, 'toarray') else np.asarray(m)
    g = torch.from_numpy(np.asarray(m, dtype=np.float32)).clone()
    for a, b in ((1, 4), (2, 5), (3, 6), (11, 14), (12, 15), (13, 16)):
        g[a, b] = 0
        g[b, a] = 0
    return g


def template_joints(J_regressor: torch.Tensor, init_vertices: torch.Tensor, num_joint: int) -> torch.Tensor:
    """GAT.py:76-88 (pelvis / neck midpoints appended for the 19-joint COCO set)."""
    tj = torch.matmul(J_regressor[None, :, :].float(), init_vertices).squeeze(0)
    if num_joint == 19:
        pelvis = ((tj[11] + tj[12]) * 0.5).reshape(1, -1)
        neck = ((tj[5] + tj[6]) * 0.5).reshape(1, -1)
        tj = torch.cat((tj, pelvis, neck), dim=0)
    return tj


def edge_lengths(graph_adj: torch.Tensor, tj: torch.Tensor) -> torch.Tensor:
    """GAT.py:96-107: bone length for every upper-triangle edge of graph_adj."""
    J = graph_adj.shape[0]
    out = torch.zeros(J, J)
    for i in range(J):
        for j in range(i + 1, J):
            if graph_adj[i][j] == 1:
                out[i][j] = math.sqrt(((tj[i] - tj[j]) ** 2).sum(0))
    return out


def _via_chain(path, i, j):
    k = int(path[i][j])
    if k == NO_VIA:
        return []
    return _via_chain(path, i, k) + [k] + _via_chain(path, k, j)


def path_edge_features(max_dist: int, path: np.ndarray, edge_feat: torch.Tensor) -> torch.Tensor:
    """modules.py:13-29 (gen_edg_input): (J,J,max_dist) bone lengths along the shortest path."""
    n = path.shape[0]
    out = torch.zeros(n, n, max_dist)
    for i in range(n):
        for j in range(n):
            if i == j or path[i][j] == NO_VIA:
                continue
            nodes = [i] + _via_chain(path, i, j) + [j]
            for k in range(len(nodes) - 1):
                out[i, j, k] = edge_feat[nodes[k], nodes[k + 1]]
    return out


def nearest_joint(joints: np.ndarray, vertices: np.ndarray) -> np.ndarray:
    """graph_utils.py:71-89 (first minimum wins, like np.argmin)."""
    d = ((vertices[:, None, :] - joints[None, :, :]) ** 2).sum(2)
    return np.argmin(d, axis=1).astype(np.int64)


def to_csr(m):
    """scipy sparse -> (rowptr int32, colidx int32, values float32, shape)."""
    import scipy.sparse
    c = scipy.sparse.csr_matrix(m)
    c.sort_indices()
    return (c.indptr.astype(np.int32), c.indices.astype(np.int32), c.data.astype(np.float32), c.shape)


def csr_to_ell4(rowptr: np.ndarray, colidx: np.ndarray, values: np.ndarray):
    """CSR -> ELL records of four (col, val) per row for gator_mesh_upsample2 (include/gator_b200.h): (width, col (rows,4)
    int32, val (rows,4) float32) with the non-zeros in CSR order, rows with fewer than four padded with val = 0 on a
    column the row already uses; None when a row is empty or has more than four non-zeros."""
    nnz = np.diff(rowptr)
    if len(nnz) == 0 or nnz.min() < 1 or nnz.max() > 4:
        return None
    k = np.arange(4)[None, :]
    valid = k < nnz[:, None]
    pos = rowptr[:-1][:, None] + np.where(valid, k, 0)
    col = np.ascontiguousarray(colidx[pos]).astype(np.int32)
    val = np.where(valid, values[pos], 0).astype(np.float32)
    return int(nnz.max()), col, val


def dense_to_csr(a: np.ndarray):
    import scipy.sparse
    return to_csr(scipy.sparse.csr_matrix(np.asarray(a)))
