"""Deterministic synthetic stand-ins for everything the reference downloads.

The reference needs files that are not in its repository (README.md:47-56,84):
``data/base_data/{smpl_mean_vertices.npy, J_regressor_h36m.npy,
mesh_downsampling.npz, shortest_path_{h36m,3dpw}.npy, path_{h36m,3dpw}.npy}``,
the licensed SMPL ``.pkl`` and the ``*.pth.tar`` checkpoints.  This module
synthesises all of them with numpy's PCG64 (bit-identical on every platform) so
that the golden-vector generator (run against the real reference, in the build
container), the tests, ``bench.py`` and ``smoke()`` (run on the GPU box, where
the reference does not exist) all see the same bytes.

Only shapes / dtypes / semantics are pinned by the reference code; the values
are ours.  Nothing in here is on the product path.
"""
from __future__ import annotations

import os
import zlib
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

V_FULL, V_MID, V_COARSE = 6890, 1723, 431

# Joint sets exactly as the reference's callers define them
# (demo/run.py:68-90, data/Human36M/dataset.py:52-75).
H36M_SKELETON = ((0, 7), (7, 8), (8, 9), (9, 10), (8, 11), (11, 12), (12, 13), (8, 14), (14, 15),
                 (15, 16), (0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6))
H36M_FLIP_PAIRS = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))
COCO_SKELETON = ((1, 2), (0, 1), (0, 2), (2, 4), (1, 3), (6, 8), (8, 10), (5, 7), (7, 9), (12, 14),
                 (14, 16), (11, 13), (13, 15), (17, 11), (17, 12), (17, 18), (18, 5), (18, 6), (18, 0))
COCO_FLIP_PAIRS = ((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16))
COCO_JOINTS_NAME = ('Nose', 'L_Eye', 'R_Eye', 'L_Ear', 'R_Ear', 'L_Shoulder', 'R_Shoulder', 'L_Elbow',
                    'R_Elbow', 'L_Wrist', 'R_Wrist', 'L_Hip', 'R_Hip', 'L_Knee', 'R_Knee', 'L_Ankle',
                    'R_Ankle', 'Pelvis', 'Neck')
# smpl_skeleton parents (data/Human36M/dataset.py:46-48); root never indexed.
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

NO_VIA = 510  # sentinel of path_*.npy: "no intermediate node" (modules.py:6-11,22)


def joint_set(name: str):
    """(J, skeleton, flip_pairs, base_data suffix) for 'human36' / 'coco'."""
    if name == 'human36':
        return 17, H36M_SKELETON, H36M_FLIP_PAIRS, 'h36m'
    if name == 'coco':
        return 19, COCO_SKELETON, COCO_FLIP_PAIRS, '3dpw'
    raise ValueError(name)


def _rng(tag: str, seed: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(tag.encode())]))


# ----------------------------------------------------------------------------------------------
# base_data
# ----------------------------------------------------------------------------------------------
def mean_vertices() -> np.ndarray:
    """(6890,3) f32 'template' ~ N(0, diag(0.25, 0.5, 0.12)^2) metres."""
    r = _rng('smpl_mean_vertices')
    return (r.standard_normal((V_FULL, 3)) * np.array([0.25, 0.5, 0.12])).astype(np.float32)


def joint_regressor(n_rows: int, tag: str, nnz: int = 6) -> np.ndarray:
    """(n_rows, 6890) f32, `nnz` non-zeros per row, rows sum to 1 (like the shipped regressors)."""
    r = _rng('J_regressor_' + tag)
    out = np.zeros((n_rows, V_FULL), np.float32)
    for j in range(n_rows):
        cols = r.choice(V_FULL, size=nnz, replace=False)
        w = r.random(nnz) + 0.1
        out[j, cols] = (w / w.sum()).astype(np.float32)
    return out


def floyd_warshall(J: int, skeleton: Sequence[Tuple[int, int]]):
    """Hop-count matrix and 'via' matrix in the format GAT.__init__ expects
    (GAT.py:89-93,109-110; modules.py:6-29): shortest (J,J) int64, path (J,J) int64 with
    510 where the pair has no intermediate node (diagonal, direct edges)."""
    INF = 10 ** 6
    d = np.full((J, J), INF, np.int64)
    via = np.full((J, J), NO_VIA, np.int64)
    for i in range(J):
        d[i, i] = 0
    for a, b in skeleton:
        d[a, b] = d[b, a] = 1
    for k in range(J):
        for i in range(J):
            for j in range(J):
                if d[i, k] + d[k, j] < d[i, j]:
                    d[i, j] = d[i, k] + d[k, j]
                    via[i, j] = k
    assert d.max() < 10, 'num_spatial=10 hop-embedding rows (GAT.py:112)'
    return d, via


def mesh_sampling_matrices():
    """A (3 adjacency), D (2 down: 1 nnz/row selection), U (2 up: 3 nnz/row barycentric, rows sum 1)
    as scipy.sparse, GraphCMR layout (mesh.py:50-58)."""
    import scipy.sparse as sp
    r = _rng('mesh_downsampling')
    sizes = [V_FULL, V_MID, V_COARSE]
    A, D, U = [], [], []
    for n in sizes:
        rows = np.repeat(np.arange(n), 3)
        cols = r.integers(0, n, size=3 * n)
        a = sp.coo_matrix((np.ones(3 * n, np.float32), (rows, cols)), shape=(n, n))
        a = ((a + a.T) > 0).astype(np.float32)
        A.append(sp.csc_matrix(a))
    for lvl in range(2):
        n_hi, n_lo = sizes[lvl], sizes[lvl + 1]
        keep = np.sort(r.choice(n_hi, size=n_lo, replace=False))
        D.append(sp.csc_matrix(sp.coo_matrix((np.ones(n_lo, np.float32), (np.arange(n_lo), keep)), shape=(n_lo, n_hi))))
        rows = np.repeat(np.arange(n_hi), 3)
        cols = np.stack([r.choice(n_lo, size=3, replace=False) for _ in range(n_hi)]).reshape(-1)
        w = r.random((n_hi, 3)) + 0.05
        w = (w / w.sum(1, keepdims=True)).astype(np.float32).reshape(-1)
        U.append(sp.csc_matrix(sp.coo_matrix((w, (rows, cols)), shape=(n_hi, n_lo))))
    return A, D, U


def write_base_data(root: str, regressor_h36m: np.ndarray | None = None) -> str:
    """Materialise ``<root>/data/base_data/*`` (the CWD-relative layout the reference hard-codes,
    GAT.py:89-93, mesh.py:62, MDR.py:16,72,85).  Returns the base_data dir."""
    d = os.path.join(root, 'data', 'base_data')
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, 'smpl_mean_vertices.npy'), mean_vertices())
    if regressor_h36m is None:
        regressor_h36m = joint_regressor(17, 'h36m')
    np.save(os.path.join(d, 'J_regressor_h36m.npy'), regressor_h36m.astype(np.float32))
    for name in ('human36', 'coco'):
        J, skel, _, suffix = joint_set(name)
        sp_, via = floyd_warshall(J, skel)
        np.save(os.path.join(d, f'shortest_path_{suffix}.npy'), sp_)
        np.save(os.path.join(d, f'path_{suffix}.npy'), via)
    A, D, U = mesh_sampling_matrices()

    def obj(lst):
        o = np.empty(len(lst), dtype=object)
        for i, m in enumerate(lst):
            o[i] = m
        return o
    np.savez(os.path.join(d, 'mesh_downsampling.npz'), A=obj(A), D=obj(D), U=obj(U))
    return d


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
# buffers whose value is fixed by the constructor, not by a checkpoint's randomness
CONSTRUCTED_BUFFERS = ('graph_adj', 'init_vertices', 'init_vertices_6890')


def synth_tensor(key: str, shape: Sequence[int], dtype: str, seed: int = 0) -> np.ndarray:
    """One deterministic tensor for state_dict entry `key` (distribution picked from the key name)."""
    r = _rng('w:' + key, seed)
    shape = tuple(shape)
    leaf = key.split('.')[-1]
    parent = key.split('.')[-2] if '.' in key else ''
    if dtype == 'int64':
        return np.array(7, dtype=np.int64).reshape(shape)
    n = lambda s: r.standard_normal(shape) * s
    u = lambda a: r.uniform(-a, a, size=shape)
    is_norm = ('norm' in parent) or key.endswith('GLinear.1.weight') or key.endswith('GLinear.1.bias')
    if leaf == 'running_mean':
        x = n(0.1)
    elif leaf == 'running_var':
        x = r.uniform(0.5, 1.5, size=shape)
    elif leaf == 'a_2' or (is_norm and leaf == 'weight'):
        x = 1.0 + n(0.1)
    elif leaf == 'b_2' or (is_norm and leaf == 'bias'):
        x = n(0.05)
    elif 'embed' in parent or parent == 'spatial_pos_encoder':
        x = n(0.5)
    elif parent == 'get_hop_path_encoding' and leaf == 'W':
        x = 1.0 + n(0.1)
    elif parent == 'gcn' and leaf == 'W':
        x = u(0.15)
    elif parent == 'gcn' and leaf == 'M':
        x = u(0.3)
    elif parent == 'gcn' and leaf == 'adj2':
        x = 1e-6 + n(0.05)
    elif key.endswith('lifter.weight'):
        x = u(300.0 / np.sqrt(shape[1]))          # pose3d is in millimetres (GATOR.py:19)
    elif key.endswith('lifter.bias'):
        x = u(50.0)
    elif leaf in ('bias', 'b'):
        x = u(0.05)
    elif leaf in ('weight', 'W') and len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        x = u(1.0 / np.sqrt(fan_in))
    else:
        x = n(0.1)
    return np.asarray(x, dtype=np.float32)


def synth_state_dict(spec: Iterable[Tuple[str, Sequence[int], str]], seed: int = 0) -> Dict[str, np.ndarray]:
    """spec = [(key, shape, 'float32'|'int64')]; constructed buffers are skipped."""
    out = {}
    for key, shape, dtype in spec:
        if key.split('.')[-1] in CONSTRUCTED_BUFFERS:
            continue
        out[key] = synth_tensor(key, shape, dtype, seed)
    return out


def load_synth_weights(module, seed: int = 0):
    """Fill a torch module (reference or replacement) with the synthetic checkpoint, strictly."""
    import torch
    sd = module.state_dict()
    spec = [(k, tuple(v.shape), 'int64' if v.dtype == torch.int64 else 'float32') for k, v in sd.items()]
    new = synth_state_dict(spec, seed)
    full = {k: (torch.from_numpy(new[k]).to(v.device) if k in new else v) for k, v in sd.items()}
    module.load_state_dict(full, strict=True)
    return module


# ----------------------------------------------------------------------------------------------
# model inputs
# ----------------------------------------------------------------------------------------------
def poses2d(batch: int, J: int, seed: int = 1) -> np.ndarray:
    """Per-sample standardised 2D joints (callers feed mean-0/std-1 coords per axis,
    data/Human36M/dataset.py:383-389, demo/run.py:130-133)."""
    r = _rng(f'pose2d:{J}', seed)
    x = r.standard_normal((batch, J, 2)).astype(np.float32)
    x = (x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)
    return x.astype(np.float32)


def coco_poses2d(base19: np.ndarray, batch: int, seed: int = 2, jitter: float = 0.05) -> np.ndarray:
    """Config 4/5 input: the (preprocessed) demo pose + N(0, jitter^2) per sample."""
    r = _rng('coco_jitter', seed)
    x = base19[None].astype(np.float32) + (r.standard_normal((batch, 19, 2)) * jitter).astype(np.float32)
    return x.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# SMPL-shaped body model (licensed pkl unavailable)
# ----------------------------------------------------------------------------------------------
def smpl_buffers() -> Dict[str, np.ndarray]:
    """Synthetic buffers with the shapes SMPL_Layer registers (smpl_layer.py:40-55)."""
    r = _rng('smpl_model')
    jreg = np.zeros((24, V_FULL), np.float32)
    for j in range(24):
        cols = r.choice(V_FULL, size=30, replace=False)
        w = r.random(30) + 0.1
        jreg[j, cols] = (w / w.sum()).astype(np.float32)
    weights = np.zeros((V_FULL, 24), np.float32)
    for v in range(V_FULL):
        cols = r.choice(24, size=4, replace=False)
        w = r.random(4) + 0.05
        weights[v, cols] = (w / w.sum()).astype(np.float32)
    faces = r.integers(0, V_FULL, size=(13776, 3)).astype(np.int64)
    return {
        'th_betas': np.zeros((1, 10), np.float32),
        'th_shapedirs': (r.standard_normal((V_FULL, 3, 10)) * 0.01).astype(np.float32),
        'th_posedirs': (r.standard_normal((V_FULL, 3, 207)) * 0.001).astype(np.float32),
        'th_v_template': mean_vertices()[None],
        'th_J_regressor': jreg,
        'th_weights': weights,
        'th_faces': faces,
    }


MANO_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
MANO_VERTS = 778


def mano_data() -> Dict[str, np.ndarray]:
    """Synthetic arrays with the shapes ManoLayer.__init__ reads from the MANO pkl (manolayer.py:62-103): a hand-sized
    template (~0.1 m), 10 shape and 135 pose blend shapes, 16-joint regressor and skinning weights, the 45 x 45 PCA
    basis of the finger poses and its mean."""
    r = _rng('mano_model')
    nv = MANO_VERTS
    jreg = np.zeros((16, nv), np.float32)
    for j in range(16):
        cols = r.choice(nv, size=12, replace=False)
        w = r.random(12) + 0.1
        jreg[j, cols] = (w / w.sum()).astype(np.float32)
    weights = np.zeros((nv, 16), np.float32)
    for v in range(nv):
        cols = r.choice(16, size=4, replace=False)
        w = r.random(4) + 0.05
        weights[v, cols] = (w / w.sum()).astype(np.float32)
    q, _ = np.linalg.qr(r.standard_normal((45, 45)))
    return {
        'betas': np.zeros(10, np.float32),
        'shapedirs': (r.standard_normal((nv, 3, 10)) * 0.003).astype(np.float32),
        'posedirs': (r.standard_normal((nv, 3, 135)) * 0.0005).astype(np.float32),
        'v_template': (r.standard_normal((nv, 3)) * np.array([0.04, 0.02, 0.06])).astype(np.float32),
        'J_regressor': jreg,
        'weights': weights,
        'f': r.integers(0, nv, size=(1538, 3)).astype(np.int64),
        'hands_components': (q * np.linspace(1.5, 0.05, 45)[:, None]).astype(np.float32),
        'hands_mean': (r.standard_normal(45) * 0.2).astype(np.float32),
        'kintree_table': np.stack([np.asarray([4294967295] + MANO_PARENTS[1:], dtype=np.int64), np.arange(16)]),
    }


def mano_inputs(batch: int, ncomps: int = 6, seed: int = 5):
    """pose coefficients (B, 3 + ncomps): root axis-angle U(-1,1) (row 0 all zero, row 1 with |a| > pi) and PCA
    coefficients N(0,1); betas N(0,1); trans N(0, 0.1)."""
    r = np.random.default_rng(1000 + seed)
    pose = np.concatenate([r.uniform(-1, 1, (batch, 3)), r.standard_normal((batch, ncomps))], 1).astype(np.float32)
    pose[0] = 0.0
    if batch > 1:
        pose[1, :3] = (2.1, -2.0, 1.9)
    return pose, r.standard_normal((batch, 10)).astype(np.float32), (r.standard_normal((batch, 3)) * 0.1).astype(np.float32)


def smpl_inputs(batch: int, seed: int = 3):
    """pose U(-0.2,0.2) incl. an all-zero row and a row with |a| > pi; betas N(0,1); trans N(0,1)."""
    r = _rng('smpl_inputs', seed)
    pose = r.uniform(-0.2, 0.2, size=(batch, 72)).astype(np.float32)
    if batch > 1:
        pose[1] = 0.0
    if batch > 2:
        pose[2, 3:6] = np.array([2.5, 2.0, -1.5], np.float32)   # |a| = 3.5 > pi
    betas = r.standard_normal((batch, 10)).astype(np.float32)
    trans = r.standard_normal((batch, 3)).astype(np.float32)
    return pose, betas, trans


def eval_inputs(batch: int, regressor: np.ndarray, seed: int = 7):
    """Inputs of the evaluation epilogue (lib/core/base.py:216-223): predicted / ground-truth meshes in metres
    and `reg_pose3d` in mm.  gt = template deformed per sample; pred = gt + a per-sample similarity perturbation
    + N(0, 15 mm) noise, so MPJPE, MPVPE and PA-MPJPE are all different and non-trivial."""
    r = _rng('eval_inputs', seed)
    tmpl = mean_vertices()
    gt = tmpl[None] * (1.0 + 0.1 * r.standard_normal((batch, 1, 3))) + 0.3 * r.standard_normal((batch, 1, 3))
    ang = 0.15 * r.standard_normal((batch, 3))
    pred = np.empty_like(gt)
    for b in range(batch):
        ax, ay, az = ang[b]
        Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        pred[b] = (1.0 + 0.05 * r.standard_normal()) * gt[b] @ (Rz @ Ry @ Rx).T + 0.05 * r.standard_normal(3)
    pred += 0.015 * r.standard_normal(pred.shape)
    gt = gt.astype(np.float32)
    pred = pred.astype(np.float32)
    gt_pose = (regressor.astype(np.float32) @ (gt * np.float32(1000.0))) + r.standard_normal((batch, len(regressor), 3)).astype(np.float32)
    return pred, gt, gt_pose.astype(np.float32)


def pixel_poses(base17: np.ndarray, batch: int, seed: int = 9) -> np.ndarray:
    """(batch, 17, 3) float64 detector-style inputs (x, y, confidence) derived from the shipped
    demo/coco_joint_input.npy: sample 0 is the file itself, then per-sample anisotropic scalings / shifts so that
    tall boxes (w < ar*h), wide boxes (w > ar*h) and small boxes all occur, plus pixel jitter."""
    r = _rng('pixel_poses', seed)
    out = np.repeat(np.asarray(base17, np.float64).reshape(1, 17, -1)[:, :, :3], batch, 0).copy()
    c = out[0, :, :2].mean(0)
    for b in range(1, batch):
        sx, sy = (3.0, 0.4) if b % 3 == 1 else ((0.5, 1.5) if b % 3 == 2 else (0.08, 0.08))
        sx, sy = sx * (1 + 0.2 * r.random()), sy * (1 + 0.2 * r.random())
        out[b, :, 0] = (out[b, :, 0] - c[0]) * sx + c[0] + 200 * r.standard_normal()
        out[b, :, 1] = (out[b, :, 1] - c[1]) * sy + c[1] + 200 * r.standard_normal()
        out[b, :, :2] += 2.0 * r.standard_normal((17, 2))
    return out


def camera_annotations(batch: int, seed: int = 13):
    """Human3.6M-style per-item annotations for ``get_smpl_coord`` (data/Human36M/dataset.py:254-262):
    SMPL pose (B,72) with a non-trivial root orientation, shape (B,10) incl. one implausible row (|beta| > 3),
    trans (B,3) metres, camera rotation R (B,3,3) and translation t (B,3) in millimetres."""
    r = _rng('camera_annotations', seed)
    pose = r.uniform(-0.3, 0.3, size=(batch, 72)).astype(np.float32)
    pose[:, :3] = r.uniform(-1.5, 1.5, size=(batch, 3)).astype(np.float32)
    shape = r.standard_normal((batch, 10)).astype(np.float32)
    if batch > 1:
        shape[1, 4] = 3.5                                    # dataset.py:266 resets the whole row to 0
    trans = r.standard_normal((batch, 3)).astype(np.float32)
    q = r.standard_normal((batch, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(batch, 3, 3)
    t = (r.standard_normal((batch, 3)) * np.array([500.0, 500.0, 1500.0]) + np.array([0.0, 0.0, 4500.0]))
    return pose, shape, trans, R.astype(np.float32), t.astype(np.float32)
