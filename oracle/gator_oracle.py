"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's pose->mesh forward.

A plain functional restatement (torch CPU ops, fp32 or fp64) of the reference algorithm, written
from the reference sources so the checker can travel to the GPU box, where /root/reference does
not exist.  It is pinned against the real reference by ``tests/golden/*.npz`` (produced by
``tests/golden/make_golden.py`` from the unmodified reference modules) - see
``tests/test_oracle_golden.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product (``gator_b200/``)
never does.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

NO_VIA = 510


# ----------------------------------------------------------------------------------------------
# init-time constants (reference constructors)
# ----------------------------------------------------------------------------------------------
def build_adj(joint_num: int, skeleton, flip_pairs) -> np.ndarray:
    """lib/graph_utils.py:60-69  skeleton U flip_pairs U I."""
    a = np.zeros((joint_num, joint_num))
    for (i, j) in skeleton:
        a[i, j] = 1
        a[j, i] = 1
    for (i, j) in flip_pairs:
        a[i, j] = 1
        a[j, i] = 1
    return a + np.eye(joint_num)


def gat_graph_adj(joint_num: int, skeleton, flip_pairs) -> torch.Tensor:
    """lib/models/GAT.py:57-65: dense adjacency with the hard-coded H36M flip edges removed
    (applied to both joint sets)."""
    g = torch.from_numpy(build_adj(joint_num, skeleton, flip_pairs).astype(np.float32)).clone()
    for (i, j) in ((1, 4), (2, 5), (3, 6), (11, 14), (12, 15), (13, 16)):
        g[i, j] = 0
        g[j, i] = 0
    return g


def _all_edges(path, i, j):
    """lib/models/backbones/modules.py:6-11."""
    k = int(path[i][j])
    if k == NO_VIA:
        return []
    return _all_edges(path, i, k) + [k] + _all_edges(path, k, j)


def gen_edg_input(max_dist: int, path: np.ndarray, edge_feat: torch.Tensor) -> torch.Tensor:
    """lib/models/backbones/modules.py:13-29."""
    n = path.shape[0]
    out = torch.zeros(n, n, max_dist)
    for i in range(n):
        for j in range(n):
            if i == j or path[i][j] == NO_VIA:
                continue
            p = [i] + _all_edges(path, i, j) + [j]
            for k in range(len(p) - 1):
                out[i, j, k] = edge_feat[p[k], p[k + 1]]
    return out


def gat_constants(joint_num: int, graph_adj: torch.Tensor, J_regressor: torch.Tensor,
                  mean_vertices: np.ndarray, shortest: np.ndarray, path: np.ndarray) -> Dict:
    """lib/models/GAT.py:74-112 (template joints, pelvis/neck for J=19, edge lengths, path tensor)."""
    init_vertices = torch.from_numpy(mean_vertices).unsqueeze(0)
    tj = torch.matmul(J_regressor[None, :, :], init_vertices).squeeze(0)
    if joint_num == 19:
        pelvis = ((tj[11, :] + tj[12, :]) * 0.5).reshape(1, -1)
        neck = ((tj[5, :] + tj[6, :]) * 0.5).reshape(1, -1)
        tj = torch.cat((tj, pelvis, neck), dim=0)
    edg_d = torch.zeros(joint_num, joint_num)
    for i in range(joint_num):
        for j in range(i + 1, joint_num):
            if graph_adj[i][j] == 1:
                edg_d[i][j] = math.sqrt(((tj[i] - tj[j]) ** 2).sum(0))
    max_dist = int(np.amax(shortest))
    edge_input = gen_edg_input(max_dist, path, edg_d)
    return {'J': joint_num, 'graph_adj': graph_adj, 'init_vertices': init_vertices,
            'edge_input': edge_input, 'spatial_pos': torch.from_numpy(shortest).long()}


def verts_joints_relation(joints: np.ndarray, vertices: np.ndarray) -> np.ndarray:
    """lib/graph_utils.py:71-89: index of the nearest joint per vertex (argmin: first minimum)."""
    out = np.zeros(vertices.shape[0], np.int64)
    for idx, v in enumerate(vertices):
        d = ((v - joints) ** 2).sum(1)
        out[idx] = int(np.argmin(d))
    return out


def mdr_constants(mean_vertices: np.ndarray, regressor_h36m: np.ndarray, D) -> Dict:
    """lib/models/MDR.py:77-90.  D = [D0 (1723x6890), D1 (431x1723)] scipy sparse."""
    v0 = torch.from_numpy(mean_vertices)
    v1 = mesh_spmm(D[0], v0)
    v2 = mesh_spmm(D[1], v1)
    jt = torch.matmul(torch.from_numpy(regressor_h36m.astype(np.float32)), v0)
    vj = verts_joints_relation(jt.numpy(), v2.numpy())
    return {'init_vertices': v2, 'init_vertices_6890': v0, 'vj_relation': vj}


# ----------------------------------------------------------------------------------------------
# Mesh re-sampling
# ----------------------------------------------------------------------------------------------
def mesh_spmm(M, x: torch.Tensor) -> torch.Tensor:
    """lib/models/backbones/graph_layers.py:105-124 / mesh.py:9-26: torch sparse COO (float32) @ dense."""
    import scipy.sparse
    m = scipy.sparse.coo_matrix(M)
    idx = torch.from_numpy(np.array([m.row, m.col])).long()
    val = torch.from_numpy(m.data.astype(np.float32)).to(x.dtype)
    sp = torch.sparse_coo_tensor(idx, val, m.shape, check_invariants=False)
    return torch.matmul(sp, x)


def mesh_downsample(D, x: torch.Tensor, n1=0, n2=1) -> torch.Tensor:
    """lib/models/backbones/mesh.py:93-108."""
    if x.ndimension() < 3:
        for i in range(n1, n2):
            x = mesh_spmm(D[i], x)
        return x
    out = []
    for b in range(x.shape[0]):
        y = x[b]
        for i in range(n1, n2):
            y = mesh_spmm(D[i], y)
        out.append(y)
    return torch.stack(out, 0)


def mesh_upsample(U, x: torch.Tensor, n1=1, n2=0) -> torch.Tensor:
    """lib/models/backbones/mesh.py:110-123."""
    if x.ndimension() < 3:
        for i in reversed(range(n2, n1)):
            x = mesh_spmm(U[i], x)
        return x
    out = []
    for b in range(x.shape[0]):
        y = x[b]
        for i in reversed(range(n2, n1)):
            y = mesh_spmm(U[i], y)
        out.append(y)
    return torch.stack(out, 0)


# ----------------------------------------------------------------------------------------------
# GAT
# ----------------------------------------------------------------------------------------------
def hop_path_encoding(sd, p, c, num_heads=8):
    """lib/models/backbones/modules.py:77-107."""
    J = c['J']
    dt = sd[p + 'W'].dtype
    spatial_pos = c['spatial_pos']
    ones = torch.ones_like(spatial_pos)
    spatial = spatial_pos - ones
    spatial = torch.where(spatial > 0, spatial, ones)
    spatial = 1.0 / spatial.expand(num_heads, -1, -1)
    spb = F.embedding(spatial_pos, sd[p + 'spatial_pos_encoder.weight']).permute(2, 0, 1)
    e = c['edge_input'].to(dt).permute(2, 0, 1)
    e = F.linear(e.reshape(-1, J * J), sd[p + 'edge_encoder.weight'], sd[p + 'edge_encoder.bias'])
    e = e.reshape(-1, num_heads, J, J).permute(1, 2, 3, 0)
    eb = torch.mul(sd[p + 'W'], e).sum(-1)
    eb = torch.mul(eb, spatial.to(dt))
    return spb + eb


def gat_block(sd, p, c, x, bias, num_heads=8):
    """lib/models/GAT.py:33-43 with Attention (modules.py:121-138), MGCN (:243-255),
    X_Feat (:158-177) and MLP (:188-196)."""
    B, N, C = x.shape
    res = x
    n = F.layer_norm(x, (C,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-5)
    # Attention
    qkv = F.linear(n, sd[p + 'attn.qkv.weight'], sd[p + 'attn.qkv.bias'])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = (attn + bias.expand(B, -1, -1, -1)).softmax(dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, C)
    a = F.linear(a, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'])
    # MGCN
    W, M = sd[p + 'gcn.W'], sd[p + 'gcn.M']
    h0 = torch.matmul(n, W[0])
    h1 = torch.matmul(n, W[1])
    adj = c['graph_adj'].to(x.dtype) + sd[p + 'gcn.adj2']
    adj = (adj.T + adj) / 2
    E = torch.eye(adj.size(0), dtype=x.dtype, device=x.device)
    g = torch.matmul(adj * E, M * h0) + torch.matmul(adj * (1 - E), M * h1) + sd[p + 'gcn.bias'].view(1, 1, -1)
    s = a + g
    # X_Feat
    sp = c['spatial_pos']
    feats = []
    for k_, lin in ((1, 'x_feat.linears.0'), (2, 'x_feat.linears.1')):
        nc = F.linear(s, sd[p + lin + '.weight'], sd[p + lin + '.bias'])
        mask = (sp <= k_) if k_ == 1 else (sp == k_)
        mask = mask.to(nc.dtype).expand(B, -1, -1)
        feats.append(torch.bmm(mask, nc))
    f = F.linear(torch.cat(feats, -1), sd[p + 'x_feat.linearback.weight'], sd[p + 'x_feat.linearback.bias'])
    x = res + f
    # MLP
    n2 = F.layer_norm(x, (C,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-5)
    h = F.gelu(F.linear(n2, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias']))
    return x + F.linear(h, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])


def gat_forward(sd, c, pose2d, prefix='pose_lifter.', depth=6, trace=None):
    """lib/models/GAT.py:133-152.  pose2d (B, 2J) or (B, J, 2) -> (x_out (B,3J), x (B,J,128))."""
    J = c['J']
    B = pose2d.shape[0]
    p = prefix
    x = pose2d.reshape(-1, J, 2).permute(0, 2, 1)
    # GraphLinear (modules.py:49-50) -> GroupNorm(4,64) -> GELU -> GraphLinear (GAT.py:69-72)
    x = torch.matmul(sd[p + 'GLinear.0.W'][None, :], x) + sd[p + 'GLinear.0.b'][None, :, None]
    x = F.group_norm(x, 4, sd[p + 'GLinear.1.weight'], sd[p + 'GLinear.1.bias'], 1e-5)
    x = F.gelu(x)
    x = torch.matmul(sd[p + 'GLinear.3.W'][None, :], x) + sd[p + 'GLinear.3.b'][None, :, None]
    x = x.permute(0, 2, 1)
    x = x + F.embedding(torch.arange(1, J + 1, device=x.device), sd[p + 'pos_id_embed.weight'])
    pos_num = sd[p + 'graph_adj'].long().sum(dim=1).view(-1)
    x = x + F.embedding(pos_num, sd[p + 'pos_num_embed.weight'])
    if trace is not None:
        trace['gat_embed'] = x
    bias = hop_path_encoding(sd, p + 'get_hop_path_encoding.', c)
    if trace is not None:
        trace['hop_path_bias'] = bias
    for i in range(depth):
        x = gat_block(sd, f'{p}blocks.{i}.', c, x, bias)
        if trace is not None:
            trace[f'gat_block{i}'] = x
    C = x.shape[-1]
    x = F.gelu(F.layer_norm(x, (C,), sd[p + 'norm.weight'], sd[p + 'norm.bias'], 1e-5))
    x_out = F.linear(x.reshape(B, -1), sd[p + 'lifter.weight'], sd[p + 'lifter.bias'])
    return x_out, x


# ----------------------------------------------------------------------------------------------
# MDR
# ----------------------------------------------------------------------------------------------
def _cross_block(sd, p, x, J, heads=2):
    """lib/models/MDR.py:64-69 (CrossAttentionBlock) + :34-46 (CrossAttention) + timm Mlp."""
    B, N, C = x.shape
    V = N - J
    y = F.layer_norm(x, (C,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-5)
    q = F.linear(y[:, :V], sd[p + 'attn.wq.weight']).reshape(B, V, heads, C // heads).permute(0, 2, 1, 3)
    k = F.linear(y[:, -J:], sd[p + 'attn.wk.weight']).reshape(B, J, heads, C // heads).permute(0, 2, 1, 3)
    v = F.linear(y[:, -J:], sd[p + 'attn.wv.weight']).reshape(B, J, heads, C // heads).permute(0, 2, 1, 3)
    attn = ((q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)).softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, V, C)
    o = F.linear(o, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'])
    x = x[:, :V] + o
    n2 = F.layer_norm(x, (C,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-5)
    h = F.gelu(F.linear(n2, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias']))
    return x + F.linear(h, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])


def _custom_ln(sd, p, x, eps=1e-6):
    """lib/models/vanilla_transformer_encoder.py:24-34: a_2 (x-mean)/(std_unbiased + eps) + b_2."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return sd[p + 'a_2'] * (x - mean) / (std + eps) + sd[p + 'b_2']


def _self_attn(sd, p, x, h=2):
    """lib/models/vanilla_transformer_encoder.py:82-94 + :36-46."""
    B, N, C = x.shape
    dk = C // h
    q, k, v = [F.linear(x, sd[f'{p}linears.{i}.weight'], sd[f'{p}linears.{i}.bias']).view(B, -1, h, dk).transpose(1, 2)
               for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    o = torch.matmul(F.softmax(scores, dim=-1), v)
    o = o.transpose(1, 2).contiguous().view(B, -1, h * dk)
    return F.linear(o, sd[p + 'linears.3.weight'], sd[p + 'linears.3.bias'])


def mdr_forward(sd, c, x, alpha: bool, prefix='pose2mesh.', trace=None):
    """lib/models/MDR.py:124-170.  x = pose_combine (B, J, 2+3+128) -> (B, 6890, 3) metres."""
    p = prefix
    B, J = x.shape[0], x.shape[1]
    iv = sd[p + 'init_vertices']
    V = iv.shape[0]
    vj = torch.from_numpy(np.asarray(c['vj_relation'])).long().to(x.device)
    verts = torch.cat([iv.unsqueeze(0).expand(B, -1, -1), x[:, vj, 2:5]], dim=2)
    joint = F.linear(x, sd[p + 'get_joint_feature.weight'], sd[p + 'get_joint_feature.bias'])
    verts = F.linear(verts, sd[p + 'get_verts_feature.weight'], sd[p + 'get_verts_feature.bias'])
    joint = joint + F.embedding(torch.arange(1, J + 1, device=x.device), sd[p + 'pos_j_id_embed.weight'])
    verts = verts + F.embedding(torch.arange(1, V + 1, device=x.device), sd[p + 'pos_v_id_embed.weight'])
    if trace is not None:
        trace['mdr_verts_embed'] = verts
        trace['mdr_joint_embed'] = joint
    for li, sfx in enumerate(('', '_1', '_2')):
        fusion = torch.cat([verts, joint], dim=1)
        verts = _cross_block(sd, f'{p}encoder{sfx}.', fusion, J)
        if trace is not None:
            trace[f'mdr_cross{li}'] = verts
        verts = _custom_ln(sd, f'{p}norm{sfx}.', verts)
        verts = verts + _self_attn(sd, f'{p}selfatt{sfx}.', verts)
        if trace is not None:
            trace[f'mdr_layer{li}'] = verts
    ac = F.linear(verts, sd[p + 'motion_linear.weight'], sd[p + 'motion_linear.bias'])
    mat_A, mat_C = ac[:, :, :20], ac[:, :, -3:]
    mat_B = F.linear(verts, sd[p + 'bias_linear.weight'], sd[p + 'bias_linear.bias'])
    if alpha:
        mat_B = F.layer_norm(mat_B, (3,), sd[p + 'bias_norm.weight'], sd[p + 'bias_norm.bias'], 1e-5)
    else:
        mat_B = F.batch_norm(mat_B, sd[p + 'bias_norm.running_mean'], sd[p + 'bias_norm.running_var'],
                             sd[p + 'bias_norm.weight'], sd[p + 'bias_norm.bias'], False, 0.1, 1e-5)
    mat_B = F.gelu(mat_B)
    mat_B = F.conv1d(mat_B, sd[p + 'bias_conv1d.weight'], sd[p + 'bias_conv1d.bias'], padding=1)
    if alpha:
        a = 1.1 ** F.linear(verts, sd[p + 'scale_linear.weight'], sd[p + 'scale_linear.bias'])
    else:
        a = 1
    coarse = a * mat_A.softmax(dim=-1).bmm(mat_B) + mat_C
    if trace is not None:
        trace['mdr_coarse'] = coarse
    out = F.conv1d(coarse, sd[p + 'upsample_conv.weight'], sd[p + 'upsample_conv.bias'], padding=1)
    return out + sd[p + 'init_vertices_6890']


def gator_forward(sd, gat_c, mdr_c, pose2d, alpha: bool, trace=None):
    """lib/models/GATOR.py:16-22.  pose2d (B,J,2) -> (cam_mesh (B,6890,3) m, pose3d (B,J,3) mm)."""
    J = gat_c['J']
    pose3d, feat = gat_forward(sd, gat_c, pose2d.reshape(len(pose2d), -1), trace=trace)
    pose3d = pose3d.reshape(-1, J, 3)
    combine = torch.cat((pose2d, pose3d / 1000, feat), dim=2)
    mesh = mdr_forward(sd, mdr_c, combine, alpha, trace=trace)
    return mesh, pose3d


def joint_regress(J_regressor: torch.Tensor, mesh: torch.Tensor) -> torch.Tensor:
    """lib/core/base.py:221, demo/run.py:142: J_regressor[None] @ pred_mesh."""
    return torch.matmul(J_regressor[None, :, :], mesh)


# ----------------------------------------------------------------------------------------------
# SMPL linear blend skinning
# ----------------------------------------------------------------------------------------------
def batch_rodrigues(aa: torch.Tensor) -> torch.Tensor:
    """smplpytorch/pytorch/rodrigues_layer.py:41-52 + quat2mat :13-38 -> (N, 9)."""
    angle = torch.norm(aa + 1e-8, p=2, dim=1).unsqueeze(-1)
    axis = aa / angle
    angle = angle * 0.5
    quat = torch.cat([torch.cos(angle), torch.sin(angle) * axis], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)


def smpl_forward(buf: Dict[str, torch.Tensor], parents: Sequence[int], pose, betas=None, trans=None,
                 center_idx=None):
    """smplpytorch/pytorch/smpl_layer.py:65-158 (+ tensutils.py:6-48).  Returns (verts, jtr, aux)."""
    B = pose.shape[0]
    dt = pose.dtype
    rot = torch.cat([batch_rodrigues(pose[:, 3 * j:3 * j + 3]) for j in range(24)], 1)   # tensutils.py:6-19
    root_rot = rot[:, :9].view(B, 3, 3)
    rot = rot[:, 9:]
    pose_map = rot - torch.eye(3, dtype=dt).view(1, 9).repeat(B, 23)                    # tensutils.py:41-48
    S, P = buf['th_shapedirs'].to(dt), buf['th_posedirs'].to(dt)
    T, Jr, W = buf['th_v_template'].to(dt), buf['th_J_regressor'].to(dt), buf['th_weights'].to(dt)
    if betas is None or bool(torch.norm(betas) == 0):
        v_shaped = T + torch.matmul(S, buf['th_betas'].to(dt).transpose(1, 0)).permute(2, 0, 1)
        th_j = torch.matmul(Jr, v_shaped).repeat(B, 1, 1)
    else:
        v_shaped = T + torch.matmul(S, betas.transpose(1, 0)).permute(2, 0, 1)
        th_j = torch.matmul(Jr, v_shaped)
    v_posed = v_shaped + torch.matmul(P, pose_map.transpose(0, 1)).permute(2, 0, 1)

    def with_zeros(t):                                                                  # tensutils.py:22-29
        pad = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=dt).view(1, 1, 4).repeat(B, 1, 1)
        return torch.cat([t, pad], 1)
    results = [with_zeros(torch.cat([root_rot, th_j[:, 0, :].view(B, 3, 1)], 2))]
    for i in range(1, 24):
        jr = rot[:, (i - 1) * 9:i * 9].contiguous().view(B, 3, 3)
        rel = with_zeros(torch.cat([jr, (th_j[:, i, :] - th_j[:, parents[i], :]).view(B, 3, 1)], 2))
        results.append(torch.matmul(results[parents[i]], rel))
    res2 = torch.zeros((B, 4, 4, 24), dtype=dt)
    for i in range(24):
        jj = torch.cat([th_j[:, i], torch.zeros(B, 1, dtype=dt)], 1)
        tmp = torch.bmm(results[i], jj.unsqueeze(2))
        res2[:, :, :, i] = results[i] - torch.cat([torch.zeros(B, 4, 3, dtype=dt), tmp], 2)  # th_pack
    th_T = torch.matmul(res2, W.transpose(0, 1))
    rest_h = torch.cat([v_posed.transpose(2, 1), torch.ones((B, 1, v_posed.shape[1]), dtype=dt)], 1)
    verts = (th_T * rest_h.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = torch.stack(results, dim=1)[:, :, :3, 3]
    if trans is None or bool(torch.norm(trans) == 0):
        if center_idx is not None:
            cj = jtr[:, center_idx].unsqueeze(1)
            jtr = jtr - cj
            verts = verts - cj
    else:
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts, jtr, {'v_posed': v_posed, 'th_j': th_j, 'rotmats': torch.cat([root_rot.reshape(B, 9), rot], 1)}


MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)
MANO_JOINT_ORDER = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)


def mano_forward(buf: Dict[str, torch.Tensor], pose_coeffs, betas=None, trans=None, *, ncomps=6, use_pca=True, side='right',
                 center_idx=None, root_palm=False, share_betas=False):
    """manopth/manopth/manolayer.py:109-273 for root_rot_mode = joint_rot_mode = 'axisang' (+ manopth/tensutils.py).
    buf: th_betas, th_shapedirs, th_posedirs, th_v_template, th_J_regressor, th_weights, th_hands_mean,
    th_selected_comps.  Returns (verts (B,778,3), jtr (B,21,3)) in millimetres."""
    B = pose_coeffs.shape[0]
    dt = pose_coeffs.dtype
    hand = pose_coeffs[:, 3:3 + ncomps]                                                  # :130-131
    full_hand = hand.mm(buf['th_selected_comps'].to(dt)) if use_pca else hand            # :132-136
    full_pose = torch.cat([pose_coeffs[:, :3], buf['th_hands_mean'].to(dt) + full_hand], 1)   # :139-142
    rot = batch_rodrigues(full_pose.contiguous().view(-1, 3)).view(B, 16 * 9)            # tensutils.py:6-12
    pose_map = (rot - torch.eye(3, dtype=dt).view(1, 9).repeat(B, 16))[:, 9:]            # tensutils.py:31-39, :147
    root_rot = rot[:, :9].view(B, 3, 3)
    rot = rot[:, 9:]
    S, P = buf['th_shapedirs'].to(dt), buf['th_posedirs'].to(dt)
    T, Jr, W = buf['th_v_template'].to(dt), buf['th_J_regressor'].to(dt), buf['th_weights'].to(dt)
    if betas is None or betas.numel() == 1:                                              # :172-178
        v_shaped = torch.matmul(S, buf['th_betas'].to(dt).transpose(1, 0)).permute(2, 0, 1) + T
        th_j = torch.matmul(Jr, v_shaped).repeat(B, 1, 1)
    else:
        if share_betas:
            betas = betas.mean(0, keepdim=True).expand(betas.shape[0], 10)
        v_shaped = torch.matmul(S, betas.transpose(1, 0)).permute(2, 0, 1) + T
        th_j = torch.matmul(Jr, v_shaped)
    v_posed = v_shaped + torch.matmul(P, pose_map.transpose(0, 1)).permute(2, 0, 1)      # :188-189

    def with_zeros(t):
        pad = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=dt).view(1, 1, 4).repeat(t.shape[0], 1, 1)
        return torch.cat([t, pad], 1)
    # :194-231: the three finger levels are the kinematic tree MANO_PARENTS walked breadth first
    results = [with_zeros(torch.cat([root_rot, th_j[:, 0, :].contiguous().view(B, 3, 1)], 2))]
    for i in range(1, 16):
        par = MANO_PARENTS[i]
        jr = rot[:, (i - 1) * 9:i * 9].contiguous().view(B, 3, 3)
        rel = with_zeros(torch.cat([jr, (th_j[:, i, :] - th_j[:, par, :]).view(B, 3, 1)], 2))
        results.append(torch.matmul(results[par], rel))
    th_results = torch.stack(results, 1)                                                 # (B,16,4,4), joint order (:229-230)
    joint_js = torch.cat([th_j, th_j.new_zeros(B, 16, 1)], 2)
    tmp2 = torch.matmul(th_results, joint_js.unsqueeze(3))
    res2 = (th_results - torch.cat([tmp2.new_zeros(B, 16, 4, 3), tmp2], 3)).permute(0, 2, 3, 1)   # :233-235
    th_T = torch.matmul(res2, W.transpose(0, 1))
    rest_h = torch.cat([v_posed.transpose(2, 1), torch.ones((B, 1, v_posed.shape[1]), dtype=dt)], 1)
    verts = (th_T * rest_h.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = th_results[:, :, :3, 3]
    tips = verts[:, [745, 317, 444 if side == 'right' else 445, 556, 673]]               # :249-252
    if root_palm:
        palm = (verts[:, 95] + verts[:, 22]).unsqueeze(1) / 2
        jtr = torch.cat([palm, jtr[:, 1:]], 1)
    jtr = torch.cat([jtr, tips], 1)[:, list(MANO_JOINT_ORDER)]                           # :256-259
    if trans is None or bool(torch.norm(trans) == 0):                                    # :261-268
        if center_idx is not None:
            cj = jtr[:, center_idx].unsqueeze(1)
            jtr = jtr - cj
            verts = verts - cj
    else:
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts * 1000, jtr * 1000                                                      # :271-272


# ----------------------------------------------------------------------------------------------
# evaluation helpers restated for the drift metric (not on the hot path)
# ----------------------------------------------------------------------------------------------
H36M_EVAL_JOINTS = (1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16)


def rigid_align(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """lib/coord_utils.py:127-149: similarity Procrustes of A onto B."""
    mu_a, mu_b = A.mean(0), B.mean(0)
    a0, b0 = A - mu_a, B - mu_b
    H = a0.T @ b0
    U, s, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        s[-1] = -s[-1]
        Vt[2] = -Vt[2]
        R = Vt.T @ U.T
    var_a = (a0 ** 2).sum() / len(A)
    c = s.sum() / len(A) / var_a
    t = mu_b - c * R @ mu_a
    return (c * (R @ A.T)).T + t


def mpjpe_pa(pred_mesh_m: np.ndarray, gt_mesh_m: np.ndarray, regressor_h36m: np.ndarray):
    """MPJPE / PA-MPJPE in mm on the 14 H36M eval joints, root-aligned
    (lib/core/base.py:219-221, data/Human36M/dataset.py:466-478)."""
    mp, pa = [], []
    for p, g in zip(pred_mesh_m, gt_mesh_m):
        jp = regressor_h36m @ (p * 1000.0)
        jg = regressor_h36m @ (g * 1000.0)
        jp, jg = jp - jp[:1], jg - jg[:1]
        jp, jg = jp[list(H36M_EVAL_JOINTS)], jg[list(H36M_EVAL_JOINTS)]
        mp.append(np.sqrt(((jp - jg) ** 2).sum(1)).mean())
        pa.append(np.sqrt(((rigid_align(jp, jg) - jg) ** 2).sum(1)).mean())
    return float(np.mean(mp)), float(np.mean(pa))


# ----------------------------------------------------------------------------------------------
# evaluation epilogue - the checker of csrc/eval.cu
# ----------------------------------------------------------------------------------------------
def eval_pred_pose(J_regressor: torch.Tensor, pred_mesh_m: torch.Tensor, scale: float = 1000.0) -> torch.Tensor:
    """lib/core/base.py:219-221: ``pred_mesh * 1000`` then the dense ``J_regressor[None] @ pred_mesh``."""
    return torch.matmul(J_regressor[None, :, :], pred_mesh_m * scale)


def compute_both_err(pred_mesh, target_mesh, pred_joint, target_joint, eval_joints=H36M_EVAL_JOINTS):
    """data/Human36M/dataset.py:466-478 (= data/PW3D/dataset.py:273-286); all arguments in mm, torch fp32.
    Returns (joint_mean_error, mesh_mean_error) as numpy float32 scalars, like the reference."""
    pred_mesh, target_mesh = pred_mesh - pred_joint[:, :1, :], target_mesh - target_joint[:, :1, :]
    pred_joint, target_joint = pred_joint - pred_joint[:, :1, :], target_joint - target_joint[:, :1, :]
    pm, tm = pred_mesh.numpy(), target_mesh.numpy()
    pj, tj = pred_joint.numpy()[:, list(eval_joints), :], target_joint.numpy()[:, list(eval_joints), :]
    mesh_err = np.sqrt(((pm - tm) ** 2).sum(axis=2)).mean()
    joint_err = np.sqrt(((pj - tj) ** 2).sum(axis=2)).mean()
    return joint_err, mesh_err


def per_sample_errors(pred_mesh, target_mesh, pred_joint, target_joint, eval_joints=H36M_EVAL_JOINTS):
    """Per-sample MPJPE / MPVPE / PA-MPJPE in the form evaluate_joint computes them
    (data/Human36M/dataset.py:480-504): root-align on joint 0, select the eval joints, mean L2, then the same
    after ``rigid_align``.  numpy arrays in mm; returns three (B,) float64 arrays."""
    ev = list(eval_joints)
    mpjpe, mpvpe, pa = [], [], []
    for n in range(len(pred_joint)):
        out, gt = pred_joint[n] - pred_joint[n][:1], target_joint[n] - target_joint[n][:1]
        out, gt = out[ev, :], gt[ev, :]
        mpjpe.append(np.sqrt(np.sum((out - gt) ** 2, 1)).mean())
        pa.append(np.sqrt(np.sum((rigid_align(out, gt) - gt) ** 2, 1)).mean())
        if target_mesh is not None:
            pm, tm = pred_mesh[n] - pred_joint[n][:1], target_mesh[n] - target_joint[n][:1]
            mpvpe.append(np.sqrt(np.sum((pm - tm) ** 2, 1)).mean())
    return np.asarray(mpjpe, np.float64), np.asarray(mpvpe, np.float64), np.asarray(pa, np.float64)


# ----------------------------------------------------------------------------------------------
# 2D-pose pre-processing - the checker of csrc/preprocess.cu
# ----------------------------------------------------------------------------------------------
def add_mid_joint(joint_coord: np.ndarray, a: int, b: int) -> np.ndarray:
    """demo/run.py:103-121 (add_pelvis / add_neck): midpoint of two joints appended; a third column
    (confidence) is multiplied instead of averaged."""
    mid = (joint_coord[a, :] + joint_coord[b, :]) * 0.5
    if joint_coord.shape[1] > 2:
        mid[2] = joint_coord[a, 2] * joint_coord[b, 2]
    return np.concatenate((joint_coord, mid.reshape(1, -1)))


def get_bbox(joint_img: np.ndarray) -> np.ndarray:
    """lib/coord_utils.py:21-39: tight box (x, y, w, h) as float32."""
    x_img, y_img = joint_img[:, 0], joint_img[:, 1]
    xmin, ymin, xmax, ymax = min(x_img), min(y_img), max(x_img), max(y_img)
    x_center = (xmin + xmax) / 2.
    width = xmax - xmin
    xmin, xmax = x_center - 0.5 * width, x_center + 0.5 * width
    y_center = (ymin + ymax) / 2.
    height = ymax - ymin
    ymin, ymax = y_center - 0.5 * height, y_center + 0.5 * height
    return np.array([xmin, ymin, xmax - xmin, ymax - ymin]).astype(np.float32)


def process_bbox(bbox: np.ndarray, aspect_ratio: float, scale: float = 1.0):
    """lib/coord_utils.py:42-66: sanitise, then grow to the aspect ratio about the centre (float32 arithmetic
    because `bbox` is float32 and Python scalars are weak)."""
    x, y, w, h = bbox
    x1, y1, x2, y2 = x, y, x + (w - 1), y + (h - 1)
    if w * h > 0 and x2 >= x1 and y2 >= y1:
        bbox = np.array([x1, y1, x2 - x1, y2 - y1])
    else:
        return None
    w, h = bbox[2], bbox[3]
    c_x, c_y = bbox[0] + w / 2., bbox[1] + h / 2.
    if w > aspect_ratio * h:
        h = w / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    bbox[2], bbox[3] = w * scale, h * scale
    bbox[0], bbox[1] = c_x - bbox[2] / 2., c_y - bbox[3] / 2.
    return bbox


def crop_affine(bbox: np.ndarray, res) -> np.ndarray:
    """j2d_processing's transform for rot = 0 (lib/aug_utils.py:51-57, get_center_scale coord_utils.py:7-18,
    get_affine_transform aug_utils.py:140-172): three float32 point pairs, solved in float64 like
    cv2.getAffineTransform."""
    x, y, w, h = bbox
    center = np.zeros(2, np.float32)
    center[0], center[1] = x + w * 0.5, y + h * 0.5
    scale = np.array([w * 1.0, h * 1.0], np.float32)
    src_w, dst_w, dst_h = scale[0], res[0], res[1]
    src_dir = [0.0, np.float64(src_w * -0.5)]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src, dst = np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32)
    src[0, :] = center
    src[1, :] = center + src_dir
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):
        direct = a - b
        return b + np.array([-direct[1], direct[0]], np.float32)
    src[2, :], dst[2, :] = third(src[0], src[1]), third(dst[0], dst[1])
    A = np.concatenate([src.astype(np.float64), np.ones((3, 1))], 1)          # (3,3) @ trans.T = dst
    return np.linalg.solve(A, dst.astype(np.float64)).T                       # (2,3)


def preprocess_pose2d(joint_input: np.ndarray, input_shape=(384, 288), mid_pairs=(), bbox_scale: float = 1.0):
    """demo/run.py:124-133 (= data/Human36M/dataset.py:383-389 after the crop): bbox -> aspect-ratio box ->
    affine into the (input_shape[1], input_shape[0]) crop -> /[W, H] -> per-axis standardisation.
    joint_input (J, >=2) pixel coordinates.  Returns (pose2d (J',2) f32, joint_img (J',2) f32, bbox (4,) f32)."""
    j = np.asarray(joint_input)
    for a, b in mid_pairs:
        j = add_mid_joint(j, a, b)
    j = j[:, :2]
    W, H = input_shape[1], input_shape[0]
    bbox = process_bbox(get_bbox(j).copy(), aspect_ratio=W / H, scale=bbox_scale)
    if bbox is None:
        return None
    trans = crop_affine(bbox, (W, H))
    kp = j.copy()
    for i in range(kp.shape[0]):
        kp[i, :2] = np.dot(trans, np.array([kp[i, 0], kp[i, 1], 1.]).T)[:2]
    kp = kp.astype('float32')
    joint_img = kp[:, :2].copy()
    ji = joint_img.copy()
    ji /= np.array([[W, H]])
    mean, std = np.mean(ji, axis=0), np.std(ji, axis=0)
    return ((ji.copy() - mean) / std).astype(np.float32), joint_img, bbox


# ----------------------------------------------------------------------------------------------
# ground-truth mesh generation - the checker of csrc/smpl_cam.cu + SMPL out_scale
# ----------------------------------------------------------------------------------------------
def axangle2mat(axis, angle, is_normalized=False):
    """transforms3d.axangles.axangle2mat (third-party, unpinned in requirements.sh, absent from the reference
    tree): published algorithm restated - Rodrigues matrix from a (normalised) axis and an angle."""
    x, y, z = axis
    if not is_normalized:
        n = math.sqrt(x * x + y * y + z * z)
        x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    xs, ys, zs = x * s, y * s, z * s
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    return np.array([[x * xC + c, xyC - zs, zxC + ys],
                     [xyC + zs, y * yC + c, yzC - xs],
                     [zxC - ys, yzC + xs, z * zC + c]])


def mat2axangle(mat, unit_thresh=1e-5):
    """transforms3d.axangles.mat2axangle restated: axis = unit eigenvector of eigenvalue 1 of the matrix (float64
    eig), angle = atan2(sin, cos) with cos = (trace - 1) / 2 and sin recovered from an off-diagonal entry."""
    M = np.asarray(mat, dtype=np.float64)
    L, W = np.linalg.eig(M.T)
    i = np.where(np.abs(L - 1.0) < unit_thresh)[0]
    if not len(i):
        raise ValueError('no unit eigenvector corresponding to eigenvalue 1')
    direction = np.real(W[:, i[-1]]).squeeze()
    cosa = (np.trace(M) - 1.0) / 2.0
    if abs(direction[2]) > 1e-8:
        sina = (M[1, 0] + (cosa - 1.0) * direction[0] * direction[1]) / direction[2]
    elif abs(direction[1]) > 1e-8:
        sina = (M[0, 2] + (cosa - 1.0) * direction[0] * direction[2]) / direction[1]
    else:
        sina = (M[2, 1] + (cosa - 1.0) * direction[1] * direction[2]) / direction[0]
    return direction, math.atan2(sina, cosa)


def h36m_smpl_coord(buf, parents, pose, shape, trans, R, t, root_idx: int = 0):
    """Human36M.get_smpl_coord (data/Human36M/dataset.py:254-298) for one item: numpy in, (mesh (6890,3),
    joints (24,3)) in millimetres, camera frame.  `buf` / `parents` describe the SMPL layer (smpl_forward)."""
    smpl_pose = torch.FloatTensor(np.asarray(pose, np.float32)).view(-1, 3).clone()
    smpl_shape = torch.FloatTensor(np.asarray(shape, np.float32)).view(1, -1).clone()
    trans = np.array(trans, dtype=np.float32).reshape(3)
    R, t = np.array(R, dtype=np.float32).reshape(3, 3), np.array(t, dtype=np.float32).reshape(3)
    smpl_shape[(smpl_shape.abs() > 3).any(dim=1)] = 0.
    root_pose = smpl_pose[root_idx, :].numpy()
    angle = np.linalg.norm(root_pose)
    root_pose = axangle2mat(root_pose / angle, angle)
    root_pose = np.dot(R, root_pose)
    axis, angle = mat2axangle(root_pose)
    smpl_pose[root_idx] = torch.from_numpy(axis * angle)
    verts, jtr, _ = smpl_forward(buf, parents, smpl_pose.view(1, -1), smpl_shape)
    mesh = verts.numpy().astype(np.float32).reshape(-1, 3)
    joints = jtr.numpy().astype(np.float32).reshape(-1, 3)
    smpl_trans = np.dot(R, trans[:, None]).reshape(1, 3) + t.reshape(1, 3) / 1000
    root = joints[root_idx].reshape(1, 3)
    smpl_trans = smpl_trans - root + np.dot(R, root.transpose(1, 0)).transpose(1, 0)
    mesh += smpl_trans
    joints += smpl_trans
    mesh *= 1000
    joints *= 1000
    return mesh, joints


def pw3d_smpl_coord(buf, parents, pose, shape, trans):
    """PW3D.get_smpl_coord (data/PW3D/dataset.py:84-102): SMPL forward with the world translation, then mm."""
    verts, jtr, _ = smpl_forward(buf, parents, torch.FloatTensor(np.asarray(pose, np.float32)).view(1, -1),
                                 torch.FloatTensor(np.asarray(shape, np.float32)).view(1, -1),
                                 torch.FloatTensor(np.asarray(trans, np.float32)).view(-1, 3))
    mesh = verts.numpy().astype(np.float32).reshape(-1, 3)
    joints = jtr.numpy().astype(np.float32).reshape(-1, 3)
    mesh *= 1000
    joints *= 1000
    return mesh, joints
