"""Drop-in for ``models.GAT`` (lib/models/GAT.py): same constructor signature, attribute names and
``state_dict`` keys (so ``*.pth.tar`` checkpoints load with strict=True), forward executed by the
sm_100a kernels behind ``gator_gat_forward`` (csrc/gat.cu).

The sub-modules below only *hold* parameters under the reference's names; none of their ``forward``
methods is ever called.  Input-independent terms are folded once in :meth:`GAT.pack`
(SURVEY.md appendix A.4): positional embeddings, the hop/path attention bias
(modules.py:98-107), the symmetrised learnable adjacency (modules.py:247-249) and the hop masks
(modules.py:165-168).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, config, graph
from ..packing import pack_umma_weight_pair, pack_umma_weight_pairs, pack_umma_blob

_BF16_GLOBAL = {'LIFT_W'}
_BF16_BLOCK = {'QKV_W', 'PROJ_W', 'GCN_W01', 'XF_W01', 'XF_WB', 'FC1_W', 'FC2_W'}

_HEADS = 8
_EMBED = 128


def _invalidate_hook(module, incompatible_keys):
    module.invalidate()


class GraphLinear(nn.Module):
    """Parameter holder for modules.py:31-50 (keys ``W``, ``b``; same U(+-1/(in*out)) init)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.W = nn.Parameter(torch.empty(out_channels, in_channels))
        self.b = nn.Parameter(torch.empty(out_channels))
        w_stdv = 1 / (in_channels * out_channels)
        self.W.data.uniform_(-w_stdv, w_stdv)
        self.b.data.uniform_(-w_stdv, w_stdv)


class HopPathEncoding(nn.Module):
    """Parameter holder for modules.py:77-107 (keys ``W``, ``spatial_pos_encoder.weight``,
    ``edge_encoder.{weight,bias}``); :meth:`bias` restates its forward with torch ops and is evaluated
    once per pack() - it does not depend on the input."""

    def __init__(self, num_heads, num_spatial, num_joint, spatial_pos, edg_adj):
        super().__init__()
        self.num_heads, self.num_joint = num_heads, num_joint
        edg_adj = edg_adj.clone()
        edg_adj[edg_adj == -1] = 0
        self.edg_adj = edg_adj                              # (J,J,D) constant
        self.spatial_pos = spatial_pos.long()               # (J,J) hop counts
        self.spatial_pos_encoder = nn.Embedding(num_spatial, num_heads, padding_idx=0)
        self.edge_encoder = nn.Linear(num_joint * num_joint, num_joint * num_joint * num_heads)
        self.W = nn.Parameter(torch.ones(num_heads, *edg_adj.shape))

    @torch.no_grad()
    def bias(self, device) -> torch.Tensor:
        J, H = self.num_joint, self.num_heads
        sp = self.spatial_pos.to(device)
        ones = torch.ones_like(sp)
        spatial = sp - ones
        spatial = torch.where(spatial > 0, spatial, ones)
        spatial = 1.0 / spatial.expand(H, -1, -1)
        spb = self.spatial_pos_encoder.weight[sp].permute(2, 0, 1)
        e = self.edg_adj.to(device=device, dtype=torch.float32).permute(2, 0, 1).reshape(-1, J * J)
        e = torch.nn.functional.linear(e, self.edge_encoder.weight, self.edge_encoder.bias)
        e = e.reshape(-1, H, J, J).permute(1, 2, 3, 0)
        eb = torch.mul(self.W, e).sum(-1)
        return (spb + torch.mul(eb, spatial)).float().contiguous()


class Attention(nn.Module):
    def __init__(self, dim, qkv_bias=True):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class MLP(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class MGCN(nn.Module):
    """Parameter holder for modules.py:213-255 (keys ``W``, ``M``, ``adj2``, ``bias``; same init)."""

    def __init__(self, in_features, out_features, adj):
        super().__init__()
        self.adj = adj
        self.W = nn.Parameter(torch.zeros(2, in_features, out_features))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.M = nn.Parameter(torch.zeros(adj.size(0), out_features))
        nn.init.xavier_uniform_(self.M.data, gain=1.414)
        self.adj2 = nn.Parameter(torch.ones_like(adj))
        nn.init.constant_(self.adj2, 1e-6)
        self.bias = nn.Parameter(torch.zeros(out_features))
        stdv = 1. / math.sqrt(self.W.size(2))
        self.bias.data.uniform_(-stdv, stdv)


class X_Feat(nn.Module):
    """Parameter holder for modules.py:140-177 with s=1, l=2, d=8: linears 128->128, 128->16; 144->128."""

    def __init__(self, input_dim, output_dim, s=1, l=2, d=8):
        super().__init__()
        self.linears = nn.ModuleList()
        c_out, total = int(input_dim), 0
        for _ in range(s, l + 1):
            total += c_out
            self.linears.append(nn.Linear(input_dim, c_out))
            c_out = int(c_out / d)
        self.linearback = nn.Linear(total, int(output_dim))


class GATBlock(nn.Module):
    """Parameter holder for GAT.py:16-43 (registration order = reference's)."""

    def __init__(self, dim, num_heads, mlp_ratio, graph_adj, qkv_bias):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, qkv_bias=qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, int(dim * mlp_ratio))
        self.adj = graph_adj
        self.gcn = MGCN(dim, dim, graph_adj)
        self.x_feat = X_Feat(dim, dim)


class GAT(nn.Module):
    def __init__(self, num_joint=17, embed_dim=256, depth=4, graph_adj=None, GCN_depth=1, J_regressor=None,
                 num_heads=8, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.4, attn_drop_rate=0.4,
                 drop_path_rate=0.2, norm_layer=nn.LayerNorm, act_layer=None, pretrained=False):
        super().__init__()
        if embed_dim != _EMBED or num_heads != _HEADS or float(mlp_ratio) != 4.0 or not qkv_bias or qk_scale is not None:
            raise NotImplementedError(
                'gator_b200 kernels are built for embed_dim=128, num_heads=8, mlp_ratio=4, qkv_bias=True '
                '(the only configuration any reference call site uses: base.py:57,59, demo/run.py:96)')
        if graph_adj is None or J_regressor is None:
            raise ValueError('graph_adj and J_regressor are required (GAT.py:57,76 dereference both)')
        if not 2 <= num_joint <= 32:
            raise NotImplementedError('num_joint must be in [2, 32]')
        self.num_joint, self.embed_dim, self.num_heads, self.depth = num_joint, embed_dim, num_heads, depth
        self.output_size = 3 * num_joint
        self.pos_id_embed = nn.Embedding(num_joint + 1, embed_dim, padding_idx=0)
        adj = graph.dense_graph_adj(graph_adj)
        self.register_buffer('graph_adj', adj)
        self.GLinear = nn.Sequential(GraphLinear(2, 64), nn.GroupNorm(64 // 16, 64), nn.GELU(),
                                     GraphLinear(64, embed_dim))
        self.pos_num_embed = nn.Embedding(num_joint, embed_dim, padding_idx=0)
        init_vertices = torch.from_numpy(np.load(config.base_data_path('smpl_mean_vertices.npy'))).unsqueeze(0)
        self.register_buffer('init_vertices', init_vertices)
        tj = graph.template_joints(torch.as_tensor(J_regressor).cpu(), init_vertices, num_joint)
        suffix = '3dpw' if num_joint == 19 else 'h36m'          # GAT.py:88-93
        shortest = np.load(config.base_data_path(f'shortest_path_{suffix}.npy'))
        path = np.load(config.base_data_path(f'path_{suffix}.npy'))
        edge_input = graph.path_edge_features(int(np.amax(shortest)), path, graph.edge_lengths(adj, tj))
        spatial_pos = torch.from_numpy(shortest)
        self.get_hop_path_encoding = HopPathEncoding(num_heads, 10, num_joint, spatial_pos, edge_input)
        block_adj = adj.clone()          # GATBlock.adj is a copy taken at construction (GAT.py:29)
        self.blocks = nn.Sequential(*[GATBlock(embed_dim, num_heads, mlp_ratio, block_adj, qkv_bias)
                                      for _ in range(depth)])
        self.gelu = nn.GELU()
        self.norm = nn.LayerNorm(embed_dim)
        self.lifter = nn.Linear(embed_dim * num_joint, 3 * num_joint)
        self._packed = None
        self._ws = None
        self.precision = _lib.PREC_FP32
        self.chunk = 0
        self.fused = True       # tensor-core precisions: all GATBlocks in one kernel (csrc/gat_chain2_umma.cu)
        self.register_load_state_dict_post_hook(_invalidate_hook)
        if pretrained:
            self._load_pretrained_model()

    # -- reference API ------------------------------------------------------------------------
    def _load_pretrained_model(self):
        """GAT.py:128-131."""
        cfg = config.get_cfg()
        print('Loading pretrained posenet...')
        checkpoint = torch.load(cfg.MODEL.posenet_path, map_location='cuda')
        self.load_state_dict(checkpoint['model_state_dict'])

    # -- packing ------------------------------------------------------------------------------
    def invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    def _need_full_pack(self):
        return not (self.fused and 2 <= self.num_joint <= 21)

    @torch.no_grad()
    def pack(self):
        """Fold constants and lay the weights out for the kernels; one device tensor per ABI slot."""
        full = self._need_full_pack()
        dev = self.lifter.weight.device
        if dev.type != 'cuda':
            raise RuntimeError('gator_b200.GAT: parameters must be on a CUDA device (no CPU fallback)')
        J = self.num_joint
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        sp = self.get_hop_path_encoding.spatial_pos.to(dev)
        deg = self.graph_adj.long().sum(dim=1).view(-1)                                 # GAT.py:143
        t = {
            'EMB_W1': f(self.GLinear[0].W), 'EMB_B1': f(self.GLinear[0].b),
            'GN_W': f(self.GLinear[1].weight), 'GN_B': f(self.GLinear[1].bias),
            'EMB_W2T': f(self.GLinear[3].W.t()), 'EMB_B2': f(self.GLinear[3].b),
            'POS_CONST': f(self.pos_id_embed.weight[1:J + 1] + self.pos_num_embed.weight[deg]),
            'ATTN_BIAS': self.get_hop_path_encoding.bias(dev),
            'HOP_MASK1': f((sp <= 1).float()), 'HOP_MASK2': f((sp == 2).float()),
            'NORM_W': f(self.norm.weight), 'NORM_B': f(self.norm.bias),
            'LIFT_W': f(self.lifter.weight), 'LIFT_B': f(self.lifter.bias),
        }
        gnames, bnames = _lib.slot_names('gat')
        t['CHAIN_BLOBS'] = t['CHAIN_PRM'] = None          # filled in after the per-block tensors exist
        eye = torch.eye(J, device=dev)
        block_dicts = []
        for blk in self.blocks:
            adj = blk.adj.to(dev) + blk.gcn.adj2                                         # modules.py:247-249
            adj = (adj.T + adj) / 2
            b = {
                'LN1_W': f(blk.norm1.weight), 'LN1_B': f(blk.norm1.bias),
                'QKV_W': f(blk.attn.qkv.weight), 'QKV_B': f(blk.attn.qkv.bias),
                'PROJ_W': f(blk.attn.proj.weight), 'PROJ_B': f(blk.attn.proj.bias),
                'GCN_W01': f(torch.cat([blk.gcn.W[0].t(), blk.gcn.W[1].t()], 0)),
                'GCN_M': f(blk.gcn.M), 'GCN_ADIAG': f(torch.diagonal(adj)), 'GCN_AOFF': f(adj * (1 - eye)),
                'GCN_BIAS': f(blk.gcn.bias),
                'XF_W01': f(torch.cat([blk.x_feat.linears[0].weight, blk.x_feat.linears[1].weight], 0)),
                'XF_B01': f(torch.cat([blk.x_feat.linears[0].bias, blk.x_feat.linears[1].bias], 0)),
                'XF_WB': f(blk.x_feat.linearback.weight), 'XF_BB': f(blk.x_feat.linearback.bias),
                'LN2_W': f(blk.norm2.weight), 'LN2_B': f(blk.norm2.bias),
                'FC1_W': f(blk.mlp.fc1.weight), 'FC1_B': f(blk.mlp.fc1.bias),
                'FC2_W': f(blk.mlp.fc2.weight), 'FC2_B': f(blk.mlp.fc2.bias),
            }
            block_dicts.append(b)
        # fused-blocks kernel (csrc/gat_chain2_umma.cu): 34 weight pieces + 14 parameter arrays per block
        keep = []
        blobs, prms = [], []
        for b in block_dicts:
            qkv, proj, gcn = b['QKV_W'], b['PROJ_W'], b['GCN_W01']
            w01 = torch.zeros(192, 128, device=dev)
            w01[:144] = b['XF_W01']
            wb = torch.zeros(128, 192, device=dev)
            wb[:, :144] = b['XF_WB']
            # 34 pieces in the order csrc/gat_chain2_umma.cu consumes them
            pieces = [qkv[0:64], qkv[64:128], qkv[128:192], qkv[256:320], qkv[192:256], qkv[320:384]]
            pieces += [proj[:, 0:64], proj[:, 64:128]]
            pieces += [gcn[0:128, 0:64], gcn[0:128, 64:128], gcn[128:256, 0:64], gcn[128:256, 64:128]]
            pieces += [w01[64 * u:64 * u + 64] for u in range(3)] + [wb[:, 64 * u:64 * u + 64] for u in range(3)]
            fc1, fc2 = b['FC1_W'], b['FC2_W']
            pieces += [fc1[64 * u:64 * u + 64] for u in range(4)] + [fc2[:, 64 * u:64 * u + 64] for u in range(4)]
            pieces += [fc1[64 * u:64 * u + 64] for u in range(4, 8)] + [fc2[:, 64 * u:64 * u + 64] for u in range(4, 8)]
            blob = pack_umma_blob(pieces)
            assert blob.numel() == len(pieces) * 2 * 64 * 128
            xfb = torch.zeros(192, device=dev)
            xfb[:144] = b['XF_B01']
            plist = [b['LN1_W'], b['LN1_B'], b['QKV_B'], b['PROJ_B'], b['GCN_M'], b['GCN_ADIAG'], b['GCN_AOFF'],
                     b['GCN_BIAS'], xfb, b['XF_BB'], b['LN2_W'], b['LN2_B'], b['FC1_B'], b['FC2_B']]
            keep += [blob, xfb]
            blobs.append(blob.data_ptr())
            prms += [x_.data_ptr() for x_ in plist]
        if self.depth > 0:
            t['CHAIN_BLOBS'] = torch.tensor(blobs, dtype=torch.int64, device=dev)
            t['CHAIN_PRM'] = torch.tensor(prms, dtype=torch.int64, device=dev)
        tensors = [t[n] for n in gnames]
        for b in block_dicts:
            tensors += [b[n] for n in bnames]
        # per-slot tcgen05 images: the lifter always; the per-block matrices only for the kernel-per-op tensor-core path
        # (fused = False or J outside the fused kernel's range) - the fused kernel reads CHAIN_BLOBS instead
        want = [(n in _BF16_GLOBAL) for n in gnames]
        for _ in block_dicts:
            want += [(n in _BF16_BLOCK) and full for n in bnames]
        pairs = iter(pack_umma_weight_pairs([tensors[i] for i, w_ in enumerate(want) if w_]))
        packed = [next(pairs) if w_ else None for w_ in want]
        tensors += keep
        table = (ctypes.c_void_p * (len(gnames) + len(bnames) * len(block_dicts)))(
            *[(t_.data_ptr() if t_ is not None else None) for t_ in tensors[:len(gnames) + len(bnames) * len(block_dicts)]])
        table16 = (ctypes.c_void_p * len(packed))(*[(t_[0].data_ptr() if t_ is not None else None) for t_ in packed])
        table16lo = (ctypes.c_void_p * len(packed))(*[(t_[1].data_ptr() if t_ is not None else None) for t_ in packed])
        self._packed = ((tensors, packed, table16, table16lo), table, dev)
        self._packed_full = full
        return self

    def _workspace(self, batch, dev):
        need = _lib.lib().gator_gat_workspace_bytes(batch, self.num_joint, self.chunk)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    # -- forward ------------------------------------------------------------------------------
    def forward(self, pose2d):
        """pose2d (B, 2J) [or (B,J,2)] -> (x_out (B,3J) mm, x (B,J,128))  (GAT.py:133-152)."""
        if self.training:
            raise NotImplementedError('gator_b200.GAT implements the eval() forward only')
        if self._packed is None or (self._need_full_pack() and not self._packed_full):
            self.pack()
        (_, _, table16, table16lo), table, dev = self._packed
        if not pose2d.is_cuda:
            raise RuntimeError('gator_b200.GAT: input must be a CUDA tensor (no CPU fallback)')
        if pose2d.device != dev:
            raise RuntimeError(f'gator_b200.GAT: input on {pose2d.device} but the packed weights are on {dev}')
        B = pose2d.shape[0]
        J = self.num_joint
        x = pose2d.detach().reshape(B, J * 2).to(torch.float32).contiguous()
        pose3d = torch.empty((B, 3 * J), dtype=torch.float32, device=dev)
        feat = torch.empty((B, J, _EMBED), dtype=torch.float32, device=dev)
        if B == 0:
            return pose3d, feat
        ws = self._workspace(B, dev)
        a = _lib.GatArgs(num_joint=J, depth=self.depth, batch=B, chunk=self.chunk, precision=self.precision,
                         reserved=0 if self.fused else 1, weights=table, weights_bf16=table16, weights_bf16_lo=table16lo, pose2d=_lib.ptr(x), pose3d=_lib.ptr(pose3d),
                         feat=_lib.ptr(feat), workspace=_lib.ptr(ws), workspace_bytes=ws.numel())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gator_gat_forward(a, _lib.stream_ptr()), 'gator_gat_forward')
        return pose3d, feat


def get_model(num_joint=17, embed_dim=256, depth=4, graph_adj=None, GCN_depth=1, J_regressor=None, pretrained=False):
    return GAT(num_joint, embed_dim, depth, graph_adj, GCN_depth, J_regressor, pretrained=pretrained)
