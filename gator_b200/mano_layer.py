"""Drop-in for ``manopth.manolayer.ManoLayer`` (manopth/manopth/manolayer.py:13-273): same constructor signature, same
registered buffers (``th_betas``, ``th_shapedirs``, ``th_posedirs``, ``th_v_template``, ``th_J_regressor``,
``th_weights``, ``th_faces``, ``th_hands_mean``, ``th_selected_comps``), same
``forward(th_pose_coeffs, th_betas, th_trans, root_palm, share_betas) -> (th_verts, th_jtr)`` in millimetres:
(B, 778, 3) vertices and (B, 21, 3) joints (16 chain joints + 5 finger tips, re-ordered).

MANO is the reference's second linear-blend-skinning model; its arithmetic is SMPL's with other
dimensions, so the forward is the generic LBS core ``gator_lbs_forward`` (csrc/smpl.cu: warp-per-chain pose kernel,
FFMA blend-shape GEMM, skinning) between two small MANO kernels (csrc/mano.cu): PCA coefficients -> axis-angle pose, and
tips / palm / re-ordering / centring / ``* 1000``.  fp32, CUDA only - no CPU fallback.  The axis-angle root and joint
rotation modes are built (``root_rot_mode='axisang'``, ``use_pca`` True or False with ``joint_rot_mode='axisang'``);
the 6-D and rotation-matrix input modes raise NotImplementedError.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.nn import Module

from . import _lib

N_JOINTS, N_VERTS, K_BLEND = 16, 778, 148       # 10 + 135 blend terms, zero-padded to a multiple of 4


class ManoLayer(Module):
    __constants__ = ['use_pca', 'rot', 'ncomps', 'ncomps', 'kintree_parents', 'check', 'side', 'center_idx', 'joint_rot_mode']

    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side='right', mano_root='mano/models', use_pca=True,
                 root_rot_mode='axisang', joint_rot_mode='axisang', robust_rot=False, mano_data=None):
        """`mano_data` (dict of numpy arrays: betas, shapedirs, posedirs, v_template, J_regressor (dense), weights, f,
        hands_components, hands_mean, kintree_table) bypasses the chumpy pkl loader, which needs the licensed MANO
        files; without it the reference's own loader is used if importable."""
        super().__init__()
        if root_rot_mode != 'axisang' or (not use_pca and joint_rot_mode != 'axisang'):
            raise NotImplementedError('gator_b200.ManoLayer: only the axis-angle rotation modes are built')
        self.center_idx = center_idx
        self.robust_rot = robust_rot
        self.rot = 3
        self.flat_hand_mean = flat_hand_mean
        self.side = side
        self.use_pca = use_pca
        self.joint_rot_mode = joint_rot_mode
        self.root_rot_mode = root_rot_mode
        self.ncomps = ncomps if use_pca else 45
        self.mano_path = os.path.join(mano_root, 'MANO_RIGHT.pkl' if side == 'right' else 'MANO_LEFT.pkl')
        if mano_data is None:
            try:
                from mano.webuser.smpl_handpca_wrapper_HAND_only import ready_arguments
            except Exception as e:   # chumpy missing
                raise RuntimeError('ManoLayer: the MANO pkl loader (mano.webuser, needs chumpy) is not importable; '
                                   'pass mano_data=...') from e
            d = ready_arguments(self.mano_path)
            r = lambda x: np.array(x.r if hasattr(x, 'r') else x)
            mano_data = {'betas': r(d['betas']), 'shapedirs': r(d['shapedirs']), 'posedirs': r(d['posedirs']),
                         'v_template': r(d['v_template']), 'J_regressor': np.array(d['J_regressor'].toarray()),
                         'weights': r(d['weights']), 'f': np.asarray(d['f']), 'hands_components': r(d['hands_components']),
                         'hands_mean': r(d['hands_mean']), 'kintree_table': np.asarray(d['kintree_table'])}
        self.smpl_data = mano_data
        T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        self.register_buffer('th_betas', T(mano_data['betas']).reshape(1, -1))
        self.register_buffer('th_shapedirs', T(mano_data['shapedirs']))
        self.register_buffer('th_posedirs', T(mano_data['posedirs']))
        self.register_buffer('th_v_template', T(mano_data['v_template']).reshape(1, -1, 3))
        self.register_buffer('th_J_regressor', T(mano_data['J_regressor']))
        self.register_buffer('th_weights', T(mano_data['weights']))
        self.register_buffer('th_faces', torch.as_tensor(np.asarray(mano_data['f']).astype(np.int32)).long())
        comps = np.asarray(mano_data['hands_components'], dtype=np.float32)
        mean = np.zeros(comps.shape[1], np.float32) if flat_hand_mean else np.asarray(mano_data['hands_mean'], np.float32).copy()
        self.register_buffer('th_hands_mean', T(mean).unsqueeze(0))
        self.register_buffer('th_selected_comps', T(comps[:ncomps]))
        self.kintree_table = np.asarray(mano_data['kintree_table'])
        self.kintree_parents = list(self.kintree_table[0].tolist())
        if len(self.kintree_parents) != N_JOINTS or self.th_v_template.shape[1] != N_VERTS:
            raise NotImplementedError('kernels are built for MANO (16 joints, 778 vertices)')
        for i in range(1, N_JOINTS):
            if not 0 <= int(self.kintree_parents[i]) < i:
                raise ValueError('kintree parents must precede their children')
        self._packed = None
        self._ws = None
        self.register_load_state_dict_post_hook(lambda module, keys: setattr(module, '_packed', None))

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    @torch.no_grad()
    def pack(self):
        dev = self.th_v_template.device
        if dev.type != 'cuda':
            raise RuntimeError('gator_b200.ManoLayer: buffers must be on a CUDA device (no CPU fallback)')
        d = torch.float64
        Jr = self.th_J_regressor.to(d)
        blend = torch.zeros(N_VERTS * 3, K_BLEND, dtype=torch.float32, device=dev)
        blend[:, :10] = self.th_shapedirs.reshape(N_VERTS * 3, 10)
        blend[:, 10:145] = self.th_posedirs.reshape(N_VERTS * 3, 135)
        W = self.th_weights
        kw = max(int((W != 0).sum(1).max().item()), 1)
        _, idx = torch.topk(W.abs(), kw, dim=1)
        idx, _ = torch.sort(idx, dim=1)
        self._packed = {
            'dev': dev, 'kw': kw,
            'parents': torch.tensor([0] + [int(p) for p in self.kintree_parents[1:]], dtype=torch.int32, device=dev),
            'j_template': (Jr @ self.th_v_template[0].to(d)).float().contiguous(),
            'j_shapedirs': torch.einsum('jv,vtk->jtk', Jr, self.th_shapedirs.to(d)).float().contiguous(),
            'default_betas': self.th_betas.reshape(-1).float().contiguous(),
            'blend_w': blend, 'v_template': self.th_v_template.reshape(-1).float().contiguous(),
            'skin_idx': idx.to(torch.int32).contiguous(), 'skin_w': torch.gather(W, 1, idx).float().contiguous(),
            'hands_mean': self.th_hands_mean.reshape(-1).float().contiguous(),
            'comps': self.th_selected_comps.float().contiguous(),
            'flag': torch.zeros(1, dtype=torch.int32, device=dev),
        }
        return self

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1), root_palm=torch.Tensor([0]),
                share_betas=torch.Tensor([0])):
        """pose coefficients (B, 3 + ncomps), betas (B,10), trans (B,3) -> (verts (B,778,3), joints (B,21,3)) in mm."""
        if self._packed is None:
            self.pack()
        p = self._packed
        dev = p['dev']
        if not th_pose_coeffs.is_cuda:
            raise RuntimeError('gator_b200.ManoLayer: inputs must be CUDA tensors (no CPU fallback)')
        if th_pose_coeffs.device != dev:
            raise RuntimeError(f'gator_b200.ManoLayer: input on {th_pose_coeffs.device} but the buffers are on {dev}')
        if th_pose_coeffs.dim() != 2 or th_pose_coeffs.shape[1] < self.rot + self.ncomps:
            raise ValueError(f'expected pose coefficients (B, >= {self.rot + self.ncomps}), got {tuple(th_pose_coeffs.shape)}')
        B = th_pose_coeffs.shape[0]
        coeffs = th_pose_coeffs.detach().float().contiguous()
        has_betas = th_betas is not None and th_betas.numel() != 1
        has_trans = th_trans is not None and th_trans.numel() != 1
        betas = None
        if has_betas:
            betas = th_betas.detach().to(dev).float()
            if bool(share_betas):                      # manolayer.py:183-184
                betas = betas.mean(0, keepdim=True).expand(betas.shape[0], 10)
            betas = betas.reshape(B, 10).contiguous()
        trans = th_trans.detach().to(dev).reshape(B, 3).float().contiguous() if has_trans else None
        verts = torch.empty((B, N_VERTS, 3), dtype=torch.float32, device=dev)
        jtr = torch.empty((B, 21, 3), dtype=torch.float32, device=dev)
        if B == 0:
            return verts, jtr
        L = _lib.lib()
        need = L.gator_lbs_workspace_bytes(B, N_VERTS, N_JOINTS)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        full_pose = torch.empty((B, 48), dtype=torch.float32, device=dev)
        jtr16 = torch.empty((B, N_JOINTS, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            _lib.check(L.gator_mano_pose(_lib.ptr(coeffs), coeffs.shape[1], self.ncomps,
                                         _lib.ptr(p['comps']) if self.use_pca else None, _lib.ptr(p['hands_mean']),
                                         _lib.ptr(full_pose), B, st), 'gator_mano_pose')
            a = _lib.LbsArgs(batch=B, n_verts=N_VERTS, n_joints=N_JOINTS, k_blend=K_BLEND, center_idx=-1,
                             has_betas=int(has_betas), has_trans=0, check_zero_norm=0, weights_per_vertex=p['kw'], out_scale=1.0,
                             parents=_lib.ptr(p['parents']), j_template=_lib.ptr(p['j_template']),
                             j_shapedirs=_lib.ptr(p['j_shapedirs']), default_betas=_lib.ptr(p['default_betas']),
                             blend_w=_lib.ptr(p['blend_w']), v_template=_lib.ptr(p['v_template']),
                             skin_idx=_lib.ptr(p['skin_idx']), skin_w=_lib.ptr(p['skin_w']), pose=_lib.ptr(full_pose),
                             betas=_lib.ptr(betas), trans=None, verts=_lib.ptr(verts), jtr=_lib.ptr(jtr16),
                             workspace=_lib.ptr(self._ws), workspace_bytes=self._ws.numel())
            _lib.check(L.gator_lbs_forward(a, st), 'gator_lbs_forward')
            tips = (745, 317, 444, 556, 673) if self.side == 'right' else (745, 317, 445, 556, 673)
            q = _lib.ManoPostArgs(batch=B, n_verts=N_VERTS, center_idx=-1 if self.center_idx is None else int(self.center_idx),
                                  has_trans=int(has_trans), check_zero_norm=int(has_trans and self.center_idx is not None),
                                  root_palm=int(bool(root_palm)), tip_verts=(_lib.C.c_int32 * 5)(*tips),
                                  palm_verts=(_lib.C.c_int32 * 2)(95, 22), scale=1000.0, reserved=0, jtr16=_lib.ptr(jtr16),
                                  trans=_lib.ptr(trans), verts=_lib.ptr(verts), jtr=_lib.ptr(jtr), flag_ws=_lib.ptr(p['flag']))
            _lib.check(L.gator_mano_post(q, st), 'gator_mano_post')
        return verts, jtr
