"""Drop-in for ``smplpytorch.pytorch.smpl_layer.SMPL_Layer``
(smplpytorch/smplpytorch/pytorch/smpl_layer.py:12-158): same constructor signature, same registered
buffers / attributes (``th_faces``, ``th_J_regressor``, ``th_weights``, ``kintree_table``,
``kintree_parents``, ``num_joints`` - read by lib/smpl.py:16-17 and the datasets), same
``forward(th_pose_axisang, th_betas, th_trans) -> (th_verts, th_jtr)``.

The forward runs three CUDA kernels behind ``gator_smpl_forward`` (csrc/smpl.cu): warp-per-sample
Rodrigues + kinematic chain, the blend-shape GEMM, and float4-vectorised skinning.  The reference's two
host synchronisations (``bool(torch.norm(x) == 0)``, smpl_layer.py:87,148) are reproduced on the device
only when they can change the result.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.nn import Module

from . import _lib
from .packing import pack_umma_weight_pair, pack_umma_wide, pack_umma_wide_a

N_JOINTS, N_VERTS, K_BLEND = 24, 6890, 220


def _invalidate_hook(module, incompatible_keys):
    module._packed = None


class SMPL_Layer(Module):
    __constants__ = ['kintree_parents', 'gender', 'center_idx', 'num_joints']

    def __init__(self, center_idx=None, gender='neutral', model_root='smpl/native/models', smpl_data=None):
        """`smpl_data` (dict of numpy arrays: betas, shapedirs, posedirs, v_template, J_regressor (dense),
        weights, f, kintree_table) bypasses the chumpy pkl loader, which needs the licensed SMPL model
        files; without it the reference's own loader is used if importable."""
        super().__init__()
        self.center_idx = center_idx
        self.gender = gender
        names = {'neutral': 'basicModel_neutral_lbs_10_207_0_v1.0.0.pkl', 'female': 'basicModel_f_lbs_10_207_0_v1.0.0.pkl',
                 'male': 'basicModel_m_lbs_10_207_0_v1.0.0.pkl'}
        self.model_path = os.path.join(model_root, names[gender])
        if smpl_data is None:
            try:
                from smplpytorch.native.webuser.serialization import ready_arguments
            except Exception as e:   # chumpy / cv2 missing
                raise RuntimeError('SMPL_Layer: the SMPL pkl loader (smplpytorch.native.webuser.serialization, '
                                   'needs chumpy) is not importable; pass smpl_data=...') from e
            d = ready_arguments(self.model_path)
            r = lambda x: np.array(x.r if hasattr(x, 'r') else x)
            smpl_data = {'betas': r(d['betas']), 'shapedirs': r(d['shapedirs']), 'posedirs': r(d['posedirs']),
                         'v_template': r(d['v_template']), 'J_regressor': np.array(d['J_regressor'].toarray()),
                         'weights': r(d['weights']), 'f': np.asarray(d['f']), 'kintree_table': np.asarray(d['kintree_table'])}
        self.smpl_data = smpl_data
        T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        self.register_buffer('th_betas', T(smpl_data['betas']).reshape(1, -1))
        self.register_buffer('th_shapedirs', T(smpl_data['shapedirs']))
        self.register_buffer('th_posedirs', T(smpl_data['posedirs']))
        self.register_buffer('th_v_template', T(smpl_data['v_template']).reshape(1, -1, 3))
        self.register_buffer('th_J_regressor', T(smpl_data['J_regressor']))
        self.register_buffer('th_weights', T(smpl_data['weights']))
        self.register_buffer('th_faces', torch.as_tensor(np.asarray(smpl_data['f']).astype(np.int32)).long())
        self.vertice_segmentation = torch.argmax(self.th_weights, dim=1)
        self.kintree_table = np.asarray(smpl_data['kintree_table'])
        self.kintree_parents = list(self.kintree_table[0].tolist())
        self.num_joints = len(self.kintree_parents)
        if self.num_joints != N_JOINTS or self.th_v_template.shape[1] != N_VERTS:
            raise NotImplementedError('kernels are built for SMPL (24 joints, 6890 vertices)')
        for i in range(1, N_JOINTS):
            if not 0 <= int(self.kintree_parents[i]) < i:
                raise ValueError('kintree parents must precede their children')
        self._packed = None
        self._ws = None
        self.precision = _lib.PREC_FP32
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def set_precision(self, precision: str):
        self.precision = _lib.PRECISIONS[precision]
        return self

    @classmethod
    def from_buffers(cls, buffers, kintree_parents, center_idx=None, gender='neutral'):
        """Build from arrays shaped like the registered buffers (synthetic / already-extracted models)."""
        parents = [int(p) for p in kintree_parents]
        table = np.stack([np.asarray([4294967295] + parents[1:], dtype=np.int64), np.arange(len(parents))])
        g = lambda k: np.asarray(buffers[k].cpu() if torch.is_tensor(buffers[k]) else buffers[k])
        data = {'betas': g('th_betas'), 'shapedirs': g('th_shapedirs'), 'posedirs': g('th_posedirs'),
                'v_template': g('th_v_template'), 'J_regressor': g('th_J_regressor'), 'weights': g('th_weights'),
                'f': g('th_faces'), 'kintree_table': table}
        return cls(center_idx=center_idx, gender=gender, smpl_data=data)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    @torch.no_grad()
    def pack(self):
        dev = self.th_v_template.device
        if dev.type != 'cuda':
            raise RuntimeError('gator_b200.SMPL_Layer: buffers must be on a CUDA device (no CPU fallback)')
        d = torch.float64
        Jr = self.th_J_regressor.to(d)
        S = self.th_shapedirs.to(d)                                         # (6890,3,10)
        P = self.th_posedirs.to(d)                                          # (6890,3,207)
        Tm = self.th_v_template[0].to(d)                                    # (6890,3)
        blend = torch.zeros(N_VERTS * 3, K_BLEND, dtype=torch.float32, device=dev)
        blend[:, :10] = self.th_shapedirs.reshape(N_VERTS * 3, 10)
        blend[:, 10:217] = self.th_posedirs.reshape(N_VERTS * 3, 207)
        W = self.th_weights
        kw = int((W != 0).sum(1).max().item())
        kw = max(kw, 1)
        vals, idx = torch.topk(W.abs(), kw, dim=1)
        idx, _ = torch.sort(idx, dim=1)
        skin_w = torch.gather(W, 1, idx)
        parents = [0] + [int(p) for p in self.kintree_parents[1:]]
        self._packed = {
            'dev': dev, 'kw': kw,
            'parents': torch.tensor(parents, dtype=torch.int32, device=dev),
            'j_template': (Jr @ Tm).float().contiguous(),                                   # (24,3)
            'j_shapedirs': torch.einsum('jv,vtk->jtk', Jr, S).float().contiguous(),          # (24,3,10)
            'default_betas': self.th_betas.reshape(-1).float().contiguous(),
            'betas_nonzero': bool((self.th_betas != 0).any().item()),
            'blend_w': blend, 'blend_w_bf16': pack_umma_weight_pair(blend), 'blend_w_wide': pack_umma_wide(blend), 'v_template': self.th_v_template.reshape(-1).float().contiguous(),
            'skin_idx': idx.to(torch.int32).contiguous(), 'skin_w': skin_w.float().contiguous(),
            'skin_w_img': pack_umma_wide_a(W.float()),
        }
        return self

    def forward(self, th_pose_axisang, th_betas=torch.zeros(1), th_trans=torch.zeros(1)):
        """pose (B,72) axis-angle, betas (B,10), trans (B,3) -> (verts (B,6890,3), joints (B,24,3)) metres."""
        return self.forward_scaled(th_pose_axisang, th_betas, th_trans, 1.0)

    def forward_scaled(self, th_pose_axisang, th_betas=torch.zeros(1), th_trans=torch.zeros(1), out_scale: float = 1.0):
        """`forward` with both outputs multiplied by `out_scale` after the translation, in the skinning kernel
        (the datasets' ``*= 1000``, data/Human36M/dataset.py:296, data/PW3D/dataset.py:99-100)."""
        if self._packed is None:
            self.pack()
        p = self._packed
        dev = p['dev']
        if not th_pose_axisang.is_cuda:
            raise RuntimeError('gator_b200.SMPL_Layer: inputs must be CUDA tensors (no CPU fallback)')
        if th_pose_axisang.device != dev:
            raise RuntimeError(f'gator_b200.SMPL_Layer: input on {th_pose_axisang.device} but the buffers are on {dev}')
        B = th_pose_axisang.shape[0]
        pose = th_pose_axisang.detach().reshape(B, 72).float().contiguous()
        has_betas = th_betas is not None and th_betas.numel() != 1      # zeros(1) is the "not given" sentinel
        has_trans = th_trans is not None and th_trans.numel() != 1
        betas = th_betas.detach().to(dev).reshape(B, 10).float().contiguous() if has_betas else None
        trans = th_trans.detach().to(dev).reshape(B, 3).float().contiguous() if has_trans else None
        # `norm(x)==0` switches (smpl_layer.py:87,148) only matter if th_betas != 0 or a centre joint is set
        check = (has_betas and p['betas_nonzero']) or (has_trans and self.center_idx is not None)
        verts = torch.empty((B, N_VERTS, 3), dtype=torch.float32, device=dev)
        jtr = torch.empty((B, N_JOINTS, 3), dtype=torch.float32, device=dev)
        if B == 0:
            return verts, jtr
        need = _lib.lib().gator_smpl_workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        a = _lib.SmplArgs(batch=B, center_idx=-1 if self.center_idx is None else int(self.center_idx),
                          has_betas=int(has_betas), has_trans=int(has_trans), check_zero_norm=int(check),
                          weights_per_vertex=p['kw'], precision=self.precision, out_scale=float(out_scale),
                          parents=_lib.ptr(p['parents']), j_template=_lib.ptr(p['j_template']),
                          j_shapedirs=_lib.ptr(p['j_shapedirs']), default_betas=_lib.ptr(p['default_betas']),
                          blend_w=_lib.ptr(p['blend_w']), blend_w_bf16=_lib.ptr(p['blend_w_bf16'][0]), blend_w_bf16_lo=_lib.ptr(p['blend_w_bf16'][1]), blend_w_wide=_lib.ptr(p['blend_w_wide']), v_template=_lib.ptr(p['v_template']),
                          skin_idx=_lib.ptr(p['skin_idx']), skin_w=_lib.ptr(p['skin_w']), skin_w_img=_lib.ptr(p['skin_w_img']),
                          pose=_lib.ptr(pose),
                          betas=_lib.ptr(betas), trans=_lib.ptr(trans), verts=_lib.ptr(verts), jtr=_lib.ptr(jtr),
                          workspace=_lib.ptr(self._ws), workspace_bytes=self._ws.numel())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gator_smpl_forward(a, _lib.stream_ptr()), 'gator_smpl_forward')
        return verts, jtr
