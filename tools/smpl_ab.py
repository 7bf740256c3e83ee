"""A/B of the SMPL skinning kernels (bf16x3): tensor-core blend (smpl_skin_umma_kernel) vs CUDA-core ELL blend
(smpl_skin_kernel); alternating rounds, CUDA events.  python tools/smpl_ab.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_smpl, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layer = build_b200_smpl(device='cuda:0').set_precision('bf16x3')
P, Bt, T = [torch.from_numpy(a).to('cuda:0') for a in synthetic.smpl_inputs(B)]
layer(P, Bt, T)
img = layer._packed['skin_w_img']
def run(use):
    layer._packed['skin_w_img'] = img if use else None
    for _ in range(2): layer(P, Bt, T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): layer(P, Bt, T)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for r in range(3):
    print('round', r, 'umma %.3f ms' % run(True), ' ell %.3f ms' % run(False), flush=True)
