"""SMPL_Layer forward at B=16384 for ncu:  python tools/profile_smpl.py [precision] [batch]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_smpl, synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
layer = build_b200_smpl(device='cuda:0').set_precision(prec)
P, Bt, T = [torch.from_numpy(a).to('cuda:0') for a in synthetic.smpl_inputs(B)]
for _ in range(2):
    layer(P, Bt, T)
torch.cuda.synchronize()
