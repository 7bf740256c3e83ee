"""SMPL_Layer forward timing (bf16x3 and fp32) at a given batch: python tools/smpl_time.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_smpl, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P, Bt, T = [torch.from_numpy(a).to('cuda:0') for a in synthetic.smpl_inputs(B)]
for prec in ('bf16x3', 'fp32'):
    layer = build_b200_smpl(device='cuda:0').set_precision(prec)
    for _ in range(3): layer(P, Bt, T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): layer(P, Bt, T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{prec} B={B}: {ms:.3f} ms  {B / ms / 1e3:.2f} M meshes/s  {B * 82680 / ms / 1e6:.0f} GB/s out', flush=True)
