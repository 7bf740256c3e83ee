"""Mesh.upsample 431 -> 1723 -> 6890: fused single launch vs two launches, CUDA events, L2 flushed between calls.
python tools/mesh_time.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import base_data_root
from gator_b200.mesh import Mesh
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
mesh = Mesh(os.path.join(base_data_root(), 'data', 'base_data', 'mesh_downsampling.npz'), device=dev)
x = torch.randn(B, 431, 3, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


t2 = timed(lambda: mesh.upsample(mesh.upsample(x, n1=2, n2=1), n1=1, n2=0))
t1 = timed(lambda: mesh.upsample(x, n1=2, n2=0))
alg = B * (431 * 12 + 6890 * 12)
print(f'B={B}: two launches {t2 * 1e3:.1f} us, fused {t1 * 1e3:.1f} us = {alg / t1 / 1e6:.0f} GB/s of the '
      f'{alg / B / 1e3:.1f} KB/mesh algorithmic traffic', flush=True)
