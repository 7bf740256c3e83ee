"""Mesh.upsample 431->1723->6890 and J-regression at B=4096 for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import base_data_root, regressor
from gator_b200.mesh import Mesh
from gator_b200.ops import JointRegressor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
mesh = Mesh(os.path.join(base_data_root(), 'data', 'base_data', 'mesh_downsampling.npz'), device=dev)
x = torch.randn(B, 431, 3, device=dev)
reg = JointRegressor(regressor('h36m'), device=dev)
for _ in range(2):
    u1 = mesh.upsample(x, n1=2, n2=1)
    u0 = mesh.upsample(u1, n1=1, n2=0)
    j = reg(u0, scale=1000.0)
    d = mesh.downsample(u0, 0, 2)
torch.cuda.synchronize()
