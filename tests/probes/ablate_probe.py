"""Development probe (accuracy of the tensor-core paths against the CPU oracle); lives under tests/ because it uses the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from gator_b200 import _lib
from helpers import build_b200_gator, oracle_setup, orc, synthetic, regressor
dev = 'cuda:0'
tag = 'h36m'
sd, gc, mc, alpha = oracle_setup(tag)
x = torch.from_numpy(synthetic.poses2d(64, 17, seed=11))
with torch.no_grad():
    ref_mesh, ref_p3 = orc.gator_forward(sd, gc, mc, x, alpha)
m = build_b200_gator(tag, dev)
def report(name):
    mesh, p3 = m(x.to(dev)); torch.cuda.synchronize()
    err = (mesh.cpu() - ref_mesh).abs()
    mp, pa = orc.mpjpe_pa(mesh.cpu().numpy(), ref_mesh.numpy(), regressor('h36m'))
    print(f'{name:32s} max-abs {err.max().item():.3e} mean-abs {err.mean().item():.3e} MPJPE {mp:.4f} PA {pa:.4f}')
m.set_precision('fp32'); report('all fp32')
m.set_precision('bf16'); m.pose2mesh.precision = 0; report('GAT bf16 only')
m.pose_lifter.precision = 0; m.pose2mesh.precision = 1
for name, mask in (('MDR layer gemms', 2), ('self-attn', 4), ('head gemm', 8), ('upsample', 16), ('jf gemm', 32), ('layers+attn', 6), ('all but upsample', 2+4+8+32), ('all but upsample+head', 2+4+32)):
    m.pose2mesh.bf16_mask = mask; report('MDR: ' + name)
