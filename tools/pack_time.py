"""pack() time of the three modules on the device (why a disk cache of packed weights is not needed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import time
import torch
from builders import build_b200_gator, build_b200_smpl
m = build_b200_gator('coco', 'cuda:0').set_precision('bf16x3')
torch.cuda.synchronize()
for name, mod in (('GAT', m.pose_lifter), ('MDR', m.pose2mesh)):
    for i in range(3):
        mod._packed = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mod.pack()
        torch.cuda.synchronize(); print(name, 'pack %.1f ms' % ((time.perf_counter() - t0) * 1e3))
s = build_b200_smpl(device='cuda:0')
for i in range(3):
    s._packed = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.pack()
    torch.cuda.synchronize(); print('SMPL pack %.1f ms' % ((time.perf_counter() - t0) * 1e3))
