"""One forward at B=4096 (J=19, alpha) for ncu launch lists:  python tools/profile_step.py [precision] [batch]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_gator, golden, synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
m = build_b200_gator('coco', 'cuda:0').set_precision(prec)
x = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], B)).to('cuda:0')
for _ in range(2):
    m(x)
torch.cuda.synchronize()
