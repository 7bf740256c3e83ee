"""Forwards at B=4096 (J=19, alpha) for ncu launch lists:  python tools/profile_step.py [precision] [batch] [forwards]
The first forward packs the weights (hundreds of small ATen launches); profile the last one, e.g. with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:'gator|kernel' --csv --log-file launches.csv python tools/profile_step.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_gator, golden, synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m = build_b200_gator('coco', 'cuda:0').set_precision(prec)
x = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], B)).to('cuda:0')
for _ in range(n):
    m(x)
torch.cuda.synchronize()
