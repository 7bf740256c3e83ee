import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from builders import build_b200_gator, golden, synthetic
m = build_b200_gator('coco', 'cuda:0').set_precision('bf16x3')
x = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], 4096)).to('cuda:0')
for chunk in (592, 1036, 1184, 1332, 1480, 2072, 4096):
    m.pose2mesh.chunk = chunk
    for _ in range(2): m(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): m(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f'MDR chunk {chunk}: {dt*1e3:.2f} ms/step  {4096/dt:.0f} meshes/s')
