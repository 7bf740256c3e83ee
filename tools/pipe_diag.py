"""Device-to-host bandwidth and HostPipeline step time for a few slicing schemes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, time
from builders import build_b200_gator, golden, synthetic
from gator_b200.pipeline import HostPipeline
m = build_b200_gator('coco', 'cuda:0').set_precision('bf16x3')
x = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], 4096)).pin_memory()
buf = torch.empty(64 << 20, dtype=torch.uint8, device='cuda:0'); host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); host.copy_(buf, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print('D2H 64 MiB: %.2f ms = %.1f GB/s' % (e0.elapsed_time(e1), 64 * 1.048576 / e0.elapsed_time(e1)))
for ss in (0, 1184, 2048):
    pipe = HostPipeline(m, 4096, slice_samples=ss)
    for _ in range(2): pipe.forward(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): pipe.forward(x)
    torch.cuda.synchronize()
    print('slice', ss, 'bounds', pipe._bounds, '%.2f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
