"""GPU probe: bf16 (tcgen05) path vs fp32 path vs CPU oracle - attention kernel error, end-to-end drift, timing."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from gator_b200 import _lib
from helpers import build_b200_gator, build_b200_smpl, oracle_setup, orc, synthetic, regressor, golden

dev = 'cuda:0'
L = _lib.lib()

# --- attention kernel alone ---
nb = 5
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(nb * 431, 192, generator=g) * 1.5).to(dev)
o32 = torch.zeros(nb * 431, 64, device=dev); o16 = torch.full((nb * 431, 64), float('nan'), device=dev)
_lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), o32.data_ptr(), nb, 0, _lib.stream_ptr()), 'sa32')
_lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), o16.data_ptr(), nb, 1, _lib.stream_ptr()), 'sa16')
o48 = torch.full((nb * 431, 64), float('nan'), device=dev)
_lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), o48.data_ptr(), nb, 2, _lib.stream_ptr()), 'sa48')
torch.cuda.synchronize()
q, k, v = [t.view(nb, 431, 2, 32).transpose(1, 2).double() for t in qkv.cpu().split(64, dim=1)]
ref = (torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, -1) @ v).transpose(1, 2).reshape(nb * 431, 64)
print('self-attn fp32 kernel max err', (o32.cpu().double() - ref).abs().max().item())
e16 = (o16.cpu().double() - ref).abs()
print('self-attn umma x3 kernel max err', (o48.cpu().double() - ref).abs().nan_to_num(9e9).max().item())
print('self-attn umma kernel max err', e16.nan_to_num(9e9).max().item(), 'mean', e16.nan_to_num(0).mean().item(), 'nan', int(torch.isnan(o16).sum()))
if e16.nan_to_num(9e9).max() > 0.05:
    bad = (e16 > 0.05)
    print('  bad rows', bad.any(1).nonzero().flatten()[:10].tolist(), 'bad cols', bad.any(0).nonzero().flatten()[:10].tolist())
    print('  o16[0,:8]', o16[0, :8].tolist()); print('  ref[0,:8]', ref[0, :8].tolist())

for nbt in (148, 296):
    qkv = torch.randn(nbt * 431, 192, device=dev); o = torch.empty(nbt * 431, 64, device=dev)
    for prec in (0, 1, 2):
        for _ in range(2): L.gator_mdr_self_attention(qkv.data_ptr(), o.data_ptr(), nbt, prec, _lib.stream_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): L.gator_mdr_self_attention(qkv.data_ptr(), o.data_ptr(), nbt, prec, _lib.stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'  self-attn nb={nbt} prec={prec}: {ms*1e3:.1f} us  {nbt*2*2*2*431*431*32/ms/1e9:.1f} TFLOP/s')

# --- end-to-end drift ---
for tag in ('h36m', 'coco'):
    sd, gc, mc, alpha = oracle_setup(tag)
    J = gc['J']
    x = torch.from_numpy(synthetic.poses2d(64, J, seed=11))
    with torch.no_grad():
        ref_mesh, ref_p3 = orc.gator_forward(sd, gc, mc, x, alpha)
    m = build_b200_gator(tag, dev)
    for prec in ('fp32', 'bf16', 'bf16x3'):
        m.set_precision(prec)
        mesh, p3 = m(x.to(dev))
        torch.cuda.synchronize()
        err = (mesh.cpu() - ref_mesh).abs()
        mp, pa = orc.mpjpe_pa(mesh.cpu().numpy(), ref_mesh.numpy(), regressor('h36m'))
        print(f'{tag} {prec}: mesh max-abs {err.max().item():.3e} m, mean-abs {err.mean().item():.3e} m, pose3d max {(p3.cpu()-ref_p3).abs().max().item():.3e} mm, '
              f'MPJPE drift {mp:.4f} mm, PA-MPJPE drift {pa:.4f} mm, nan {int(torch.isnan(mesh).sum())}')
    if tag == 'coco':
        xb = torch.from_numpy(synthetic.coco_poses2d(golden('fixtures')['demo_pose19'], 4096)).to(dev)
        for prec in ('fp32', 'bf16', 'bf16x3'):
            m.set_precision(prec)
            for _ in range(2): m(xb)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5): m(xb)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
            print(f'   B=4096 {prec}: {dt*1e3:.2f} ms/step  {4096/dt:.0f} meshes/s')
            # per-stage
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); p3, feat = m.pose_lifter(xb.view(4096, -1)); e[1].record()
            m.pose2mesh.forward_parts(xb, p3.view(4096, -1, 3), feat); e[2].record(); torch.cuda.synchronize()
            print(f'      GAT {e[0].elapsed_time(e[1]):.2f} ms, MDR {e[1].elapsed_time(e[2]):.2f} ms')

# --- SMPL ---
buf = {k: torch.from_numpy(v) for k, v in synthetic.smpl_buffers().items()}
pose, betas, trans = [torch.from_numpy(a) for a in synthetic.smpl_inputs(64)]
rv, rj, _ = orc.smpl_forward(buf, synthetic.SMPL_PARENTS, pose, betas, trans)
layer = build_b200_smpl(device=dev)
for prec in ('fp32', 'bf16', 'bf16x3'):
    layer.set_precision(prec)
    v, j = layer(pose.to(dev), betas.to(dev), trans.to(dev))
    print(f'SMPL {prec}: verts max-abs {(v.cpu()-rv).abs().max().item():.3e} m, jtr {(j.cpu()-rj).abs().max().item():.3e}')
    B = 16384
    P, Bt, T = [torch.from_numpy(a).to(dev) for a in synthetic.smpl_inputs(B)]
    for _ in range(2): layer(P, Bt, T)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): layer(P, Bt, T)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f'   SMPL B={B} {prec}: {dt*1e3:.2f} ms  {B/dt:.0f} meshes/s  out {B*82680/dt/1e9:.0f} GB/s')
