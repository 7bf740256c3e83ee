"""CPU probe (no GPU): end-to-end vertex error when selected GEMMs run with operands rounded to fp16 / bf16 (single term,
fp32 accumulate), emulated inside the fp32 oracle forward.  Decides which products need the 3-term split on the GPU.
Lives under tests/ because it uses the oracle."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from helpers import oracle_setup, orc, synthetic, regressor, to_dtype

GROUPS = {
    'mdr.so': lambda k: 'selfatt' in k and 'linears.3' in k,
    'mdr.qkv': lambda k: 'selfatt' in k and ('linears.0' in k or 'linears.1' in k or 'linears.2' in k),
    'mdr.wq': lambda k: 'pose2mesh.encoder' in k and 'attn.wq' in k,
    'mdr.wkv': lambda k: 'pose2mesh.encoder' in k and ('attn.wk' in k or 'attn.wv' in k),
    'mdr.proj': lambda k: 'pose2mesh.encoder' in k and 'attn.proj' in k,
    'mdr.fc1': lambda k: 'pose2mesh.encoder' in k and 'mlp.fc1' in k,
    'mdr.fc2': lambda k: 'pose2mesh.encoder' in k and 'mlp.fc2' in k,
    'mdr.head': lambda k: any(s in k for s in ('motion_linear', 'bias_linear', 'scale_linear')),
    'mdr.embed': lambda k: 'get_joint_feature' in k or 'get_verts_feature' in k,
    'gat.blocks': lambda k: 'pose_lifter.blocks' in k,
    'gat.lifter': lambda k: 'pose_lifter.lifter' in k,
}


def run(tag, nb=32):
    sd, gc, mc, alpha = oracle_setup(tag)
    J = gc['J']
    x = torch.from_numpy(synthetic.poses2d(nb, J, seed=11))
    sd64 = to_dtype(sd, torch.float64)
    cv = lambda d: {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
    with torch.no_grad():
        ref, _ = orc.gator_forward(sd64, cv(gc), cv(mc), x.double(), alpha)
    ref = ref.float()
    real_linear = F.linear
    real_conv = F.conv1d

    def measure(name, sel, dt, conv_dt=None):
        ids = {id(v) for k, v in sd.items() if k.endswith('weight') and any(GROUPS[g](k) for g in sel)}

        def lin(inp, w, b=None):
            if id(w) in ids:
                return real_linear(inp.to(dt).float(), w.to(dt).float(), b)
            return real_linear(inp, w, b)

        def conv(inp, w, b=None, **kw):
            if conv_dt is not None and w.shape[0] == 6890:
                return real_conv(inp.to(conv_dt).float(), w.to(conv_dt).float(), b, **kw)
            return real_conv(inp, w, b, **kw)
        shim = types.SimpleNamespace(**{n: getattr(F, n) for n in dir(F) if not n.startswith('__')})
        shim.linear = lin
        shim.conv1d = conv
        orc.F = shim
        try:
            with torch.no_grad():
                mesh, _ = orc.gator_forward(sd, gc, mc, x, alpha)
        finally:
            orc.F = F
        err = (mesh - ref).abs()
        mp, pa = orc.mpjpe_pa(mesh.numpy(), ref.numpy(), regressor('h36m'))
        print(f'{tag:5s} {name:44s} max-abs {err.max().item():.3e} m  mean {err.mean().item():.3e}  MPJPE {mp:.4f} PA {pa:.4f} mm', flush=True)

    hf, bf = torch.float16, torch.bfloat16
    measure('fp32', [], hf)
    for g in GROUPS:
        measure(f'fp16 {g}', [g], hf)
    layer = ['mdr.so', 'mdr.qkv', 'mdr.wq', 'mdr.proj', 'mdr.fc1', 'mdr.fc2']
    measure('fp16 all MDR layer gemms (row chain)', layer, hf)
    measure('fp16 MDR layer + wkv', layer + ['mdr.wkv'], hf)
    measure('fp16 MDR layer + wkv + head + embed', layer + ['mdr.wkv', 'mdr.head', 'mdr.embed'], hf)
    measure('fp16 GAT blocks + lifter', ['gat.blocks', 'gat.lifter'], hf)
    measure('fp16 everything (linear)', list(GROUPS), hf)
    measure('fp16 upsample_conv only', [], hf, conv_dt=hf)
    measure('fp16 everything + upsample', list(GROUPS), hf, conv_dt=hf)
    measure('bf16 all MDR layer gemms', layer, bf)


for tag in ('coco', 'h36m'):
    run(tag)
