"""CPU probe (no GPU): how much vertex error does a reduced-precision self-attention core cost end to end?

Emulates operand rounding of the 431x431 self-attention core inside the otherwise fp32 oracle forward:
q/k/v and P = exp(.) are rounded to bf16 / fp16 (one term) or kept as a hi+lo pair (== the 3-term split to first order).
Prints max-abs vertex error [m] and MPJPE drift [mm] against the fp64 oracle.  Lives under tests/ because it uses the oracle."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from helpers import oracle_setup, orc, synthetic, regressor, to_dtype


def rnd(x, dt):
    return x if dt is None else x.to(dt).to(x.dtype)


def make_attn(qk_dt, p_dt, v_dt, shift_bound=False):
    def _self_attn(sd, p, x, h=2):
        B, N, C = x.shape
        dk = C // h
        q, k, v = [F.linear(x, sd[f'{p}linears.{i}.weight'], sd[f'{p}linears.{i}.bias']).view(B, -1, h, dk).transpose(1, 2)
                   for i in range(3)]
        s = torch.matmul(rnd(q, qk_dt), rnd(k, qk_dt).transpose(-2, -1)) / math.sqrt(dk)
        if shift_bound:   # Cauchy-Schwarz bound instead of the row max
            b = (q.norm(dim=-1, keepdim=True) * k.norm(dim=-1).amax(-1, keepdim=True)[..., None]) / math.sqrt(dk)
            pexp = torch.exp(s - b)
        else:
            pexp = torch.exp(s - s.amax(-1, keepdim=True))
        l = pexp.sum(-1, keepdim=True)
        o = torch.matmul(rnd(pexp, p_dt), rnd(v, v_dt)) / l
        o = o.transpose(1, 2).contiguous().view(B, -1, h * dk)
        return F.linear(o, sd[p + 'linears.3.weight'], sd[p + 'linears.3.bias'])
    return _self_attn


bf, hf = torch.bfloat16, torch.float16
variants = [('fp32 attention', None, None, None, False),
            ('bf16 qk, bf16 p, bf16 v', bf, bf, bf, False),
            ('bf16 qk only', bf, None, None, False),
            ('bf16 p+v only', None, bf, bf, False),
            ('bf16 p only', None, bf, None, False),
            ('bf16 v only', None, None, bf, False),
            ('fp16 qk, fp16 p, fp16 v', hf, hf, hf, False),
            ('fp16 qk only', hf, None, None, False),
            ('fp16 p+v only', None, hf, hf, False),
            ('fp16 all, bound shift', hf, hf, hf, True),
            ('bf16 all, bound shift', bf, bf, bf, True)]

orig = orc._self_attn
for tag in ('coco', 'h36m'):
    sd, gc, mc, alpha = oracle_setup(tag)
    J = gc['J']
    x = torch.from_numpy(synthetic.poses2d(32, J, seed=11))
    with torch.no_grad():
        sd64 = to_dtype(sd, torch.float64)
        gc64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in gc.items()}
        mc64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in mc.items()}
        try:
            ref, _ = orc.gator_forward(sd64, gc64, mc64, x.double(), alpha)
            ref = ref.float()
        except Exception as e:   # fall back to fp32 reference
            print('fp64 oracle failed:', e)
            ref, _ = orc.gator_forward(sd, gc, mc, x, alpha)
        for name, a, b, c, sh in variants:
            orc._self_attn = make_attn(a, b, c, sh)
            mesh, _ = orc.gator_forward(sd, gc, mc, x, alpha)
            err = (mesh - ref).abs()
            mp, pa = orc.mpjpe_pa(mesh.numpy(), ref.numpy(), regressor('h36m'))
            print(f'{tag:5s} {name:28s} max-abs {err.max().item():.3e} m  mean-abs {err.mean().item():.3e}  MPJPE {mp:.4f} mm  PA {pa:.4f} mm')
    orc._self_attn = orig
