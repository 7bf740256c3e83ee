import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gator_b200 import _lib
dev = 'cuda:0'; L = _lib.lib()
nb = 5
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(nb * 431, 192, generator=g) * 1.5).to(dev)
q, k, v = [t.view(nb, 431, 2, 32).transpose(1, 2).double() for t in qkv.cpu().split(64, dim=1)]
ref = (torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, -1) @ v).transpose(1, 2).reshape(nb * 431, 64)
for prec in (0, 1, 2):
    o = torch.full((nb * 431, 64), float('nan'), device=dev)
    _lib.check(L.gator_mdr_self_attention(qkv.data_ptr(), o.data_ptr(), nb, prec, _lib.stream_ptr()), 'sa')
    torch.cuda.synchronize()
    e = (o.cpu().double() - ref).abs()
    print(f'prec {prec}: max err {e.nan_to_num(9e9).max().item():.3e} mean {e.nan_to_num(0).mean().item():.3e} mean signed {(o.cpu().double()-ref).mean().item():.2e}')
nbt = 296
qkv = torch.randn(nbt * 431, 192, device=dev); o = torch.empty(nbt * 431, 64, device=dev)
for prec in (0, 1, 2):
    for _ in range(3): L.gator_mdr_self_attention(qkv.data_ptr(), o.data_ptr(), nbt, prec, _lib.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.gator_mdr_self_attention(qkv.data_ptr(), o.data_ptr(), nbt, prec, _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print(f'  nb={nbt} prec={prec}: {e0.elapsed_time(e1)/10*1e3:.1f} us')
