"""GPU probe: fp16 self-attention core (gator_mdr_self_attention_f16) against fp64, and its timing with / without the image pack."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gator_b200 import _lib
dev = 'cuda:0'; L = _lib.lib()


def ref_attn(qkv, nb):
    q, k, v = [t.view(nb, 431, 2, 32).transpose(1, 2).double() for t in qkv.cpu().split(64, dim=1)]
    return (torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, -1) @ v).transpose(1, 2).reshape(nb * 431, 64)


def run_f16(qkv, nb):
    img = torch.empty(L.gator_mdr_self_attention_image_bytes(nb), dtype=torch.uint8, device=dev)
    o = torch.full((nb * 431, 64), float('nan'), device=dev)
    _lib.check(L.gator_mdr_self_attention_f16(qkv.data_ptr(), img.data_ptr(), o.data_ptr(), nb, _lib.stream_ptr()), 'sa_f16')
    torch.cuda.synchronize()
    return o


for nb, scale, tail in ((1, 1.0, 1.0), (5, 0.3, 1.0), (5, 1.5, 1.0), (3, 4.0, 1.0), (3, 1.0, 6.0), (301, 1.0, 1.0)):
    g = torch.Generator().manual_seed(nb)
    qkv = torch.randn(nb * 431, 192, generator=g) * scale
    if tail != 1.0:   # later keys get much larger logits: exercises the running-max update + O rescale
        qkv.view(nb, 431, 192)[:, 300:, 64:128] *= tail
    qkv = qkv.to(dev)
    ref = ref_attn(qkv, nb)
    o = run_f16(qkv, nb)
    e = (o.cpu().double() - ref).abs()
    print(f'nb={nb} scale={scale} tail={tail}: f16 max err {e.nan_to_num(9e9).max().item():.3e} mean {e.nan_to_num(0).mean().item():.3e} nan {int(torch.isnan(o).sum())}', flush=True)
    if e.nan_to_num(9e9).max() > 0.05:
        bad = (e.nan_to_num(9e9) > 0.05)
        print('   bad rows', bad.any(1).nonzero().flatten()[:12].tolist(), 'bad cols', bad.any(0).nonzero().flatten()[:12].tolist())
        print('   o[0,:6]', o[0, :6].tolist(), 'ref', ref[0, :6].tolist())

for nbt in (296, 4096):
    qkv = torch.randn(nbt * 431, 192, device=dev); o = torch.empty(nbt * 431, 64, device=dev)
    img = torch.empty(L.gator_mdr_self_attention_image_bytes(nbt), dtype=torch.uint8, device=dev)
    def t(fn, n=10):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    us = t(lambda: L.gator_mdr_self_attention_core(img.data_ptr(), o.data_ptr(), nbt, _lib.stream_ptr()))
    print(f'  nb={nbt} f16 core alone: {us:.1f} us  ({nbt*47.6e6/us/1e6:.1f} TFLOP/s algorithmic)')
    us = t(lambda: L.gator_mdr_self_attention_f16(qkv.data_ptr(), img.data_ptr(), o.data_ptr(), nbt, _lib.stream_ptr()))
    print(f'  nb={nbt} f16 kernel incl. image pack: {us:.1f} us  ({nbt*47.6e6/us/1e6:.1f} TFLOP/s algorithmic)')
