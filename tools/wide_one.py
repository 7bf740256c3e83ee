"""One wide-GEMM launch at the SMPL blend-shape shape, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from gator_b200 import _lib
from gator_b200.packing import pack_umma_weight_pair, pack_umma_wide
dev = torch.device('cuda:0'); L = _lib.lib()
M, N, K, ldc = 1024, 20670, 220, 20672
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; bias = torch.randn(N, device=dev)
Cb = torch.zeros(M, ldc, device=dev)
Ww = pack_umma_wide(W); ws = torch.empty(L.gator_umma_wide_a_bytes(M, K), dtype=torch.uint8, device=dev)
a = _lib.GemmArgs(M=M, N=N, K=K, lda=K, ldw=K, ldc=ldc, ldr=0, act=0, bias_period=0, precision=2, A=_lib.ptr(A), W=None, W_lo=None,
                  W_wide=_lib.ptr(Ww), a_image=_lib.ptr(ws), a_image_bytes=ws.numel(), bias=_lib.ptr(bias), bias_rows=None, R=None, C=_lib.ptr(Cb))
for _ in range(3):
    _lib.check(L.gator_gemm(a, _lib.stream_ptr()), 'g')
torch.cuda.synchronize()
