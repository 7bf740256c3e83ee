"""Correctness and timing of the wide-N tcgen05 GEMM (csrc/umma_gemm_wide.cu) against the kernel it replaces,
at the two shapes of the path: upsample_conv (M = 3B, K = 1296, N = 6890) and SMPL blend shapes (K = 220, N = 20670)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gator_b200 import _lib, build
from gator_b200.packing import pack_umma_weight_pair, pack_umma_wide

build.build()
dev = torch.device('cuda:0')
L = _lib.lib()


def run(M, N, K, wide, iters=10, check=True, ldc=None):
    g = torch.Generator(device='cpu').manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ldc = ldc or N
    Cb = torch.zeros(M, ldc, device=dev)
    hi, lo = pack_umma_weight_pair(W)
    Ww = pack_umma_wide(W) if wide else None
    ws = torch.empty(L.gator_umma_wide_a_bytes(M, K), dtype=torch.uint8, device=dev) if wide else None
    a = _lib.GemmArgs(M=M, N=N, K=K, lda=K, ldw=K, ldc=ldc, ldr=0, act=0, bias_period=0, precision=2,
                      A=_lib.ptr(A), W=_lib.ptr(hi), W_lo=_lib.ptr(lo), W_wide=_lib.ptr(Ww), a_image=_lib.ptr(ws),
                      a_image_bytes=ws.numel() if wide else 0, bias=_lib.ptr(bias), bias_rows=None, R=None, C=_lib.ptr(Cb))
    fn = lambda: _lib.check(L.gator_gemm(a, _lib.stream_ptr()), 'gator_gemm')
    fn()
    torch.cuda.synchronize()
    err = None
    if check:
        rows = torch.randint(0, M, (64,), generator=g).tolist() + [0, M - 1]
        ref = A[rows].double() @ W.double().t() + bias.double()
        err = float((Cb[rows, :N].double() - ref).abs().max())
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({'M': M, 'N': N, 'K': K, 'wide': wide, 'ms': round(ms, 4), 'max_err': err,
                      'TFLOPs_x3': round(3 * 2.0 * M * N * K / ms / 1e9, 1)}), flush=True)


for M, N, K in ((130, 300, 72), (1000, 700, 220), (4096 * 3, 6890, 1296), (8192 * 3, 6890, 1296), (1024, 20670, 220)):
    for wide in (False, True):
        run(M, N, K, wide, ldc=20672 if N == 20670 else None)
