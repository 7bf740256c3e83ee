"""GPU probe: which tcgen05 operand forms behave as documented on this part (see csrc/probe.cu)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gator_b200 import _lib
dev = 'cuda:0'
L = _lib.lib()
L.gator_umma_forms_probe.restype = C.c_int
L.gator_umma_forms_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
names = {1: 'fp16', 2: 'B MN-major', 4: 'A in TMEM', 8: 'swap LBO/SBO'}
for (N, K) in ((32, 48), (96, 32), (64, 64)):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    for mode in (0, 3, 4, 5, 6, 7):
        D = torch.full((128, N), float('nan'), device=dev)
        st = L.gator_umma_forms_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mode, _lib.stream_ptr())
        torch.cuda.synchronize()
        dt = torch.float16 if mode & 1 else torch.bfloat16
        ref = A.to(dt).double() @ B.to(dt).double().t()
        err = (D.double() - ref).abs().nan_to_num(9e9).max().item()
        desc = ' + '.join(v for k, v in names.items() if mode & k) or 'bf16 SS K-major'
        print(f'N={N} K={K} mode={mode:2d} [{desc}]: status {st} max err {err:.3e} {"OK" if err < 1e-3 else "MISMATCH"}', flush=True)
