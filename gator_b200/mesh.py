"""Drop-in for ``models.backbones.mesh.Mesh`` (lib/models/backbones/mesh.py:60-123).

Same constructor signature and methods; ``downsample`` / ``upsample`` run the batched CSR SpMM CUDA
kernel (csrc/sparse.cu) over the whole batch in one launch instead of a Python loop of torch.sparse
products per sample (mesh.py:93-123, graph_layers.py:105-124).  CUDA only - no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, graph


def _load_npz(filename):
    """mesh.py:50-58: pickled object arrays A, U, D of scipy sparse matrices (GraphCMR format)."""
    data = np.load(filename, encoding='latin1', allow_pickle=True)
    return list(data['A']), list(data['U']), list(data['D'])


class _Csr:
    """One sparse operator resident on the device as CSR (int32 / fp32)."""

    def __init__(self, m, device):
        rp, ci, va, shape = graph.to_csr(m)
        self.shape = shape
        self.rowptr = torch.from_numpy(rp).to(device)
        self.colidx = torch.from_numpy(ci).to(device)
        self.values = torch.from_numpy(va).to(device)
        self.device = torch.device(device)
        # ELL copy (records of four (col, val) per row) for the fused two-level upsampling kernel; rows with fewer
        # non-zeros are padded with weight 0 on a column the row already uses
        ell = graph.csr_to_ell4(rp, ci, va)
        self.ell_width = ell[0] if ell else 0
        self.ell_col = torch.from_numpy(ell[1]).to(device) if ell else None
        self.ell_val = torch.from_numpy(ell[2]).to(device) if ell else None

    def apply(self, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """y[..., r, :] = scale * sum_k val_k x[..., col_k, :] for x of shape (N,F) or (B,N,F)."""
        if not x.is_cuda:
            raise RuntimeError('gator_b200.Mesh: CUDA tensors only (no CPU fallback)')
        if x.device != self.rowptr.device:
            raise RuntimeError(f'gator_b200.Mesh: input on {x.device} but the operator is on {self.rowptr.device}')
        if x.dtype != torch.float32:
            raise TypeError('gator_b200.Mesh: float32 only')
        squeeze = x.dim() == 2
        xb = x.unsqueeze(0) if squeeze else x
        if xb.dim() != 3 or xb.shape[1] != self.shape[1]:
            raise ValueError(f'expected (*, {self.shape[1]}, F), got {tuple(x.shape)}')
        xb = xb.contiguous()
        B, _, F = xb.shape
        y = torch.empty((B, self.shape[0], F), dtype=torch.float32, device=x.device)
        a = _lib.CsrArgs(batch=B, rows=self.shape[0], cols=self.shape[1], feat=F, scale=scale, reserved=0,
                         rowptr=_lib.ptr(self.rowptr), colidx=_lib.ptr(self.colidx), values=_lib.ptr(self.values),
                         x=_lib.ptr(xb), y=_lib.ptr(y))
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().gator_csr_spmm(a, _lib.stream_ptr()), 'gator_csr_spmm')
        return y[0] if squeeze else y


def _fusable(first: _Csr, second: _Csr, x) -> bool:
    """Two consecutive operators the fused kernel (gator_mesh_upsample2) takes: xyz features, ELL width <= 4, both
    coarse levels of four samples in 112 KB of shared memory."""
    return (x.dim() in (2, 3) and x.shape[-1] == 3 and first.ell_col is not None and second.ell_col is not None
            and second.shape[1] == first.shape[0] and (first.shape[1] + first.shape[0]) * 48 <= 112 * 1024)


def _apply2(first: _Csr, second: _Csr, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """second(first(x)) in one launch; same checks and bit-identical results as two _Csr.apply calls."""
    if not x.is_cuda:
        raise RuntimeError('gator_b200.Mesh: CUDA tensors only (no CPU fallback)')
    if x.device != first.rowptr.device:
        raise RuntimeError(f'gator_b200.Mesh: input on {x.device} but the operator is on {first.rowptr.device}')
    if x.dtype != torch.float32:
        raise TypeError('gator_b200.Mesh: float32 only')
    squeeze = x.dim() == 2
    xb = x.unsqueeze(0) if squeeze else x
    if xb.shape[1] != first.shape[1]:
        raise ValueError(f'expected (*, {first.shape[1]}, 3), got {tuple(x.shape)}')
    xb = xb.contiguous()
    B = xb.shape[0]
    y = torch.empty((B, second.shape[0], 3), dtype=torch.float32, device=x.device)
    a = _lib.Upsample2Args(batch=B, cols=first.shape[1], rows1=first.shape[0], rows2=second.shape[0],
                           width1=first.ell_width, width2=second.ell_width, scale=scale, reserved=0,
                           col1=_lib.ptr(first.ell_col), val1=_lib.ptr(first.ell_val),
                           col2=_lib.ptr(second.ell_col), val2=_lib.ptr(second.ell_val), x=_lib.ptr(xb), y=_lib.ptr(y))
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().gator_mesh_upsample2(a, _lib.stream_ptr()), 'gator_mesh_upsample2')
    return y[0] if squeeze else y


def adjmat_sparse(adjmat, nsize=1):
    """Row-normalised neighbourhood operator of a mesh level as a torch sparse COO tensor - what mesh.py:29-48 builds:
    pattern of A^nsize with unit weights, self loops added, every row divided by its number of entries."""
    import scipy.sparse
    pattern = scipy.sparse.csr_matrix(adjmat)
    reach = pattern.copy()
    for _ in range(nsize - 1):
        reach = reach @ pattern                       # walks of length <= nsize
    reach = (reach + scipy.sparse.identity(reach.shape[0], format='csr', dtype=reach.dtype)).tocsr()
    reach.sum_duplicates()
    reach.eliminate_zeros()
    counts = np.diff(reach.indptr)                    # entries per row, self loop included
    rows = np.repeat(np.arange(reach.shape[0]), counts)
    vals = (1.0 / counts)[rows]
    idx = torch.from_numpy(np.stack([rows, reach.indices]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(vals).float(), reach.shape, check_invariants=False)


class Mesh(object):
    """Mesh object that is used for handling certain graph operations (same API as the reference)."""

    def __init__(self, filename='data/base_data/mesh_downsampling.npz', num_downsampling=1, nsize=1,
                 device=torch.device('cuda')):
        A, U, D = _load_npz(filename)
        self._A = [adjmat_sparse(a, nsize=nsize) for a in A]
        self.device = torch.device(device)
        self._U_host, self._D_host = U, D          # scipy, for init-time host products
        self._U = [_Csr(u, self.device) for u in U] if self.device.type == 'cuda' else None
        self._D = [_Csr(d, self.device) for d in D] if self.device.type == 'cuda' else None
        self.num_downsampling = num_downsampling

    @property
    def adjmat(self):
        """Return the graph adjacency matrix at the specified subsampling level."""
        return self._A[self.num_downsampling].float()

    def _ops(self, which):
        ops = self._U if which == 'U' else self._D
        if ops is None:
            raise RuntimeError('gator_b200.Mesh was built with a CPU device: re-sampling kernels are CUDA only')
        return ops

    def downsample(self, x, n1=0, n2=None):
        """Downsample mesh: x (N,3) or (B,N,3) through D[n1..n2-1]."""
        if n2 is None:
            n2 = self.num_downsampling
        for i in range(n1, n2):
            x = self._ops('D')[i].apply(x)
        return x

    def upsample(self, x, n1=1, n2=0):
        """Upsample mesh: x (N,3) or (B,N,3) through U[n1-1..n2] (coarse to fine)."""
        ops = [self._ops('U')[i] for i in reversed(range(n2, n1))]
        i = 0
        while i < len(ops):
            if i + 1 < len(ops) and _fusable(ops[i], ops[i + 1], x):
                x = _apply2(ops[i], ops[i + 1], x)       # intermediate level stays in shared memory
                i += 2
            else:
                x = ops[i].apply(x)
                i += 1
        return x

    # init-time helper used by MDR.__init__ (MDR.py:79-81): same product as the reference's
    # torch.sparse COO @ dense, evaluated once on the host in float32.
    def downsample_host(self, x: torch.Tensor, n1=0, n2=None) -> torch.Tensor:
        import scipy.sparse
        if n2 is None:
            n2 = self.num_downsampling
        for i in range(n1, n2):
            d = scipy.sparse.coo_matrix(self._D_host[i])
            sp = torch.sparse_coo_tensor(torch.from_numpy(np.array([d.row, d.col])).long(),
                                         torch.from_numpy(d.data.astype(np.float32)), d.shape, check_invariants=False)
            x = torch.matmul(sp, x)
        return x
