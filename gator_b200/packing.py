"""Weight packing for the tcgen05 kernels (run once per pack(), on the device, with torch ops)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def umma_weight_layout(N: int, K: int):
    """(BN, n_tiles, K_pad) as chosen by the library (csrc/umma_gemm.cu: umma_weight_layout)."""
    bn, nt, kp = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().gator_umma_weight_layout(N, K, C.byref(bn), C.byref(nt), C.byref(kp)), 'gator_umma_weight_layout')
    return bn.value, nt.value, kp.value


@torch.no_grad()
def pack_umma_weight(W: torch.Tensor) -> torch.Tensor:
    """W (N,K) float -> bf16 [N_pad/8, K_pad/8, 8, 8]: 8x8 core matrices, K-major, no swizzle - exactly the
    shared-memory image tcgen05.mma reads, so the kernel stages it with linear 16-byte copies."""
    N, K = W.shape
    BN, n_tiles, K_pad = umma_weight_layout(N, K)
    N_pad = BN * n_tiles
    Wp = torch.zeros((N_pad, K_pad), dtype=torch.float32, device=W.device)
    Wp[:N, :K] = W.float()
    Wp = Wp.reshape(N_pad // 8, 8, K_pad // 8, 8).permute(0, 2, 1, 3).contiguous()
    return Wp.to(torch.bfloat16).contiguous()


@torch.no_grad()
def pack_umma_weight_pair(W: torch.Tensor):
    """(hi, lo) packed bf16 tensors with W ~= hi + lo (lo = bf16 of the rounding residual), for the 3-term
    split product of GATOR_PREC_BF16X3."""
    Wf = W.float()
    hi = Wf.to(torch.bfloat16).float()
    return pack_umma_weight(hi), pack_umma_weight(Wf - hi)


def _pack_groups(ws):
    """Stack matrices of equal shape and pack each stack with one set of device ops.
    Yields (indices, both) with both = bf16 (n, 2 (hi|lo), N_pad/8, K_pad/8, 8, 8)."""
    groups = {}
    for i, w in enumerate(ws):
        groups.setdefault(tuple(w.shape), []).append(i)
    for (N, K), idx in groups.items():
        BN, n_tiles, K_pad = umma_weight_layout(N, K)
        N_pad = BN * n_tiles
        W = torch.stack([ws[i].float() for i in idx])                       # (n, N, K)
        if (N_pad, K_pad) != (N, K):
            Wp = torch.zeros((len(idx), N_pad, K_pad), dtype=torch.float32, device=W.device)
            Wp[:, :N, :K] = W
            W = Wp
        hi = W.to(torch.bfloat16)
        lo = (W - hi.float()).to(torch.bfloat16)
        both = torch.stack([hi, lo], 1)                                     # (n, 2, N_pad, K_pad)
        yield idx, both.reshape(len(idx), 2, N_pad // 8, 8, K_pad // 8, 8).permute(0, 1, 2, 4, 3, 5).contiguous()


@torch.no_grad()
def pack_umma_weight_pairs(ws):
    """[W_i (N_i, K_i)] -> [(hi_i, lo_i)], the same images as pack_umma_weight_pair, but matrices of equal shape are stacked
    and packed together (pack() of a model issues a few dozen launches instead of several thousand)."""
    out = [None] * len(ws)
    for idx, both in _pack_groups(ws):
        for j, i in enumerate(idx):
            out[i] = (both[j, 0], both[j, 1])
    return out


@torch.no_grad()
def pack_umma_blob(ws) -> torch.Tensor:
    """[W_i] -> one contiguous bf16 tensor [hi_0 | lo_0 | hi_1 | lo_1 | ...] (the weight-unit blobs of the fused kernels);
    every W_i must pack to the same number of elements."""
    blob = None
    for idx, both in _pack_groups(ws):
        flat = both.reshape(len(idx), 2, -1)
        if blob is None:
            blob = torch.empty((len(ws), 2, flat.shape[2]), dtype=torch.bfloat16, device=flat.device)
        blob[torch.tensor(idx, device=flat.device)] = flat
    return blob.reshape(-1)


@torch.no_grad()
def pack_umma_wide(W: torch.Tensor) -> torch.Tensor:
    """W (N,K) float -> bf16 [n_tiles, kblocks, 2 (hi|lo), 32, 4, 8, 8] for the wide-N kernel (csrc/umma_gemm_wide.cu):
    256-row tiles, 32-wide K blocks, 8x8 core matrices; one (tile, block) is a contiguous 32 KB chunk."""
    N, K = W.shape
    nt, kb = C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().gator_umma_wide_layout(N, K, C.byref(nt), C.byref(kb)), 'gator_umma_wide_layout')
    nt, kb = nt.value, kb.value
    Wp = torch.zeros((nt * 256, kb * 32), dtype=torch.float32, device=W.device)
    Wp[:N, :K] = W.float()
    hi = Wp.to(torch.bfloat16)
    lo = (Wp - hi.float()).to(torch.bfloat16)
    both = torch.stack([hi, lo])                                            # (2, N_pad, K_pad)
    both = both.reshape(2, nt, 32, 8, kb, 4, 8)                             # (h, tile, rg, r, block, kc, e)
    return both.permute(1, 4, 0, 2, 5, 3, 6).contiguous()                   # (tile, block, h, rg, kc, r, e)


@torch.no_grad()
def pack_umma_wide_a(W: torch.Tensor) -> torch.Tensor:
    """W (M,K) float -> bf16 [m_tiles, kblocks, 2 (hi|lo), 16, 4, 8, 8]: the 128-row (M-side) tile image of the wide
    kernels - the layout csrc/umma_gemm_wide.cu's wide_a_image_kernel produces for activations, here for a constant
    operand (the SMPL skinning weights of csrc/smpl_skin_umma.cu)."""
    M, K = W.shape
    mt, kb = (M + 127) // 128, (K + 31) // 32
    Wp = torch.zeros((mt * 128, kb * 32), dtype=torch.float32, device=W.device)
    Wp[:M, :K] = W.float()
    hi = Wp.to(torch.bfloat16)
    lo = (Wp - hi.float()).to(torch.bfloat16)
    both = torch.stack([hi, lo]).reshape(2, mt, 16, 8, kb, 4, 8)            # (h, tile, rg, r, block, kc, e)
    return both.permute(1, 4, 0, 2, 5, 3, 6).contiguous()                   # (tile, block, h, rg, kc, r, e)
