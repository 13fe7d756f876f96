"""Thin diagnostics over single kernels of the hot path (used by tests; not a user-facing op library)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_ACT = {None: 0, 'none': 0, 'relu': 1, 'gelu': 2}


def linear(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None, act: str | None = None, precision: str = 'fp32'):
    """Y = act(A @ W.T + bias) through libsherf_b200.so (`sherf_debug_linear`); A[M,K], W[N,K] fp32 CUDA tensors."""
    from .renderer import PRECISIONS
    lib = _lib.load()
    assert A.is_cuda and W.is_cuda and A.dtype == torch.float32 and W.dtype == torch.float32
    A, W = A.contiguous(), W.contiguous()
    M, K = A.shape
    N = W.shape[0]
    assert K % 4 == 0 or True
    lda = (K + 3) // 4 * 4
    if lda != K:
        Ap = A.new_zeros(M, lda)
        Ap[:, :K] = A
        A = Ap
    Y = torch.empty(M, N, device=A.device, dtype=torch.float32)
    scratch = torch.empty(8 * 272 * 272 * 4 + 1024, dtype=torch.uint8, device=A.device)
    b = bias.contiguous() if bias is not None else None
    rc = lib.sherf_debug_linear(PRECISIONS[precision], A.data_ptr(), lda, W.data_ptr(), b.data_ptr() if b is not None else None,
                                Y.data_ptr(), N, M, N, K, _ACT[act], scratch.data_ptr(), scratch.numel(),
                                torch.cuda.current_stream(A.device).cuda_stream)
    _lib.check(rc)
    return Y
