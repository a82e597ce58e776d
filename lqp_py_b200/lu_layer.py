"""Cached-factor LU solve as an autograd op -- drop-in for ``lqp_py/lu_layer.py``.

``TorchLU(A=..., LU=None, P=None)`` factorises once (reference :7-12) and
``TorchLU.forward(A, b)`` / ``TorchLULayer.apply(A, b, LU, P)`` solve ``A x = b`` from the cached
factors (:14-38).  The backward re-uses the same factors (:41-58):
``dx = A^-1 (-dl_dx)``, ``dl_dA = dx x^T``, ``dl_db = -dx`` (valid for symmetric ``A``, as in the
reference).  Factorisation, solves and the outer product are the sm_100a kernels of
``csrc/lu.cu`` behind ``lqpb_lu_factor / lqpb_lu_solve / lqpb_outer`` (include/lqpb.h).
``LU`` is LAPACK-packed and ``P`` holds 1-based int32 pivots, like ``torch.linalg.lu_factor``.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _abi


def _cuda(t):
    if t.is_cuda:
        return t.detach().contiguous()
    if not torch.cuda.is_available():
        raise RuntimeError("lqp_py_b200 runs on a CUDA device (B200, sm_100a) only; there is no CPU fallback")
    return t.detach().to(torch.device("cuda", torch.cuda.current_device())).contiguous()


def lu_factor(A):
    """Batched partial-pivoting LU: returns ``(LU, pivots)`` on the CUDA device."""
    L = _abi.lib()
    Ad = _cuda(A)
    squeeze = Ad.dim() == 2
    if squeeze:
        Ad = Ad.unsqueeze(0)
    B, N = Ad.shape[0], Ad.shape[1]
    LU = torch.empty_like(Ad)
    piv = torch.empty((B, N), dtype=torch.int32, device=Ad.device)
    with torch.cuda.device(Ad.device):
        stream = torch.cuda.current_stream(Ad.device).cuda_stream
        rc = getattr(L, f"lqpb_lu_factor_{_abi.suffix(Ad.dtype)}")(B, N, _abi.ptr(Ad), _abi.ptr(LU), _abi.ptr(piv),
                                                                   C.c_void_p(stream))
    _abi.check(rc, "lqpb_lu_factor")
    return (LU[0], piv[0]) if squeeze else (LU, piv)


def lu_solve(LU, P, b, negate=False):
    """Solve with cached factors; ``b`` is ``(B, N, nrhs)`` (or ``(N, nrhs)`` for a single matrix)."""
    L = _abi.lib()
    LUd, Pd, bd = _cuda(LU), _cuda(P), _cuda(b)
    squeeze = LUd.dim() == 2
    if squeeze:
        LUd, Pd, bd = LUd.unsqueeze(0), Pd.unsqueeze(0), bd.unsqueeze(0)
    B, N, nrhs = bd.shape[0], bd.shape[1], bd.shape[2]
    x = torch.empty_like(bd)
    with torch.cuda.device(bd.device):
        stream = torch.cuda.current_stream(bd.device).cuda_stream
        rc = getattr(L, f"lqpb_lu_solve_{_abi.suffix(bd.dtype)}")(B, N, nrhs, _abi.ptr(LUd), _abi.ptr(Pd), _abi.ptr(bd),
                                                                  _abi.ptr(x), 1 if negate else 0, C.c_void_p(stream))
    _abi.check(rc, "lqpb_lu_solve")
    return x[0] if squeeze else x


def _outer(a, b):
    """(B,N,1) x (B,M,1) -> (B,N,M) = a b^T"""
    L = _abi.lib()
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    out = torch.empty((B, N, M), dtype=a.dtype, device=a.device)
    with torch.cuda.device(a.device):
        stream = torch.cuda.current_stream(a.device).cuda_stream
        rc = getattr(L, f"lqpb_outer_{_abi.suffix(a.dtype)}")(B, N, M, _abi.ptr(a.contiguous()), _abi.ptr(b.contiguous()),
                                                              _abi.ptr(out), C.c_void_p(stream))
    _abi.check(rc, "lqpb_outer")
    return out


class TorchLU(nn.Module):
    def __init__(self, A=None, LU=None, P=None):
        super().__init__()
        if LU is None or P is None:
            LU, P = lu_factor(A)
        self.LU = LU
        self.P = P

    def forward(self, A, b):
        return TorchLULayer.apply(A, b, self.LU, self.P)


class TorchLULayer(torch.autograd.Function):
    """Forward solve / backward solve with the same cached factors (reference :18-58)."""

    @staticmethod
    def forward(ctx, A, b, LU=None, P=None):
        if LU is None or P is None:
            LU, P = lu_factor(A)
        x = lu_solve(LU, P, b)
        ctx.save_for_backward(_cuda(LU), _cuda(P), x)
        ctx.b_device = b.device
        ctx.A_device = A.device
        return x.to(b.device)

    @staticmethod
    def backward(ctx, dl_dx):
        LU, P, x = ctx.saved_tensors
        dx = lu_solve(LU, P, _cuda(dl_dx), negate=True)          # :52
        dl_dA = dl_db = None
        if ctx.needs_input_grad[0]:
            if x.dim() == 3 and x.shape[2] == 1:
                dl_dA = _outer(dx, x).to(ctx.A_device)           # :53
            else:
                dl_dA = torch.matmul(dx, x.transpose(-1, -2)).to(ctx.A_device)
        if ctx.needs_input_grad[1]:
            dl_db = (-dx).to(ctx.b_device)                       # :54
        return dl_dA, dl_db, None, None
