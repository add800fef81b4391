"""Shared helpers for the GPU parity tests."""
from __future__ import annotations

import ctypes as C

import torch


def stats(name: str, got: torch.Tensor, want: torch.Tensor) -> str:
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    diff = (got - want).abs()
    scale = want.abs().max().item() + 1e-30
    return (f"{name}: max_abs={diff.max().item():.3e} rel_to_max={diff.max().item() / scale:.3e} "
            f"ref_std={want.std().item():.3e} nan={int(torch.isnan(got).sum())}")


def rvq_mismatches(codes_got: torch.Tensor, codes_want: torch.Tensor, margins: torch.Tensor,
                   n_semantic: int = 1, tol: float = 1e-4):
    """Margin-aware comparison of RVQ indices ([B, K, T] each).

    A mismatch at level k is excused iff the oracle's own relative gap between the best and the
    second-best squared distance at that level is below ``tol`` (a rounding-order tie) or an earlier
    level of the same quantizer already flipped (the residual, hence every later index, differs).
    Returns (n_mismatch, n_unexcused).
    """
    got, want = codes_got.cpu(), codes_want.cpu()
    B, K, T = want.shape
    bad = unexcused = 0
    for b in range(B):
        for t in range(T):
            for lo, hi in ((0, n_semantic), (n_semantic, K)):
                flipped = False
                for k in range(lo, hi):
                    if got[b, k, t] != want[b, k, t]:
                        bad += 1
                        if not flipped and margins[b, k, t] > tol:
                            unexcused += 1
                        flipped = True
    return bad, unexcused


def cptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
