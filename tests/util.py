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


def greedy_unexcused(got: torch.Tensor, want: torch.Tensor, logits_oracle: torch.Tensor, tol: float) -> tuple[int, int]:
    """Margin-aware comparison of greedy ids ([N] each) against the oracle's fp32-viewed logits ([N, card]): a mismatch is
    excused iff the oracle's own top-2 logit gap is below ``2 * tol`` (two logits that each moved by <= tol can swap).
    Returns (n_mismatch, n_unexcused)."""
    top2 = logits_oracle.float().topk(2, dim=-1).values
    gap = top2[:, 0] - top2[:, 1]
    bad = got.cpu() != want.cpu()
    return int(bad.sum()), int((bad & (gap > 2 * tol)).sum())


def sample_is_stable(logits: torch.Tensor, temp: float, top_k: int, noise: torch.Tensor, tol: float) -> torch.Tensor:
    """For every row of oracle logits ([N, card], fp32 view of the bf16 tensor) with its Exp(1) noise ([N, k]): is the
    sampled token (sampling.py:86-106: softmax(l / temp) -> top-k -> argmax(p / q), noise indexed by RANK) the same for
    every logit vector within ``tol`` (max-norm) of this one?  Conservative test:
      * a candidate's log-probability moves by at most 2 * tol / temp (its logit and the normaliser);
      * candidates whose logits lie in one run of the sorted order with adjacent gaps < 2 * tol can exchange ranks, i.e. any
        of them can receive any noise value of that run (this also covers exact ties, whose torch.topk order is unspecified,
        and candidates that can cross the top-k boundary).
    Stable iff the winner's worst-case score exceeds every other candidate's best-case score.  Returns bool [N]."""
    logits = logits.float().cpu()
    noise = noise.float().cpu()
    N, card = logits.shape
    k = min(top_k, card)
    out = torch.zeros(N, dtype=torch.bool)
    for n in range(N):
        srt, _ = torch.sort(logits[n], descending=True, stable=True)
        logp = torch.log_softmax(srt / temp, dim=-1)
        q = torch.cat([noise[n, :k], torch.full((card - k,), float("inf"))])     # ranks beyond k never win
        score = logp[:k] - torch.log(noise[n, :k])
        w = int(score.argmax())
        # runs of the sorted order whose adjacent gaps are all < 2 * tol
        brk = torch.cat([torch.tensor([True]), (srt[:-1] - srt[1:]) >= 2 * tol])
        run_id = torch.cumsum(brk.int(), 0) - 1
        n_runs = int(run_id[-1]) + 1
        qmin = torch.full((n_runs,), float("inf")).scatter_reduce(0, run_id, q, reduce="amin")
        qmax = torch.full((n_runs,), 0.0).scatter_reduce(0, run_id, q, reduce="amax")
        slack = 2 * tol / temp
        best_case = logp + slack - torch.log(qmin[run_id])          # every candidate, with the smallest noise it could be dealt
        worst_w = logp[w] - slack - torch.log(qmax[run_id[w]])      # the winner, with the largest noise it could be dealt
        best_case[w] = -float("inf")
        if int((run_id == run_id[w]).sum()) > 1:                    # the winner itself can be re-ranked: not stable
            continue
        out[n] = bool(worst_w > best_case.max())
    return out


def to_candle_layout(cfg, sd: dict) -> dict:
    """The Rust / candle checkpoint layout of the LM (``scripts/import_rust.py:45-113``, what ``rust/moshi-core`` loads): the
    temporal transformer under the reference's names, the depformer per codebook step as ``depformer.<k>.*``."""
    out = {k: v for k, v in sd.items() if k.startswith(("text_emb", "text_linear", "out_norm", "emb.", "transformer.", "extra_heads."))}
    for k in range(cfg.dep_q):
        base = f"depformer.{k}."
        out[base + "linear_in.weight"] = sd[f"depformer_in.{k}.weight"]
        out[base + "linear_out.weight"] = sd[f"linears.{k}.weight"]
        out[base + "emb.weight"] = sd["depformer_text_emb.weight"] if k == 0 else sd[f"depformer_emb.{k - 1}.weight"]
        for layer in range(cfg.depformer_num_layers):
            src, dst = f"depformer.layers.{layer}.", base + f"transformer.layers.{layer}."
            out[dst + "self_attn.in_proj_weight"] = sd[src + f"self_attn.in_projs.{k}.weight"]
            out[dst + "self_attn.out_proj.weight"] = sd[src + f"self_attn.out_projs.{k}.weight"]
            out[dst + "norm1.alpha"] = sd[src + "norm1.alpha"]
            out[dst + "norm2.alpha"] = sd[src + "norm2.alpha"]
            out[dst + "gating.linear_in.weight"] = sd[src + f"gating.{k}.linear_in.weight"]
            out[dst + "gating.linear_out.weight"] = sd[src + f"gating.{k}.linear_out.weight"]
    return out
