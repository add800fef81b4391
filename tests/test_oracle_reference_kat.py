"""Pins the LM oracle against the reference's own known-answer vectors.

``tests/golden/reference_test_lm/*.safetensors`` are the unmodified assets of the reference's
``moshi/tests/test_lm.py`` (model weights, codes, expected logits/masks of the teacher-forced
``LMModel.forward``).  The oracle has no teacher-forced path: it is driven step by step through its
*streaming* ``forward_text`` / ``forward_depformer`` with the delayed codes, which is exactly what
``LMGen`` does at inference, and must reproduce the stored logits.
"""
import torch
from safetensors.torch import load_file

from oracle import transformer as tr
from oracle.lm import LMOracle, LMSpec


def _spec() -> LMSpec:
    # moshi/tests/test_lm.py:13-36
    return LMSpec(dim=16, text_card=48, n_q=3, dep_q=3, card=32, num_heads=1, num_layers=2,
                  hidden_scale=1, context=4, delays=[0, 1, 2, 4], norm="layer_norm", gating="none",
                  positional_embedding="sin", depformer_dim=16, depformer_num_heads=1,
                  depformer_num_layers=2, depformer_gating="silu", depformer_pos_emb="sin",
                  depformer_multi_linear=True, depformer_weights_per_step=True,
                  depformer_schedule=[0, 1, 1], depformer_low_rank=8)


def _split_packed(sd):
    """The reference's load hook (transformer.py:422-446): packed attention weights of shape
    [mult * rows, in] become ``in_projs.{i}.weight`` / ``out_projs.{i}.weight``."""
    out = {}
    for key, w in sd.items():
        if key.endswith("self_attn.in_proj_weight"):
            mult = w.shape[0] // (3 * w.shape[1])
            for i, part in enumerate(w.view(mult, -1, w.shape[1])):
                out[key.replace("in_proj_weight", f"in_projs.{i}.weight")] = part
        elif key.endswith("self_attn.out_proj.weight"):
            mult = w.shape[0] // w.shape[1]
            for i, part in enumerate(w.view(mult, -1, w.shape[1])):
                out[key.replace("out_proj.weight", f"out_projs.{i}.weight")] = part
        else:
            out[key] = w
    return out


def _delay(delays, codes, initial):
    # lm_utils.py:10-21
    rows = []
    for k, d in enumerate(delays):
        line = codes[:, k].roll(d, dims=1)
        if d > 0:
            line[:, :d] = initial[:, k]
        rows.append(line)
    return torch.stack(rows, 1)


def _undelay(delays, x):
    # lm_utils.py:24-40
    B, K, T = x.shape[:3]
    mask = torch.ones(B, K, T, dtype=torch.bool)
    rows = []
    for k, d in enumerate(delays):
        line = x[:, k].roll(-d, dims=1)
        if d > 0:
            line[:, -d:] = float("nan")
            mask[:, k, -d:] = False
        rows.append(line)
    return torch.stack(rows, 1), mask


def _ce(logits, targets, mask):
    # utils/utils.py cross_entropy semantics: mean over valid positions per codebook
    lp = torch.log_softmax(torch.nan_to_num(logits.float()), dim=-1)
    nll = -lp.gather(-1, targets.clamp(min=0)[..., None])[..., 0]
    return (nll * mask).sum(dim=(0, 2)) / mask.sum(dim=(0, 2))


@torch.no_grad()
def test_streaming_oracle_reproduces_reference_lm_kat(golden_dir):
    d = golden_dir / "reference_test_lm"
    sd = _split_packed(load_file(d / "test_lm_model.safetensors"))
    codes = load_file(d / "test_lm_codes.safetensors")["codes"]
    ref = load_file(d / "test_lm_out.safetensors")
    spec = _spec()
    orc = LMOracle(sd, spec)
    B, K, T = codes.shape
    assert K == spec.num_codebooks
    orc.streaming(B)
    initial = orc.initial.expand(B, -1, -1)
    delayed = torch.cat([initial, _delay(spec.delays, codes.clone(), initial)], dim=2)
    text_logits, dep_logits = [], []
    for t in range(T):
        out, tl = orc.forward_text(delayed[:, :, t:t + 1])
        text_logits.append(tl)
        st = tr.init_state(orc.dep_spec, B, orc.dtype)
        per_k = []
        for k in range(spec.dep_q):
            per_k.append(orc.forward_depformer(k, delayed[:, k, t + 1:t + 2], out, st))
        dep_logits.append(torch.cat(per_k, dim=1))
    text_logits = torch.cat(text_logits, dim=2)          # [B, 1, T, text_card]
    dep_logits = torch.cat(dep_logits, dim=2)            # [B, dep_q, T, card]

    logits, mask = _undelay(spec.delays[1:1 + spec.dep_q], dep_logits)
    mask &= codes[:, 1:1 + spec.dep_q] != -1
    tlogits, tmask = _undelay(spec.delays[:1], text_logits)
    tmask &= codes[:, :1] != -1
    assert (mask == ref["mask"]).all()
    assert (tmask == ref["text_mask"]).all()
    # direct comparison of every valid logit, then the reference test's own CE criterion
    torch.testing.assert_close(logits[mask], ref["logits"][mask], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(tlogits[tmask], ref["text_logits"][tmask], rtol=1e-4, atol=1e-5)
    ce, ce_ref = _ce(logits, codes[:, 1:], mask), _ce(ref["logits"], codes[:, 1:], mask)
    assert ((ce - ce_ref).abs() / ce_ref).amax() <= 1e-5
    ce, ce_ref = _ce(tlogits, codes[:, :1], tmask), _ce(ref["text_logits"], codes[:, :1], tmask)
    assert ((ce - ce_ref).abs() / ce_ref).amax() <= 1e-5
