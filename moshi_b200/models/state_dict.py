"""Checkpoint key normalisation: the on-disk names the reference accepts map to the module names.

* packed attention weights ``self_attn.in_proj_weight`` / ``in_proj.weight`` / ``out_proj.weight``
  of shape ``[mult * rows, in]`` become ``in_projs.{i}.weight`` / ``out_projs.{i}.weight``
  (load hook ``moshi/moshi/modules/transformer.py:422-446``);
* legacy codebook buffers ``inited`` / ``cluster_size`` / ``embed_avg`` / ``embed_sum``
  become ``_initialized`` / ``cluster_usage`` / ``embedding_sum`` (``core_vq.py:162-176``).
"""
from __future__ import annotations

import typing as tp

import torch

_CODEBOOK_RENAMES = {"inited": "_initialized", "cluster_size": "cluster_usage",
                     "embed_avg": "embedding_sum", "embed_sum": "embedding_sum"}


def _split_attention(sd: tp.Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    out: dict[str, torch.Tensor] = {}
    for key, w in sd.items():
        if key.endswith("self_attn.in_proj_weight") or key.endswith("self_attn.in_proj.weight"):
            base = key[: key.rindex("in_proj")]
            mult = max(1, w.shape[0] // (3 * w.shape[1]))
            for i, part in enumerate(w.view(mult, -1, w.shape[1])):
                out[f"{base}in_projs.{i}.weight"] = part
        elif key.endswith("self_attn.out_proj.weight"):
            base = key[: key.rindex("out_proj")]
            mult = max(1, w.shape[0] // w.shape[1])
            for i, part in enumerate(w.view(mult, -1, w.shape[1])):
                out[f"{base}out_projs.{i}.weight"] = part
        else:
            out[key] = w
    return out


def normalize_mimi_state_dict(sd: tp.Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    out = {}
    for key, v in _split_attention(sd).items():
        head, _, leaf = key.rpartition(".")
        if head.endswith("_codebook") and leaf in _CODEBOOK_RENAMES:
            key = f"{head}.{_CODEBOOK_RENAMES[leaf]}"
        out[key] = v
    return out


def normalize_lm_state_dict(sd: tp.Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    return _split_attention(sd)
