"""Oracle (test infrastructure): streaming transformer with a ring KV cache, functional form.

Restates ``moshi/moshi/modules/transformer.py`` (StreamingTransformer :809-929, layer :752-802,
attention :533-597, RingKVCache.complete :236-288, apply_weights_per_step :291-318, norms :45-134),
``rope.py:45-82`` and ``gating.py:13-22`` of the reference.  One implementation serves Mimi's
bottleneck transformers (fp32, LayerNorm, GELU FFN, LayerScale, 2 tokens/frame), Moshi's Temporal
transformer (bf16, RMSNorm-f32, gated SiLU, RoPE, ring of 3000) and the Depformer (per-step weights,
capacity ``dep_q``, no positional embedding).
"""
from __future__ import annotations

import math
import typing as tp
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class TransformerSpec:
    d_model: int
    num_heads: int
    num_layers: int
    context: tp.Optional[int]
    norm: str = "layer_norm"              # layer_norm | rms_norm | rms_norm_f32
    gating: str = "none"                  # none (=> linear1/gelu/linear2) | silu
    positional_embedding: str = "rope"    # rope | sin | sin_rope | none
    max_period: float = 10000.0
    layer_scale: bool = False
    weights_per_step: int = 0
    schedule: tp.Optional[tp.List[int]] = None
    positional_scale: float = 1.0
    quantize: bool = False                # every nn.Linear replaced by QLinear (transformer.py:885-888, utils/quantize.py)
    kv_quant: str = ""                    # NOT a reference option: restates the opt-in 8-bit KV rings of the CUDA path
                                          # ("fp8_e4m3" | "int8")


@dataclass
class LayerState:
    k: torch.Tensor            # [B, H, cap, D]
    v: torch.Tensor
    end_offset: torch.Tensor   # [B] (or [1] when every row advances together)
    offset: torch.Tensor       # [B] rope position per row
    offset_cpu: int = 0        # weights-per-step index


@dataclass
class TransformerState:
    layers: tp.List[LayerState]
    offsets: torch.Tensor      # [B] (sin embedding position)
    exec_mask: torch.Tensor    # [B] bool
    capacity: int = 0
    per_row: bool = True


def init_state(spec: TransformerSpec, batch: int, dtype: torch.dtype) -> TransformerState:
    """``_MHAState`` / ``RingKVCache.__init__`` (transformer.py:208-227, 448-480)."""
    if spec.context is None:
        assert spec.weights_per_step, "need a context or weights_per_step to size the cache"
        capacity = spec.weights_per_step
    else:
        capacity = spec.context
    per_row = not spec.weights_per_step
    hd = spec.d_model // spec.num_heads
    if spec.kv_quant:
        dtype = torch.float32            # the ring holds 8-bit values times an fp32 scale; kept dequantised here
    layers = []
    for _ in range(spec.num_layers):
        layers.append(LayerState(
            k=torch.zeros(batch, spec.num_heads, capacity, hd, dtype=dtype),
            v=torch.zeros(batch, spec.num_heads, capacity, hd, dtype=dtype),
            end_offset=torch.zeros(batch if per_row else 1, dtype=torch.long),
            offset=torch.zeros(batch, dtype=torch.long)))
    return TransformerState(layers, torch.zeros(batch, dtype=torch.long),
                            torch.ones(batch, dtype=torch.bool), capacity, per_row)


def reset_state(st: TransformerState, reset_mask: torch.Tensor) -> None:
    """``_MHAState.reset`` / ``_LayerState.reset`` / ``_TransformerState.reset`` / ``State.reset``."""
    st.exec_mask |= reset_mask
    st.offsets[reset_mask] = 0
    for ls in st.layers:
        ls.offset[reset_mask] = 0
        if st.per_row:
            ls.end_offset[reset_mask] = 0
        elif bool(reset_mask.any()):
            ls.end_offset.zero_()
        ls.offset_cpu = 0


# ---------------------------------------------------------------------------------------------
# elementary ops
# ---------------------------------------------------------------------------------------------

def rms_norm(x: torch.Tensor, alpha: torch.Tensor, eps: float, f32: bool) -> torch.Tensor:
    """transformer.py:45-58: var = eps + mean(x^2); y = x * (alpha * rsqrt(var)); cast back."""
    dt = x.dtype
    xf = x.float() if f32 else x
    var = eps + torch.mean(xf ** 2, dim=2, keepdim=True)
    return (xf * (alpha.to(var) * torch.rsqrt(var))).to(dt)


def apply_norm(kind: str, x: torch.Tensor, sd: dict, prefix: str) -> torch.Tensor:
    """``create_norm_fn`` dispatch, transformer.py:113-134."""
    if kind == "layer_norm":
        return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)
    if kind == "rms_norm_f32":
        return rms_norm(x, sd[prefix + ".alpha"], 1e-8, True)
    if kind == "rms_norm":
        return rms_norm(x, sd[prefix + ".alpha"], 1e-5, False)
    raise ValueError(kind)


def rope(q: torch.Tensor, k: torch.Tensor, offset: torch.Tensor, max_period: float):
    """Interleaved rotary embedding on [B, H, T, D] tensors, fp32 math (rope.py:45-82)."""
    B, H, T, D = q.shape
    half = torch.arange(D // 2, dtype=torch.float32)
    freqs = torch.exp(half * (-math.log(max_period) * 2 / D))
    ts = offset.float().view(-1, 1) + torch.arange(T, dtype=torch.float32)
    ang = freqs * ts.view(B, 1, T, 1)
    c, s = torch.cos(ang), torch.sin(ang)

    def rot(x):
        xp = x.view(B, H, T, D // 2, 2)
        a, b = xp[..., 0].float(), xp[..., 1].float()
        ra = a * c - b * s
        rb = a * s + b * c
        return torch.stack([ra.to(x.dtype), rb.to(x.dtype)], dim=-1).view(B, H, T, D)

    return rot(q), rot(k)


def sin_embedding(positions: torch.Tensor, dim: int, max_period: float, dtype) -> torch.Tensor:
    """transformer.py:137-165."""
    half = dim // 2
    positions = positions.to(dtype)
    adim = torch.arange(half, dtype=dtype).view(1, 1, -1)
    mp = torch.full([], max_period, dtype=dtype)
    phase = positions / (mp ** (adim / (half - 1)))
    return torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1)


def ring_append(ls: LayerState, k: torch.Tensor, v: torch.Tensor, exec_mask: torch.Tensor,
                capacity: int, per_row: bool) -> torch.Tensor:
    """Write T new keys/values, return the absolute position stored in every slot (-1 = empty).

    Follows ``RingKVCache.complete`` (transformer.py:236-288) step by step, including the fact
    that the write happens for every row (masked rows simply do not advance ``end_offset``).
    """
    B, H, T, D = k.shape
    slots = (torch.arange(T) + ls.end_offset.view(-1, 1)) % capacity          # [B or 1, T]
    if per_row:
        idx = slots.view(B, 1, T, 1).expand(-1, H, T, D)
        ls.k.scatter_(2, idx, k)
        ls.v.scatter_(2, idx, v)
    else:
        ls.k.index_copy_(2, slots[0], k)
        ls.v.index_copy_(2, slots[0], v)
    all_slots = torch.arange(capacity)
    last = ls.end_offset.view(-1, 1) + T - 1
    delta = all_slots - (last % capacity)
    positions = torch.where(delta <= 0, last + delta, last + delta - capacity)
    if per_row:
        ls.end_offset[:] = torch.where(exec_mask, ls.end_offset + T, ls.end_offset)
    else:
        ls.end_offset.add_(T)
    empty = all_slots >= ls.end_offset.view(-1, 1)
    return torch.where(empty, torch.full_like(positions, -1), positions)


def fp8_roundtrip(t: torch.Tensor) -> torch.Tensor:
    """What the opt-in fp8 ring stores for one key / value row of ``D`` entries, dequantised to fp32:
    ``e4m3(t * (448 / absmax)) * (absmax / 448)``, round to nearest even (csrc/lm_kernels.cuh, attn_step_f8_kernel)."""
    t = t.float()
    amax = t.abs().amax(dim=-1, keepdim=True)
    inv = torch.where(amax > 0, torch.full_like(amax, 448.0) / amax, torch.zeros_like(amax))   # (scalar / tensor is reciprocal * scalar in torch)
    q = (t * inv).to(torch.float8_e4m3fn).float()
    return q * (amax * (1.0 / 448.0))


def int8_roundtrip(t: torch.Tensor) -> torch.Tensor:
    """The opt-in int8 ring: ``clamp(round(t * (127 / absmax)), -127, 127) * (absmax / 127)`` per row of ``D`` entries, ties to
    even (csrc/lm_kernels.cuh, attn_step_i8_kernel)."""
    t = t.float()
    amax = t.abs().amax(dim=-1, keepdim=True)
    inv = torch.where(amax > 0, torch.full_like(amax, 127.0) / amax, torch.zeros_like(amax))
    q = torch.round(t * inv).clamp_(-127, 127)
    return q * (amax * (1.0 / 127.0))


def kv_roundtrip(t: torch.Tensor, fmt: str) -> torch.Tensor:
    return {"fp8_e4m3": fp8_roundtrip, "int8": int8_roundtrip}[fmt](t)


def linear(x: torch.Tensor, w: torch.Tensor, quantize: bool = False) -> torch.Tensor:
    """``nn.Linear`` (bias-free on this path) or, for a quantised LM, ``QLinear.forward`` (oracle/quant.py)."""
    if quantize:
        from .quant import qlinear
        return qlinear(x, w)
    return F.linear(x, w)


def _per_step_linear(sd: dict, fmt: str, spec: TransformerSpec, x: torch.Tensor, offset_cpu: int):
    """``apply_weights_per_step`` (transformer.py:291-318) for F.linear weights named by ``fmt``."""
    if not spec.weights_per_step:
        return linear(x, sd[fmt.format(i=0)], spec.quantize)
    outs = []
    for t in range(x.shape[1]):
        i = t + offset_cpu
        if spec.schedule is not None:
            i = spec.schedule[i]
        outs.append(linear(x[:, t:t + 1], sd[fmt.format(i=i)], spec.quantize))
    return torch.cat(outs, 1)


def attention(sd: dict, p: str, spec: TransformerSpec, x: torch.Tensor, ls: LayerState,
              st: TransformerState) -> torch.Tensor:
    """``StreamingMultiheadAttention.forward`` self-attention branch (transformer.py:533-597)."""
    B, T, C = x.shape
    H = spec.num_heads
    proj = _per_step_linear(sd, p + ".in_projs.{i}.weight", spec, x, ls.offset_cpu)
    qkv = proj.view(B, T, 3, H, C // H).permute(2, 0, 3, 1, 4)               # p b h t d
    q, k, v = qkv[0], qkv[1], qkv[2]
    if spec.positional_embedding in ("rope", "sin_rope"):
        q, k = rope(q, k, ls.offset, spec.max_period)
    if spec.kv_quant:
        k, v = kv_roundtrip(k, spec.kv_quant), kv_roundtrip(v, spec.kv_quant)
    pos_k = ring_append(ls, k.contiguous(), v.contiguous(), st.exec_mask, st.capacity, st.per_row)
    pos_k = pos_k[:, None]                                                     # [B|1, 1, cap]
    pos_q = ls.offset.view(-1, 1, 1) + torch.arange(T).view(-1, 1)            # [B, T, 1]
    delta = pos_q - pos_k
    allowed = (pos_k >= 0) & (delta >= 0)
    if spec.context is not None:
        allowed = allowed & (delta < spec.context)
    if spec.kv_quant:
        out = F.scaled_dot_product_attention(q.float(), ls.k, ls.v, allowed[:, None], dropout_p=0.0).to(x.dtype)
    else:
        out = F.scaled_dot_product_attention(q, ls.k, ls.v, allowed[:, None], dropout_p=0.0)
    out = out.transpose(1, 2).reshape(B, T, C)
    out = _per_step_linear(sd, p + ".out_projs.{i}.weight", spec, out, ls.offset_cpu)
    ls.offset[:] = torch.where(st.exec_mask, ls.offset + T, ls.offset)
    return out


def gated_ffn(w_in: torch.Tensor, w_out: torch.Tensor, x: torch.Tensor, quantize: bool = False) -> torch.Tensor:
    """gating.py:13-22 with SiLU: W_out (silu(h[:H]) * h[H:]), h = W_in x."""
    h = linear(x, w_in, quantize)
    B, T, _ = h.shape
    h = h.view(B, T, 2, -1)
    return linear(F.silu(h[..., 0, :]) * h[..., 1, :], w_out, quantize)


def feed_forward(sd: dict, p: str, spec: TransformerSpec, x: torch.Tensor, offset_cpu: int):
    """``_ff_block`` body without the residual (transformer.py:752-768)."""
    if spec.gating == "none":
        return F.linear(F.gelu(F.linear(x, sd[p + ".linear1.weight"])), sd[p + ".linear2.weight"])
    if not spec.weights_per_step:
        return gated_ffn(sd[p + ".gating.linear_in.weight"], sd[p + ".gating.linear_out.weight"], x, spec.quantize)
    outs = []
    for t in range(x.shape[1]):
        i = t + offset_cpu
        if spec.schedule is not None:
            i = spec.schedule[i]
        outs.append(gated_ffn(sd[f"{p}.gating.{i}.linear_in.weight"],
                              sd[f"{p}.gating.{i}.linear_out.weight"], x[:, t:t + 1], spec.quantize))
    return torch.cat(outs, 1)


def forward(sd: dict, prefix: str, spec: TransformerSpec, x: torch.Tensor,
            st: TransformerState) -> torch.Tensor:
    """``StreamingTransformer.forward`` (transformer.py:894-929) on [B, T, C]."""
    B, T, C = x.shape
    if spec.positional_embedding in ("sin", "sin_rope"):
        positions = torch.arange(T).view(1, -1, 1) + st.offsets.view(-1, 1, 1)
        x = x + spec.positional_scale * sin_embedding(positions, C, spec.max_period, x.dtype)
    for li, ls in enumerate(st.layers):
        p = f"{prefix}.layers.{li}"
        upd = attention(sd, p + ".self_attn", spec, apply_norm(spec.norm, x, sd, p + ".norm1"), ls, st)
        if spec.layer_scale:
            upd = sd[p + ".layer_scale_1.scale"] * upd
        x = x.to(upd) + upd
        upd = feed_forward(sd, p, spec, apply_norm(spec.norm, x, sd, p + ".norm2"), ls.offset_cpu)
        if spec.layer_scale:
            upd = sd[p + ".layer_scale_2.scale"] * upd
        x = x.to(upd) + upd
        ls.offset_cpu += T
    st.offsets[:] = torch.where(st.exec_mask, st.offsets + T, st.offsets)
    return x
