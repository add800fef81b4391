"""Oracle (test infrastructure): Moshi LM streaming decode step on the CPU.

Restates, with explicit state, the inference path of the reference:

* ``ScaledEmbedding.forward``      – ``models/lm_utils.py:102-124`` (token -1 -> zero vector, low-rank)
* ``LMModel.forward_text``         – ``models/lm.py:379-408``
* ``LMModel.forward_depformer``    – ``models/lm.py:450-493``
* ``LMGen._step`` / ``depformer_step`` – ``models/lm.py:668-783, 809-850``
* ``sample_token`` / top-k / multinomial-by-exponential – ``utils/sampling.py:44-106``

The spec is wider than the 7B family (LayerNorm, sinusoidal positions, plain GELU FFN, low-rank
depformer embeddings, weight-sharing schedule) so that the oracle can be pinned against the
reference's own golden LM vectors (``moshi/tests/test_lm.py`` + ``tests/assets``), which use those.
"""
from __future__ import annotations

import typing as tp
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import transformer as tr


@dataclass
class LMSpec:
    dim: int
    text_card: int
    n_q: int
    dep_q: int
    card: int
    num_heads: int
    num_layers: int
    hidden_scale: float
    context: tp.Optional[int]
    delays: tp.List[int]
    norm: str = "layer_norm"
    gating: str = "none"
    positional_embedding: str = "sin"
    max_period: float = 10000.0
    depformer_dim: int = 256
    depformer_num_heads: int = 8
    depformer_num_layers: int = 6
    depformer_gating: str = "none"
    depformer_pos_emb: str = "sin"
    depformer_max_period: float = 10000.0
    depformer_multi_linear: bool = False
    depformer_weights_per_step: bool = False
    depformer_schedule: tp.Optional[tp.List[int]] = None
    depformer_low_rank: tp.Optional[int] = None
    existing_text_padding_id: int = 3
    extra_heads_num_heads: int = 0        # lm.py:224-226: nn.Linear(dim, extra_heads_dim) on transformer_out (the STT models)
    extra_heads_dim: int = 6

    @staticmethod
    def from_config(cfg) -> "LMSpec":
        """From ``moshi_b200.config.LMConfig`` (the 7B family)."""
        return LMSpec(
            dim=cfg.dim, text_card=cfg.text_card, n_q=cfg.n_q, dep_q=cfg.dep_q, card=cfg.card,
            num_heads=cfg.num_heads, num_layers=cfg.num_layers, hidden_scale=cfg.hidden_scale,
            context=cfg.context, delays=list(cfg.delays), norm=cfg.norm, gating=cfg.gating,
            positional_embedding=cfg.positional_embedding, max_period=cfg.max_period,
            depformer_dim=cfg.depformer_dim, depformer_num_heads=cfg.depformer_num_heads,
            depformer_num_layers=cfg.depformer_num_layers, depformer_gating=cfg.depformer_gating,
            depformer_pos_emb=cfg.depformer_pos_emb, depformer_max_period=cfg.depformer_max_period,
            depformer_multi_linear=cfg.depformer_multi_linear,
            depformer_weights_per_step=cfg.depformer_weights_per_step,
            existing_text_padding_id=cfg.existing_text_padding_id)

    @property
    def num_codebooks(self) -> int:
        return self.n_q + 1

    @property
    def max_delay(self) -> int:
        return max(self.delays)


def scaled_embedding(sd: dict, prefix: str, tokens: torch.Tensor) -> torch.Tensor:
    """lm_utils.py:102-124 without the demux / norm variants (unused on this path)."""
    is_zero = tokens == -1
    y = F.embedding(tokens.clamp(min=0), sd[prefix + ".weight"])
    y = torch.where(is_zero[..., None], torch.zeros(1, dtype=y.dtype), y)
    low = sd.get(prefix + ".low_rank.weight")
    if low is not None:
        y = F.linear(y, low)
    return y


def sample_token(logits: torch.Tensor, use_sampling: bool, temp: float, top_k: int,
                 noise: torch.Tensor | None = None, tie_break: str = "torch") -> torch.Tensor:
    """sampling.py:86-106 (top-k branch).  ``noise`` replaces the Exp(1) draw of sampling.py:44
    ([N, k], one row per flattened leading index); when None it is drawn from torch's global
    generator exactly like the reference, so that a seeded reference run is reproduced.

    ``tie_break``: the reference ranks candidates with ``torch.topk`` (sampling.py:62), whose order
    among *equal* probabilities is unspecified (and differs between torch's CPU and CUDA kernels);
    the noise is indexed by rank, so tied bf16 logits make the sample depend on that order.
    "torch" keeps torch.topk (bit-exact with a CPU reference run); "index" ranks ties by ascending
    token id, which is the order the CUDA sampler defines (tolerances)."""
    if not (use_sampling and temp > 0.0):
        return torch.argmax(logits, dim=-1)
    probs = torch.softmax(logits / temp, dim=-1)
    k = min(top_k, probs.shape[-1])
    if tie_break == "index":
        idx = torch.sort(logits, dim=-1, descending=True, stable=True).indices[..., :k]
        top = probs.gather(-1, idx)
    else:
        top, idx = torch.topk(probs, k, dim=-1)
    flat = top.reshape(-1, k)
    q = torch.empty_like(flat).exponential_(1) if noise is None else noise.reshape(-1, k).to(flat)
    choice = (flat / q).argmax(dim=-1, keepdim=True).reshape(*top.shape[:-1], 1)
    return idx.gather(-1, choice)[..., 0]


class LMOracle:
    def __init__(self, sd: tp.Dict[str, torch.Tensor], spec: LMSpec, use_sampling: bool = True,
                 temp: float = 0.8, temp_text: float = 0.7, top_k: int = 250, top_k_text: int = 25,
                 tie_break: str = "torch", quantize: bool = False, kv_quant: str = "", cfg_coef: float = 1.0,
                 cfg_is_no_text: bool = False, cfg_is_masked_until: tp.Sequence[int] | None = None,
                 condition_sum: torch.Tensor | None = None):
        # classifier-free guidance without a conditioner (lm.py:596-604, 646-662, 714-732, 820-833): the model runs on 2B
        # rows, the second half with the text stream zeroed (cfg_is_no_text) or every stream zeroed until a per-row step
        # (cfg_is_masked_until); logits are logits_null + (logits - logits_null) * cfg_coef
        self.cfg_coef, self.cfg_is_no_text = cfg_coef, cfg_is_no_text
        self.cfg_is_masked_until = None if cfg_is_masked_until is None else torch.tensor(list(cfg_is_masked_until), dtype=torch.long)
        # fuser.get_sum(condition_tensors) cast to the model dtype (lm.py:616-626): [B (2B with CFG), 1, dim], added to the
        # summed input embeddings every step (lm.py:398-399).  Evaluating the conditioners themselves happens once per
        # session outside the step and is not restated here.
        self.condition_sum = condition_sum
        if cfg_coef != 1.0:
            assert cfg_is_no_text or cfg_is_masked_until is not None or condition_sum is not None, \
                "CFG needs condition tensors or one of the two masks"
        self.tie_break = tie_break
        self.quantize = quantize          # LMModel(quantize=True): every nn.Linear is a QLinear (lm.py:242-243)
        self.sd = sd
        self.spec = spec
        self.use_sampling, self.temp, self.temp_text = use_sampling, temp, temp_text
        self.top_k, self.top_k_text = top_k, top_k_text
        self.dtype = sd["text_emb.weight"].dtype
        s = spec
        self.main_spec = tr.TransformerSpec(
            d_model=s.dim, num_heads=s.num_heads, num_layers=s.num_layers, context=s.context,
            norm=s.norm, gating=s.gating, positional_embedding=s.positional_embedding,
            max_period=s.max_period, quantize=quantize, kv_quant=kv_quant)
        self.dep_spec = tr.TransformerSpec(
            d_model=s.depformer_dim, num_heads=s.depformer_num_heads,
            num_layers=s.depformer_num_layers, context=None, norm=s.norm, gating=s.depformer_gating,
            positional_embedding=s.depformer_pos_emb, max_period=s.depformer_max_period,
            weights_per_step=s.dep_q if s.depformer_weights_per_step else 0,
            schedule=s.depformer_schedule, quantize=quantize)
        self.batch: int | None = None

    # ------------------------------------------------------------------ state (lm.py:604-666)
    def streaming(self, batch: int) -> None:
        s = self.spec
        self.batch = batch
        self.cache = torch.full((batch, s.num_codebooks, s.max_delay + 2), -2, dtype=torch.long)
        self.offsets = torch.zeros(batch, dtype=torch.long)
        self.offset_cpu = 0
        self.exec_mask = torch.ones(batch, dtype=torch.bool)
        self.model_batch = batch * 2 if self.cfg_coef != 1.0 else batch             # lm.py:646-647
        self.main_state = tr.init_state(self.main_spec, self.model_batch, self.dtype)
        self.initial = torch.cat([torch.full((1, 1, 1), s.text_card, dtype=torch.long),
                                  torch.full((1, s.n_q, 1), s.card, dtype=torch.long)], dim=1)
        self.delays = torch.tensor(s.delays, dtype=torch.long)

    def _model_rows(self, mask: torch.Tensor) -> torch.Tensor:
        return mask.repeat(2) if self.cfg_coef != 1.0 else mask                     # lm.py:653-661

    def set_exec_mask(self, mask: torch.Tensor) -> None:
        self.exec_mask[:] = mask
        self.main_state.exec_mask[:] = self._model_rows(mask)

    def reset_streaming(self, reset_mask: torch.Tensor | None = None) -> None:
        """``_LMGenState.reset`` (lm.py:537-542) + the LM's own streaming reset."""
        if reset_mask is None:
            reset_mask = torch.ones(self.batch, dtype=torch.bool)
        self.exec_mask |= reset_mask
        self.offsets[reset_mask] = 0
        self.offset_cpu = 0
        tr.reset_state(self.main_state, self._model_rows(reset_mask))

    # ------------------------------------------------------------------ model
    def forward_text(self, tokens: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """tokens [B, 1 + n_q, S] -> (transformer_out [B, S, dim], text_logits [B, 1, S, text_card])."""
        s = self.spec
        x = None
        for k in range(s.n_q):
            e = scaled_embedding(self.sd, f"emb.{k}", tokens[:, k + 1])
            x = e if x is None else x + e
        t = scaled_embedding(self.sd, "text_emb", tokens[:, 0])
        x = t if x is None else x + t
        if self.condition_sum is not None:
            x = x + self.condition_sum.to(x)
        out = tr.forward(self.sd, "transformer", self.main_spec, x, self.main_state)
        out = tr.apply_norm(s.norm, out, self.sd, "out_norm")
        if self.quantize:
            logits = tr.linear(out, self.sd["text_linear.weight"], True)
        else:
            logits = F.linear(out, self.sd["text_linear.weight"], self.sd.get("text_linear.bias"))
        return out, logits[:, None]

    def forward_depformer(self, k: int, prev: torch.Tensor, transformer_out: torch.Tensor,
                          dep_state: tr.TransformerState) -> torch.Tensor:
        """One Depformer sub-step, lm.py:450-493.  prev [B, 1] -> logits [B, 1, 1, card]."""
        s = self.spec
        in_idx = 0
        if s.depformer_multi_linear:
            in_idx = k if s.depformer_schedule is None else s.depformer_schedule[k]
        x = tr.linear(transformer_out, self.sd[f"depformer_in.{in_idx}.weight"], self.quantize)
        emb = scaled_embedding(self.sd, "depformer_text_emb" if k == 0 else f"depformer_emb.{k - 1}", prev)
        x = x + emb
        y = tr.forward(self.sd, "depformer", self.dep_spec, x, dep_state)
        if self.quantize:
            logits = tr.linear(y, self.sd[f"linears.{k}.weight"], True)
        else:
            logits = F.linear(y, self.sd[f"linears.{k}.weight"], self.sd.get(f"linears.{k}.bias"))
        return logits[:, None]

    def depformer_step(self, text_token: torch.Tensor, transformer_out: torch.Tensor,
                       noise: tp.Sequence[torch.Tensor] | None = None,
                       logits_out: list | None = None) -> torch.Tensor:
        """lm.py:809-850 – fresh Depformer state every frame, dep_q sequential sub-steps."""
        B = text_token.shape[0]
        cfg = self.cfg_coef != 1.0
        st = tr.init_state(self.dep_spec, 2 * B if cfg else B, self.dtype)
        prev = text_token
        toks = []
        for k in range(self.spec.dep_q):
            logits = self.forward_depformer(k, (prev.repeat(2) if cfg else prev)[:, None], transformer_out, st)
            if cfg:                                                                 # lm.py:828-833
                lg, lg_null = logits.chunk(2)
                logits = lg_null + (lg - lg_null) * self.cfg_coef
            if logits_out is not None:
                logits_out.append(logits)
            nxt = sample_token(logits.float(), self.use_sampling, self.temp, self.top_k,
                               None if noise is None else noise[k], self.tie_break)[:, 0, 0]
            toks.append(nxt)
            prev = nxt
        return torch.stack(toks, dim=1)

    @torch.no_grad()
    def step(self, input_tokens: torch.Tensor, noise_text: torch.Tensor | None = None,
             noise_audio: tp.Sequence[torch.Tensor] | None = None,
             debug: dict | None = None, support_out_of_sync: bool = False,
             depformer_replace_tokens: torch.Tensor | None = None) -> torch.Tensor | None:
        """``LMGen._step`` (lm.py:668-783) -> [B, 1 + dep_q, 1] or None while warming up."""
        if self.batch is None:
            raise RuntimeError("call streaming(B) first")
        s = self.spec
        B, Ki, S = input_tokens.shape
        assert B == self.batch and S == 1
        needed = s.num_codebooks - s.dep_q - 1
        assert Ki >= needed
        input_tokens = input_tokens[:, :needed]
        CT = self.cache.shape[2]
        em = self.exec_mask[:, None, None]

        # 1. user codes go into the ring at (offset + delay) % CT   (lm.py:691-696)
        d_in = self.delays[s.dep_q + 1:]
        wpos = (self.offsets[:, None, None] + d_in[:, None]) % CT
        view = self.cache[:, s.dep_q + 1:]
        old = view.gather(-1, wpos)
        view.scatter_(-1, wpos, torch.where(em, input_tokens, old))

        # 2. model input = ring[offset % CT], initial tokens while offset <= delay (lm.py:698-702)
        is_init = (self.offsets[:, None, None] <= self.delays[:, None]) | ~em
        pos = (self.offsets % CT)[:, None, None].expand_as(is_init)
        inp = torch.where(is_init, self.initial, self.cache.gather(2, pos))

        if self.cfg_coef != 1.0:                                                    # lm.py:714-726
            zero = torch.full((1,), -1, dtype=torch.long)
            if self.cfg_is_masked_until is not None:
                limit = self.delays[:, None] + self.cfg_is_masked_until.view(-1, 1, 1)
                is_zeroed = self.offsets[:, None, None] <= limit
                inp = torch.cat([inp, torch.where(is_zeroed & ~is_init, zero, inp)], dim=0)
            else:
                inp = inp.repeat(2, 1, 1)
            if self.cfg_is_no_text:
                inp[B:, :1] = torch.where(~is_init[:, :1], zero, inp[B:, :1])
        # 3./4. temporal transformer + text sampling (lm.py:734-747)
        transformer_out, text_logits = self.forward_text(inp)
        if self.cfg_coef != 1.0:                                                    # lm.py:728-732
            logits, logits_null = text_logits.chunk(2)
            text_logits = logits if self.cfg_is_no_text else logits_null + (logits - logits_null) * self.cfg_coef
        text_token = sample_token(text_logits.float(), self.use_sampling, self.temp_text,
                                  self.top_k_text, noise_text, self.tie_break)[:, 0, 0]
        # 5. depformer (absent when dep_q == 0: "No-Depformer --- e.g., an ASR model", lm.py:219-222)
        dep_logits: list = []
        self.last_transformer_out = transformer_out
        if s.dep_q == 0:
            audio = None
        elif depformer_replace_tokens is None:
            audio = self.depformer_step(text_token, transformer_out, noise_audio, dep_logits)
        else:                                      # lm.py:751-755: the caller forces this frame's audio tokens
            assert depformer_replace_tokens.dim() == 3
            audio = depformer_replace_tokens.squeeze(-1)
        if debug is not None:
            debug.update(input=inp, transformer_out=transformer_out, text_logits=text_logits,
                         text_token=text_token, audio_tokens=audio, dep_logits=dep_logits)

        # 6. advance and store (lm.py:759-772)
        self.offsets = torch.where(self.exec_mask, self.offsets + 1, self.offsets)
        self.offset_cpu += 1
        pos = (self.offsets % CT)[:, None, None]
        tv = self.cache[:, :1]
        tv.scatter_(-1, pos, torch.where(em, text_token[:, None, None], tv.gather(-1, pos)))
        if audio is not None:
            av = self.cache[:, 1:s.dep_q + 1]
            apos = pos.expand(-1, s.dep_q, -1)
            av.scatter_(-1, apos, torch.where(em, audio[:, :, None], av.gather(-1, apos)))

        # 7. re-aligned output (lm.py:774-783)
        if not support_out_of_sync and self.offset_cpu <= s.max_delay:
            return None
        gd = self.delays[:s.dep_q + 1]
        index = (self.offsets[:, None, None] - s.max_delay + gd[:, None]) % CT
        out = self.cache.gather(2, index)
        out[(self.offsets <= s.max_delay) | ~self.exec_mask] = -2
        return out

    @torch.no_grad()
    def step_with_extra_heads(self, input_tokens: torch.Tensor, noise_text: torch.Tensor | None = None,
                              noise_audio: tp.Sequence[torch.Tensor] | None = None, **kw):
        """``LMGen.step_with_extra_heads`` (lm.py:793-807): the step's tokens plus ``softmax(extra_head(transformer_out))``
        for every extra head ([B, 1, extra_heads_dim] each, in the model dtype like the reference)."""
        out = self.step(input_tokens, noise_text, noise_audio, **kw)
        if out is None:
            return None
        heads = [torch.softmax(F.linear(self.last_transformer_out, self.sd[f"extra_heads.{i}.weight"]), dim=-1)
                 for i in range(self.spec.extra_heads_num_heads)]
        return out, heads
