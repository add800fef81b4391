"""``LMModel`` / ``LMGen`` with the reference's public surface, backed by the sm_100a library.

Mirrors ``moshi/moshi/models/lm.py``: ``LMModel`` exposes the attributes callers read
(``dep_q, n_q, card, text_card, delays, device, dtype, num_codebooks, ...``, lm.py:248-295) and owns
the weights on the device; ``LMGen`` (lm.py:555-850) owns the streaming state and runs one
``step`` per 80 ms frame.  The Exp(1) noise of ``sampling.py:44`` is drawn here with torch, in the
reference's order and shapes, and handed to the fused sampler, so that a run with the same seed
reproduces the reference's token stream (up to logit ties).
"""
from __future__ import annotations

import ctypes as C
import typing as tp
from contextlib import ExitStack

import torch

from .. import _lib
from ..config import LMConfig
from .state_dict import normalize_lm_state_dict


def _config_struct(cfg: LMConfig) -> _lib.LMConfigC:
    c = _lib.LMConfigC()
    c.dim, c.text_card, c.n_q, c.dep_q, c.card = cfg.dim, cfg.text_card, cfg.n_q, cfg.dep_q, cfg.card
    c.num_heads, c.num_layers, c.ffn_hidden, c.context = cfg.num_heads, cfg.num_layers, cfg.ffn_hidden, cfg.context
    c.max_period = cfg.max_period
    c.depformer_dim, c.depformer_num_heads = cfg.depformer_dim, cfg.depformer_num_heads
    c.depformer_num_layers, c.depformer_ffn_hidden = cfg.depformer_num_layers, cfg.depformer_ffn_hidden
    for i, d in enumerate(cfg.delays):
        c.delays[i] = d
    c.quantize = int(bool(cfg.quantize))
    return c


class LMModel:
    """Moshi Temporal + Depth transformer weights on one B200.  Construct via ``loaders.get_moshi_lm``."""

    def __init__(self, cfg: LMConfig, tensors: tp.Iterable[tuple[str, torch.Tensor]] | tp.Mapping[str, torch.Tensor],
                 device: torch.device | str = "cuda", dtype: torch.dtype = torch.bfloat16):
        cfg.check_supported()
        if dtype != torch.bfloat16:
            raise ValueError("the B200 LM path computes in bfloat16 (reference default, loaders.py:370)")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("moshi_b200.LMModel runs on a CUDA device only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        # attributes read by callers (lm.py:119-160, 248-295)
        self.n_q, self.dep_q, self.card, self.text_card = cfg.n_q, cfg.dep_q, cfg.card, cfg.text_card
        self.delays = list(cfg.delays)
        self.dim = cfg.dim
        self.context = cfg.context
        self.existing_text_padding_id = cfg.existing_text_padding_id
        self.condition_provider = None
        self.fuser = None
        self.depformer = True
        self.extra_heads: list = []
        self.training = False
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        items = tensors.items() if isinstance(tensors, tp.Mapping) else tensors
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200_lm_create(C.byref(_config_struct(cfg)), C.byref(self._h)))
            for name, t in items:
                for n2, t2 in normalize_lm_state_dict({name: t}).items():
                    t2 = t2.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
                    shape = (C.c_int64 * t2.dim())(*t2.shape)
                    _lib.check(self._lib.b200_lm_load_tensor(self._h, n2.encode(), _lib.ptr(t2), _lib.B200_BF16,
                                                             t2.dim(), shape))
                    del t2
            _lib.check(self._lib.b200_lm_finalize(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.b200_lm_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    @property
    def num_codebooks(self) -> int:
        return self.n_q + 1

    @property
    def num_audio_codebooks(self) -> int:
        return self.n_q

    @property
    def audio_offset(self) -> int:
        return 1

    @property
    def initial_token_id(self) -> int:
        return self.card

    @property
    def text_initial_token_id(self) -> int:
        return self.text_card

    @property
    def text_padding_token_id(self) -> int:
        return self.existing_text_padding_id

    @property
    def zero_token_id(self) -> int:
        return -1

    @property
    def ungenerated_token_id(self) -> int:
        return -2

    def eval(self) -> "LMModel":
        return self


class LMGen:
    """``LMGen`` (lm.py:555-850) for the options of the 7B dialogue model (no CFG, no conditioners)."""

    def __init__(self, lm_model: LMModel, use_sampling: bool = True, temp: float = 0.8, temp_text: float = 0.7,
                 top_k: int = 250, top_k_text: int = 25, cfg_coef: float = 1., check: bool = False,
                 condition_tensors=None, on_text_hook=None, on_text_logits_hook=None, on_audio_hook=None,
                 support_out_of_sync: bool = False, cfg_is_masked_until=None, cfg_is_no_text: bool = False):
        if cfg_coef != 1. or condition_tensors or cfg_is_masked_until or cfg_is_no_text:
            raise ValueError("CFG / conditioning are outside the B200 hot path (SURVEY.md 8(f) item 2)")
        self.lm_model = lm_model
        self.use_sampling, self.temp, self.temp_text = use_sampling, temp, temp_text
        self.top_k, self.top_k_text = top_k, top_k_text
        self.cfg_coef = cfg_coef
        self.check = check
        self.max_delay = max(lm_model.delays)
        self.delays_cuda = torch.tensor(lm_model.delays, device=lm_model.device, dtype=torch.long)
        self.on_text_hook, self.on_text_logits_hook, self.on_audio_hook = on_text_hook, on_text_logits_hook, on_audio_hook
        self.support_out_of_sync = support_out_of_sync
        self._lib = lm_model._lib
        self._h = lm_model._h
        self._batch: int | None = None
        self._streaming_state = None      # truthy while streaming (callers test `lm_gen._streaming_state`)
        self.use_graph = True
        # "bf16" = the reference's ring; "fp8_e4m3" / "int8" = opt-in extensions outside the reference's numerics (half the ring)
        self.kv_dtype = "bf16"

    # ---- streaming protocol ----------------------------------------------------------------------
    @property
    def is_streaming(self) -> bool:
        return self._batch is not None

    def _start(self, batch_size: int) -> None:
        assert self._batch is None, "lm_gen is already streaming!"
        dev = self.lm_model.device
        with torch.cuda.device(dev):
            _lib.check(self._lib.b200_lm_set_sampling(self._h, int(self.use_sampling), float(self.temp),
                                                      float(self.temp_text), int(self.top_k), int(self.top_k_text)))
            _lib.check(self._lib.b200_lm_set_graph(self._h, int(self.use_graph)))
            kinds = {"bf16": 0, "fp8_e4m3": 1, "int8": 2}
            if self.kv_dtype not in kinds:
                raise ValueError(f"kv_dtype {self.kv_dtype!r}: expected one of {sorted(kinds)}")
            _lib.check(self._lib.b200_lm_set_kv_dtype(self._h, kinds[self.kv_dtype]))
            _lib.check(self._lib.b200_lm_streaming_begin(self._h, int(batch_size), _lib.current_stream(dev)))
        self._batch = int(batch_size)
        self._noise_per_row = int(self._lib.b200_lm_noise_per_row(self._h))
        self._kt = min(self.top_k_text, self.lm_model.text_card)
        self._ka = min(self.top_k, self.lm_model.card)
        self._streaming_state = self

    def _stop(self) -> None:
        if self._batch is not None:
            self._lib.b200_lm_streaming_end(self._h)
            self._batch = None
            self._streaming_state = None

    def streaming(self, batch_size: int) -> ExitStack:
        stack = ExitStack()
        self._start(batch_size)
        stack.callback(self._stop)
        return stack

    def streaming_forever(self, batch_size: int) -> None:
        self._start(batch_size)

    def _mask(self, mask: torch.Tensor) -> torch.Tensor:
        mask = mask.to(device=self.lm_model.device, dtype=torch.bool).contiguous()
        assert mask.shape == (self._batch,)
        return mask

    def reset_streaming(self, reset_mask: torch.Tensor | None = None) -> None:
        assert self._batch is not None, "Trying to reset streaming, but lm_gen wasn't streaming."
        m = None if reset_mask is None else self._mask(reset_mask)
        _lib.check(self._lib.b200_lm_reset(self._h, _lib.ptr(m)))

    def set_exec_mask(self, exec_mask: torch.Tensor) -> None:
        assert self._batch is not None
        m = self._mask(exec_mask)
        _lib.check(self._lib.b200_lm_set_exec_mask(self._h, _lib.ptr(m)))

    def get_streaming_state(self) -> dict:
        """Snapshot of the generation state (token ring, offsets, KV rings) as one device blob."""
        assert self._batch is not None, "lm_gen is not streaming"
        n = int(self._lib.b200_lm_state_bytes(self._h))
        blob = torch.empty(n, dtype=torch.uint8, device=self.lm_model.device)
        _lib.check(self._lib.b200_lm_get_state(self._h, _lib.ptr(blob), n))
        return {"batch_size": self._batch, "blob": blob}

    def set_streaming_state(self, state: dict) -> None:
        assert self._batch is not None, "lm_gen is not streaming"
        assert state["batch_size"] == self._batch, "snapshot was taken with another batch size"
        blob = state["blob"].to(self.lm_model.device).contiguous()
        _lib.check(self._lib.b200_lm_set_state(self._h, _lib.ptr(blob), blob.numel()))

    # ---- sampling noise --------------------------------------------------------------------------
    def draw_noise(self) -> torch.Tensor | None:
        """Exp(1) draws of one step in the reference's order (lm.py:736 then lm.py:836 x dep_q),
        each ``torch.empty(B, k).exponential_(1)`` on the model device, packed as [B, kt + dep_q*ka]."""
        if not (self.use_sampling and (self.temp > 0 or self.temp_text > 0)):
            return None
        B, dev = self._batch, self.lm_model.device
        parts = [torch.empty(B, self._kt, device=dev, dtype=torch.float32).exponential_(1)]
        for _ in range(self.lm_model.dep_q):
            parts.append(torch.empty(B, self._ka, device=dev, dtype=torch.float32).exponential_(1))
        return torch.cat(parts, dim=1).contiguous()

    def pack_noise(self, noise_text: torch.Tensor, noise_audio: tp.Sequence[torch.Tensor]) -> torch.Tensor:
        dev = self.lm_model.device
        return torch.cat([noise_text.to(dev)] + [n.to(dev) for n in noise_audio], dim=1).float().contiguous()

    def read_buffer(self, name: str, dtype: torch.dtype, shape: tuple[int, ...]) -> torch.Tensor:
        out = torch.empty(shape, device=self.lm_model.device, dtype=dtype)
        n = C.c_int64()
        _lib.check(self._lib.b200_lm_read_buffer(self._h, name.encode(), _lib.ptr(out),
                                                 out.numel() * out.element_size(), C.byref(n)))
        assert n.value == out.numel() * out.element_size(), (name, n.value, shape)
        return out

    # ---- the step ---------------------------------------------------------------------------------
    @torch.no_grad()
    def _step(self, input_tokens: torch.Tensor, depformer_replace_tokens=None, noise: torch.Tensor | None = None):
        if self._batch is None:
            raise RuntimeError("You should wrap those calls with a `with lm_gen.streaming(): ...`.")   # lm.py:673-676
        lm = self.lm_model
        assert input_tokens.dim() == 3, "Shape should be [B, K, T]."
        B, Ki, S = input_tokens.shape
        assert B == self._batch, f"Got a batch size {B}, expected {self._batch}"
        assert S == 1, "Only support being given steps one by one."
        needed = lm.num_codebooks - lm.dep_q - 1
        assert Ki >= needed, f"We expect {needed} tokens from the user stream, got {Ki}."
        codes = input_tokens[:, :needed, 0].to(device=lm.device, dtype=torch.int64).contiguous()
        if noise is None:
            noise = self.draw_noise()
        out = torch.empty(B, lm.dep_q + 1, device=lm.device, dtype=torch.int64)
        ready = C.c_int(0)
        replace = None
        if depformer_replace_tokens is not None:            # lm.py:751-755
            assert depformer_replace_tokens.dim() == 3
            replace = depformer_replace_tokens.squeeze(-1).to(device=lm.device, dtype=torch.int64).contiguous()
            assert replace.shape == (B, lm.dep_q), f"expected [{B}, {lm.dep_q}, 1] replacement tokens"
        _lib.check(self._lib.b200_lm_step_ex(self._h, _lib.ptr(codes), needed, _lib.ptr(noise), _lib.ptr(replace), _lib.ptr(out),
                                             int(self.support_out_of_sync), C.byref(ready)))
        if self.on_text_logits_hook is not None:
            tl = self.read_buffer("text_logits", torch.bfloat16, (B, lm.text_card))
            self.on_text_logits_hook(tl[:, None, None, :])
        if self.on_text_hook is not None:
            self.on_text_hook(self.read_buffer("text_token", torch.int64, (B,)))
        if self.on_audio_hook is not None:
            self.on_audio_hook(self.read_buffer("audio_tokens", torch.int64, (lm.dep_q, B)).t().contiguous())
        if not ready.value:
            return None
        return out[:, :, None]

    def step(self, input_tokens: torch.Tensor, depformer_replace_tokens=None,
             noise: torch.Tensor | None = None) -> torch.Tensor | None:
        return self._step(input_tokens, depformer_replace_tokens, noise)

    def step_with_extra_heads(self, input_tokens: torch.Tensor, depformer_replace_tokens=None):
        out = self._step(input_tokens, depformer_replace_tokens)
        if out is None:
            return None
        return out, []   # the 7B dialogue model has no extra heads (lm.py:224-226)

    def step_host(self, codes_cpu: torch.Tensor, noise_cpu: torch.Tensor | None, out_cpu: torch.Tensor) -> bool:
        """Host-buffer step (H2D of codes/noise and D2H of tokens inside): i64 [B, n_in] -> i64 [B, dep_q+1]."""
        ready = C.c_int(0)
        _lib.check(self._lib.b200_lm_step_host(self._h, _lib.ptr(codes_cpu), codes_cpu.shape[1], _lib.ptr(noise_cpu),
                                               _lib.ptr(out_cpu), int(self.support_out_of_sync), C.byref(ready)))
        return bool(ready.value)

    def assume_fill(self, fill: int) -> None:
        _lib.check(self._lib.b200_lm_assume_fill(self._h, int(fill)))

    def algorithmic_bytes(self, kv_fill: int) -> int:
        return int(self._lib.b200_lm_algorithmic_bytes(self._h, int(kv_fill)))
