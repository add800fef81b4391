"""``MimiModel`` with the reference's public surface, backed by the sm_100a library.

Mirrors ``moshi/moshi/models/compression.py:97-433`` (properties, ``encode`` / ``decode`` /
``encode_to_latent`` / ``decode_latent``) and the ``StreamingModule`` protocol of
``moshi/moshi/modules/streaming.py:54-211`` (``streaming`` / ``streaming_forever`` /
``reset_streaming`` / ``set_exec_mask``).  torch tensors are only owners of device memory here; all
arithmetic happens in ``libmoshi_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import typing as tp
from contextlib import ExitStack

import torch

from .. import _lib
from ..config import MimiConfig
from .state_dict import normalize_mimi_state_dict


def _config_struct(cfg: MimiConfig) -> _lib.MimiConfigC:
    c = _lib.MimiConfigC()
    c.sample_rate, c.frame_rate, c.channels = cfg.sample_rate, cfg.frame_rate, cfg.channels
    c.dimension, c.n_filters, c.n_residual_layers = cfg.dimension, cfg.n_filters, cfg.n_residual_layers
    c.n_ratios = len(cfg.ratios)
    for i, r in enumerate(cfg.ratios):
        c.ratios[i] = r
    c.kernel_size, c.residual_kernel_size = cfg.kernel_size, cfg.residual_kernel_size
    c.last_kernel_size, c.dilation_base, c.compress = cfg.last_kernel_size, cfg.dilation_base, cfg.compress
    c.tr_d_model, c.tr_num_heads, c.tr_num_layers = cfg.tr_d_model, cfg.tr_num_heads, cfg.tr_num_layers
    c.tr_dim_feedforward, c.tr_context, c.tr_max_period = cfg.tr_dim_feedforward, cfg.tr_context, cfg.tr_max_period
    c.q_dimension, c.q_bins, c.q_n_q, c.q_n_semantic = cfg.q_dimension, cfg.q_bins, cfg.q_n_q, cfg.q_n_semantic
    c.num_codebooks = cfg.num_codebooks
    return c


class MimiModel:
    """Mimi on one B200.  Construct through ``loaders.get_mimi``."""

    def __init__(self, cfg: MimiConfig, state_dict: tp.Mapping[str, torch.Tensor],
                 device: torch.device | str = "cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("moshi_b200.MimiModel runs on a CUDA device only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        self._batch: int | None = None
        self._num_codebooks = cfg.num_codebooks
        self.use_graph = True      # replay each one-frame encode / decode as a CUDA graph
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200_mimi_create(C.byref(_config_struct(cfg)), C.byref(self._h)))
            for name, t in normalize_mimi_state_dict(state_dict).items():
                if not t.dtype.is_floating_point:
                    continue
                t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(self._lib.b200_mimi_load_tensor(self._h, name.encode(), _lib.ptr(t), _lib.B200_F32,
                                                           t.dim(), shape))
            _lib.check(self._lib.b200_mimi_finalize(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.b200_mimi_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ---- properties (compression.py:240-275) ----------------------------------------------------
    @property
    def channels(self) -> int:
        return self.cfg.channels

    @property
    def frame_rate(self) -> float:
        return self.cfg.frame_rate

    @property
    def sample_rate(self) -> int:
        return self.cfg.sample_rate

    @property
    def frame_size(self) -> int:
        return self.cfg.frame_size

    @property
    def dimension(self) -> int:
        return self.cfg.dimension

    @property
    def total_codebooks(self) -> int:
        return self.cfg.q_n_q

    @property
    def num_codebooks(self) -> int:
        return self._num_codebooks

    @property
    def cardinality(self) -> int:
        return self.cfg.q_bins

    def set_num_codebooks(self, n: int) -> None:
        _lib.check(self._lib.b200_mimi_set_num_codebooks(self._h, int(n)))
        self._num_codebooks = int(n)

    def eval(self) -> "MimiModel":
        return self

    # ---- streaming protocol (streaming.py:131-211) ----------------------------------------------
    @property
    def is_streaming(self) -> bool:
        return self._batch is not None

    def _start(self, batch_size: int) -> None:
        assert self._batch is None, "mimi is already streaming!"          # streaming.py:112
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200_mimi_set_graph(self._h, int(self.use_graph)))
            _lib.check(self._lib.b200_mimi_streaming_begin(self._h, int(batch_size), _lib.current_stream(self.device)))
        self._batch = int(batch_size)

    def _stop(self) -> None:
        if self._batch is not None:
            self._lib.b200_mimi_streaming_end(self._h)
            self._batch = None

    def streaming(self, batch_size: int) -> ExitStack:
        stack = ExitStack()
        self._start(batch_size)
        stack.callback(self._stop)
        return stack

    def streaming_forever(self, batch_size: int) -> None:
        self._start(batch_size)

    def _mask(self, mask: torch.Tensor) -> torch.Tensor:
        assert self._batch is not None
        mask = mask.to(device=self.device, dtype=torch.bool).contiguous()
        assert mask.shape == (self._batch,), f"mask must have shape ({self._batch},)"
        return mask

    def reset_streaming(self, reset_mask: torch.Tensor | None = None) -> None:
        assert self._batch is not None, "Trying to reset streaming, but mimi wasn't streaming."
        m = None if reset_mask is None else self._mask(reset_mask)
        _lib.check(self._lib.b200_mimi_reset(self._h, _lib.ptr(m)))

    def set_exec_mask(self, exec_mask: torch.Tensor) -> None:
        assert self._batch is not None
        m = self._mask(exec_mask)
        _lib.check(self._lib.b200_mimi_set_exec_mask(self._h, _lib.ptr(m)))

    def get_streaming_state(self) -> dict:
        """Snapshot of all streaming state (streaming.py:158-170) as one device blob."""
        assert self._batch is not None, "mimi is not streaming"
        n = int(self._lib.b200_mimi_state_bytes(self._h))
        blob = torch.empty(n, dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.b200_mimi_get_state(self._h, _lib.ptr(blob), n))
        return {"batch_size": self._batch, "blob": blob}

    def set_streaming_state(self, state: dict) -> None:
        """Restore a snapshot taken by ``get_streaming_state`` (streaming.py:172-181)."""
        assert self._batch is not None, "mimi is not streaming"
        assert state["batch_size"] == self._batch, "snapshot was taken with another batch size"
        blob = state["blob"].to(self.device).contiguous()
        _lib.check(self._lib.b200_mimi_set_state(self._h, _lib.ptr(blob), blob.numel()))

    # ---- data path ------------------------------------------------------------------------------
    def _check_pcm(self, x: torch.Tensor) -> torch.Tensor:
        assert x.dim() == 3, f"expected audio of shape [B, C, T] but got {tuple(x.shape)}"
        assert x.shape[1] == self.channels
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def _with_session(self, batch: int):
        """Non-streaming calls run on a throw-away streaming session (fresh zero state), which is
        what the reference's ``state is None`` branch computes (conv.py:249-251, transformer.py:538-545)."""
        stack = ExitStack()
        if self._batch is None:
            stack.enter_context(self.streaming(batch))
        return stack

    def _encode_to_unquantized_latent(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_pcm(x)
        fs = self.frame_size
        if self._batch is None:
            extra = (-x.shape[-1]) % fs                                   # pad_for_conv1d, compression.py:358
            if extra:
                x = torch.nn.functional.pad(x, (0, extra))
        elif x.shape[-1] % fs != 0 or x.shape[-1] == 0:
            raise RuntimeError(
                f"Invalid input x of length {x.shape[-1]}. The length must be "
                f"a positive multiple of the frame size {fs}. "
                "You are responsible for buffering accordingly before feeding audio to Mimi.")
        n = x.shape[-1] // fs
        with self._with_session(x.shape[0]):
            assert x.shape[0] == self._batch, f"Got a batch size {x.shape[0]}, expected {self._batch}"
            out = torch.empty(x.shape[0], self.dimension, n, device=self.device, dtype=torch.float32)
            _lib.check(self._lib.b200_mimi_encode_to_latent(self._h, _lib.ptr(x), n, _lib.ptr(out)))
        return out

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """f32 [B, 1, T] -> i64 [B, K, T / frame_size]  (compression.py:376-388)."""
        x = self._check_pcm(x)
        fs = self.frame_size
        if self._batch is None:
            extra = (-x.shape[-1]) % fs
            if extra:
                x = torch.nn.functional.pad(x, (0, extra))
        elif x.shape[-1] % fs != 0 or x.shape[-1] == 0:
            raise RuntimeError(
                f"Invalid input x of length {x.shape[-1]}. The length must be "
                f"a positive multiple of the frame size {fs}. "
                "You are responsible for buffering accordingly before feeding audio to Mimi.")
        n = x.shape[-1] // fs
        with self._with_session(x.shape[0]):
            assert x.shape[0] == self._batch, f"Got a batch size {x.shape[0]}, expected {self._batch}"
            codes = torch.empty(x.shape[0], self.num_codebooks, n, device=self.device, dtype=torch.int64)
            _lib.check(self._lib.b200_mimi_encode(self._h, _lib.ptr(x), n, _lib.ptr(codes)))
        return codes

    def encode_to_latent(self, x: torch.Tensor, quantize: bool = True) -> torch.Tensor:
        emb = self._encode_to_unquantized_latent(x)
        if not quantize:
            return emb
        with self._with_session(emb.shape[0]):
            codes = torch.empty(emb.shape[0], self.num_codebooks, emb.shape[-1], device=self.device, dtype=torch.int64)
            _lib.check(self._lib.b200_mimi_quantize(self._h, _lib.ptr(emb), emb.shape[-1], _lib.ptr(codes)))
        return self.decode_latent(codes)

    def _check_codes(self, codes: torch.Tensor) -> torch.Tensor:
        assert codes.dim() == 3, f"expected codes of shape [B, K, T] but got {tuple(codes.shape)}"
        assert not codes.dtype.is_floating_point, f"Codes should be integers, got {codes.dtype}"
        return codes.to(device=self.device, dtype=torch.int64).contiguous()

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """i64 [B, K, T] -> f32 [B, 1, T * frame_size]  (compression.py:406-429)."""
        codes = self._check_codes(codes)
        B, K, n = codes.shape
        with self._with_session(B):
            assert B == self._batch, f"Got a batch size {B}, expected {self._batch}"
            out = torch.empty(B, 1, n * self.frame_size, device=self.device, dtype=torch.float32)
            _lib.check(self._lib.b200_mimi_decode(self._h, _lib.ptr(codes), K, n, _lib.ptr(out)))
        return out

    def decode_latent(self, codes: torch.Tensor) -> torch.Tensor:
        codes = self._check_codes(codes)
        B, K, n = codes.shape
        with self._with_session(B):
            out = torch.empty(B, self.dimension, n, device=self.device, dtype=torch.float32)
            _lib.check(self._lib.b200_mimi_decode_latent(self._h, _lib.ptr(codes), K, n, _lib.ptr(out)))
        return out

    # ---- host-buffer entry points (bench.py e2e) --------------------------------------------------
    def encode_host(self, pcm_cpu: torch.Tensor, codes_cpu: torch.Tensor) -> None:
        n = pcm_cpu.shape[-1] // self.frame_size
        _lib.check(self._lib.b200_mimi_encode_host(self._h, _lib.ptr(pcm_cpu), n, _lib.ptr(codes_cpu)))

    def decode_host(self, codes_cpu: torch.Tensor, pcm_cpu: torch.Tensor) -> None:
        B, K, n = codes_cpu.shape
        _lib.check(self._lib.b200_mimi_decode_host(self._h, _lib.ptr(codes_cpu), K, n, _lib.ptr(pcm_cpu)))

    # ---- debugging / measurement -------------------------------------------------------------------
    def debug_buffer(self, name: str) -> torch.Tensor:
        """Copy of a named intermediate of the last encode/decode call (flat fp32)."""
        n = C.c_int64()
        _lib.check(self._lib.b200_mimi_read_buffer(self._h, name.encode(), None, 0, C.byref(n)))
        out = torch.empty(n.value, device=self.device, dtype=torch.float32)
        _lib.check(self._lib.b200_mimi_read_buffer(self._h, name.encode(), _lib.ptr(out), n.value, C.byref(n)))
        return out

    def algorithmic_bytes(self) -> int:
        return int(self._lib.b200_mimi_algorithmic_bytes(self._h))
