"""GGUF checkpoints (the quantised form the Rust stack loads: ``rust/moshi-core/src/nn.rs:9-116``, candle's
``quantized_var_builder::VarBuilder::from_gguf``).

Reads GGUF v2 / v3 files with ``F32``, ``F16``, ``BF16`` and ``Q8_0`` tensors (ggml ``block_q8_0``: 32 weights as one fp16 scale
``d`` followed by 32 int8, value = ``d * q``; rows are whole numbers of blocks).  A ``Q8_0`` tensor is *dequantised at load* to
the bf16 values ``MaybeQuantizedWeight::to_tensor`` (``nn.rs:16-22``) would give and then takes the bf16 path: weight-only
quantisation, i.e. the file format is read, candle's q8_0 x q8_1 integer matmul is not reproduced (for int8 arithmetic use the
reference's own ``quantize=True`` / ``model.q8.safetensors`` path).  Tensor names are candle's (``scripts/import_rust.py``
layout), which ``state_dict.normalize_lm_state_dict`` maps back onto the module names.

Host-side format code only (numpy); ``write_gguf`` exists for the tests and for producing such files offline.
"""
from __future__ import annotations

import struct
import typing as tp
from pathlib import Path

import numpy as np
import torch

GGUF_MAGIC = b"GGUF"
GGML_F32, GGML_F16, GGML_Q8_0, GGML_BF16 = 0, 1, 8, 30
Q8_BLOCK = 32
_Q8_DTYPE = np.dtype([("d", "<f2"), ("qs", "i1", (Q8_BLOCK,))])          # 34 bytes

# metadata value types (gguf.md): scalar formats by id; 8 = string, 9 = array
_SCALARS = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}


class _Reader:
    def __init__(self, buf: memoryview):
        self.buf, self.pos = buf, 0

    def take(self, fmt: str):
        v = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += struct.calcsize(fmt)
        return v[0] if len(v) == 1 else v

    def string(self) -> str:
        n = self.take("<Q")
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8")
        self.pos += n
        return s

    def value(self, vtype: int):
        if vtype in _SCALARS:
            return self.take(_SCALARS[vtype])
        if vtype == 8:
            return self.string()
        if vtype == 9:
            etype, count = self.take("<I"), self.take("<Q")
            return [self.value(etype) for _ in range(count)]
        raise ValueError(f"gguf: unknown metadata value type {vtype}")


def dequantize_q8_0(raw: np.ndarray, shape: tp.Sequence[int]) -> np.ndarray:
    """``block_q8_0`` bytes -> float32 of ``shape`` (innermost dimension = whole blocks): ``d * qs`` like ggml's dequantize_row_q8_0."""
    blocks = raw.view(_Q8_DTYPE)
    out = blocks["d"].astype(np.float32)[:, None] * blocks["qs"].astype(np.float32)
    return out.reshape(tuple(shape))


def quantize_q8_0(w: np.ndarray) -> np.ndarray:
    """float32 ``[..., n]`` with ``n % 32 == 0`` -> ``block_q8_0`` bytes (ggml's quantize_row_q8_0_ref: ``d = amax / 127``,
    ``q = round(x / d)``, the scale stored as fp16)."""
    if w.shape[-1] % Q8_BLOCK:
        raise ValueError("gguf Q8_0: the innermost dimension must be a multiple of 32")
    x = np.ascontiguousarray(w, dtype=np.float32).reshape(-1, Q8_BLOCK)
    amax = np.abs(x).max(axis=1)
    d = (amax / 127.0).astype(np.float32)
    inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0).astype(np.float32)
    blocks = np.zeros(x.shape[0], dtype=_Q8_DTYPE)
    blocks["d"] = d.astype(np.float16)
    blocks["qs"] = np.clip(np.rint(x * inv[:, None]), -127, 127).astype(np.int8)
    return blocks.view(np.uint8).reshape(-1)


def read_gguf(path: str | Path) -> tuple[dict[str, tp.Any], tp.Iterator[tuple[str, torch.Tensor]]]:
    """Returns ``(metadata, tensors)``; ``tensors`` yields ``(name, tensor)`` in file order with the PyTorch shape (ggml stores
    dimensions innermost-first): F32 stays float32, F16 / BF16 / Q8_0 come back as bfloat16 values (Q8_0 dequantised)."""
    data = np.memmap(str(path), dtype=np.uint8, mode="r")
    r = _Reader(memoryview(data))
    if bytes(r.buf[:4]) != GGUF_MAGIC:
        raise ValueError(f"{path}: not a GGUF file")
    r.pos = 4
    version = r.take("<I")
    if version not in (2, 3):
        raise ValueError(f"{path}: GGUF version {version} (2 and 3 are read)")
    n_tensors, n_kv = r.take("<Q"), r.take("<Q")
    meta: dict[str, tp.Any] = {}
    for _ in range(n_kv):
        key = r.string()
        meta[key] = r.value(r.take("<I"))
    infos = []
    for _ in range(n_tensors):
        name = r.string()
        nd = r.take("<I")
        ne = [r.take("<Q") for _ in range(nd)]
        ttype, off = r.take("<I"), r.take("<Q")
        infos.append((name, tuple(reversed(ne)), ttype, off))
    align = int(meta.get("general.alignment", 32))
    base = (r.pos + align - 1) // align * align

    def tensors():
        for name, shape, ttype, off in infos:
            n = int(np.prod(shape)) if shape else 1
            start = base + off
            if ttype == GGML_F32:
                t = torch.from_numpy(np.array(data[start:start + 4 * n]).view("<f4").reshape(shape))
            elif ttype == GGML_F16:
                t = torch.from_numpy(np.array(data[start:start + 2 * n]).view("<f2").reshape(shape)).to(torch.bfloat16)
            elif ttype == GGML_BF16:
                t = torch.from_numpy(np.array(data[start:start + 2 * n]).view("<i2").reshape(shape)).view(torch.bfloat16)
            elif ttype == GGML_Q8_0:
                if not shape or shape[-1] % Q8_BLOCK:
                    raise ValueError(f"{path}: Q8_0 tensor {name!r} of shape {shape}: rows must be whole blocks of 32")
                raw = np.array(data[start:start + n // Q8_BLOCK * _Q8_DTYPE.itemsize])
                t = torch.from_numpy(dequantize_q8_0(raw, shape)).to(torch.bfloat16)
            else:
                raise ValueError(f"{path}: tensor {name!r} has ggml type {ttype}; F32, F16, BF16 and Q8_0 are read")
            yield name, t
    return meta, tensors()


def write_gguf(path: str | Path, tensors: tp.Mapping[str, torch.Tensor], q8_0: tp.Callable[[str, torch.Tensor], bool] | None = None,
               metadata: tp.Mapping[str, tp.Any] | None = None, alignment: int = 32) -> None:
    """Writes a GGUF v3 file: tensors for which ``q8_0(name, tensor)`` is true as ``Q8_0``, float32 tensors as ``F32``, the rest as
    ``BF16``.  Metadata values may be str, int, float or bool."""
    def gstr(s: str) -> bytes:
        b = s.encode("utf-8")
        return struct.pack("<Q", len(b)) + b
    kv = dict(metadata or {})
    kv.setdefault("general.alignment", alignment)
    blobs, infos, off = [], [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        if q8_0 is not None and q8_0(name, t):
            raw, ttype = quantize_q8_0(t.float().numpy()).tobytes(), GGML_Q8_0
        elif t.dtype == torch.float32:
            raw, ttype = t.numpy().astype("<f4").tobytes(), GGML_F32
        else:
            raw, ttype = t.to(torch.bfloat16).view(torch.int16).numpy().astype("<i2").tobytes(), GGML_BF16
        infos.append((name, tuple(reversed(t.shape)), ttype, off))
        pad = (-len(raw)) % alignment
        blobs.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = bytearray(GGUF_MAGIC + struct.pack("<IQQ", 3, len(infos), len(kv)))
    for k, v in kv.items():
        head += gstr(k)
        if isinstance(v, bool):
            head += struct.pack("<I?", 7, v)
        elif isinstance(v, int):
            head += struct.pack("<IQ", 10, v) if v >= 0 else struct.pack("<Iq", 11, v)
        elif isinstance(v, float):
            head += struct.pack("<Id", 12, v)
        else:
            head += struct.pack("<I", 8) + gstr(str(v))
    for name, ne, ttype, toff in infos:
        head += gstr(name) + struct.pack("<I", len(ne)) + b"".join(struct.pack("<Q", d) for d in ne) + struct.pack("<IQ", ttype, toff)
    head += b"\0" * ((-len(head)) % alignment)
    with open(path, "wb") as f:
        f.write(bytes(head))
        for b in blobs:
            f.write(b)
