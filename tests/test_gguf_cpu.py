"""CPU: the GGUF reader / writer (moshi_b200/models/gguf.py; the quantised checkpoint form of rust/moshi-core/src/nn.rs:9-116)."""
import struct

import numpy as np
import torch

from moshi_b200.models import gguf


def test_q8_0_block_known_answer():
    """ggml block_q8_0: fp16 scale d = amax / 127, q = round(x / d) (quantize_row_q8_0_ref); value = d * q."""
    x = np.zeros((1, 32), dtype=np.float32)
    x[0, :4] = [127.0, -63.5, 1.0, 0.49]
    raw = gguf.quantize_q8_0(x)
    assert raw.dtype == np.uint8 and raw.size == 34
    d = np.frombuffer(raw[:2].tobytes(), dtype="<f2")[0]
    q = np.frombuffer(raw[2:].tobytes(), dtype=np.int8)
    assert d == np.float16(1.0) and q[:4].tolist() == [127, -64, 1, 0] and not q[4:].any()      # -63.5 rounds to even: -64
    back = gguf.dequantize_q8_0(raw, (1, 32))
    assert back[0, :4].tolist() == [127.0, -64.0, 1.0, 0.0]
    # a block of zeros stays zero (d = 0)
    assert not gguf.dequantize_q8_0(gguf.quantize_q8_0(np.zeros((2, 64), np.float32)), (2, 64)).any()
    # error bound: half a step of each block's own scale (+ the fp16 rounding of the scale)
    g = np.random.default_rng(0)
    w = g.standard_normal((8, 96)).astype(np.float32)
    deq = gguf.dequantize_q8_0(gguf.quantize_q8_0(w), w.shape)
    step = np.abs(w.reshape(-1, 32)).max(axis=1) / 127
    assert (np.abs(deq - w).reshape(-1, 32).max(axis=1) <= step * 0.51 + 1e-6).all()


def test_gguf_round_trip_and_header_layout(tmp_path):
    g = torch.Generator().manual_seed(1)
    tensors = {
        "transformer.layers.0.gating.linear_in.weight": torch.randn(24, 64, generator=g).bfloat16(),      # -> Q8_0
        "out_norm.alpha": torch.randn(1, 1, 64, generator=g).bfloat16(),                                   # -> BF16
        "some.f32": torch.randn(3, 5, generator=g),                                                        # -> F32
    }
    path = tmp_path / "m.gguf"
    gguf.write_gguf(path, tensors, q8_0=lambda n, t: n.endswith("linear_in.weight"), metadata={"general.architecture": "moshi"})
    raw = path.read_bytes()
    assert raw[:4] == b"GGUF" and struct.unpack_from("<IQQ", raw, 4) == (3, 3, 2)
    meta, it = gguf.read_gguf(path)
    got = dict(it)
    assert meta["general.architecture"] == "moshi" and meta["general.alignment"] == 32
    assert list(got) == list(tensors)
    assert got["some.f32"].dtype == torch.float32 and torch.equal(got["some.f32"], tensors["some.f32"])
    assert got["out_norm.alpha"].dtype == torch.bfloat16 and torch.equal(got["out_norm.alpha"], tensors["out_norm.alpha"])
    w = tensors["transformer.layers.0.gating.linear_in.weight"]
    want = torch.from_numpy(gguf.dequantize_q8_0(gguf.quantize_q8_0(w.float().numpy()), w.shape)).bfloat16()
    assert got["transformer.layers.0.gating.linear_in.weight"].shape == w.shape
    assert torch.equal(got["transformer.layers.0.gating.linear_in.weight"], want)
    assert (want.float() - w.float()).abs().max() < 0.05
