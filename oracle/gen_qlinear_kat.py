"""Generates ``tests/golden/qlinear_kat.json``: a known-answer test for the int8 ``QLinear`` restatement (``oracle/quant.py``).

The reference's arithmetic for this path lives in bitsandbytes (``moshi/moshi/utils/quantize.py:13-40`` calls
``bnbF.int8_vectorwise_quant`` and ``bnb.matmul``), which is neither installed nor vendored, and no reference test or fixture
touches it: the parity of BASELINE config 5 therefore stays *unpinned against the reference*.  What this KAT
pins is the restatement itself, against an independent implementation of the published algorithm:

  * quantiser: numpy float32, ``q = rint(v * (127 / absmax_row))`` (``np.rint`` rounds half to even), weights through float16
    first (quantize.py:20), rows of zeros stay zero;
  * product: exact Python integers;
  * dequantisation: float64 ``acc * (sa * sw) / 127**2``, which the fp32 restatement must match to fp32 rounding.

Inputs include exact ties (x.5 after scaling), a zero row, a row whose absmax is negative, and values that change when they go
through float16.  Run with ``python -m oracle.gen_qlinear_kat`` (numpy only; no reference needed).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def quant_rows(t: np.ndarray):
    t = t.astype(np.float32)
    absmax = np.abs(t).max(axis=-1)
    inv = np.where(absmax > 0, np.float32(127.0) / np.where(absmax > 0, absmax, 1).astype(np.float32), np.float32(0)).astype(np.float32)
    q = np.clip(np.rint((t * inv[:, None]).astype(np.float32)), -127, 127).astype(np.int64)
    return q, absmax.astype(np.float32)


def main() -> None:
    rng = np.random.default_rng(12345)
    K = 16
    x = rng.standard_normal((4, K)).astype(np.float32)
    x[1] = 0.0                                                   # a row of zeros
    x[2, :6] = np.array([127.0, 0.5, 1.5, -0.5, -2.5, 63.5], dtype=np.float32)   # absmax 127: scale 1, exact ties
    x[2, 6:] = 0.25
    x[3] = -np.abs(x[3])
    x[3, 0] = -3.0                                               # the absmax element is negative
    # bf16-representable inputs (the LM's activations are bf16): keep 8 mantissa bits
    x = (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    w = (rng.standard_normal((5, K)) * 0.1).astype(np.float32)
    w[0, 0] = 0.100036621                                        # not a float16 value: the weight path rounds it first
    w[4] = 0.0
    w = (w.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    qw, sw = quant_rows(w.astype(np.float16).astype(np.float32))
    qx, sa = quant_rows(x)
    acc = [[int(sum(int(a) * int(b) for a, b in zip(qx[m], qw[n]))) for n in range(w.shape[0])] for m in range(x.shape[0])]
    y64 = [[acc[m][n] * (float(sa[m]) * float(sw[n])) / 16129.0 for n in range(w.shape[0])] for m in range(x.shape[0])]
    out = {"generated_by": "oracle/gen_qlinear_kat.py", "x": x.tolist(), "w": w.tolist(), "qx": qx.tolist(), "sa": sa.tolist(),
           "qw": qw.tolist(), "sw": sw.tolist(), "acc": acc, "y_float64": y64}
    (ROOT / "tests" / "golden" / "qlinear_kat.json").write_text(json.dumps(out, indent=1))
    print("ties row:", qx[2][:6].tolist(), "| acc[0]:", acc[0])


if __name__ == "__main__":
    main()
