#!/usr/bin/env python
"""Per-phase timing of the fused depformer kernel (diagnostics): with B200_DEP_TRACE=1 CTA 0 records %globaltimer after every
grid barrier; this prints the time between consecutive barriers grouped by the phase that ran in between.

    B200_DEP_TRACE=1 python tools/dep_trace.py --B 104 > gpurun_out/dep_trace_b104.json
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["B200_DEP_TRACE"] = "1"

import torch  # noqa: E402


def phase_names(dep_q: int, L: int) -> list[str]:
    """Phase that ENDS at barrier e (1-based), in the kernel's order (dep_fused.cu)."""
    names = []
    for k in range(dep_q):
        names.append("input+norm" if k == 0 else "sample+input+norm")
        for _ in range(L):
            names += ["gemm.in_proj", "attn", "gemm.out_proj", "row(res+norm)", "gemm.lin_in", "gate", "gemm.lin_out", "row(res+norm)"]
        names.append("gemm.head")
    return names


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=104)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    from moshi_b200.config import MOSHI_7B
    from moshi_b200.models import LMGen, loaders
    lm = loaders.get_moshi_lm(None, device="cuda")
    gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7)
    cfg = MOSHI_7B
    g = torch.Generator().manual_seed(1)
    codes = torch.randint(0, cfg.card, (args.B, 8, 1), generator=g).cuda()
    names = phase_names(cfg.dep_q, cfg.depformer_num_layers)
    with gen.streaming(args.B):
        gen.assume_fill(200)
        for _ in range(args.steps):
            gen.step(codes)
        torch.cuda.synchronize()
        tr = gen.read_buffer("dep_trace", torch.int64, (512,)).cpu().tolist()
    n = len(names)
    ts = tr[:n + 1]
    d = [(ts[i + 1] - ts[i]) / 1e3 for i in range(n)]
    agg = {}
    for nm, v in zip(names, d):
        a = agg.setdefault(nm, [0, 0.0, 1e9, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    out = {"kernel": "dep_fused", "B": args.B, "total_us": (ts[n] - ts[0]) / 1e3, "barriers": n,
           "phases": {k: {"count": a[0], "sum_us": round(a[1], 1), "avg_us": round(a[1] / a[0], 2), "min_us": round(a[2], 2), "max_us": round(a[3], 2)}
                      for k, a in agg.items()},
           "substep0_us": [round(v, 2) for v in d[:names.index("gemm.head") + 1]]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
