#!/usr/bin/env python
"""A few Moshi 7B `LMGen.step` calls at one batch size (for `ncu` launch lists / captures of the LM's kernels).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lm_launches_b1.csv \\
        python tools/lm_frames.py --B 1 --steps 3 --no-graph
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--fill", type=int, default=200)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    from moshi_b200.config import MOSHI_7B
    from moshi_b200.models import LMGen, loaders
    lm = loaders.get_moshi_lm(None, device="cuda", synth_device="cuda")
    gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7)
    gen.use_graph = not args.no_graph
    g = torch.Generator().manual_seed(1)
    codes = torch.randint(0, MOSHI_7B.card, (args.B, 8, 1), generator=g).cuda()
    with gen.streaming(args.B):
        gen.assume_fill(args.fill)
        for _ in range(args.steps):
            gen.step(codes)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
