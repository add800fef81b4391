#!/usr/bin/env python
"""A few Mimi streaming encode + decode frames at one batch size (for `ncu` launch lists / captures of the codec's kernels).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mimi_launches.csv \\
        python tools/mimi_frames.py --B 104 --frames 3
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=104)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    from moshi_b200.models import loaders
    mimi = loaders.get_mimi(None, device="cuda", num_codebooks=8)
    mimi.use_graph = not args.no_graph
    g = torch.Generator().manual_seed(4242)
    pcm = (0.1 * torch.randn(args.B, 1, 1920, generator=g)).cuda()
    with mimi.streaming(args.B), torch.no_grad():
        for _ in range(args.frames):
            mimi.decode(mimi.encode(pcm))
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
