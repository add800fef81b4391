#!/usr/bin/env python
"""BASELINE.json configs 2-5 through the reference-shaped API, one JSON object per line.

    python tools/configs.py --config 2              # Mimi streaming encode+decode, B = 1 .. 256
    python tools/configs.py --config 3              # Moshi 7B LMGen.step, B = 1: p50/p90 latency, xRT (fill 200 and full ring)
    python tools/configs.py --config 4 [--sessions N]   # one GPU's shard of the 512-session config, all rows / 25 % masked
    python tools/configs.py --config 4 --sessions -512  # 512 / N sessions on one GPU (N = 4, 2) with the rings sized to what fits
    python tools/configs.py --config 5 [--kv-dtype int8]   # int8 (QLinear) weights: sessions swept up to what HBM holds
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

PEAK = 6576.1
pk = ROOT / "MEASURED_PEAKS.json"
if pk.exists():
    PEAK = float(json.loads(pk.read_text())["hbm_gbs"])


def config2(batches):
    from moshi_b200.models import loaders
    mimi = loaders.get_mimi(None, device="cuda", num_codebooks=8)
    g = torch.Generator().manual_seed(4242)
    for B in batches:
        pcm = (0.1 * torch.randn(B, 1, 1920, generator=g)).cuda()
        with mimi.streaming(B), torch.no_grad():
            codes = None
            for _ in range(255):                       # >= 250 warm-up frames: both transformer KV rings full
                codes = mimi.encode(pcm)
                mimi.decode(codes)
            torch.cuda.synchronize()
            n = 100
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                codes = mimi.encode(pcm)
                mimi.decode(codes)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            alg = mimi.algorithmic_bytes()
        print(json.dumps({"config": 2, "what": "Mimi streaming encode+decode, 8 codebooks, rings full", "B": B,
                          "ms_per_frame_pair": round(ms, 4), "frames_per_s": round(B * 1e3 / ms, 1),
                          "algorithmic_GBps": round(alg / ms / 1e6, 1), "frac_of_hbm_peak": round(alg / ms / 1e6 / PEAK, 4),
                          "fp32_TFLOPs": round(B * 0.9 / ms, 2)}), flush=True)


def _lm():
    from moshi_b200.config import MOSHI_7B
    from moshi_b200.models import loaders
    return loaders.get_moshi_lm(None, MOSHI_7B.to_reference_kwargs(), device="cuda", synth_device="cuda"), MOSHI_7B


def config3():
    from moshi_b200.models import LMGen
    lm, cfg = _lm()
    g = torch.Generator().manual_seed(4242)
    codes = torch.randint(0, 2048, (1, 8, 1), generator=g).cuda()
    for fill in (200, cfg.context):
        gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7, top_k=250, top_k_text=25)
        out_host = torch.empty(1, 9, 1, dtype=torch.int64).pin_memory()
        with gen.streaming(1), torch.no_grad():
            gen.assume_fill(fill)
            for _ in range(20):
                gen.step(codes)
            torch.cuda.synchronize()
            dev, wall = [], []
            for _ in range(200):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                o = gen.step(codes)
                e1.record()
                out_host.copy_(o, non_blocking=True)
                torch.cuda.synchronize()
                wall.append((time.perf_counter() - t0) * 1e3)
                dev.append(e0.elapsed_time(e1))
            alg = gen.algorithmic_bytes(fill)
        q = lambda v, p: sorted(v)[int(p * (len(v) - 1))]
        p50 = q(dev, 0.5)
        print(json.dumps({"config": 3, "what": "Moshi 7B bf16 LMGen.step, B=1", "kv_fill": fill,
                          "p50_ms_device": round(p50, 3), "p90_ms_device": round(q(dev, 0.9), 3),
                          "p50_ms_wall_incl_d2h": round(q(wall, 0.5), 3), "p90_ms_wall_incl_d2h": round(q(wall, 0.9), 3),
                          "xRT": round(80.0 / p50, 1), "algorithmic_GBps": round(alg / p50 / 1e6, 1),
                          "frac_of_hbm_peak": round(alg / p50 / 1e6 / PEAK, 3)}), flush=True)


def config4(sessions: int):
    from moshi_b200.models import LMGen, loaders
    lm, cfg = _lm()
    mimi = loaders.get_mimi(None, device="cuda", num_codebooks=8)
    free, _ = torch.cuda.mem_get_info()
    cap = int((free - 6e9) // (524288 * cfg.context + 40e6))
    B = min(sessions or cap, cap)
    g = torch.Generator().manual_seed(4242)
    pcm = (0.1 * torch.randn(B, 1, 1920, generator=g)).cuda()
    gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7)
    with mimi.streaming(B), gen.streaming(B), torch.no_grad():
        gen.assume_fill(cfg.context)
        for masked in (0.0, 0.25):
            m = torch.ones(B, dtype=torch.bool)
            m[: int(B * masked)] = False
            mimi.set_exec_mask(m)
            gen.set_exec_mask(m)
            times = []
            for i in range(13):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                codes = mimi.encode(pcm)
                toks = gen.step(codes)
                audio = toks[:, 1:].clamp(min=0) if toks is not None else torch.zeros(B, 8, 1, dtype=torch.int64, device="cuda")
                mimi.decode(audio)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    times.append(e0.elapsed_time(e1))
            ms = statistics.mean(times)
            print(json.dumps({"config": 4, "what": "one GPU's shard of the 512-session config (512 sessions need >= 5 GPUs at "
                              "full context: 1.573 GB of bf16 KV per session)", "sessions_on_this_gpu": B,
                              "max_admissible_sessions_per_gpu": cap, "rows_masked": masked, "kv_fill": cfg.context,
                              "ms_per_step": round(ms, 2), "p_max_ms": round(max(times), 2), "real_time": ms <= 80.0}), flush=True)


def config4_shards():
    """The 512-session configuration as written (512 / N sessions per GPU): which history the bf16 rings can hold when that many
    sessions share one GPU ("the fill level at which 512 fit"), with the rings sized to it (`kv_capacity`,
    reference numerics until a session is that old) — N = 2 (256 sessions per GPU) and N = 4 (128)."""
    from moshi_b200.models import LMGen, loaders
    lm, cfg = _lm()
    mimi = loaders.get_mimi(None, device="cuda", num_codebooks=8)
    free, _ = torch.cuda.mem_get_info()
    for n_gpus in (4, 2):
        B = 512 // n_gpus
        slots = int(((free - 6e9) / B - 40e6) // 524288)
        slots = min(slots, cfg.context)
        fill = slots - 64
        g = torch.Generator().manual_seed(4242)
        pcm = (0.1 * torch.randn(B, 1, 1920, generator=g)).cuda()
        gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7)
        gen.kv_capacity = slots if slots < cfg.context else None
        with mimi.streaming(B), gen.streaming(B), torch.no_grad():
            gen.assume_fill(fill)
            times = []
            for i in range(13):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                toks = gen.step(mimi.encode(pcm))
                audio = toks[:, 1:].clamp(min=0) if toks is not None else torch.zeros(B, 8, 1, dtype=torch.int64, device="cuda")
                mimi.decode(audio)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    times.append(e0.elapsed_time(e1))
            flags = gen.error_flags()
        ms = statistics.mean(times)
        print(json.dumps({"config": 4, "what": "512 sessions on %d GPUs: one GPU's shard with the bf16 rings sized to what fits" % n_gpus,
                          "n_gpus": n_gpus, "sessions_on_this_gpu": B, "kv_capacity_slots": slots,
                          "history_each_session_can_hold_s": round(slots * 0.08, 1), "kv_fill": fill,
                          "ms_per_step": round(ms, 2), "p_max_ms": round(max(times), 2), "real_time": ms <= 80.0,
                          "stepped_past_capacity": bool(flags & 4)}), flush=True)
        del gen
        torch.cuda.empty_cache()


def config5(kv_dtype: str):
    """int8 weights (row-wise absmax QLinear), 1 GPU: sweep the sessions upward until the step exceeds 80 ms or HBM is full."""
    from moshi_b200.config import MOSHI_7B
    from moshi_b200.models import LMGen, loaders
    kw = MOSHI_7B.to_reference_kwargs()
    kw["quantize"] = True
    lm = loaders.get_moshi_lm(None, kw, device="cuda", synth_device="cuda")
    mimi = loaders.get_mimi(None, device="cuda", num_codebooks=8)
    cfg = MOSHI_7B
    free, _ = torch.cuda.mem_get_info()
    kv_step = 524288 if kv_dtype == "bf16" else 32 * 2 * (4096 + 32 * 4)
    cap = min(int((free - 6e9) // (kv_step * cfg.context + 40e6)), 256)
    g = torch.Generator().manual_seed(4242)
    for B in sorted({b for b in (16, 32, 64, 96, 128, 160, 192, 224) if b < cap} | {cap}):
        pcm = (0.1 * torch.randn(B, 1, 1920, generator=g)).cuda()
        gen = LMGen(lm, use_sampling=True, temp=0.8, temp_text=0.7)
        gen.kv_dtype = kv_dtype
        with mimi.streaming(B), gen.streaming(B), torch.no_grad():
            gen.assume_fill(cfg.context)
            times = []
            for i in range(28):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                codes = mimi.encode(pcm)
                toks = gen.step(codes)
                audio = toks[:, 1:].clamp(min=0) if toks is not None else torch.zeros(B, 8, 1, dtype=torch.int64, device="cuda")
                mimi.decode(audio)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    times.append(e0.elapsed_time(e1))
        times.sort()
        print(json.dumps({"config": 5, "what": "Moshi 7B int8 linears (W8A8 QLinear) + Mimi, full 3000-frame rings", "kv_ring": kv_dtype,
                          "sessions": B, "max_sessions_hbm": cap, "ms_per_step_mean": round(statistics.mean(times), 2),
                          "ms_per_step_p99": round(times[-1], 2), "real_time": times[-1] <= 80.0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--sessions", type=int, default=0)
    ap.add_argument("--kv-dtype", default="bf16", choices=["bf16", "int8", "fp8_e4m3"])
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.config == 2:
        config2([int(v) for v in args.batches.split(",")])
    elif args.config == 3:
        config3()
    elif args.config == 4:
        if args.sessions == -512:
            config4_shards()
        else:
            config4(args.sessions)
    elif args.config == 5:
        config5(args.kv_dtype)


if __name__ == "__main__":
    main()
