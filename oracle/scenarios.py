"""Seeded inputs and mask/reset schedules shared by the golden generator and the tests.

Any object with ``set_exec_mask`` / ``reset_streaming`` works (reference modules, the oracle, the
B200 shims): the same schedule drives all three.
"""
from __future__ import annotations

import math

import torch

MIMI_SEED = 1234
MIMI_MASK_B = 3
MIMI_MASK_FRAMES = 8
LM_SEED = 4242
LM_NOISE_SEED = 99
LM_B = 3
LM_STEPS = 30


def sine_1s() -> torch.Tensor:
    """BASELINE.json configs[0] / SURVEY.md 8(d) config 1: 0.5*sin(2*pi*440*t), 1 s at 24 kHz."""
    t = torch.arange(24000, dtype=torch.float32) / 24000
    x = 0.5 * torch.sin(2 * math.pi * 440 * t)
    return x[None, None, :23040].contiguous()


def sine_1s_full() -> torch.Tensor:
    t = torch.arange(24000, dtype=torch.float32) / 24000
    return (0.5 * torch.sin(2 * math.pi * 440 * t))[None, None]


def mimi_noise(batch: int, frames: int, seed: int = 4242) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(batch, 1, 1920 * frames, generator=g)


def mimi_mask_events(model, frame: int, batch: int) -> None:
    """Row 1 is paused for frames 3-4, row 2 is recycled (reset) before frame 6."""
    if frame == 3:
        m = torch.ones(batch, dtype=torch.bool)
        m[1] = False
        model.set_exec_mask(m)
    if frame == 5:
        model.set_exec_mask(torch.ones(batch, dtype=torch.bool))
    if frame == 6:
        r = torch.zeros(batch, dtype=torch.bool)
        r[2] = True
        model.reset_streaming(r)


def lm_input_codes(cfg, batch: int, steps: int, seed: int = 4242) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, cfg.card, (steps, batch, cfg.n_q - cfg.dep_q, 1), generator=g)


def lm_mask_events(gen, step: int, batch: int) -> None:
    if step == 10:
        m = torch.ones(batch, dtype=torch.bool)
        m[1] = False
        gen.set_exec_mask(m)
    if step == 12:
        gen.set_exec_mask(torch.ones(batch, dtype=torch.bool))
    if step == 20:
        r = torch.zeros(batch, dtype=torch.bool)
        r[1] = True
        gen.reset_streaming(r)


def lm_noise(cfg, batch: int, top_k_text: int = 25, top_k: int = 250):
    """The Exp(1) draws of one ``LMGen.step`` in the order and shapes the reference makes them
    (sampling.py:44 via lm.py:736 and lm.py:836): [B, k_text] then dep_q x [B, k]."""
    kt = min(top_k_text, cfg.text_card)
    ka = min(top_k, cfg.card)
    nt = torch.empty(batch, kt).exponential_(1)
    na = [torch.empty(batch, ka).exponential_(1) for _ in range(cfg.dep_q)]
    return nt, na
