"""GPU: ``DialogueService.step`` / ``b200_frame_step`` (host buffers in and out, one wait per frame) against the same frame
driven call by call through ``MimiModel.encode`` -> ``LMGen.step`` -> ``MimiModel.decode`` on twin models — the loop of
``server.py:120-147`` with the slot bookkeeping of ``batched_asr.py:138-215``.  Bit-exact: both paths launch the same
kernels in the same order on the same inputs."""
import numpy as np
import pytest
import torch

from moshi_b200.config import MimiConfig, tiny_lm_config
from moshi_b200.synth import synth_lm_state_dict, synth_mimi_state_dict
from oracle import scenarios

pytestmark = pytest.mark.gpu


def _models():
    from moshi_b200.models import LMModel, MimiModel
    cfg = tiny_lm_config(card=2048)             # the LM reads and writes the codec's 2048-entry codebooks
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    msd = synth_mimi_state_dict(MimiConfig(), seed=scenarios.MIMI_SEED)
    return cfg, LMModel(cfg, sd, device="cuda"), MimiModel(MimiConfig(), msd, device="cuda")


@pytest.mark.parametrize("sampling", [False, True])
def test_frame_service_equals_call_by_call(sampling):
    from moshi_b200.models import LMGen
    from moshi_b200.serving import ACTIVE, NODATA, RESET, DialogueService
    B, steps = 4, 12
    cfg, lm_a, mimi_a = _models()
    _, lm_b, mimi_b = _models()
    svc = DialogueService(B, lm_a, mimi_a, use_sampling=sampling)
    gen = LMGen(lm_b, use_sampling=sampling, support_out_of_sync=True)
    gen.streaming_forever(B)
    mimi_b.streaming_forever(B)
    g = torch.Generator().manual_seed(7)
    active = torch.ones(B, dtype=torch.bool)
    npr = svc._lib.b200_lm_noise_per_row(lm_a._h)
    produced = 0
    for i in range(steps):
        pcm = 0.1 * torch.randn(B, 1, 1920, generator=g)
        noise = torch.empty(B, npr).exponential_(1, generator=g) if sampling else None
        updates = [9] * B                      # a positive marker: slot keeps its state (batched_asr.py:161-170)
        if i == 0:
            updates = [ACTIVE] * B
        if i == 3:
            updates[1] = NODATA
        if i == 6:
            updates[1] = ACTIVE
            updates[2] = RESET
        reset = torch.tensor([u == RESET for u in updates])
        for b, u in enumerate(updates):
            if u == NODATA:
                active[b] = False
            elif u in (ACTIVE, RESET):
                active[b] = True
        # ---- service
        pcm_out = np.zeros((B, 1920), dtype=np.float32)
        tok_out = np.zeros((B, cfg.dep_q + 1), dtype=np.int64)
        flags = np.zeros(B, dtype=np.uint8)
        svc.step(pcm.reshape(-1).numpy().copy(), pcm_out, tok_out, updates=updates, flags_out=flags,
                 noise=None if noise is None else noise.numpy().copy())
        # ---- call by call (what server.py / batched_asr.py do around the same handles)
        if reset.any():
            gen.reset_streaming(reset.cuda())
            mimi_b.reset_streaming(reset.cuda())
        gen.set_exec_mask(active.cuda())
        mimi_b.set_exec_mask(active.cuda())
        codes = mimi_b.encode(pcm.cuda())
        toks = gen.step(codes, noise=None if noise is None else noise.cuda())
        assert toks is not None
        ready = active.cuda() & (toks[:, :, 0] >= 0).all(dim=1)
        mimi_b.set_exec_mask(ready)
        want_pcm = mimi_b.decode(toks[:, 1:].clamp(min=0) * ready[:, None, None])
        ready = ready.cpu()
        assert np.array_equal(flags.astype(bool), ready.numpy()), (i, flags, ready)
        want_tok = toks[:, :, 0].cpu()
        for b in range(B):
            if ready[b]:
                assert np.array_equal(tok_out[b], want_tok[b].numpy()), (i, b)
                assert np.array_equal(pcm_out[b], want_pcm[b, 0].cpu().numpy()), (i, b)
                produced += 1
            else:
                assert (pcm_out[b] == 0).all()
                assert (tok_out[b] < 0).any() or not active[b]
    assert produced >= B * (steps - 4)
    # warm-up rows: a reset slot is silent for max_delay frames again (lm.py:779-782)
    svc.close()


def test_frame_service_rejects_bad_buffers():
    from moshi_b200.serving import DialogueService
    cfg, lm, mimi = _models()
    svc = DialogueService(2, lm, mimi, use_sampling=False)
    good = np.zeros(2 * 1920, dtype=np.float32)
    with pytest.raises(AssertionError):
        svc.step(good[:100], np.zeros((2, 1920), np.float32), np.zeros((2, 9), np.int64))
    with pytest.raises(ValueError):
        svc.step(good, np.zeros((2, 1920), np.float32), np.zeros((2, 9), np.int64), updates=[-7, 0])
    svc.close()
