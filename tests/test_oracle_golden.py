"""CPU: the oracle reproduces the fixtures that oracle/gen_golden.py recorded from the unmodified
reference (bit-exact on the same torch build; tolerances only guard against a different CPU BLAS)."""
import json

import pytest
import torch
from safetensors.torch import load_file

from moshi_b200.config import MimiConfig, tiny_lm_config
from moshi_b200.synth import synth_lm_state_dict, synth_mimi_state_dict
from oracle import scenarios
from oracle.lm import LMOracle, LMSpec
from oracle.mimi import MimiOracle


@pytest.fixture(scope="module")
def mimi_sd():
    return synth_mimi_state_dict(MimiConfig(), seed=scenarios.MIMI_SEED)


def test_manifest_says_oracle_was_pinned(golden_dir):
    m = json.loads((golden_dir / "MANIFEST.json").read_text())
    assert m["mimi_sine"]["oracle_bit_exact"] and m["mimi_masked"]["oracle_bit_exact"]
    assert m["lm_tiny_sampled"]["oracle_bit_exact_tokens"] and m["lm_tiny_greedy"]["oracle_bit_exact_tokens"]


@torch.no_grad()
def test_mimi_sine_roundtrip(golden_dir, mimi_sd):
    gold = load_file(golden_dir / "mimi_sine.safetensors")
    cfg = MimiConfig()
    orc = MimiOracle(mimi_sd, cfg)
    orc.streaming(1)
    sine = scenarios.sine_1s()
    codes, pcm = [], []
    for f in range(sine.shape[-1] // cfg.frame_size):
        c = orc.encode(sine[..., f * 1920:(f + 1) * 1920])
        codes.append(c)
        pcm.append(orc.decode(c))
    codes, pcm = torch.cat(codes, -1), torch.cat(pcm, -1)
    assert codes.dtype == torch.int64 and codes.shape == (1, 8, 12)
    assert (codes == gold["codes_stream"]).all()
    torch.testing.assert_close(pcm, gold["pcm_stream"], rtol=0, atol=1e-5)
    # the reference's non-streaming call (13 padded frames) agrees with streaming on the common part
    assert (gold["codes_batch"][..., :12] == codes).all()


@torch.no_grad()
def test_mimi_masked_rows_and_reset(golden_dir, mimi_sd):
    gold = load_file(golden_dir / "mimi_masked.safetensors")
    B, frames = scenarios.MIMI_MASK_B, scenarios.MIMI_MASK_FRAMES
    pcm = scenarios.mimi_noise(B, frames)
    orc = MimiOracle(mimi_sd, MimiConfig())
    orc.streaming(B)
    for f in range(frames):
        scenarios.mimi_mask_events(orc, f, B)
        c = orc.encode(pcm[..., f * 1920:(f + 1) * 1920])
        assert (c == gold["codes"][f]).all(), f
        torch.testing.assert_close(orc.decode(c), gold["pcm"][f], rtol=0, atol=1e-5)


def test_mimi_rejects_partial_frames(mimi_sd):
    orc = MimiOracle(mimi_sd, MimiConfig())
    orc.streaming(1)
    with pytest.raises(RuntimeError):
        orc.encode(torch.zeros(1, 1, 1000))


@pytest.mark.parametrize("name,sampling", [("lm_tiny_sampled", True), ("lm_tiny_greedy", False)])
def test_lm_tiny_steps(golden_dir, name, sampling):
    gold = load_file(golden_dir / f"{name}.safetensors")
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    B, steps = scenarios.LM_B, scenarios.LM_STEPS
    codes = scenarios.lm_input_codes(cfg, B, steps)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=sampling)
    orc.streaming(B)
    torch.manual_seed(scenarios.LM_NOISE_SEED)
    for i in range(steps):
        scenarios.lm_mask_events(orc, i, B)
        nt, na = scenarios.lm_noise(cfg, B) if sampling else (None, None)
        dbg = {}
        out = orc.step(codes[i], nt, na, debug=dbg)
        want = gold["tokens"][i]
        if out is None:
            assert (want == -3).all(), i
        else:
            assert (out == want).all(), i
        torch.testing.assert_close(dbg["text_logits"].float()[:, 0, 0], gold["text_logits"][i], rtol=0, atol=0)


def test_step_outside_streaming_raises():
    cfg = tiny_lm_config()
    orc = LMOracle(synth_lm_state_dict(cfg), LMSpec.from_config(cfg))
    with pytest.raises(RuntimeError):
        orc.step(torch.zeros(1, 8, 1, dtype=torch.long))


def test_lm_stt_extra_heads(golden_dir):
    """STT-style member of the family (no depformer, ``extra_heads``; ``LMGen.step_with_extra_heads``, lm.py:793-807 — what
    ``rust/moshi-server/batched_asr.py:197`` calls): the oracle reproduces the fixture recorded from the unmodified reference
    bit for bit, including the recycled slot.  Groundwork for SURVEY.md 8(f) item 2; the CUDA path does not build it yet."""
    import json

    info = json.loads((golden_dir / "lm_stt_tiny.json").read_text())
    assert info["oracle_bit_exact"] is True
    gold = load_file(golden_dir / "lm_stt_tiny.safetensors")
    sd = scenarios.stt_state_dict()
    codes = scenarios.stt_input_codes()
    orc = LMOracle(sd, scenarios.stt_spec(), use_sampling=False)
    orc.streaming(scenarios.STT_B)
    for i in range(scenarios.STT_STEPS):
        if i == 6:
            orc.reset_streaming(torch.tensor([False, True]))
        got = orc.step_with_extra_heads(codes[i])
        if (gold["tokens"][i] == info["none_marker"]).all():
            assert got is None
            continue
        toks, heads = got
        assert torch.equal(toks, gold["tokens"][i])
        assert toks.shape == (scenarios.STT_B, 1, 1)                  # text stream only
        assert torch.equal(torch.stack([h[:, 0].float() for h in heads]), gold["extra_heads"][i])
        assert torch.allclose(torch.stack(heads).float().sum(-1), torch.ones(2, scenarios.STT_B, 1), atol=2e-2)


def test_lm_delay2_pattern(golden_dir):
    """The 2B configuration's delay pattern (acoustic streams delayed by up to 2 steps, ``configs/moshi_dev_2b.json``) on a
    tiny member of the family: ``None`` for the first max_delay steps, then the oracle reproduces the reference's tokens."""
    info = json.loads((golden_dir / "lm_tiny_delay2.json").read_text())
    assert info["oracle_bit_exact_tokens"] is True
    gold = load_file(golden_dir / "lm_tiny_delay2.safetensors")["tokens"]
    cfg = scenarios.delay2_config()
    assert cfg.max_delay == 2
    sd = synth_lm_state_dict(cfg, seed=scenarios.DELAY2_SEED)
    codes = scenarios.lm_input_codes(cfg, scenarios.DELAY2_B, scenarios.DELAY2_STEPS, seed=scenarios.DELAY2_SEED)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False)
    orc.streaming(scenarios.DELAY2_B)
    for i in range(scenarios.DELAY2_STEPS):
        out = orc.step(codes[i])
        if i < cfg.max_delay:
            assert out is None and (gold[i] == info["none_marker"]).all()
        else:
            assert torch.equal(out, gold[i]), i


@pytest.mark.parametrize("mode", ["no_text", "masked_until", "both"])
def test_lm_cfg_without_conditioner(golden_dir, mode):
    """Classifier-free guidance without a conditioner (``cfg_coef != 1`` with ``cfg_is_no_text`` / ``cfg_is_masked_until``,
    lm.py:596-604, 646-662, 714-732, 820-833): 2B model rows, guided text and depformer logits; the oracle reproduces the
    fixture recorded from the unmodified reference, including a slot reset.  Groundwork for SURVEY.md 8(f) item 2."""
    info = json.loads((golden_dir / "lm_tiny_cfg.json").read_text())
    assert info["modes"][mode]["oracle_bit_exact_tokens"] is True
    gold = load_file(golden_dir / "lm_tiny_cfg.safetensors")[mode]
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    codes = scenarios.lm_input_codes(cfg, scenarios.CFG_B, scenarios.CFG_STEPS, seed=scenarios.CFG_SEED)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, **scenarios.CFG_MODES[mode])
    orc.streaming(scenarios.CFG_B)
    for i in range(scenarios.CFG_STEPS):
        if i == scenarios.CFG_RESET_STEP:
            orc.reset_streaming(torch.tensor([True, False]))
        out = orc.step(codes[i])
        if (gold[i] == info["none_marker"]).all():
            assert out is None
        else:
            assert torch.equal(out, gold[i]), i


@pytest.mark.parametrize("mode", ["sum", "sum_cfg"])
def test_lm_condition_sum(golden_dir, mode):
    """The 2B configuration's conditioning (LUT conditioner fused by ``sum``, ``configs/moshi_dev_2b.json``): ``condition_sum``
    (computed once per session by the reference's conditioner, stored in the fixture) is added to the summed input embeddings
    every step (lm.py:398-399); with ``cfg_coef != 1`` the second half of the 2B model rows carries the null condition."""
    info = json.loads((golden_dir / "lm_tiny_cond.json").read_text())
    assert info["modes"][mode]["oracle_bit_exact_tokens"] is True
    gold = load_file(golden_dir / "lm_tiny_cond.safetensors")
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    codes = scenarios.lm_input_codes(cfg, scenarios.CFG_B, scenarios.CFG_STEPS, seed=scenarios.CFG_SEED)
    csum = gold[mode + ".condition_sum"]
    assert csum.shape == (scenarios.CFG_B * (2 if mode == "sum_cfg" else 1), 1, cfg.dim)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, cfg_coef=info["modes"][mode]["cfg_coef"], condition_sum=csum)
    orc.streaming(scenarios.CFG_B)
    for i in range(scenarios.CFG_STEPS):
        out = orc.step(codes[i])
        want = gold[mode + ".tokens"][i]
        if (want == info["none_marker"]).all():
            assert out is None
        else:
            assert torch.equal(out, want), i
