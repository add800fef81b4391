"""GPU: other members of the model family behind the same kernels."""
import pytest
import torch
from safetensors.torch import load_file

from moshi_b200.synth import synth_lm_state_dict
from oracle import scenarios
from oracle.lm import LMOracle, LMSpec

pytestmark = pytest.mark.gpu


def test_delay2_pattern_matches_oracle_and_reference(golden_dir):
    """The 2B configuration's delay pattern (acoustic delay 2, ``configs/moshi_dev_2b.json``) on a tiny member of the family:
    ``None`` for the first two steps, then greedy tokens against the oracle (teacher-synchronised) and the reference fixture."""
    from moshi_b200.models import LMGen, LMModel
    cfg = scenarios.delay2_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.DELAY2_SEED)
    gold = load_file(golden_dir / "lm_tiny_delay2.safetensors")["tokens"]
    lm = LMModel(cfg, sd, device="cuda")
    B, steps = scenarios.DELAY2_B, scenarios.DELAY2_STEPS
    codes = scenarios.lm_input_codes(cfg, B, steps, seed=scenarios.DELAY2_SEED)
    gen = LMGen(lm, use_sampling=False)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, tie_break="index")
    orc.streaming(B)
    on_ref = torch.ones(B, dtype=torch.bool)
    agree = total = 0
    with gen.streaming(B):
        for i in range(steps):
            dbg = {}
            want = orc.step(codes[i], None, None, debug=dbg)
            got = gen.step(codes[i].cuda())
            assert (got is None) == (want is None) == (i < cfg.max_delay), i
            tt = gen.read_buffer("text_token", torch.int64, (B,)).cpu()
            at = gen.read_buffer("audio_tokens", torch.int64, (cfg.dep_q, B)).cpu()
            on_ref &= (tt == dbg["text_token"]) & (at.t() == dbg["audio_tokens"]).all(dim=1)
            if got is not None:
                agree += int((got.cpu() == want).sum())
                total += want.numel()
                assert torch.equal(got.cpu()[on_ref], gold[i][on_ref]), i       # rows still on the reference trajectory
            pos = (orc.offsets % orc.cache.shape[2])
            for b in range(B):
                orc.cache[b, 0, pos[b]] = tt[b]
                orc.cache[b, 1:cfg.dep_q + 1, pos[b]] = at[:, b]
    assert agree / total > 0.95
