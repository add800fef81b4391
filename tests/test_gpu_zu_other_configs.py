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


# ---------------------------------------------------------------------------------------------------------------------------
# shared teacher-synchronised runner: the oracle is kept on the GPU's token trajectory, rows that have not left the
# reference's trajectory must reproduce the reference-recorded fixture exactly
# ---------------------------------------------------------------------------------------------------------------------------
LOGIT_ATOL = 0.08


def _run_greedy(cfg, sd, gold, steps, B, codes, gen_kw=None, orc_kw=None, reset_at=None, reset_mask=None):
    from moshi_b200.models import LMGen, LMModel
    lm = LMModel(cfg, sd, device="cuda")
    gen = LMGen(lm, use_sampling=False, **(gen_kw or {}))
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, tie_break="index", **(orc_kw or {}))
    orc.streaming(B)
    on_ref = torch.ones(B, dtype=torch.bool)
    agree = total = gold_rows = 0
    worst = 0.0
    with gen.streaming(B):
        for i in range(steps):
            if reset_at is not None and i == reset_at:
                gen.reset_streaming(reset_mask)
                orc.reset_streaming(reset_mask)
                on_ref |= reset_mask                 # a recycled row starts a fresh trajectory (the fixture's row does too)
            dbg = {}
            want = orc.step(codes[i], None, None, debug=dbg)
            got = gen.step(codes[i].cuda())
            assert (got is None) == (want is None), i
            tl = gen.read_buffer("text_logits", torch.bfloat16, (B, cfg.text_card)).float().cpu()
            worst = max(worst, (tl - dbg["text_logits"].float()[:, 0, 0]).abs().max().item())
            tt = gen.read_buffer("text_token", torch.int64, (B,)).cpu()
            at = gen.read_buffer("audio_tokens", torch.int64, (cfg.dep_q, B)).cpu()
            same_text = tt == dbg["text_token"]
            if same_text.any():
                dl = gen.read_buffer("dep_logits", torch.bfloat16, (cfg.dep_q, B, cfg.card)).float().cpu()
                worst = max(worst, (dl[0] - dbg["dep_logits"][0].float()[:, 0, 0])[same_text].abs().max().item())
            on_ref &= same_text & (at.t() == dbg["audio_tokens"]).all(dim=1)
            if got is not None:
                agree += int((got.cpu() == want).sum())
                total += want.numel()
                if gold is not None:
                    assert torch.equal(got.cpu()[on_ref], gold[i][on_ref]), i
                    gold_rows += int(on_ref.sum())
            pos = (orc.offsets % orc.cache.shape[2])
            for b in range(B):
                orc.cache[b, 0, pos[b]] = tt[b]
                orc.cache[b, 1:cfg.dep_q + 1, pos[b]] = at[:, b]
    del gen, lm
    return agree, total, worst, gold_rows


@pytest.mark.parametrize("mode", list(scenarios.CFG_MODES))
def test_classifier_free_guidance_matches_oracle_and_reference(golden_dir, mode):
    """``LMGen(cfg_coef, cfg_is_no_text, cfg_is_masked_until)`` (lm.py:596-604, 646-662, 714-732, 820-833): the model runs on
    2B rows, the guided logits feed both samplers; against the oracle and the fixture recorded from the reference."""
    from moshi_b200.config import tiny_lm_config
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    gold = load_file(golden_dir / "lm_tiny_cfg.safetensors")[mode]
    B, steps = scenarios.CFG_B, scenarios.CFG_STEPS
    codes = scenarios.lm_input_codes(cfg, B, steps, seed=scenarios.CFG_SEED)
    kw = scenarios.CFG_MODES[mode]
    agree, total, worst, gold_rows = _run_greedy(cfg, sd, gold, steps, B, codes, gen_kw=kw, orc_kw=kw,
                                                 reset_at=scenarios.CFG_RESET_STEP, reset_mask=torch.tensor([True, False]))
    print(f"cfg[{mode}]: tokens equal to oracle {agree}/{total}, worst guided-logit diff {worst:.3e}, fixture rows checked {gold_rows}")
    # guided logits are null + (cond - null) * coef: the bf16 noise of two model rows, amplified by coef (<= 3)
    assert worst < LOGIT_ATOL * 2 * max(kw["cfg_coef"], 1.0)
    assert agree / total > 0.9 and gold_rows > 0


@pytest.mark.parametrize("mode", ["sum", "sum_cfg"])
def test_condition_sum_matches_oracle_and_reference(golden_dir, mode):
    """The 2B configuration's conditioning (LUT conditioner fused by sum, lm.py:398-399, 616-626), with and without CFG:
    the conditioner runs in ``moshi_b200.conditioners`` (pinned on the CPU), its sum is added by the embedding kernel."""
    from moshi_b200.conditioners import ConditionAttributes
    from moshi_b200.config import tiny_lm_config
    from moshi_b200.models import LMGen, LMModel
    from oracle.gen_golden_cond import COND_CFG
    gold = load_file(golden_dir / "lm_tiny_cond.safetensors")
    cfg = tiny_lm_config(conditioners=COND_CFG["conditioners"], fuser=COND_CFG["fuser"])
    sd = synth_lm_state_dict(tiny_lm_config(), seed=scenarios.LM_SEED)
    sd.update({k: v for k, v in gold.items() if k.startswith("condition_provider.")})
    coef = 2.0 if mode == "sum_cfg" else 1.0
    B, steps = scenarios.CFG_B, scenarios.CFG_STEPS
    codes = scenarios.lm_input_codes(cfg, B, steps, seed=scenarios.CFG_SEED)
    lm = LMModel(cfg, sd, device="cuda")
    assert lm.condition_provider is not None and lm.fuser is not None
    texts = ["very_good"] * B + (["very_bad"] * B if coef != 1.0 else [])
    ct = lm.condition_provider(lm.condition_provider.prepare([ConditionAttributes(text={"description": t}) for t in texts]))
    assert torch.equal(lm.fuser.get_sum(ct).to(torch.bfloat16).cpu(), gold[mode + ".condition_sum"])
    gen = LMGen(lm, use_sampling=False, cfg_coef=coef, condition_tensors=ct)
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, tie_break="index", cfg_coef=coef,
                   condition_sum=gold[mode + ".condition_sum"])
    orc.streaming(B)
    on_ref = torch.ones(B, dtype=torch.bool)
    agree = total = checked = 0
    with gen.streaming(B):
        for i in range(steps):
            dbg = {}
            want = orc.step(codes[i], None, None, debug=dbg)
            got = gen.step(codes[i].cuda())
            assert (got is None) == (want is None), i
            tt = gen.read_buffer("text_token", torch.int64, (B,)).cpu()
            at = gen.read_buffer("audio_tokens", torch.int64, (cfg.dep_q, B)).cpu()
            on_ref &= (tt == dbg["text_token"]) & (at.t() == dbg["audio_tokens"]).all(dim=1)
            if got is not None:
                agree += int((got.cpu() == want).sum())
                total += want.numel()
                assert torch.equal(got.cpu()[on_ref], gold[mode + ".tokens"][i][on_ref]), i
                checked += int(on_ref.sum())
            pos = (orc.offsets % orc.cache.shape[2])
            for b in range(B):
                orc.cache[b, 0, pos[b]] = tt[b]
                orc.cache[b, 1:cfg.dep_q + 1, pos[b]] = at[:, b]
    print(f"condition_sum[{mode}]: tokens equal to oracle {agree}/{total}, fixture rows checked {checked}")
    assert agree / total > 0.9 and checked > 0
    with pytest.raises(AssertionError):                     # lm.py:602-603: CFG without masks needs condition tensors
        LMGen(lm, cfg_coef=2.0)


def test_stt_configuration_without_depformer_and_with_extra_heads(golden_dir):
    """STT-style member of the family: ``dep_q = 0`` (no depformer, every codebook comes from the user, lm.py:219-222) and
    ``extra_heads`` read by ``LMGen.step_with_extra_heads`` (lm.py:224-226, 793-807), with a slot recycled mid-stream like
    ``ASRService.step`` does (batched_asr.py:154-158); against the oracle and the fixture recorded from the reference."""
    from moshi_b200.config import LMConfig
    from moshi_b200.models import LMGen, LMModel
    gold = load_file(golden_dir / "lm_stt_tiny.safetensors")
    cfg = LMConfig.from_dict(scenarios.stt_reference_kwargs())
    sd = scenarios.stt_state_dict()
    lm = LMModel(cfg, sd, device="cuda")
    assert lm.depformer is None and len(lm.extra_heads) == 2
    B, steps = scenarios.STT_B, scenarios.STT_STEPS
    codes = scenarios.stt_input_codes()
    gen = LMGen(lm, use_sampling=False, temp=0.0, temp_text=0.0)
    orc = LMOracle(sd, scenarios.stt_spec(), use_sampling=False, tie_break="index")
    orc.streaming(B)
    on_ref = torch.ones(B, dtype=torch.bool)
    agree = total = checked = 0
    worst_head = worst_logit = 0.0
    with gen.streaming(B):
        for i in range(steps):
            if i == 6:
                r = torch.tensor([False, True])
                gen.reset_streaming(r)
                orc.reset_streaming(r)
                on_ref |= r
            dbg = {}
            want = orc.step_with_extra_heads(codes[i], debug=dbg)
            got = gen.step_with_extra_heads(codes[i].cuda())
            assert (got is None) == (want is None), i
            tl = gen.read_buffer("text_logits", torch.bfloat16, (B, cfg.text_card)).float().cpu()
            worst_logit = max(worst_logit, (tl - dbg["text_logits"].float()[:, 0, 0]).abs().max().item())
            tt = gen.read_buffer("text_token", torch.int64, (B,)).cpu()
            on_ref &= tt == dbg["text_token"]
            if got is not None:
                toks, heads = got
                assert toks.shape == (B, 1, 1) and len(heads) == 2 and heads[0].shape == (B, 1, 6)
                agree += int((toks.cpu() == want[0]).sum())
                total += B
                assert torch.equal(toks.cpu()[on_ref], gold["tokens"][i][on_ref]), i
                checked += int(on_ref.sum())
                hs = torch.stack([h[:, 0].float().cpu() for h in heads])             # [2, B, 6]
                worst_head = max(worst_head, (hs - torch.stack([w[:, 0].float() for w in want[1]])).abs().max().item())
                if on_ref.all():
                    # softmax probabilities in bf16: one ulp at p ~ 0.2 is 1e-3; the temporal output feeding the heads carries
                    # bf16 accumulation-order noise
                    torch.testing.assert_close(hs, gold["extra_heads"][i], rtol=0, atol=8e-3)
            pos = (orc.offsets % orc.cache.shape[2])
            for b in range(B):
                orc.cache[b, 0, pos[b]] = tt[b]
    print(f"stt: text tokens equal to oracle {agree}/{total}, worst text-logit diff {worst_logit:.3e}, worst extra-head "
          f"probability diff {worst_head:.3e}, fixture rows checked {checked}")
    assert worst_logit < LOGIT_ATOL and worst_head < 8e-3
    assert agree / total > 0.9 and checked > 0


def test_2b_shape_family_member_against_the_oracle():
    """``configs/moshi_dev_2b.json``'s shape of the step on a tiny member of the family: 32 codebooks of which 16 are
    generated, acoustic delay 2, RoPE period 100000: exercises 33-wide token rings, 16 depformer sub-steps and 16 keys in the
    depformer attention (the 7B model stops at 8)."""
    from moshi_b200.config import tiny_lm_config
    delays = [0, 0] + [2] * 15 + [0] + [2] * 15
    cfg = tiny_lm_config(n_q=32, dep_q=16, delays=delays, depformer_context=16, max_period=100000.0)
    sd = synth_lm_state_dict(cfg, seed=77)
    B, steps = 3, 12
    codes = scenarios.lm_input_codes(cfg, B, steps, seed=5)
    agree, total, worst, _ = _run_greedy(cfg, sd, None, steps, B, codes)
    print(f"2B-shaped tiny model: tokens equal to oracle {agree}/{total}, worst logit diff {worst:.3e}")
    assert worst < LOGIT_ATOL and agree / total > 0.95
    # batch >= 3 runs the fused depformer kernel; B = 1 takes the GEMV launch chain
    agree, total, worst, _ = _run_greedy(cfg, sd, None, steps, 1, codes[:, :1])
    assert worst < LOGIT_ATOL and agree / total > 0.95


def test_out_of_range_tokens_raise_a_flag_instead_of_reading_out_of_bounds():
    """token ids outside the embedding tables (the reference would hit a device assert in F.embedding) embed as
    the zero row and raise the device error flag; ``check=True`` surfaces it like lm.py:704-711."""
    from moshi_b200.config import tiny_lm_config
    from moshi_b200.models import LMGen, LMModel
    cfg = tiny_lm_config()
    lm = LMModel(cfg, synth_lm_state_dict(cfg, seed=1), device="cuda")
    gen = LMGen(lm, use_sampling=False)
    with gen.streaming(2):
        ok = torch.zeros(2, 8, 1, dtype=torch.long, device="cuda")
        gen.step(ok)
        gen.step(ok)
        assert gen.error_flags() == 0
        bad = ok.clone()
        bad[1, 3, 0] = cfg.card + 7                       # past the table
        gen.step(bad)
        gen.step(ok)                                      # delay 1: the bad code reaches the model on the next step at the latest
        assert gen.error_flags() == 1
        assert gen.error_flags() == 0                     # cleared by the read
    chk = LMGen(lm, use_sampling=False, check=True)
    with chk.streaming(2):
        chk.step(ok)
        with pytest.raises(AssertionError):
            chk.step(bad)
            chk.step(ok)
