"""CPU: ``moshi_b200.conditioners`` (LUT conditioner + sum fuser, ``configs/moshi_dev_2b.json``) pinned on a fixture recorded
from the unmodified reference (``oracle/gen_golden_cond.py``: the conditioner's weights, its ``ConditionType`` output and
``fuser.get_sum`` for the conditioned and the CFG-doubled batch)."""
import pytest
import torch
from safetensors.torch import load_file

from moshi_b200.conditioners import ConditionAttributes, ConditionFuser, build_conditioning
from oracle.gen_golden_cond import COND_CFG


@pytest.fixture(scope="module")
def gold(golden_dir):
    return load_file(golden_dir / "lm_tiny_cond.safetensors")


def test_lut_conditioner_and_sum_fuser_match_the_reference(gold):
    provider, fuser = build_conditioning(COND_CFG["conditioners"], COND_CFG["fuser"], 256, gold, "cpu")
    assert provider.text_conditions == ["description"] and fuser.has_conditions
    B = 2
    for name, texts in (("sum", ["very_good"] * B), ("sum_cfg", ["very_good"] * B + ["very_bad"] * B)):
        ct = provider(provider.prepare([ConditionAttributes(text={"description": t}) for t in texts]))
        assert torch.equal(ct["description"].condition, gold[name + ".condition"])
        assert torch.equal(ct["description"].mask.to(torch.uint8), gold[name + ".mask"])
        assert torch.equal(fuser.get_sum(ct).to(torch.bfloat16), gold[name + ".condition_sum"])


def test_missing_attribute_uses_the_learnt_padding(gold):
    provider, _ = build_conditioning(COND_CFG["conditioners"], COND_CFG["fuser"], 256, gold, "cpu")
    ct = provider(provider.prepare([ConditionAttributes(text={"description": None})]))
    assert not ct["description"].mask.any()
    assert torch.equal(ct["description"].condition, gold["condition_provider.conditioners.description.learnt_padding"])


def test_errors_like_the_reference(gold):
    provider, _ = build_conditioning(COND_CFG["conditioners"], COND_CFG["fuser"], 256, gold, "cpu")
    with pytest.raises(ValueError):
        provider.prepare([ConditionAttributes(text={"description": "excellent"})])        # not in possible_values
    with pytest.raises(RuntimeError):
        provider.prepare([ConditionAttributes(text={})])                                  # conditioner without input
    with pytest.raises(ValueError):
        ConditionFuser({"cross": ["description"]})                                        # outside the step path
