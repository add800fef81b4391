"""CPU: the margin-aware sampled-token comparison itself (tests/test_gpu_lm.py::_peaked_sampling_run), with a second oracle
standing in for the device: identical logits, so every decision equals the teacher's, and every disagreement with the reference's
recorded decisions must be one the stability test flags (exact bf16 ties ranked by token id instead of torch.topk's order)."""
from safetensors.torch import load_file

from moshi_b200.config import tiny_lm_config
from oracle import scenarios


def test_gate_logic_with_an_oracle_as_the_device(golden_dir):
    from tests.test_gpu_lm import _OracleAsDevice, _peaked_sampling_run
    cfg = tiny_lm_config()
    sd = scenarios.peaked_state_dict(cfg)
    gold = load_file(golden_dir / "lm_tiny_sampled_peaked.safetensors")
    c = _peaked_sampling_run(_OracleAsDevice(sd, cfg, scenarios.LM_B), sd, cfg, gold)
    print(c)
    assert c["decisions"] == c["equal"] and c["unexcused"] == 0 and c["worst"] == 0.0
    assert c["ref_unexcused"] == 0 and c["ref_decisions"] > 30
    assert c["stable"] > 0.4 * c["decisions"]
