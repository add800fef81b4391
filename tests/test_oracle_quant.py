"""CPU: the int8 QLinear restatement (oracle/quant.py) — quantiser properties and faithfulness to the float linear."""
import torch

from oracle import quant


def test_rowwise_quantiser_properties():
    g = torch.Generator().manual_seed(0)
    t = torch.randn(6, 320, generator=g)
    t[2] = 0.0
    q, s = quant.quantize_rows(t)
    assert q.dtype == torch.int8 and s.shape == (6,)
    assert int(q.abs().max()) == 127 and (q[2] == 0).all() and s[2] == 0
    # every row attains +-127 at its absmax element, and dequantisation is within half a step
    for r in (0, 1, 3, 4, 5):
        assert int(q[r].abs().max()) == 127
        assert ((q[r].float() * s[r] / 127) - t[r]).abs().max() <= s[r] / 254 * 1.0001
    # ties go to even (torch.round): 0.5 * step -> 0, 1.5 * step -> 2
    row = torch.tensor([[127.0, 0.5, 1.5, -0.5, -2.5]])
    assert quant.quantize_rows(row)[0].tolist() == [[127, 0, 2, 0, -2]]


def test_qlinear_is_a_faithful_linear_and_exact_in_integers():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 512, generator=g).bfloat16()
    w = (torch.randn(64, 512, generator=g) / 512 ** 0.5).bfloat16()
    y = quant.qlinear_f32(x, w)
    ref = x.float() @ w.float().t()
    assert (y - ref).abs().max() < 0.03 * ref.abs().max()
    qw, sw = quant.quantize_weight(w)
    qx, sa = quant.quantize_rows(x)
    acc = (qx.long() @ qw.long().t())
    assert torch.equal(acc.double(), qx.double() @ qw.double().t())
    assert torch.equal(y, acc.float() * ((sa[:, None] * sw[None, :]) * torch.tensor(float(quant.INV_127_SQ))))
    assert quant.qlinear(x, w).dtype == torch.bfloat16


def test_quantized_lm_oracle_steps_close_to_bf16():
    """LMOracle(quantize=True) routes every linear through QLinear; its first-step logits stay close to the bf16 model's
    (later steps feed back each model's own greedy tokens and drift apart)."""
    from moshi_b200.config import tiny_lm_config
    from moshi_b200.synth import synth_lm_state_dict
    from oracle import scenarios
    from oracle.lm import LMOracle, LMSpec
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    codes = scenarios.lm_input_codes(cfg, 2, 3)
    logits = []
    for q in (False, True):
        orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, quantize=q)
        orc.streaming(2)
        dbg = {}
        orc.step(codes[0][:2], None, None, debug=dbg)
        logits.append(dbg["text_logits"].float())
    rel = (logits[0] - logits[1]).abs().max() / logits[0].abs().max()
    assert 0 < rel < 0.08, rel


def test_qlinear_known_answers(golden_dir):
    """The KAT SURVEY.md 8(c) asks for: ``oracle/quant.py`` against an independent implementation of the published QLinear
    algorithm (numpy float32 quantiser with half-to-even rounding, exact Python-integer products, float64 dequantisation;
    ``oracle/gen_qlinear_kat.py``).  It pins the restatement, not bitsandbytes itself (absent: parity stays unpinned)."""
    import json
    kat = json.loads((golden_dir / "qlinear_kat.json").read_text())
    x = torch.tensor(kat["x"], dtype=torch.float32).bfloat16()
    w = torch.tensor(kat["w"], dtype=torch.float32).bfloat16()
    assert torch.equal(x.float(), torch.tensor(kat["x"])) and torch.equal(w.float(), torch.tensor(kat["w"]))   # bf16-exact inputs
    qx, sa = quant.quantize_rows(x)
    qw, sw = quant.quantize_weight(w)
    assert qx.tolist() == kat["qx"] and qw.tolist() == kat["qw"]
    assert qx[2][:6].tolist() == [127, 0, 2, 0, -2, 64]              # ties to even
    assert torch.equal(sa, torch.tensor(kat["sa"])) and torch.equal(sw, torch.tensor(kat["sw"]))
    acc = qx.long() @ qw.long().t()
    assert acc.tolist() == kat["acc"]
    y = quant.qlinear_f32(x, w)
    want = torch.tensor(kat["y_float64"], dtype=torch.float64)
    assert (y.double() - want).abs().max() <= 2.0 ** -22 * want.abs().max()      # fp32 rounding of two multiplies
    assert (y[1] == 0).all() and (y[:, 4] == 0).all()                # zero activation row, zero weight row
