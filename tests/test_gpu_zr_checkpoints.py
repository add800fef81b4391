"""GPU: on-disk checkpoints through the reference-shaped loaders (``loaders.get_mimi(filename)`` / ``get_moshi_lm(filename)``,
loaders.py:323-446): safetensors with the reference's key names incl. the legacy packed attention / codebook names, and the
pre-quantised ``model.q8.safetensors`` layout (``weight`` int8 + ``weight_scb`` float32, utils/quantize.py:13-22), and the Rust
stack's GGUF form (candle tensor names, Q8_0 blocks; rust/moshi-core/src/nn.rs:9-116)."""
import pytest
import torch
from safetensors.torch import save_file

from moshi_b200.config import MimiConfig, tiny_lm_config
from moshi_b200.synth import synth_lm_state_dict, synth_mimi_state_dict
from oracle import quant, scenarios

pytestmark = pytest.mark.gpu


def _legacy_lm_names(sd: dict) -> dict:
    """The names older checkpoints carry: one packed ``in_proj_weight`` / ``out_proj.weight`` per attention, all the depformer's
    per-step projections concatenated (split back by ``mult`` in the load hook, transformer.py:422-446)."""
    out, groups = {}, {}
    for k, v in sd.items():
        if ".self_attn.in_projs." in k or ".self_attn.out_projs." in k:
            base, idx = k.rsplit(".", 2)[0], int(k.rsplit(".", 2)[1])
            groups.setdefault(base, {})[idx] = v
        else:
            out[k] = v
    for base, parts in groups.items():
        cat = torch.cat([parts[i] for i in range(len(parts))], dim=0)
        if base.endswith("in_projs"):
            out[base[:-len("in_projs")] + "in_proj_weight"] = cat
        else:
            out[base[:-len("out_projs")] + "out_proj.weight"] = cat
    return out


def _steps(lm, cfg, n=4, B=3):
    from moshi_b200.models import LMGen
    gen = LMGen(lm, use_sampling=False)
    codes = scenarios.lm_input_codes(cfg, B, n).cuda()
    logits, toks = [], []
    with gen.streaming(B):
        for i in range(n):
            gen.step(codes[i])
            logits.append(gen.read_buffer("text_logits", torch.bfloat16, (B, cfg.text_card)).cpu())
            toks.append(gen.read_buffer("audio_tokens", torch.int64, (cfg.dep_q, B)).cpu())
    return logits, toks


def test_lm_safetensors_round_trip_with_packed_attention_names(tmp_path):
    from moshi_b200.models import LMModel, loaders
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    path = tmp_path / "model.safetensors"
    save_file({k: v.contiguous() for k, v in _legacy_lm_names(sd).items()}, str(path))
    assert any(k.endswith("in_proj_weight") for k in _legacy_lm_names(sd))
    from_disk = loaders.get_moshi_lm(path, cfg.to_reference_kwargs(), device="cuda")
    direct = LMModel(cfg, sd, device="cuda")
    a, b = _steps(from_disk, cfg), _steps(direct, cfg)
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert torch.equal(x, y)


def test_lm_q8_checkpoint_loads_without_requantising(tmp_path):
    """``model.q8.safetensors``: every nn.Linear as int8 ``weight`` + float32 ``weight_scb`` (quantised here on the CPU by the
    oracle's restatement of bitsandbytes' row-wise quantiser, exactly what ``QLinear.__init__`` stores), everything else bf16.
    The library tiles the int8 weights as they are: the model must be bit-identical to ``LMModel(quantize=True)`` built from
    the bf16 weights (whose in-library quantiser is bit-exact with the same restatement, tests/test_gpu_zx_int8.py)."""
    from moshi_b200.models import LMModel, loaders
    cfg = tiny_lm_config(quantize=True)
    sd = synth_lm_state_dict(tiny_lm_config(), seed=scenarios.LM_SEED)
    q8 = {}
    linear = lambda k: (k.endswith(".weight") and not k.startswith(("emb.", "text_emb.", "depformer_emb.", "depformer_text_emb.")))
    for k, v in sd.items():
        if linear(k):
            cb, scb = quant.quantize_weight(v)
            q8[k], q8[k + "_scb"] = cb.contiguous(), scb.contiguous()
        else:
            q8[k] = v.contiguous()
    assert q8["transformer.layers.0.gating.linear_in.weight"].dtype == torch.int8
    path = tmp_path / "model.q8.safetensors"
    save_file(q8, str(path))
    from_disk = loaders.get_moshi_lm(path, cfg.to_reference_kwargs(), device="cuda")
    direct = LMModel(cfg, sd, device="cuda")
    a, b = _steps(from_disk, cfg), _steps(direct, cfg)
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert torch.equal(x, y)
    # a q8 checkpoint into a model built without quantize=True is an error, and so is a scale that lost its dtype
    with pytest.raises(ValueError):
        loaders.get_moshi_lm(path, tiny_lm_config().to_reference_kwargs(), device="cuda")
    bad = dict(q8)
    del bad["text_linear.weight_scb"]
    save_file(bad, str(tmp_path / "bad.safetensors"))
    with pytest.raises(RuntimeError):
        loaders.get_moshi_lm(tmp_path / "bad.safetensors", cfg.to_reference_kwargs(), device="cuda")


def test_mimi_safetensors_round_trip_with_legacy_names(tmp_path):
    """``get_mimi(filename)``: packed attention names and the legacy codebook buffer names (core_vq.py:162-176)."""
    from moshi_b200.models import MimiModel, loaders
    cfg = MimiConfig()
    sd = synth_mimi_state_dict(cfg, seed=scenarios.MIMI_SEED)
    legacy = {}
    for k, v in sd.items():
        k = k.replace("self_attn.in_projs.0.weight", "self_attn.in_proj_weight").replace("self_attn.out_projs.0.weight", "self_attn.out_proj.weight")
        k = k.replace("_codebook._initialized", "_codebook.inited").replace("_codebook.cluster_usage", "_codebook.cluster_size")
        k = k.replace("_codebook.embedding_sum", "_codebook.embed_sum")
        legacy[k] = v.contiguous()
    path = tmp_path / "tokenizer.safetensors"
    save_file(legacy, str(path))
    from_disk = loaders.get_mimi(path, device="cuda", num_codebooks=8)
    direct = MimiModel(cfg, sd, device="cuda")
    pcm = scenarios.mimi_noise(2, 3, seed=3).cuda()
    outs = []
    for m in (from_disk, direct):
        with m.streaming(2):
            codes = m.encode(pcm)
            outs.append((codes.cpu(), m.decode(codes).cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert from_disk.num_codebooks == 8 and from_disk.total_codebooks == 32


def test_lm_gguf_checkpoint_in_the_candle_layout(tmp_path):
    """A ``.gguf`` file as the Rust stack loads it: candle's tensor names (``scripts/import_rust.py``), every linear weight as
    ``Q8_0`` blocks, norms and embeddings bf16.  ``get_moshi_lm`` dequantises the blocks and maps the names back: the model must be
    bit-identical to one built directly from the dequantised weights under the reference's names."""
    from moshi_b200.models import LMModel, gguf, loaders
    from tests.util import to_candle_layout
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    candle = to_candle_layout(cfg, sd)
    q8 = lambda k, t: t.dim() == 2 and "emb" not in k and t.shape[-1] % 32 == 0       # every nn.Linear weight; norms are [1, 1, C]
    path = tmp_path / "model.q8_0.gguf"
    gguf.write_gguf(path, candle, q8_0=q8, metadata={"general.architecture": "moshi"})
    n_q8 = sum(1 for k, t in candle.items() if q8(k, t))
    assert n_q8 > 20
    # the same dequantised values under the reference's names
    deq = {}
    from moshi_b200.models.state_dict import normalize_lm_state_dict
    back = {k2: k for k in candle for k2 in normalize_lm_state_dict({k: candle[k]})}
    for name, t in sd.items():
        src = back[name]
        if q8(src, candle[src]):
            t = torch.from_numpy(gguf.dequantize_q8_0(gguf.quantize_q8_0(t.float().numpy()), t.shape)).bfloat16()
        deq[name] = t
    from_disk = loaders.get_moshi_lm(path, cfg.to_reference_kwargs(), device="cuda")
    direct = LMModel(cfg, deq, device="cuda")
    a, b = _steps(from_disk, cfg), _steps(direct, cfg)
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert torch.equal(x, y)
    # and close to the unquantised model (weight-only Q8_0: ~0.4 % relative error per weight)
    c = _steps(LMModel(cfg, sd, device="cuda"), cfg)
    assert max((x.float() - y.float()).abs().max().item() for x, y in zip(a[0], c[0])) < 0.25
