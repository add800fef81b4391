"""CPU: configuration records, checkpoint-key normalisation, synthetic checkpoints, session sharding."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from moshi_b200.config import LMConfig, MimiConfig, MOSHI_7B, tiny_lm_config
from moshi_b200.models.state_dict import normalize_lm_state_dict, normalize_mimi_state_dict
from moshi_b200.serving import shard_sessions
from moshi_b200.synth import lm_tensor_specs, seanet_layout, synth_lm_state_dict, synth_mimi_state_dict

ROOT = Path(__file__).resolve().parent.parent


def test_7b_config_matches_reference_json():
    # values of configs/moshi_7b_202409.json
    c = MOSHI_7B
    assert (c.dim, c.num_layers, c.num_heads, c.context) == (4096, 32, 32, 3000)
    assert c.ffn_hidden == 11264 and c.depformer_ffn_hidden == 2816
    assert c.max_delay == 1 and c.num_codebooks == 17
    total = sum(int(torch.tensor(shape).prod()) for _, shape, _ in lm_tensor_specs(c))
    assert abs(total / 1e9 - 7.688) < 0.01            # 7.688 B parameters


def test_lm_config_rejects_options_outside_hot_path():
    d = MOSHI_7B.to_reference_kwargs()
    d["depformer_causal"] = True                       # deprecated key is accepted and dropped
    LMConfig.from_dict(d)
    with pytest.raises(ValueError):
        LMConfig.from_dict({**d, "norm": "layer_norm"})
    with pytest.raises(ValueError):
        LMConfig.from_dict({**d, "conditioners": {}})


def test_mimi_config_roundtrip_and_layout():
    cfg = MimiConfig()
    assert cfg.frame_size == 1920 and cfg.hop_length == 960 and cfg.resample_stride == 2
    assert MimiConfig.from_reference_dict(cfg.to_reference_dict()) == cfg
    enc, dec = seanet_layout(cfg)
    assert [l[1] for l in enc if l[0] == "conv"] == [0, 3, 6, 9, 12, 14]
    assert [l[1] for l in dec if l[0] == "convtr"] == [2, 5, 8, 11]
    with pytest.raises(ValueError):
        bad = cfg.to_reference_dict()
        bad["seanet"]["pad_mode"] = "reflect"
        MimiConfig.from_reference_dict(bad)


def test_synth_checkpoints_have_reference_keys():
    sd = synth_mimi_state_dict(MimiConfig(), seed=1)
    assert sd["encoder.model.12.conv.conv.weight"].shape == (1024, 512, 16)
    assert sd["decoder.model.2.convtr.convtr.weight"].shape == (1024, 512, 16)
    assert sd["quantizer.rvq_rest.vq.layers.30._codebook.embedding_sum"].shape == (2048, 256)
    assert sd["upsample.convtr.convtr.convtr.weight"].shape == (512, 1, 4)
    assert len(sd) == 318
    a = synth_lm_state_dict(tiny_lm_config(), seed=7)
    b = synth_lm_state_dict(tiny_lm_config(), seed=7)
    assert all(torch.equal(a[k], b[k]) for k in a) and len(a) == len(lm_tensor_specs(tiny_lm_config()))


def test_legacy_checkpoint_names_are_normalised():
    w = torch.arange(2 * 3 * 4 * 4, dtype=torch.float32).view(2 * 12, 4)
    sd = normalize_lm_state_dict({"depformer.layers.0.self_attn.in_proj_weight": w,
                                  "depformer.layers.0.self_attn.out_proj.weight": torch.zeros(8, 4)})
    assert torch.equal(sd["depformer.layers.0.self_attn.in_projs.1.weight"], w[12:])
    assert sd["depformer.layers.0.self_attn.out_projs.1.weight"].shape == (4, 4)
    m = normalize_mimi_state_dict({"quantizer.rvq_first.vq.layers.0._codebook.embed_sum": torch.zeros(2, 2),
                                   "quantizer.rvq_first.vq.layers.0._codebook.cluster_size": torch.ones(2)})
    assert set(m) == {"quantizer.rvq_first.vq.layers.0._codebook.embedding_sum",
                      "quantizer.rvq_first.vq.layers.0._codebook.cluster_usage"}


def test_session_sharding_is_a_partition():
    for total in (1, 7, 512):
        for world in (1, 2, 4, 8):
            parts = [shard_sessions(total, world, r) for r in range(world)]
            flat = [s for p in parts for s in range(p.start, p.stop)]
            assert flat == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_replicas_over_gloo(tmp_path):
    """world_size 2 on CPU: each rank owns a disjoint shard of the sessions, the only exchange is the
    timing reduction bench.py performs (max over ranks) — same code path as the NCCL launch."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from moshi_b200.serving import shard_sessions, init_distributed, max_over_ranks, sum_over_ranks\n"
        "rank, world = init_distributed(backend='gloo')\n"
        "mine = shard_sessions(11, world, rank)\n"
        "t = max_over_ranks(float(rank + 1))\n"
        "n = sum_over_ranks(float(len(mine)))\n"
        "if rank == 0: print(json.dumps({'t': t, 'n': n, 'world': world}))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line) == {"t": 2.0, "n": 11.0, "world": 2}
