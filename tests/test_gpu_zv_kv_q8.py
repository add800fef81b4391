"""GPU: the opt-in 8-bit temporal KV rings (one byte per element + one fp32 scale per key row): e4m3 and int8.

This is NOT the reference's numerics (its ring is bf16, transformer.py:196-288); it exists to halve the 1.57 GB a
session's rings take.  Two levels:
  * the fused attention step kernel against an fp32 emulation that stores ``e4m3(x * 448/absmax) * absmax/448`` (or
    ``round(x * 127/absmax) * absmax/127``) for the appended key / value row and attends over the dequantised ring: the
    kernel must agree with that to bf16 rounding, and the bytes + scales it wrote must be the emulation's;
  * the LM step against the oracle with the same emulation in its ring, and the *cost* of the option: logit deviation
    of the fp8-ring model from the bf16-ring model on the same token stream (reported and bounded).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle.transformer import kv_roundtrip
from tests.test_gpu_ops import _rope_ref
from tests.util import cptr

pytestmark = pytest.mark.gpu


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


FMT = {"fp8_e4m3": (1, 448.0), "int8": (2, 127.0)}


def _encode(t, amax, fmt):
    """Row-scaled bytes as the kernel stores them."""
    qmax = FMT[fmt][1]
    inv = torch.where(amax > 0, torch.full_like(amax, qmax) / amax, torch.zeros_like(amax))     # true division, like the kernel
    y = t.float() * inv[..., None]
    if fmt == "fp8_e4m3":
        return y.to(torch.float8_e4m3fn).view(torch.uint8)
    return (torch.round(y).clamp_(-127, 127) + 128).to(torch.uint8)


def _decode(u8, scale, fmt):
    if fmt == "fp8_e4m3":
        return u8.view(torch.float8_e4m3fn).float() * scale[..., None]
    return (u8.float() - 128.0) * scale[..., None]


@pytest.mark.parametrize("fmt", ["int8", "fp8_e4m3"])
@pytest.mark.parametrize("B,H,cap,nsplit,steps", [(3, 2, 12, 0, 30), (2, 32, 300, 4, 6), (5, 4, 3000, 0, 3), (1, 32, 3000, 16, 2)])
def test_q8_attention_step_matches_emulation(B, H, cap, nsplit, steps, fmt):
    from moshi_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(B * 17 + cap)
    Cd = H * 128
    # start from a ring that already holds `pos0` keys written by the kernel's own format
    pos0 = torch.tensor([0, cap - 2, cap + 7, 5, 2 * cap + 1][:B], dtype=torch.int64)
    hist_k = torch.randn(B, H, cap, 128, generator=g).bfloat16()
    hist_v = torch.randn(B, H, cap, 128, generator=g).bfloat16() * torch.rand(B, H, cap, 1, generator=g) * 4
    kind, qmax = FMT[fmt]
    ring_k, ring_v = kv_roundtrip(hist_k, fmt), kv_roundtrip(hist_v, fmt)          # dequantised fp32 view of the ring
    amax_k, amax_v = hist_k.float().abs().amax(-1), hist_v.float().abs().amax(-1)
    k8, v8 = _encode(hist_k, amax_k, fmt).cuda(), _encode(hist_v, amax_v, fmt).cuda()
    ks, vs = (amax_k * (1.0 / qmax)).cuda(), (amax_v * (1.0 / qmax)).cuda()
    assert torch.equal(_decode(k8.cpu(), ks.cpu(), fmt), ring_k)
    pos = pos0.clone()
    out = torch.empty(B, Cd, dtype=torch.bfloat16, device="cuda")
    worst = 0.0
    for i in range(steps):
        qkv = torch.randn(B, 3 * Cd, generator=g).bfloat16()
        mask = torch.ones(B, dtype=torch.bool)
        if B > 1 and i % 3 == 1:
            mask[1] = False
        q_in, k_in, v_in = (qkv[:, j * Cd:(j + 1) * Cd].reshape(B, H, 128) for j in range(3))
        q_rot, k_rot = _rope_ref(q_in, pos), _rope_ref(k_in, pos)
        k_new, v_new = kv_roundtrip(k_rot, fmt), kv_roundtrip(v_in, fmt)
        for b in range(B):
            if mask[b]:
                ring_k[b, :, pos[b] % cap] = k_new[b]
                ring_v[b, :, pos[b] % cap] = v_new[b]
        n_valid = (pos + mask.long()).clamp(max=cap)
        allowed = torch.arange(cap)[None, :] < n_valid[:, None]
        want = F.scaled_dot_product_attention(q_rot.float()[:, :, None], ring_k, ring_v, allowed[:, None, None, :])[:, :, 0].reshape(B, Cd)
        qd, pd, md = qkv.cuda(), pos.cuda(), mask.cuda()
        _lib.check(lib.b200_op_attn_step_q8(cptr(qd), cptr(k8), cptr(v8), cptr(ks), cptr(vs), cptr(out), cptr(pd), cptr(md),
                                            B, H, cap, nsplit, 10000.0, kind, _stream()))
        torch.cuda.synchronize()
        rows = n_valid > 0
        worst = max(worst, (out.float().cpu()[rows] - want[rows]).abs().max().item())
        # a bf16 ulp of sincos in the rotated key can move it (or its row's absmax) across a rounding boundary of the 8-bit
        # grid: rare entries differ by one quantisation step of a key the softmax weights heavily; the bulk agrees to
        # bf16 rounding
        err = (out.float().cpu()[rows] - want[rows]).abs()
        assert err.max() < 0.1 and err.mean() < 4e-3, (err.max(), err.mean())
        pos = pos + mask.long()
    print(f"{fmt} attn step B={B} H={H} cap={cap}: worst |d| vs emulation {worst:.3e}")
    # the ring the kernel maintained == the emulation's (keys: one grid step where sincos moved a bf16 ulp of the key)
    deq_k, deq_v = _decode(k8.cpu(), ks.cpu(), fmt), _decode(v8.cpu(), vs.cpu(), fmt)
    assert torch.equal(deq_v, ring_v)
    torch.testing.assert_close(deq_k, ring_k, rtol=0.13, atol=0.05)
    assert (deq_k != ring_k).float().mean() < 0.02


@pytest.mark.parametrize("fmt,use_graph", [("int8", False), ("int8", True), ("fp8_e4m3", True)])
def test_q8_ring_lm_matches_emulating_oracle(fmt, use_graph):
    """LM steps with an 8-bit ring against the oracle whose ring holds the same rounded rows (teacher-synchronised)."""
    from moshi_b200.config import tiny_lm_config
    from moshi_b200.models import LMGen, LMModel
    from moshi_b200.synth import synth_lm_state_dict
    from oracle import scenarios
    from oracle.lm import LMOracle, LMSpec
    cfg = tiny_lm_config()
    sd = synth_lm_state_dict(cfg, seed=scenarios.LM_SEED)
    lm = LMModel(cfg, sd, device="cuda")
    B, steps = scenarios.LM_B, 20
    codes = scenarios.lm_input_codes(cfg, B, steps)
    gen = LMGen(lm, use_sampling=False)
    gen.kv_dtype = fmt
    gen.use_graph = use_graph
    orc = LMOracle(sd, LMSpec.from_config(cfg), use_sampling=False, tie_break="index", kv_quant=fmt)
    orc.streaming(B)
    worst = 0.0
    agree = total = 0
    with gen.streaming(B):
        for i in range(steps):
            dbg = {}
            orc.step(codes[i], None, None, debug=dbg)
            gen.step(codes[i].cuda())
            tl = gen.read_buffer("text_logits", torch.bfloat16, (B, cfg.text_card)).float().cpu()
            worst = max(worst, (tl - dbg["text_logits"].float()[:, 0, 0]).abs().max().item())
            tt = gen.read_buffer("text_token", torch.int64, (B,)).cpu()
            at = gen.read_buffer("audio_tokens", torch.int64, (cfg.dep_q, B)).cpu()
            agree += int((tt == dbg["text_token"]).sum())
            total += B
            pos = (orc.offsets % orc.cache.shape[2])        # keep the oracle on the GPU's token trajectory
            for b in range(B):
                orc.cache[b, 0, pos[b]] = tt[b]
                orc.cache[b, 1:cfg.dep_q + 1, pos[b]] = at[:, b]
    print(f"{fmt} ring LM vs emulating oracle (graph={use_graph}): worst text-logit diff {worst:.3e}, greedy text tokens equal {agree}/{total}")
    assert worst < 0.12
    assert agree / total > 0.9


@pytest.mark.parametrize("fmt,outlier,bound", [("int8", 1.0, 0.02), ("fp8_e4m3", 1.0, 0.05), ("int8", 8.0, 0.08), ("fp8_e4m3", 8.0, 0.06)])
def test_q8_ring_cost_against_exact_ring(fmt, outlier, bound):
    """What the option costs: attention over a full 3000-slot 8-bit ring against fp32 attention over the unquantised
    keys / values (keys with a few large channels, as rotary models have).  Reported as relative RMS error of the head
    outputs; the bf16 ring's own error on the same data is printed beside it."""
    from moshi_b200 import _lib
    lib = _lib.lib()
    kind, qmax = FMT[fmt]
    B, H, cap = 2, 8, 3000
    g = torch.Generator().manual_seed(5)
    Cd = H * 128
    chan = torch.ones(128)
    chan[torch.randperm(128, generator=g)[:6]] = outlier                  # outlier channels
    hist_k = (torch.randn(B, H, cap, 128, generator=g) * chan).bfloat16()
    hist_v = torch.randn(B, H, cap, 128, generator=g).bfloat16()
    qkv = torch.randn(B, 3 * Cd, generator=g).bfloat16()
    pos = torch.full((B,), 2 * cap + 11, dtype=torch.int64)
    mask = torch.ones(B, dtype=torch.bool)
    q_in, k_in, v_in = (qkv[:, j * Cd:(j + 1) * Cd].reshape(B, H, 128) for j in range(3))
    q_rot, k_rot = _rope_ref(q_in, pos), _rope_ref(k_in, pos)
    ek, ev = hist_k.float().clone(), hist_v.float().clone()
    for b in range(B):
        ek[b, :, pos[b] % cap] = k_rot[b].float()
        ev[b, :, pos[b] % cap] = v_in[b].float()
    exact = F.scaled_dot_product_attention(q_rot.float()[:, :, None], ek, ev)[:, :, 0].reshape(B, Cd)
    amax_k, amax_v = hist_k.float().abs().amax(-1), hist_v.float().abs().amax(-1)
    k8, v8 = _encode(hist_k, amax_k, fmt).cuda(), _encode(hist_v, amax_v, fmt).cuda()
    ks, vs = (amax_k * (1.0 / qmax)).cuda(), (amax_v * (1.0 / qmax)).cuda()
    out = torch.empty(B, Cd, dtype=torch.bfloat16, device="cuda")
    qd, pd, md = qkv.cuda(), pos.cuda(), mask.cuda()
    _lib.check(lib.b200_op_attn_step_q8(cptr(qd), cptr(k8), cptr(v8), cptr(ks), cptr(vs), cptr(out), cptr(pd), cptr(md),
                                        B, H, cap, 0, 10000.0, kind, _stream()))
    kb, vb, out_b = hist_k.cuda(), hist_v.cuda(), torch.empty_like(out)
    _lib.check(lib.b200_op_attn_step(cptr(qd), cptr(kb), cptr(vb), cptr(out_b), cptr(pd), cptr(md), B, H, cap, 0, 10000.0, _stream()))
    torch.cuda.synchronize()
    rms = exact.pow(2).mean().sqrt()
    e_q8 = ((out.float().cpu() - exact).pow(2).mean().sqrt() / rms).item()
    e_bf = ((out_b.float().cpu() - exact).pow(2).mean().sqrt() / rms).item()
    print(f"{fmt} ring, key outlier channels x{outlier:g}: relative RMS error of the attention output {e_q8:.3%} (bf16 ring on the same data: {e_bf:.3%}); "
          f"ring bytes per session at context 3000, 32 layers: {32 * 2 * 3000 * 32 * (128 + 4) / 1e9:.3f} GB vs 1.573 GB")
    assert e_q8 < bound
