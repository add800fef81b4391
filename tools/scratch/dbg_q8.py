import ctypes as C, torch, sys
sys.path.insert(0, '.')
from moshi_b200 import _lib
from oracle.transformer import kv_roundtrip
from tests.test_gpu_zv_kv_q8 import _encode, _decode, FMT
from tests.test_gpu_ops import _rope_ref
from tests.util import cptr
lib = _lib.lib()
for fmt in ("int8", "fp8_e4m3"):
    kind, qmax = FMT[fmt]
    B, H, cap, steps = 1, 32, 3000, 2
    g = torch.Generator().manual_seed(B * 17 + cap)
    Cd = H * 128
    pos = torch.tensor([0], dtype=torch.int64)
    hist_k = torch.randn(B, H, cap, 128, generator=g).bfloat16()
    hist_v = torch.randn(B, H, cap, 128, generator=g).bfloat16() * torch.rand(B, H, cap, 1, generator=g) * 4
    ring_v = kv_roundtrip(hist_v, fmt)
    amax_k, amax_v = hist_k.float().abs().amax(-1), hist_v.float().abs().amax(-1)
    k8, v8 = _encode(hist_k, amax_k, fmt).cuda(), _encode(hist_v, amax_v, fmt).cuda()
    ks, vs = (amax_k * (1.0 / qmax)).cuda(), (amax_v * (1.0 / qmax)).cuda()
    print(fmt, "initial V equal:", torch.equal(_decode(v8.cpu(), vs.cpu(), fmt), ring_v))
    out = torch.empty(B, Cd, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(steps):
        qkv = torch.randn(B, 3 * Cd, generator=g).bfloat16()
        mask = torch.ones(B, dtype=torch.bool)
        v_in = qkv[:, 2 * Cd:].reshape(B, H, 128)
        v_new = kv_roundtrip(v_in, fmt)
        ring_v[0, :, pos[0] % cap] = v_new[0]
        qd, pd, md = qkv.cuda(), pos.cuda(), mask.cuda()
        _lib.check(lib.b200_op_attn_step_q8(cptr(qd), cptr(k8), cptr(v8), cptr(ks), cptr(vs), cptr(out), cptr(pd), cptr(md), B, H, cap, 16, 10000.0, kind, st))
        torch.cuda.synchronize()
        deq = _decode(v8.cpu(), vs.cpu(), fmt)
        bad = (deq != ring_v).nonzero()
        print(fmt, "step", i, "mismatches", bad.shape[0])
        if bad.shape[0]:
            slots = bad[:, 2].unique()
            print("  slots", slots[:10].tolist(), "heads", bad[:, 1].unique()[:10].tolist())
            b, h, s, d = bad[0].tolist()
            print("  first", (b, h, s, d), "gpu byte", v8.cpu()[b, h, s, d].item(), "gpu scale", vs.cpu()[b, h, s].item(), "deq", deq[b, h, s, d].item(), "want", ring_v[b, h, s, d].item(),
                  "x", v_in[b, h, d].float().item() if s == pos[0] % cap else None, "amax", v_in[b, h].float().abs().max().item())
        pos = pos + 1
