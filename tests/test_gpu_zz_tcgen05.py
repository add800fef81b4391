"""GPU: the tcgen05 / TMA GEMM against fp32 math and against the SIMT kernel (runs last: a faulting
tensor-core kernel poisons the CUDA context for everything after it)."""
import ctypes as C

import pytest
import torch

from tests.util import cptr, stats

pytestmark = pytest.mark.gpu

SHAPES = [(1, 512, 256), (3, 1024, 4096), (8, 12288, 4096), (16, 128, 64), (17, 4096, 11264), (96, 4096, 4096),
          (128, 2048, 1024), (200, 1024, 2816), (256, 32000, 4096), (96, 704, 256), (5, 200, 72)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tcgen05_linear(M, N, K):
    from moshi_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    y = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    y1 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.b200_op_linear_bf16(cptr(x), cptr(w), cptr(y), M, N, K, 2, st))
    _lib.check(lib.b200_op_linear_bf16(cptr(x), cptr(w), cptr(y1), M, N, K, 1, st))
    torch.cuda.synchronize()
    want = x.float() @ w.float().t()
    print(stats(f"tcgen05 {M}x{N}x{K}", y, want), "| vs simt:", stats("", y, y1))
    assert not torch.isnan(y.float()).any()
    torch.testing.assert_close(y.float(), want, rtol=1e-2, atol=2e-2)
    ulp_off = ((y.float() - want.bfloat16().float()).abs() > 0).float().mean().item()
    assert ulp_off < 0.05, ulp_off
