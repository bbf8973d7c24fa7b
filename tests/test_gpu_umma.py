"""Bring-up test of the tcgen05 building blocks (umma.cuh): a 128 x N x 64 bf16 tile product staged exactly
like the fused MC-forward kernel stages its operands, against torch on the same bf16-rounded inputs."""

import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [16, 64, 128, 192, 256])
def test_umma_tile(N):
    from bayesformer_b200 import _lib

    lib = _lib.load()
    fn = lib.bfb200_debug_umma_tile
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N)
    A = torch.randn(128, 64, generator=g).to(dev)
    W = torch.randn(N, 64, generator=g).to(dev)
    out = torch.full((128, N), float("nan"), device=dev)
    _lib.check(fn(A.data_ptr(), W.data_ptr(), N, out.data_ptr(), None))
    torch.cuda.synchronize()
    want = A.bfloat16().float() @ W.bfloat16().float().t()
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-4)
