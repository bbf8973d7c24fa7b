"""TMA-fed tcgen05 GEMM (bfb200_linear_tc) against torch on the same 16-bit operands (fp32 accumulation)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,relu,bias", [(1000, 1536, 512, False, False), (129, 114, 512, False, True),
                                             (4096, 2048, 512, True, True), (4096, 512, 2048, False, True),
                                             (1, 8, 64, False, True), (300, 200, 72, True, False),
                                             (20000, 512, 512, False, False)])
def test_linear_tc_matches_torch(dtype, M, N, K, relu, bias):
    from bayesformer_b200 import engine

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev, dtype)
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev, dtype)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    want = a.float() @ w.float().t()
    if bias:
        want = want + b
    if relu:
        want = torch.relu(want)
    got32 = engine.linear_tc(a, w, b, relu, out_dtype=torch.float32)
    torch.testing.assert_close(got32, want, rtol=1e-4, atol=1e-3)
    got16 = engine.linear_tc(a, w, b, relu)
    torch.testing.assert_close(got16.float(), want.to(dtype).float(), rtol=2e-2 if dtype == torch.bfloat16 else 2e-3,
                               atol=2e-2 if dtype == torch.bfloat16 else 2e-3)


def test_linear_tc_strided_rows_pick_last_positions():
    from bayesformer_b200 import engine

    dev = torch.device("cuda:0")
    S, L, K, N = 37, 16, 512, 114
    x = torch.randn(S, L, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * 0.05).half()
    last = x[:, -1, :].contiguous()
    want = last.float() @ w.float().t()
    got = engine.linear_tc(x.view(-1)[(L - 1) * K:], w, out_dtype=torch.float32, a_pitch=L * K, rows=S)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_seq,length,heads", [(3, 256, 8), (2, 200, 8), (5, 128, 2), (4, 40, 2), (7, 7, 1), (1, 129, 4),
                                                (2, 16, 8), (3, 1, 2)])
def test_attention_tc_matches_torch(dtype, n_seq, length, heads):
    """Fused causal attention (tcgen05 QK^T and PV) against the reference formula (transformer.py:58-65) evaluated in
    fp32 on the same 16-bit Q, K, V."""
    import math

    from bayesformer_b200 import engine

    dev = torch.device("cuda:0")
    d = heads * 64
    g = torch.Generator().manual_seed(n_seq * 1000 + length)
    qkv = (torch.randn(n_seq * length, 3 * d, generator=g) * 0.7).to(dev, dtype)
    got = engine.attention_tc(qkv, n_seq, length, heads).float()
    x = qkv.float().view(n_seq, length, 3, heads, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))            # (S, H, L, 64)
    scores = q @ k.transpose(-1, -2) / math.sqrt(64)
    scores = scores + torch.triu(torch.full((length, length), float("-inf"), device=dev), diagonal=1)
    want = (torch.softmax(scores, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * length, d)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    torch.testing.assert_close(got, want, rtol=tol, atol=tol)
