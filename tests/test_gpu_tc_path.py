"""Layered tensor-core path (TMA-fed tcgen05 GEMMs + tcgen05 attention, fp32 residual stream) — the scaled
configuration's route (d_model 512, 8 heads of 64, F 2048, 6 layers, 256 tokens).  16-bit tolerance
class: probabilities within 1e-2 of the fp32 reference."""

import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engine():
    from bayesformer_b200 import _lib, engine as e

    _lib.load()
    return e


def _bits(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


def _probs_close(got_logp, want_logp, rtol, atol=1e-4):
    got, want = np.exp(np.asarray(got_logp, np.float64)), np.exp(np.asarray(want_logp, np.float64))
    err = np.abs(got - want) - (rtol * want + atol)
    assert err.max() <= 0, f"probabilities off by {np.abs(got - want).max():.3e} (worst excess {err.max():.3e})"


@pytest.mark.parametrize("compute,rtol", [("fp16", 1e-2), ("bf16", 5e-2)])
def test_scaled_c4_forward_against_reference_outputs(engine, dev, compute, rtol):
    """The reference's stored log-probs for the C4 model (weights / masks regenerated from their seeds and pinned by
    the fp32 module forward, as in test_gpu_parity.test_scaled_c4_forward)."""
    from bayesformer_b200.model.transformer import MyTransformer
    from oracle import ref_shim

    z = H.load("scaled_c4.npz")
    V, d, Hh, F, L = (int(v) for v in z["dims"])
    p = float(z["p"])
    torch.manual_seed(int(z["weight_seed"]))
    model = MyTransformer(V, d, Hh, F, L, bayes_dropout=p, bayes=True).eval()
    src = torch.from_numpy(z["src"])
    log = []
    torch.manual_seed(int(z["mask_seed"]))
    with torch.no_grad(), ref_shim.record_dropout(log):
        cpu_out = model._module_forward(src)
    np.testing.assert_allclose(cpu_out[-1].numpy(), z["logp_last"], rtol=1e-4, atol=1e-5)
    pm = R.PackedMasks(1, L, src.shape[1], src.shape[0], d)
    pm.set(0, R.SITE_EMBED, log[0].numpy())
    for i in range(L):
        pm.set(0, R.site_layer_out(i), log[1 + i].numpy())
    model = model.to(dev)
    pack = model.native_pack(dev, compute)
    assert pack.path == 2
    with torch.no_grad():
        got = engine.transformer_forward(pack, src.to(dev), mask_bits=_bits(pm.bits, dev)).cpu().numpy()
    _probs_close(got[-1], z["logp_last"], rtol)
    _probs_close(got[::37], z["logp_rows"], rtol)


@pytest.mark.parametrize("P,N,B,T", [(9, 4, 21, 5), (40, 2, 7, 3), (130, 1, 3, 2), (256, 1, 2, 2), (1, 3, 5, 4)])
def test_tc_path_matches_exact_path(engine, dev, P, N, B, T):
    """Same seed => same Philox masks; the fp32 path's tokens are forced into the tensor-core run."""
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(P)
    model = MyTransformer(50, 128, 2, 72, 2, bayes_dropout=0.2, bayes=True).eval().to(dev)
    prompts = torch.randint(0, 50, (P, B), device=dev)
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, T, "dropout", "entropy", seed=7,
                                        return_logp=True, return_tokens=True)
    forced = exact.tokens[:, P:].t().contiguous()
    pack = model.native_pack(dev, "fp16")
    assert pack.path == 2
    got = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "entropy", seed=7, forced_tokens=forced,
                                      return_logp=True)
    _probs_close(got.logp.cpu().numpy(), exact.logp.cpu().numpy(), 1e-2)
    np.testing.assert_allclose(got.step_scores.cpu().numpy(), exact.step_scores.cpu().numpy(), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("compute,fp16_residual,rtol,atol", [("fp16", False, 1e-2, 1e-4), ("fp16", True, 1e-2, 1e-4),
                                                             ("bf16", False, 6e-2, 2e-3)])
def test_tc_path_fused_epilogues_many_tiles_per_cluster(engine, dev, monkeypatch, compute, fp16_residual, rtol, atol):
    """d_model 256: every product has more than 128 output columns, so the two-SM GEMM with its fused epilogues runs —
    LayerNorm1 folded into the FFN1 product, residual + row statistics in the W_out epilogue (and, with
    BFB200_TC_FP16_RESIDUAL=1, in the FFN2 epilogue + ln16_apply) — on enough rows (31 k) that every cluster walks
    several tiles and the residual slots wrap around the operand ring.  F = 136 leaves a column tail in FFN1."""
    from bayesformer_b200.model.transformer import MyTransformer

    if fp16_residual:
        monkeypatch.setenv("BFB200_TC_FP16_RESIDUAL", "1")
    else:
        monkeypatch.delenv("BFB200_TC_FP16_RESIDUAL", raising=False)

    torch.manual_seed(11)
    model = MyTransformer(50, 256, 4, 136, 2, bayes_dropout=0.2, bayes=True).eval().to(dev)
    P, N, B, T = 130, 2, 40, 6
    prompts = torch.randint(0, 50, (P, B), device=dev)
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, T, "dropout", "entropy", seed=3,
                                        return_logp=True, return_tokens=True)
    forced = exact.tokens[:, P:].t().contiguous()
    pack = model.native_pack(dev, compute)
    assert pack.path == 2
    with engine.profile(dev) as prof:
        got = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "entropy", seed=3, forced_tokens=forced,
                                          return_logp=True)
    assert "gemm_tc_kernel[LayerNorm folded]" in prof.kernels and "gemm_tc_kernel[+residual,+row stats]" in prof.kernels
    assert ("ln16_apply_kernel" in prof.kernels) == fp16_residual
    _probs_close(got.logp.cpu().numpy(), exact.logp.cpu().numpy(), rtol, atol)
    np.testing.assert_allclose(got.step_scores.cpu().numpy(), exact.step_scores.cpu().numpy(), rtol=rtol, atol=1e-3)


def test_tc_path_latent_sites_and_chunking(engine, dev):
    import ctypes as C

    from bayesformer_b200 import _lib
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(3)
    p = 0.25
    model = MyTransformer(40, 64, 1, 32, 2, bayes_dropout=p, bayes=True).eval()
    for layer in model.layers:
        layer.bayes, layer.bayes_dropout = True, p
        layer.ff.bayes_dropout = p
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
    model = model.to(dev)
    prompts = torch.randint(0, 40, (20, 33), device=dev)
    N, T = 3, 4
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, T, "dropout", "margin", seed=11,
                                        return_logp=True, return_tokens=True)
    forced = exact.tokens[:, 20:].t().contiguous()
    pack = model.native_pack(dev, "bf16")   # d_model 64 with ONE head of width 64: tensor-core path, not the fused class
    assert pack.path == 2
    got = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=11, forced_tokens=forced,
                                      return_logp=True)
    _probs_close(got.logp.cpu().numpy(), exact.logp.cpu().numpy(), 6e-2, atol=2e-3)
    args = _lib.McArgs(n_items=33, prompt_len=20, new_tokens=N, n_draws=T)
    small = _lib.load().bfb200_mc_workspace_bytes(C.byref(pack.desc), C.byref(args), 4)
    chunked = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=11, forced_tokens=forced,
                                          workspace_bytes=small)
    assert torch.equal(chunked.step_scores, got.step_scores)


def test_auto_picks_the_tensor_core_path_for_width_64_heads(engine, dev):
    from bayesformer_b200.model.transformer import MyTransformer

    big = MyTransformer(114, 512, 8, 2048, 1, bayes_dropout=0.3, bayes=True).eval().to(dev)
    pack = engine.scoring_pack(big, dev, 256, 1, "auto")
    assert pack.compute == "fp16" and pack.path == 2
    assert engine.scoring_pack(big, dev, 300, 1, "auto").compute == "fp32"


def test_scaled_c4_full_size_properties(engine, dev):
    """BASELINE configs[3] per-GPU size (8,192 items x 50 draws x 256 tokens, d = 512, F = 2048, 6 layers, entropy
    score), through properties that need no oracle run: two shards with their global index bases reproduce the
    unsharded scores bit for bit (whatever the chunking) and acquire the same 200 items; scores lie in [0, ln V]."""
    import math

    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(4)
    V, B, T, k = 114, 8192, 50, 200
    model = MyTransformer(V, 512, 8, 2048, 6, bayes_dropout=0.3, bayes=True).eval().to(dev)
    prompts = torch.randint(0, V, (256, B), generator=torch.Generator().manual_seed(4)).to(dev)
    pack = engine.scoring_pack(model, dev, 256, 1, "auto")
    assert pack.path == 2
    whole = engine.mc_score_transformer(pack, prompts, 1, T, "dropout", "entropy", seed=42).pool_scores
    assert bool(torch.isfinite(whole).all()) and float(whole.min()) >= 0.0 and float(whole.max()) <= math.log(V) + 1e-4
    keys = []
    for lo, hi in ((0, 3000), (3000, B)):
        part = engine.mc_score_transformer(pack, prompts[:, lo:hi].contiguous(), 1, T, "dropout", "entropy", seed=42,
                                           index_base=lo).pool_scores
        assert torch.equal(part, whole[lo:hi])
        keys.append(engine.topk_keys(part, k, index_base=lo))
    _, got = engine.unpack_keys(engine.merge_topk_keys(torch.cat(keys), k))
    _, want = engine.unpack_keys(engine.topk_keys(whole, k))
    assert torch.equal(got, want)


@pytest.mark.parametrize("compute", ["fp32", "fp16", "bf16"])
def test_layered_paths_are_deterministic_and_shard_invariant(engine, dev, compute):
    """A model outside the fused kernel's class (d = 128, two heads of width 64) on a 20,000-item pool: the layered
    fp32 path and the layered tensor-core path are deterministic and reproduce the unsharded scores bit for bit from
    shards with their global index bases (chunk boundaries, GEMM tile tails and acquisition differ between the runs)."""
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(12)
    model = MyTransformer(50, 128, 2, 72, 2, bayes_dropout=0.2, bayes=True).eval().to(dev)
    n = 20000
    prompts = torch.randint(0, 50, (12, n), generator=torch.Generator().manual_seed(12)).to(dev)
    pack = model.native_pack(dev, compute)
    whole = engine.mc_score_transformer(pack, prompts, 3, 10, "dropout", "margin", seed=5).pool_scores
    assert torch.equal(whole, engine.mc_score_transformer(pack, prompts, 3, 10, "dropout", "margin", seed=5).pool_scores)
    for lo, hi in ((0, 4097), (4097, 13000), (13000, n)):
        part = engine.mc_score_transformer(pack, prompts[:, lo:hi].contiguous(), 3, 10, "dropout", "margin", seed=5,
                                           index_base=lo).pool_scores
        assert torch.equal(part, whole[lo:hi])
