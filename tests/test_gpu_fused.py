"""GPU parity of the fused tcgen05 MC-forward kernel (compute='fp16', bayesformer_b200/csrc/mc_fused.cu).

16-bit tolerance class ("1e-2 in bf16") of BASELINE.json's north_star: probabilities within 1e-2 relative of the fp32 reference
(equivalently |dlogp| <= 1e-2), scores within 1e-2, integer outputs identical wherever the oracle's decision gap
exceeds that tolerance.  Checked against (1) the golden vectors recorded from the unmodified reference with its own
dropout masks injected, (2) the CPU oracle under the library's Philox counter scheme, (3) the exact fp32 CUDA path
under the same seed (same masks), which also covers the latent dropout sites and the 16- and 40-token variants.
"""

import ctypes as C

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import restatement as R
from tests import helpers as H

pytestmark = pytest.mark.gpu

RTOL = 1e-2
ATOL_P = 1e-4       # probabilities
ATOL_SCORE = 1e-3


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engine():
    from bayesformer_b200 import _lib, engine as e

    _lib.load()
    return e


@pytest.fixture(scope="module")
def c2():
    z = H.load("c2_transformer.npz")
    return z, H.state_dict_from(z)


def _bits(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


def _assert_probs_close(got_logp, want_logp, rtol=RTOL, atol=ATOL_P):
    got, want = np.exp(np.asarray(got_logp, dtype=np.float64)), np.exp(np.asarray(want_logp, dtype=np.float64))
    err = np.abs(got - want) - (rtol * want + atol)
    assert err.max() <= 0, f"probabilities off by {np.abs(got - want).max():.3e} (worst excess {err.max():.3e})"


def _decisive(logp: torch.Tensor, tol: float) -> torch.Tensor:
    top2 = torch.topk(logp, 2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > 2 * tol  # both candidates may move by tol


@pytest.mark.parametrize("mode", H.MODES)
def test_teacher_forced_with_reference_masks(engine, dev, c2, mode):
    z, sd = c2
    N, T, L, p = int(z["N"]), int(z["T"]), 2, float(z["p"])
    B = z["prompts"].shape[1]
    bits = H.expand_live_bits(z["dropout.keep_bits"], L)
    forced = torch.from_numpy(z["dropout.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)
    model = H.build_model(sd, p, True).to(dev)
    pack = model.native_pack(dev, "fp16")
    res = engine.mc_score_transformer(pack, torch.from_numpy(z["prompts"]).to(dev), N, T, "dropout", mode,
                                      mask_bits=_bits(bits, dev), mask_max_len=bits.shape[3], forced_tokens=forced,
                                      return_logp=True)
    np.testing.assert_allclose(res.step_scores.cpu().numpy(), z[f"dropout.scores.{mode}"], rtol=RTOL, atol=ATOL_SCORE)
    lp = res.logp.cpu().view(N, B, T, -1).permute(2, 0, 1, 3)[:6].numpy()
    _assert_probs_close(lp, z["dropout.logp_first_draws"])


def test_free_running_with_reference_masks(engine, dev, c2):
    z, sd = c2
    P, N, T, L, d, p = int(z["P"]), int(z["N"]), int(z["T"]), 2, 64, float(z["p"])
    B = z["prompts"].shape[1]
    bits = H.expand_live_bits(z["dropout.keep_bits"], L)
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev, "fp16"), torch.from_numpy(z["prompts"]).to(dev), N, T,
                                      "dropout", "entropy", mask_bits=_bits(bits, dev), mask_max_len=bits.shape[3],
                                      return_logp=True, return_tokens=True)
    pm = H.packed_masks_from_bits(bits, L, d)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", "entropy", p,
                       lambda step: pm.mask_fn(step, R.default_live_sites(L, True)))
    got_tok = res.tokens.cpu().long()
    decisive = _decisive(want["logp"], RTOL).all(dim=0)
    assert decisive.float().mean() > 0.7
    assert torch.equal(got_tok[decisive], want["tokens"][decisive])
    same = (got_tok == want["tokens"]).all(dim=1).numpy()
    _assert_probs_close(res.logp.cpu().numpy()[:, same], want["logp"].numpy()[:, same])
    np.testing.assert_allclose(res.step_scores[0].cpu().numpy(), z["dropout.scores.entropy"][0], rtol=RTOL, atol=ATOL_SCORE)
    items_same = torch.from_numpy(same).view(B, T).all(dim=1).numpy()
    np.testing.assert_allclose(res.pool_scores.cpu().numpy()[items_same],
                               z["dropout.scores.entropy"].mean(0)[items_same], rtol=RTOL, atol=ATOL_SCORE)


@pytest.mark.parametrize("mode", H.MODES)
def test_greedy_and_sampling_schemes(engine, dev, c2, mode):
    z, sd = c2
    N, T = int(z["N"]), int(z["T"])
    B = z["prompts"].shape[1]
    model = H.build_model(sd, None, False).to(dev)
    pack = model.native_pack(dev, "fp16")
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    res = engine.mc_score_transformer(pack, prompts, N, 1, "greedy", mode, return_logp=True)
    want_logp = torch.from_numpy(z["greedy.logp"])
    ok = _decisive(want_logp, RTOL).all(dim=0).numpy()
    _assert_probs_close(res.logp.cpu().numpy()[:, ok], want_logp.numpy()[:, ok])
    np.testing.assert_allclose(res.step_scores.cpu().numpy()[:, ok], z[f"greedy.scores.{mode}"][:, ok], rtol=RTOL,
                               atol=ATOL_SCORE)
    forced = torch.from_numpy(z["sampling.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)
    res = engine.mc_score_transformer(pack, prompts, N, T, "sampling", mode, temperature=0.8, forced_tokens=forced)
    np.testing.assert_allclose(res.step_scores.cpu().numpy(), z[f"sampling.scores.{mode}"], rtol=RTOL, atol=ATOL_SCORE)


def test_philox_mode_matches_oracle(engine, dev, c2):
    z, sd = c2
    N, T, L, d, p, seed, base = int(z["N"]), 5, 2, 64, float(z["p"]), 0xC0FFEE1234, 4000
    B = z["prompts"].shape[1]
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev, "fp16"), torch.from_numpy(z["prompts"]).to(dev), N, T,
                                      "dropout", "margin", seed=seed, index_base=base, return_logp=True,
                                      return_tokens=True)
    items, draws = np.repeat(np.arange(base, base + B), T), np.tile(np.arange(T), B)
    live = R.default_live_sites(L, True)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", "margin", p,
                       lambda step: R.philox_mask_fn(seed, p, items, draws, step, d, live))
    same = (res.tokens.cpu().long() == want["tokens"]).all(dim=1).numpy()
    assert same.mean() > 0.75
    _assert_probs_close(res.logp.cpu().numpy()[:, same], want["logp"].numpy()[:, same])
    np.testing.assert_allclose(res.step_scores[0].cpu().numpy(), want["step_scores"][0].numpy(), rtol=RTOL, atol=ATOL_SCORE)


def test_sampling_philox_tokens(engine, dev, c2):
    z, sd = c2
    N, T, seed = int(z["N"]), 6, 777
    B = z["prompts"].shape[1]
    model = H.build_model(sd, None, False).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev, "fp16"), torch.from_numpy(z["prompts"]).to(dev), N, T,
                                      "sampling", "entropy", temperature=0.8, seed=seed, index_base=50,
                                      return_tokens=True)
    items, draws = np.repeat(np.arange(50, 50 + B), T), np.tile(np.arange(T), B)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "sampling", "entropy", 0.0, lambda step: None,
                       temperature=0.8, sampling_u=lambda step: philox.sampling_uniform(seed, items, draws, step))
    same = (res.tokens.cpu().long() == want["tokens"]).all(dim=1)
    assert same.float().mean() > 0.8


def _fused_vs_exact(engine, dev, model, prompts, N, T, scheme, mode, seed, rtol=RTOL, atol=ATOL_P):
    """Same seed => same Philox masks in both CUDA paths; the fp32 path's tokens are forced into the fused run."""
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, T, scheme, mode, seed=seed,
                                        return_logp=True, return_tokens=True)
    P = prompts.shape[0]
    forced = exact.tokens[:, P:].t().contiguous()
    fused = engine.mc_score_transformer(model.native_pack(dev, "fp16"), prompts, N, T, scheme, mode, seed=seed,
                                        forced_tokens=forced, return_logp=True)
    _assert_probs_close(fused.logp.cpu().numpy(), exact.logp.cpu().numpy(), rtol, atol)
    np.testing.assert_allclose(fused.step_scores.cpu().numpy(), exact.step_scores.cpu().numpy(), rtol=rtol,
                               atol=ATOL_SCORE)
    return exact, fused


@pytest.mark.parametrize("P,N,B,T", [(4, 5, 257, 20), (1, 3, 33, 4), (2, 2, 1, 1), (3, 6, 70, 7), (7, 2, 19, 20),
                                     (9, 4, 45, 5), (12, 5, 23, 3), (16, 1, 9, 2), (17, 3, 11, 3), (32, 5, 29, 4),
                                     (40, 1, 7, 2), (25, 6, 10, 5), (60, 3, 5, 3), (127, 2, 3, 2)])
def test_fused_matches_exact_path_all_shapes(engine, dev, c2, P, N, B, T):
    """Tile tails, 1..128-token sequences, every mode, uniformly random token prompts (far
    outside the trained model's input distribution), at the stated bound: 1e-2 relative + 1e-4 absolute on every
    probability (tools/fused_accuracy.py: the worst of 10^7 probabilities sits at 0.5 of it)."""
    z, sd = c2
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    g = torch.Generator().manual_seed(P * 100 + N)
    prompts = torch.randint(0, 114, (P, B), generator=g).to(dev)
    for mode in H.MODES:
        _fused_vs_exact(engine, dev, model, prompts, N, T, "dropout", mode, seed=1000 + P)


def test_fused_serves_the_char_level_configuration(engine, dev):
    """SURVEY C1 shape: MyTransformer(67, 64, 8, 8, 2), 32-character prompts, 5 decode steps (sequences of 32..36
    tokens) — the long-sequence variant of the fused kernel, against the exact fp32 path under the same seed, with the
    latent dropout sites switched on as well."""
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(67)
    model = MyTransformer(67, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval()
    prompts = torch.randint(0, 65, (32, 50), generator=torch.Generator().manual_seed(5)).to(dev)
    assert engine.scoring_pack(model.to(dev), dev, 32, 5, "auto").compute == "fp16"
    for mode in H.MODES:
        _fused_vs_exact(engine, dev, model.to(dev), prompts, 5, 6, "dropout", mode, seed=4242)
    for layer in model.layers:
        layer.bayes, layer.bayes_dropout = True, 0.3
        layer.ff.bayes_dropout = 0.3
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, 0.3
    _fused_vs_exact(engine, dev, model.to(dev), prompts, 5, 4, "sampling", "entropy", seed=77)


@pytest.mark.parametrize("compute,rtol,atol", [("fp16", RTOL, ATOL_P), ("fp32", 1e-3, 1e-5)])
@pytest.mark.parametrize("mode", H.MODES)
def test_char_level_c1_with_reference_masks(engine, dev, mode, compute, rtol, atol):
    """SURVEY C1 / BASELINE configs[0] against the REFERENCE itself: the char-level model (V = 67) on windows of
    data/shakespeare.txt, sequences of 32..36 tokens, the reference's own dropout masks injected and its tokens forced
    (tests/golden/c1_char_level.npz, recorded from the unmodified reference by oracle/make_golden.py)."""
    z = H.load("c1_char_level.npz")
    sd = H.state_dict_from(z)
    N, T, L, p = int(z["N"]), int(z["T"]), 2, float(z["p"])
    B = z["prompts"].shape[1]
    bits = H.expand_live_bits(z["dropout.keep_bits"], L)
    forced = torch.from_numpy(z["dropout.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev, compute), torch.from_numpy(z["prompts"]).to(dev), N, T,
                                      "dropout", mode, mask_bits=_bits(bits, dev), mask_max_len=bits.shape[3],
                                      forced_tokens=forced, return_logp=True)
    np.testing.assert_allclose(res.step_scores.cpu().numpy(), z[f"dropout.scores.{mode}"], rtol=rtol, atol=ATOL_SCORE * rtol / RTOL)
    lp = res.logp.cpu().view(N, B, T, -1).permute(2, 0, 1, 3).numpy()
    _assert_probs_close(lp, z["dropout.logp"], rtol, atol)


def test_fused_latent_sites_match_exact_path(engine, dev, c2):
    z, sd = c2
    p = 0.25
    model = H.build_model(sd, p, True)
    for layer in model.layers:
        layer.bayes, layer.bayes_dropout = True, p
        layer.ff.bayes_dropout = p
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
    model = model.to(dev)
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    _fused_vs_exact(engine, dev, model, prompts, int(z["N"]), 6, "dropout", "entropy", seed=31337)


def test_fused_is_invariant_to_chunking_and_sharding(engine, dev, c2):
    from bayesformer_b200 import _lib

    z, sd = c2
    N, T, seed = int(z["N"]), 20, 99
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    B = prompts.shape[1]
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    pack = model.native_pack(dev, "fp16")
    whole = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=seed)
    lib = _lib.load()
    args = _lib.McArgs(n_items=B, prompt_len=prompts.shape[0], new_tokens=N, n_draws=T)
    small = lib.bfb200_mc_workspace_bytes(C.byref(pack.desc), C.byref(args), 5)
    chunked = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=seed, workspace_bytes=small)
    assert torch.equal(whole.step_scores, chunked.step_scores)
    shard = engine.mc_score_transformer(pack, prompts[:, 7:23].contiguous(), N, T, "dropout", "margin", seed=seed,
                                        index_base=7)
    assert torch.equal(shard.step_scores, whole.step_scores[:, 7:23])


def test_fused_other_model_shapes(engine, dev):
    """F = 64 (four K steps in the second FFN product), V = 67 (Shakespeare alphabet), one and two layers."""
    from bayesformer_b200.model.transformer import MyTransformer

    for (V, F, L) in ((67, 64, 1), (128, 16, 2), (20, 8, 2)):
        torch.manual_seed(V + F)
        model = MyTransformer(V, 64, 8, F, L, bayes_dropout=0.3, bayes=True).eval().to(dev)
        prompts = torch.randint(0, V, (5, 100), device=dev)
        _fused_vs_exact(engine, dev, model, prompts, 4, 6, "dropout", "entropy", seed=5)


def test_fp16_outside_the_fused_class_is_an_error(engine, dev):
    from bayesformer_b200.model.transformer import MyTransformer

    model = MyTransformer(50, 128, 8, 64, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
    with pytest.raises(NotImplementedError):
        model.native_pack(dev, "fp16")
    # 'auto' falls to the exact fp32 CUDA path (still native, still on the GPU)
    pack = engine.scoring_pack(model, dev, 4, 5, "auto")
    assert pack.compute == "fp32"
    small = MyTransformer(114, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
    assert engine.scoring_pack(small, dev, 4, 5, "auto").compute == "fp16"
    assert engine.scoring_pack(small, dev, 32, 5, "auto").compute == "fp16"   # 36-token sequences
    assert engine.scoring_pack(small, dev, 124, 5, "auto").compute == "fp16"  # 128-token sequences: one per tile
    assert engine.scoring_pack(small, dev, 125, 5, "auto").compute == "fp32"  # 129 tokens: outside the class
    with pytest.raises(NotImplementedError):
        engine.mc_score_transformer(small.native_pack(dev, "fp16"), torch.zeros(125, 4, dtype=torch.int64, device=dev),
                                    5, 2, "dropout", "max")


def test_select_samples_uses_fused_kernel_and_agrees_with_exact_ranking(engine, dev, c2):
    import random

    from bayesformer_b200 import _lib
    from bayesformer_b200.libs import active_learning as al, preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    tok = ReversedPairingTokenizer()
    random.seed(5)
    pool = preprocessing.create_dataset(3000, 3)
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    lib = _lib.load()
    with engine.profile(dev) as prof:
        picked = al.select_samples(model, tok, pool, dev, nb_samples=200, compute="dropout", mode="entropy")
    assert "mc_forward_kernel" in prof.kernels and "sgemm_tn_kernel" not in prof.kernels
    assert len(set(picked)) == 200
    model.native_compute = "fp32"
    exact = al.select_samples(model, tok, pool, dev, nb_samples=200, compute="dropout", mode="entropy")
    # different Philox streams per call: the overlap is bounded by the MC noise floor (~150/200 for
    # entropy on trained weights), far above chance (200*200/3000 = 13)
    assert len(set(picked) & set(exact)) >= 110


def test_weight_updates_invalidate_the_packed_images(engine, dev, c2):
    """Active-learning rounds fine-tune the model in place between acquisitions (reference utils.train,
    src/libs/utils.py:68-78): the fp32 pack and the fp16 weight image must follow the parameters."""
    import random

    from bayesformer_b200.libs import preprocessing, utils
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    N = int(z["N"])
    before = engine.mc_score_transformer(engine.scoring_pack(model, dev, 4, N), prompts, N, 4, "dropout", "entropy", seed=5)
    pack_before = engine.scoring_pack(model, dev, 4, N)
    assert engine.scoring_pack(model, dev, 4, N) is pack_before          # cached while the weights are unchanged
    tok = ReversedPairingTokenizer()
    random.seed(0)
    data = preprocessing.create_dataset(64, 3)
    with torch.enable_grad():
        utils.train(model, data, data[:32], 1, 16, 1e-2, tok.ntokens, tok, dev)   # eager autograd, in-place AdamW steps
    model.eval()
    pack_after = engine.scoring_pack(model, dev, 4, N)
    assert pack_after is not pack_before
    after = engine.mc_score_transformer(pack_after, prompts, N, 4, "dropout", "entropy", seed=5)
    assert not torch.allclose(after.step_scores, before.step_scores)
    exact = engine.mc_score_transformer(model.native_pack(dev, "fp32"), prompts, N, 4, "dropout", "entropy", seed=5,
                                        return_tokens=True)
    forced = exact.tokens[:, 4:].t().contiguous()
    again = engine.mc_score_transformer(pack_after, prompts, N, 4, "dropout", "entropy", seed=5, forced_tokens=forced)
    np.testing.assert_allclose(again.step_scores.cpu().numpy(), exact.step_scores.cpu().numpy(), rtol=RTOL, atol=ATOL_SCORE)
    # a deep copy (active_learning.ipynb cell 24) gets its own packs
    import copy
    clone = copy.deepcopy(model)
    assert engine.scoring_pack(clone, dev, 4, N) is not pack_after


def test_sharded_selection_equals_single_device_selection(engine, dev, c2):
    """a pool scored as G shards (global index bases, same seed) and merged through the candidate keys
    acquires exactly the indices of the unsharded run — 1-GPU and 8-GPU runs agree bit for bit."""
    import random

    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    tok = ReversedPairingTokenizer()
    random.seed(9)
    pool = preprocessing.create_dataset(4003, 3)
    X, _, P, A, _, _ = preprocessing.get_batch(pool, tok, 0, len(pool))
    X = X.to(dev)
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    pack = engine.scoring_pack(model, dev, P, A + 1)
    k, seed, G = 200, 1234, 8
    whole = engine.mc_score_transformer(pack, X, A + 1, 20, "dropout", "entropy", seed=seed)
    _, want = engine.unpack_keys(engine.topk_keys(whole.pool_scores, k))
    keys = []
    n = X.shape[1]
    for g in range(G):
        lo, hi = n * g // G, n * (g + 1) // G
        part = engine.mc_score_transformer(pack, X[:, lo:hi].contiguous(), A + 1, 20, "dropout", "entropy", seed=seed,
                                           index_base=lo)
        assert torch.equal(part.pool_scores, whole.pool_scores[lo:hi])
        keys.append(engine.topk_keys(part.pool_scores, k, index_base=lo))
    _, got = engine.unpack_keys(engine.merge_topk_keys(torch.cat(keys), k))
    assert torch.equal(got, want)


def test_full_size_pool_properties(engine, dev, c2):
    """BASELINE configs[1] at its full size (100,000 pool items, T = 20, N = 5 decode steps, margin and max scores),
    through properties that need no oracle run: a re-run is bit-identical, eight shards with their global index bases
    reproduce the unsharded scores bit for bit and acquire the same 200 items, scores stay in their range, the
    acquired list is sorted (descending score, ties by lowest index) and holds the 200 largest scores."""
    import random

    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    random.seed(42)
    pool = preprocessing.create_dataset(100000, 3)
    X, _, P, A, _, _ = preprocessing.get_batch(pool, ReversedPairingTokenizer(), 0, len(pool))
    X = X.to(dev)
    n, k, T, seed, G = X.shape[1], 200, 20, 42, 8
    assert n == 100000
    model = H.build_model(sd, float(z["p"]), True).to(dev)
    pack = engine.scoring_pack(model, dev, P, A + 1)
    for mode in ("margin", "max"):
        whole = engine.mc_score_transformer(pack, X, A + 1, T, "dropout", mode, seed=seed).pool_scores
        again = engine.mc_score_transformer(pack, X, A + 1, T, "dropout", mode, seed=seed).pool_scores
        assert torch.equal(whole, again)
        assert bool(torch.isfinite(whole).all()) and float(whole.min()) >= 0.0 and float(whole.max()) <= 1.0
        vals, idx = engine.unpack_keys(engine.topk_keys(whole, k))
        v, i = vals.cpu().numpy(), idx.cpu().numpy()
        assert len(set(i.tolist())) == k
        assert all(v[j] > v[j + 1] or (v[j] == v[j + 1] and i[j] < i[j + 1]) for j in range(k - 1))
        np.testing.assert_array_equal(v, whole[idx].cpu().numpy())
        rest = whole.clone()
        rest[idx] = -1.0
        assert float(rest.max()) <= float(v[-1])
        keys = []
        for g in range(G):
            lo, hi = n * g // G, n * (g + 1) // G
            part = engine.mc_score_transformer(pack, X[:, lo:hi].contiguous(), A + 1, T, "dropout", mode, seed=seed,
                                               index_base=lo).pool_scores
            assert torch.equal(part, whole[lo:hi])
            keys.append(engine.topk_keys(part, k, index_base=lo))
        _, got = engine.unpack_keys(engine.merge_topk_keys(torch.cat(keys), k))
        assert torch.equal(got, idx)


def test_full_size_greedy_and_sampling_are_shard_invariant(engine, dev, c2):
    """The standard Transformer's two schemes (greedy_uncertainty, sampling_uncertainty: active_learning.py:38-115) on the
    100,000-item pool: deterministic, and shards with global index bases reproduce the unsharded scores bit for bit
    (sampling draws its uniforms from the Philox stream keyed by the global item index)."""
    import random

    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    random.seed(43)
    pool = preprocessing.create_dataset(100000, 3)
    X, _, P, A, _, _ = preprocessing.get_batch(pool, ReversedPairingTokenizer(), 0, len(pool))
    X = X.to(dev)
    n = X.shape[1]
    model = H.build_model(sd, None, False).to(dev)
    pack = engine.scoring_pack(model, dev, P, A + 1)
    for scheme, T, kw in (("greedy", 1, {}), ("sampling", 20, {"temperature": 0.8})):
        whole = engine.mc_score_transformer(pack, X, A + 1, T, scheme, "entropy", seed=7, **kw).pool_scores
        again = engine.mc_score_transformer(pack, X, A + 1, T, scheme, "entropy", seed=7, **kw).pool_scores
        assert torch.equal(whole, again)
        assert bool(torch.isfinite(whole).all()) and float(whole.min()) >= 0.0
        for lo, hi in ((0, 12345), (12345, 70001), (70001, n)):
            part = engine.mc_score_transformer(pack, X[:, lo:hi].contiguous(), A + 1, T, scheme, "entropy", seed=7,
                                               index_base=lo, **kw).pool_scores
            assert torch.equal(part, whole[lo:hi])


def test_char_level_full_size_pool_properties(engine, dev):
    """BASELINE configs[0] size (10,000 windows of 32 characters, 5 decode steps, T = 20, entropy; V = 67): the
    long-sequence variant of the fused kernel is deterministic and invariant to sharding, bit for bit."""
    import math

    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(67)
    model = MyTransformer(67, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).eval().to(dev)
    n, k = 10000, 200
    prompts = torch.randint(0, 65, (32, n), generator=torch.Generator().manual_seed(6)).to(dev)
    pack = engine.scoring_pack(model, dev, 32, 5, "auto")
    assert pack.compute == "fp16"
    whole = engine.mc_score_transformer(pack, prompts, 5, 20, "dropout", "entropy", seed=42).pool_scores
    again = engine.mc_score_transformer(pack, prompts, 5, 20, "dropout", "entropy", seed=42).pool_scores
    assert torch.equal(whole, again)
    assert bool(torch.isfinite(whole).all()) and float(whole.min()) >= 0.0 and float(whole.max()) <= math.log(67) + 1e-4
    keys = []
    for lo, hi in ((0, 1), (1, 2500), (2500, 7777), (7777, n)):
        part = engine.mc_score_transformer(pack, prompts[:, lo:hi].contiguous(), 5, 20, "dropout", "entropy", seed=42,
                                           index_base=lo).pool_scores
        assert torch.equal(part, whole[lo:hi])
        keys.append(engine.topk_keys(part, min(k, hi - lo), index_base=lo))
    _, got = engine.unpack_keys(engine.merge_topk_keys(torch.cat(keys), k))
    _, want = engine.unpack_keys(engine.topk_keys(whole, k))
    assert torch.equal(got, want)


def test_select_samples_on_a_char_level_pool_runs_the_fused_kernel(engine, dev):
    """SURVEY C1 through the unchanged list-of-pairs API: 32-character windows + 4-character continuations under a
    character tokenizer (V = 67), MC-dropout entropy acquisition — tokenised by the vectorised char encoder, scored by
    the long-sequence variant of the fused kernel; the exact fp32 path ranks the same pool consistently."""
    import random
    import string

    from bayesformer_b200.libs import active_learning as al
    from bayesformer_b200.libs.tokenizer import CharVocabTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    rnd = random.Random(3)
    alphabet = (string.ascii_letters + " ,.;:!?'\n-&3$")[:65]
    tok = CharVocabTokenizer(alphabet)
    assert tok.ntokens == 67
    text = "".join(rnd.choice(alphabet) for _ in range(36 * 600))
    pool = [(text[36 * i:36 * i + 32], text[36 * i + 32:36 * i + 36]) for i in range(600)]
    torch.manual_seed(1)
    model = MyTransformer(67, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
    with engine.profile(dev) as prof:
        picked = al.select_samples(model, tok, pool, dev, nb_samples=50, compute="dropout", mode="entropy")
    assert "mc_forward_kernel" in prof.kernels and "sgemm_tn_kernel" not in prof.kernels
    assert len(set(picked)) == 50 and all(0 <= i < 600 for i in picked)


@pytest.mark.parametrize("bayes", [True, False])
def test_specialised_kernels_equal_the_generic_kernel_bit_for_bit(engine, dev, c2, bayes):
    """The FAST instantiations (default sites / no site, Philox masks, greedy decoding, no side channels) compile the
    other branches out; asking for the log-probs routes the same call through the generic instantiation.  Same seed
    => identical masks => the scores and the decoded tokens must agree exactly, for short and long sequences
    (one streamed pair phase serves every length up to 128)."""
    z, sd = c2
    model = H.build_model(sd, float(z["p"]) if bayes else None, bayes).to(dev)
    pack = model.native_pack(dev, "fp16")
    for P, N in ((4, 5), (2, 3), (11, 4), (20, 3)):
        prompts = torch.randint(0, 114, (P, 77), generator=torch.Generator().manual_seed(P)).to(dev)
        for scheme, T in ((("dropout", 6),) if bayes else (("greedy", 1), ("sampling", 4))):
            for mode in H.MODES:
                fast = engine.mc_score_transformer(pack, prompts, N, T, scheme, mode, seed=123, return_tokens=True)
                slow = engine.mc_score_transformer(pack, prompts, N, T, scheme, mode, seed=123, return_tokens=True,
                                                   return_logp=True)
                assert torch.equal(fast.step_scores, slow.step_scores)
                assert torch.equal(fast.tokens, slow.tokens)
