"""GPU parity tests: the CUDA path, called through the C ABI (bayesformer_b200.engine -> ctypes ->
libbayesformer_b200.so), against the CPU oracle and the golden vectors recorded from the unmodified
reference.  Tolerances (BASELINE.json north_star): 1e-3 relative for fp32 log-probs / probabilities /
scores; integer outputs (tokens, top-k indices) bit-exact wherever the oracle's decision gap exceeds
that tolerance.
"""

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import restatement as R
from tests import helpers as H

pytestmark = pytest.mark.gpu

RTOL = 1e-3          # north_star fp32 tolerance
ATOL_LOGP = 2e-4     # absolute floor for log-probs near 0
ATOL_SCORE = 2e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engine():
    from bayesformer_b200 import _lib, engine as e

    _lib.load()  # fail loudly if the CUDA library is missing
    return e


def _bits(a: np.ndarray, dev) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


# ----------------------------------------------------------------------------------------------
# compute_uncertainty / acquire
# ----------------------------------------------------------------------------------------------


@pytest.mark.parametrize("mode", H.MODES)
def test_compute_uncertainty_golden(engine, dev, mode):
    z = H.load("uncertainty.npz")
    for key, suffix in (("probs", ""), ("probs2", "2")):
        got = engine.compute_uncertainty(torch.from_numpy(z[key]).to(dev), mode).cpu().numpy()
        np.testing.assert_allclose(got, z[mode + suffix], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n_classes", [1, 2, 3, 31, 32, 33, 67, 114, 128, 129, 300, 1024])
def test_compute_uncertainty_shapes(engine, dev, n_classes):
    g = torch.Generator().manual_seed(n_classes)
    probs = torch.softmax(torch.randn(257, n_classes, generator=g) * 3.0, dim=-1)
    for mode in H.MODES:
        if n_classes == 1 and mode == "margin":
            continue  # the reference indexes sorted[:, 1] -> IndexError
        got = engine.compute_uncertainty(probs.to(dev), mode).cpu().numpy()
        np.testing.assert_allclose(got, R.compute_uncertainty(probs, mode).numpy(), rtol=1e-5, atol=1e-6)


def test_compute_uncertainty_edge_cases(engine, dev):
    with pytest.raises(ValueError):
        engine.compute_uncertainty(torch.rand(4, 3, device=dev), "bogus")
    assert engine.compute_uncertainty(torch.empty(0, 5, device=dev), "max").shape == (0,)
    # ties and one-hot rows
    p = torch.tensor([[0.5, 0.5, 0.0], [1.0, 0.0, 0.0], [0.25, 0.25, 0.5]], device=dev)
    assert engine.compute_uncertainty(p, "margin").cpu().tolist() == [0.0, 1.0, 0.25]
    assert engine.compute_uncertainty(p, "max").cpu().tolist() == [0.5, 1.0, 0.5]
    np.testing.assert_allclose(engine.compute_uncertainty(p, "entropy").cpu().numpy(),
                               R.compute_uncertainty(p.cpu(), "entropy").numpy(), rtol=1e-6, atol=1e-7)
    with pytest.raises(RuntimeError):
        engine.compute_uncertainty(torch.rand(4, 3), "max")  # host tensor: no CPU path


@pytest.mark.parametrize("is_logp", [False, True])
@pytest.mark.parametrize("shape", [(20, 100, 114), (1, 7, 2), (50, 33, 67), (3, 5, 300),
                                   # >= 128 items: TMA-staged kernel on the full tiles + warp-per-item tail
                                   (20, 1000, 114), (5, 132, 67), (7, 300, 2), (3, 256, 128), (4, 384, 33), (2, 130, 114),
                                   (3, 129, 67)])
def test_acquire_matches_oracle(engine, dev, is_logp, shape):
    g = torch.Generator().manual_seed(sum(shape))
    logits = torch.randn(*shape, generator=g) * 2.0
    x = torch.log_softmax(logits, -1) if is_logp else torch.softmax(logits, -1)
    got = engine.acquire(x.to(dev), is_logp=is_logp)
    want = R.acquire(x, is_logp)
    np.testing.assert_allclose(got.mean_p.cpu().numpy(), want["mean_p"].numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got.score_of_mean.cpu().numpy(), want["score_of_mean"].numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got.mean_of_scores.cpu().numpy(), want["mean_of_scores"].numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("is_logp", [False, True])
def test_acquire_full_size_properties(engine, dev, is_logp):
    """(T, B, C) = (20, 100000, 114) — one decode step of BASELINE configs[1]: tile-aligned slices of the pool reproduce the
    whole reduction bit for bit, arbitrary slices to rounding, the predictive mean is a distribution, scores stay in range, and the torch expression of
    compute_uncertainty (active_learning.py:26-31) on the device agrees."""
    import math

    g = torch.Generator(device=dev).manual_seed(3)
    T, B, Cn = 20, 100000, 114
    logp = torch.log_softmax(torch.randn(T, B, Cn, generator=g, device=dev) * 2.0, -1)
    x = logp if is_logp else logp.exp()
    whole = engine.acquire(x, is_logp=is_logp)
    assert float((whole.mean_p.sum(-1) - 1.0).abs().max()) < 1e-5
    p = logp.exp()
    ent = -(p * (p + 1e-10).log()).sum(-1).mean(0)
    top2 = p.topk(2, dim=-1).values
    want = torch.stack([top2[..., 0].mean(0), (top2[..., 0] - top2[..., 1]).mean(0), ent])
    assert float((whole.mean_of_scores - want).abs().max()) < 2e-5
    assert float(whole.mean_of_scores[2].max()) <= math.log(Cn) + 1e-4 and float(whole.mean_of_scores.min()) >= 0.0
    # slices that start on a 128-item tile boundary: bit for bit
    for lo, hi in ((0, 768), (768, 50048), (50048, B)):
        part = engine.acquire(x[:, lo:hi].contiguous(), is_logp=is_logp)
        assert torch.equal(part.mean_of_scores, whole.mean_of_scores[:, lo:hi])
        assert torch.equal(part.mean_p, whole.mean_p[lo:hi])
        assert torch.equal(part.score_of_mean, whole.score_of_mean[:, lo:hi])
    # any other slice: the class sums of an item are taken in an order that depends on its place in the tile (the
    # bank-conflict-free rotation of the staged rows) or by the warp-per-item tail kernel — equal to rounding only
    # (the scoring paths reduce per sequence and ARE slice-invariant bit for bit)
    for lo, hi in ((0, 777), (777, 50001), (50001, B)):
        part = engine.acquire(x[:, lo:hi].contiguous(), is_logp=is_logp)
        assert float((part.mean_of_scores - whole.mean_of_scores[:, lo:hi]).abs().max()) < 2e-6
        assert float((part.mean_p - whole.mean_p[lo:hi]).abs().max()) < 1e-7
        assert float((part.score_of_mean - whole.score_of_mean[:, lo:hi]).abs().max()) < 2e-6


# ----------------------------------------------------------------------------------------------
# top-k: bit-exact
# ----------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,k", [(1, 1), (7, 7), (200, 200), (1000, 200), (8192, 200), (8193, 200), (100000, 200),
                                 (1000000, 200), (300000, 4096), (20000, 1)])
def test_topk_bit_exact(engine, dev, n, k):
    rng = np.random.default_rng(n + k)
    scores = rng.standard_normal(n).astype(np.float32)
    scores[rng.integers(0, n, size=max(1, n // 10))] = scores[0]  # plenty of exact ties
    vals, idx = engine.topk(torch.from_numpy(scores).to(dev), k)
    want = R.topk_lowest_index(scores, k)
    assert idx.cpu().numpy().tolist() == want.tolist()
    assert np.array_equal(vals.cpu().numpy(), scores[want])


def test_topk_special_values_and_errors(engine, dev):
    s = np.array([1.0, 5.0, 5.0, 5.0, 2.0, -0.0, 0.0, -3.0, np.inf, -np.inf], dtype=np.float32)
    vals, idx = engine.topk(torch.from_numpy(s).to(dev), 10)
    # -0.0 orders below +0.0 in the key space; torch.topk treats them as equal: accept either order there
    got = idx.cpu().tolist()
    assert got[:5] == [8, 1, 2, 3, 4] and set(got[5:8]) == {0, 5, 6} and got[8:] == [7, 9]
    with pytest.raises(RuntimeError):
        engine.topk(torch.zeros(5, device=dev), 6)  # torch.topk: "selected index k out of range"
    all_equal = torch.full((50000,), 0.25, device=dev)
    assert engine.topk(all_equal, 200)[1].cpu().tolist() == list(range(200))


@pytest.mark.parametrize("kind", ["normal", "all_equal", "wide", "few_values", "narrow"])
@pytest.mark.parametrize("n,k,base", [(16385, 200, 0), (1000000, 200, 7), (2000003, 1000, 123456), (40000, 4096, 0)])
def test_topk_one_launch_path(engine, dev, kind, n, k, base):
    """Scores beyond one 16,384-key chunk take the cooperative one-launch radix select over global histograms
    (csrc/topk.cu topk_global_kernel): bit-exact against the rule of torch.topk with ties towards the lowest index, on
    distributions that stress each stage — every key equal (all six rounds, down to the index bits), values spread over
    the whole exponent range (thousands of non-empty bins per CTA), a handful of distinct values (huge tie bins), and a
    narrow band (candidates cut only by the low mantissa bits)."""
    rng = np.random.default_rng(n + k)
    if kind == "normal":
        scores = rng.standard_normal(n).astype(np.float32)
    elif kind == "all_equal":
        scores = np.full(n, 0.69314718, dtype=np.float32)
    elif kind == "wide":
        scores = (rng.standard_normal(n) * np.exp(rng.uniform(-40, 40, n))).astype(np.float32)
    elif kind == "few_values":
        scores = rng.integers(0, 5, n).astype(np.float32) * 0.25
    else:
        scores = (1.0 + rng.random(n) * 1e-4).astype(np.float32)
    keys = engine.topk_keys(torch.from_numpy(scores).to(dev), k, index_base=base)
    vals, idx = engine.unpack_keys(keys)
    want = R.topk_lowest_index(scores, k)
    assert (idx.cpu().numpy() - base).tolist() == want.tolist()
    assert np.array_equal(vals.cpu().numpy(), scores[want])


def test_topk_and_scoring_on_two_devices_in_one_process(engine):
    """Kernel attributes (dynamic shared memory opt-in, cooperative launch, SM count) belong to a device, not to the
    process: the same process drives cuda:0 and then cuda:1.  Needs two GPUs (skipped otherwise)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    rng = np.random.default_rng(4)
    scores = rng.standard_normal(300000).astype(np.float32)
    want = R.topk_lowest_index(scores, 4096)
    small = scores[:9000]
    want_small = R.topk_lowest_index(small, 300)
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        _, idx = engine.topk(torch.from_numpy(scores).to(d), 4096)          # cooperative one-launch path
        assert idx.cpu().numpy().tolist() == want.tolist()
        _, idx = engine.topk(torch.from_numpy(small).to(d), 300)            # one-block path (128 KB of dynamic smem)
        assert idx.cpu().numpy().tolist() == want_small.tolist()
    z = H.load("c2_transformer.npz")
    sd = H.state_dict_from(z)
    outs = []
    for d in ("cuda:0", "cuda:1"):
        model = H.build_model(sd, float(z["p"]), True).to(d)
        prompts = torch.from_numpy(z["prompts"]).to(d)
        res = engine.mc_score_transformer(engine.scoring_pack(model, torch.device(d), prompts.shape[0], 5), prompts, 5, 4,
                                          "dropout", "margin", seed=3)
        outs.append(res.pool_scores.cpu())
    assert torch.equal(outs[0], outs[1])


def test_topk_sharded_merge_equals_global(engine, dev):
    """per-shard keys with global index bases + merge == single-device top-k (the multi-GPU exchange)."""
    rng = np.random.default_rng(5)
    n, k, G = 100003, 200, 8
    scores = rng.standard_normal(n).astype(np.float32)
    scores[::97] = 0.5
    bounds = [n * g // G for g in range(G + 1)]
    keys = [engine.topk_keys(torch.from_numpy(scores[a:b]).to(dev), k, index_base=a)
            for a, b in zip(bounds[:-1], bounds[1:])]
    merged = engine.merge_topk_keys(torch.cat(keys), k)
    vals, idx = engine.unpack_keys(merged)
    want = R.topk_lowest_index(scores, k)
    assert idx.cpu().tolist() == want.tolist()
    assert np.array_equal(vals.cpu().numpy(), scores[want])
    want_keys = R.pack_key(scores[want], want)
    assert np.array_equal(merged.cpu().numpy().view(np.uint64), want_keys)


# ----------------------------------------------------------------------------------------------
# toy classifiers
# ----------------------------------------------------------------------------------------------


def test_mlp_mc_golden_masks(engine, dev):
    z = H.load("classifiers.npz")
    w = [torch.from_numpy(z["mlp.sd." + k]).to(dev) for k in
         ("hidden_layer.weight", "hidden_layer.bias", "output_layer.weight", "output_layer.bias")]
    out = engine.mlp_mc_score(torch.from_numpy(z["x"]).to(dev), *w, float(z["mlp.p"]), 20,
                              mask_bits=_bits(z["mlp.keep_bits"], dev))
    np.testing.assert_allclose(out.probs.cpu().numpy(), z["mlp.probs"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.mean_p.cpu().numpy(), z["mlp.mean"], rtol=1e-4, atol=1e-6)
    want = R._classifier_reduce(torch.from_numpy(z["mlp.probs"]))
    np.testing.assert_allclose(out.score_of_mean.cpu().numpy(), want["score_of_mean"].numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(out.mean_of_scores.cpu().numpy(), want["mean_of_scores"].numpy(), rtol=1e-3, atol=1e-5)


def test_mlp_full_size_pool_properties(engine, dev):
    """BASELINE configs[2] size for the classifier (1,000,000 items, T = 20), through properties that hold at any size:
    a re-run is bit-identical, eight shards with global index bases reproduce the unsharded draws bit for bit and
    acquire the same items, the predictive mean is the mean of the draws, and the two-class scores follow
    compute_uncertainty (active_learning.py:26-31) applied to [1 - p, p]."""
    z = H.load("classifiers.npz")
    w = [torch.from_numpy(z["mlp.sd." + k]).to(dev) for k in
         ("hidden_layer.weight", "hidden_layer.bias", "output_layer.weight", "output_layer.bias")]
    n, T, p, seed, G, k = 1000000, 20, 0.5, 42, 8, 200
    x = (torch.randn(n, 2, generator=torch.Generator().manual_seed(42)) * 0.8).to(dev)
    full = engine.mlp_mc_score(x, *w, p, T, seed=seed)
    again = engine.mlp_mc_score(x, *w, p, T, seed=seed)
    assert torch.equal(full.probs, again.probs) and torch.equal(full.mean_of_scores, again.mean_of_scores)
    assert float(full.probs.min()) >= 0.0 and float(full.probs.max()) <= 1.0
    pm = full.probs.double().mean(0)
    assert float((full.mean_p.double() - pm).abs().max()) < 1e-6
    two = torch.stack([1.0 - pm, pm], dim=1)
    want = torch.stack([two.max(1).values, (two[:, 0] - two[:, 1]).abs(), -(two * (two + 1e-10).log()).sum(1)])
    assert float((full.score_of_mean.double() - want).abs().max()) < 2e-6
    keys = []
    for g in range(G):
        lo, hi = n * g // G, n * (g + 1) // G
        part = engine.mlp_mc_score(x[lo:hi], *w, p, T, seed=seed, index_base=lo)
        assert torch.equal(part.probs, full.probs[:, lo:hi])
        assert torch.equal(part.mean_of_scores, full.mean_of_scores[:, lo:hi])
        keys.append(engine.topk_keys(part.mean_of_scores[1].contiguous(), k, index_base=lo))
    _, got = engine.unpack_keys(engine.merge_topk_keys(torch.cat(keys), k))
    _, idx = engine.unpack_keys(engine.topk_keys(full.mean_of_scores[1].contiguous(), k))
    assert torch.equal(got, idx)


def test_mlp_mc_philox_matches_oracle_and_is_shard_invariant(engine, dev):
    z = H.load("classifiers.npz")
    w_cpu = [torch.from_numpy(z["mlp.sd." + k]) for k in
             ("hidden_layer.weight", "hidden_layer.bias", "output_layer.weight", "output_layer.bias")]
    w = [t.to(dev) for t in w_cpu]
    x = torch.from_numpy(z["x"])
    B, T, p, seed = x.shape[0], 20, 0.5, 0x1234ABCD5678
    keep = torch.from_numpy(philox.mlp_keep(seed, p, np.arange(1000, 1000 + B), T, 50))
    want = R.mlp_mc(x, *w_cpu, p, keep)
    full = engine.mlp_mc_score(x.to(dev), *w, p, T, seed=seed, index_base=1000)
    np.testing.assert_allclose(full.probs.cpu().numpy(), want["probs"].numpy(), rtol=1e-4, atol=1e-6)
    part = engine.mlp_mc_score(x[100:180].to(dev), *w, p, T, seed=seed, index_base=1100)
    assert torch.equal(part.probs, full.probs[:, 100:180])  # bit-identical across shards


def test_vmlp_mc_golden(engine, dev):
    z = H.load("classifiers.npz")
    args = [torch.from_numpy(z["vmlp.sd." + k]).to(dev) for k in
            ("hidden_layer.w_mu", "hidden_layer.w_rho", "hidden_layer.b_mu", "output_layer.w_mu",
             "output_layer.w_rho", "output_layer.b_mu")]
    out = engine.vmlp_mc_score(torch.from_numpy(z["x"]).to(dev), *args, torch.from_numpy(z["vmlp.eps"]).to(dev))
    np.testing.assert_allclose(out.probs.cpu().numpy(), z["vmlp.probs"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(out.mean_p.cpu().numpy(), z["vmlp.mean"], rtol=1e-3, atol=1e-5)


def test_vmlp_full_size_is_shard_invariant(engine, dev):
    """VariationalMLP at 1,000,000 items: every shard uses the same weight draws (bayesian_classification.py:76 samples
    once per call for the whole batch), so shards reproduce the unsharded probabilities bit for bit."""
    z = H.load("classifiers.npz")
    args = [torch.from_numpy(z["vmlp.sd." + k]).to(dev) for k in
            ("hidden_layer.w_mu", "hidden_layer.w_rho", "hidden_layer.b_mu", "output_layer.w_mu",
             "output_layer.w_rho", "output_layer.b_mu")]
    eps = torch.from_numpy(z["vmlp.eps"]).to(dev)
    n = 1000000
    x = (torch.randn(n, 2, generator=torch.Generator().manual_seed(7)) * 0.8).to(dev)
    full = engine.vmlp_mc_score(x, *args, eps)
    assert torch.equal(full.probs, engine.vmlp_mc_score(x, *args, eps).probs)
    assert float(full.probs.min()) >= 0.0 and float(full.probs.max()) <= 1.0
    assert float((full.mean_p.double() - full.probs.double().mean(0)).abs().max()) < 1e-6
    for lo, hi in ((0, 333333), (333333, 333334), (333334, n)):
        part = engine.vmlp_mc_score(x[lo:hi], *args, eps)
        assert torch.equal(part.probs, full.probs[:, lo:hi]) and torch.equal(part.mean_p, full.mean_p[lo:hi])


def test_classifier_modules_route_to_native(engine, dev):
    from bayesformer_b200.model.bayesian_classification import MLP

    torch.manual_seed(0)
    m = MLP(2, 50, 1, dropout=True, dropout_rate=0.5).to(dev).eval()
    x = torch.randn(4096, 2, device=dev)
    with torch.no_grad():
        res = m.mc_predict(x, nsamples=20, return_scores=True)
    assert res.mean_p.shape == (4096,) and bool(((res.mean_p >= 0) & (res.mean_p <= 1)).all())
    # statistical check against the deterministic network: E[dropout(h)] = h
    with torch.no_grad():
        res_many = m.mc_predict(x[:256], nsamples=2000, return_scores=True)
        h = torch.relu(m.hidden_layer(x[:256]))
        lo = torch.sigmoid(m.output_layer(h)).squeeze(-1)
    assert float((res_many.mean_p - lo).abs().mean()) < 0.05


# ----------------------------------------------------------------------------------------------
# transformer forward (mask injection)
# ----------------------------------------------------------------------------------------------


def _native_forward(engine, dev, sd, p, bayes, src, bits, latent=False):
    model = H.build_model(sd, p, bayes)
    if latent:
        for layer in model.layers:
            layer.bayes, layer.bayes_dropout = True, p
            layer.ff.bayes_dropout = p
            layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
    model = model.to(dev)
    pack = model.native_pack(dev)
    with torch.no_grad():
        return engine.transformer_forward(pack, torch.from_numpy(src).to(dev),
                                          mask_bits=None if bits is None else _bits(bits, dev)).cpu().numpy()


def test_forward_mid_model(engine, dev):
    z = H.load("mid_model.npz")
    got = _native_forward(engine, dev, H.state_dict_from(z), float(z["p"]), True, z["src"], z["bits"])
    np.testing.assert_allclose(got, z["logp"], rtol=RTOL, atol=ATOL_LOGP)


def test_forward_latent_sites(engine, dev):
    z = H.load("latent_sites.npz")
    got = _native_forward(engine, dev, H.state_dict_from(z), float(z["p"]), True, z["src"], z["bits"], latent=True)
    np.testing.assert_allclose(got, z["logp"], rtol=RTOL, atol=ATOL_LOGP)


def test_forward_head_in_sites(engine, dev):
    """Head.bayes (reference transformer.py:50-53): three independent (l, B, d) input masks per head — the reference's
    own masks injected (tests/golden/head_in_sites.npz), together with the head-output and default sites."""
    z = H.load("head_in_sites.npz")
    sd, p = H.state_dict_from(z), float(z["p"])
    model = H.build_model(sd, p, True)
    for layer in model.layers:
        layer.self_attn.bayes, layer.self_attn.bayes_dropout = True, p
        for head in layer.self_attn.heads:
            head.bayes, head.bayes_dropout = True, p
    model = model.to(dev)
    pack = model.native_pack(dev)          # fp32: the only path that serves this site
    with torch.no_grad():
        got = engine.transformer_forward(pack, torch.from_numpy(z["src"]).to(dev), mask_bits=_bits(z["bits"], dev)).cpu().numpy()
    np.testing.assert_allclose(got, z["logp"], rtol=RTOL, atol=ATOL_LOGP)
    with pytest.raises(NotImplementedError):
        model.native_pack(dev, "fp16")
    # Philox mode: the library's counter scheme for the new sites against the oracle run with the same masks
    V, d, Hh, F, L = (int(v) for v in z["dims"])
    src = torch.from_numpy(z["src"])
    seed, draw, base = 0xABCDEF12345, 3, 70
    with torch.no_grad():
        got = engine.transformer_forward(pack, src.to(dev), seed=seed, draw=draw, index_base=base).cpu().numpy()
    live = ({R.SITE_EMBED} | {R.site_layer_out(i) for i in range(L)} | {R.site_head_out(i) for i in range(L)}
            | {R.site_head_in(i, h, c, L, Hh) for i in range(L) for h in range(Hh) for c in range(3)})
    B = src.shape[1]
    want = R.transformer_forward(sd, src, p, R.philox_mask_fn(seed, p, np.arange(base, base + B), np.full(B, draw), 0, d, live))
    np.testing.assert_allclose(got, want.numpy(), rtol=RTOL, atol=ATOL_LOGP)


def test_forward_deterministic_model_matches_golden_greedy_step0(engine, dev):
    z = H.load("c2_transformer.npz")
    got = _native_forward(engine, dev, H.state_dict_from(z), None, False, z["prompts"], None)
    np.testing.assert_allclose(got[-1], z["greedy.logp"][0], rtol=RTOL, atol=ATOL_LOGP)
    # rows of the full (len, B, V) output are log-distributions
    np.testing.assert_allclose(np.exp(got).sum(-1), 1.0, rtol=1e-4)


def test_module_forward_dispatches_native_and_rejects_nothing_silently(engine, dev):
    z = H.load("c2_transformer.npz")
    sd = H.state_dict_from(z)
    model = H.build_model(sd, None, False).to(dev)
    src = torch.from_numpy(z["prompts"]).to(dev)
    lib = __import__("bayesformer_b200._lib", fromlist=["load"]).load()
    lib.bfb200_reset_launch_count()
    with torch.no_grad():
        out = model(src)
    assert lib.bfb200_launch_count() > 0, "model(src) under no_grad on CUDA must run the native library"
    want = R.transformer_forward(sd, torch.from_numpy(z["prompts"]))
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL_LOGP)


def test_scaled_c4_forward(engine, dev):
    """d=512, H=8 (dh=64), F=2048, L=6, len=256 against the reference's stored outputs; weights and masks
    are regenerated from their seeds with this repo's module (same constructor RNG order as the reference,
    pinned by the stored weight probe and — through the masks — by the outputs themselves)."""
    from bayesformer_b200.model.transformer import MyTransformer
    from oracle import ref_shim

    z = H.load("scaled_c4.npz")
    V, d, Hh, F, L = (int(v) for v in z["dims"])
    p = float(z["p"])
    torch.manual_seed(int(z["weight_seed"]))
    model = MyTransformer(V, d, Hh, F, L, bayes_dropout=p, bayes=True).eval()
    np.testing.assert_array_equal(model.layers[5].ff.net[2].weight.detach().numpy()[:4, :8], z["probe"])
    src = torch.from_numpy(z["src"])
    log = []
    torch.manual_seed(int(z["mask_seed"]))
    with torch.no_grad(), ref_shim.record_dropout(log):
        cpu_out = model._module_forward(src)
    np.testing.assert_allclose(cpu_out[-1].numpy(), z["logp_last"], rtol=1e-4, atol=1e-5)  # masks regenerated OK
    pm = R.PackedMasks(1, L, src.shape[1], src.shape[0], d)
    pm.set(0, R.SITE_EMBED, log[0].numpy())
    for i in range(L):
        pm.set(0, R.site_layer_out(i), log[1 + i].numpy())
    model = model.to(dev)
    with torch.no_grad():
        got = engine.transformer_forward(model.native_pack(dev), src.to(dev), mask_bits=_bits(pm.bits, dev)).cpu().numpy()
    np.testing.assert_allclose(got[-1], z["logp_last"], rtol=RTOL, atol=ATOL_LOGP)
    np.testing.assert_allclose(got[::37], z["logp_rows"], rtol=RTOL, atol=ATOL_LOGP)


# ----------------------------------------------------------------------------------------------
# the MC scoring loop
# ----------------------------------------------------------------------------------------------


@pytest.fixture(scope="module")
def c2(dev):
    z = H.load("c2_transformer.npz")
    sd = H.state_dict_from(z)
    return z, sd


def _decisive(logp: torch.Tensor, tol: float) -> torch.Tensor:
    """rows whose top-2 log-prob gap exceeds tol * |logp| (greedy argmax is then tolerance-proof)."""
    top2 = torch.topk(logp, 2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > tol * top2[..., 0].abs().clamp_min(1e-3)


@pytest.mark.parametrize("mode", H.MODES)
def test_dropout_scheme_with_reference_masks(engine, dev, c2, mode):
    z, sd = c2
    P, N, T, L, d, p = int(z["P"]), int(z["N"]), int(z["T"]), 2, 64, float(z["p"])
    B = z["prompts"].shape[1]
    bits = H.expand_live_bits(z["dropout.keep_bits"], L)
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, T, "dropout",
                                      mode, mask_bits=_bits(bits, dev), mask_max_len=bits.shape[3],
                                      return_logp=True, return_tokens=True)
    pm = H.packed_masks_from_bits(bits, L, d)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", mode, p,
                       lambda step: pm.mask_fn(step, R.default_live_sites(L, True)))
    got_tok = res.tokens.cpu().long()
    # trajectories: identical wherever every decision of the oracle was decisive at 1e-3
    decisive = _decisive(want["logp"], RTOL).all(dim=0)  # (S,)
    assert decisive.float().mean() > 0.8
    assert torch.equal(got_tok[decisive], want["tokens"][decisive])
    # log-probs of sequences whose trajectory matched: 1e-3
    same = (got_tok == want["tokens"]).all(dim=1)
    np.testing.assert_allclose(res.logp.cpu().numpy()[:, same.numpy()], want["logp"].numpy()[:, same.numpy()],
                               rtol=RTOL, atol=ATOL_LOGP)
    # scores: step 0 is trajectory independent; all steps agree when all trajectories agree
    np.testing.assert_allclose(res.step_scores[0].cpu().numpy(), z[f"dropout.scores.{mode}"][0], rtol=RTOL, atol=ATOL_SCORE)
    items_same = same.view(B, T).all(dim=1).numpy()
    assert items_same.mean() > 0.8
    np.testing.assert_allclose(res.step_scores.cpu().numpy()[:, items_same], z[f"dropout.scores.{mode}"][:, items_same],
                               rtol=RTOL, atol=ATOL_SCORE)
    np.testing.assert_allclose(res.pool_scores.cpu().numpy()[items_same],
                               z[f"dropout.scores.{mode}"].mean(0)[items_same], rtol=RTOL, atol=ATOL_SCORE)


@pytest.mark.parametrize("mode", H.MODES)
def test_dropout_scheme_teacher_forced(engine, dev, c2, mode):
    """Token-injection mode: every appended token forced to the reference's -> all steps comparable."""
    z, sd = c2
    P, N, T, L, p = int(z["P"]), int(z["N"]), int(z["T"]), 2, float(z["p"])
    B = z["prompts"].shape[1]
    bits = H.expand_live_bits(z["dropout.keep_bits"], L)
    forced = torch.from_numpy(z["dropout.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)  # (N, item*T+draw)
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, T, "dropout",
                                      mode, mask_bits=_bits(bits, dev), mask_max_len=bits.shape[3],
                                      forced_tokens=forced, return_logp=True)
    np.testing.assert_allclose(res.step_scores.cpu().numpy(), z[f"dropout.scores.{mode}"], rtol=RTOL, atol=ATOL_SCORE)
    lp = res.logp.cpu().view(N, B, T, -1).permute(2, 0, 1, 3)[:6].numpy()
    np.testing.assert_allclose(lp, z["dropout.logp_first_draws"], rtol=RTOL, atol=ATOL_LOGP)


@pytest.mark.parametrize("mode", H.MODES)
def test_greedy_scheme(engine, dev, c2, mode):
    z, sd = c2
    N = int(z["N"])
    model = H.build_model(sd, None, False).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, 1, "greedy",
                                      mode, return_logp=True)
    want_logp = torch.from_numpy(z["greedy.logp"])
    got_logp = res.logp.cpu()
    ok = _decisive(want_logp, RTOL).all(dim=0)
    np.testing.assert_allclose(got_logp.numpy()[:, ok.numpy()], want_logp.numpy()[:, ok.numpy()], rtol=RTOL, atol=ATOL_LOGP)
    np.testing.assert_allclose(res.step_scores.cpu().numpy()[:, ok.numpy()], z[f"greedy.scores.{mode}"][:, ok.numpy()],
                               rtol=RTOL, atol=ATOL_SCORE)


@pytest.mark.parametrize("mode", H.MODES)
def test_sampling_scheme_forced_tokens(engine, dev, c2, mode):
    z, sd = c2
    N, T = int(z["N"]), int(z["T"])
    B = z["prompts"].shape[1]
    forced = torch.from_numpy(z["sampling.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)
    model = H.build_model(sd, None, False).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, T, "sampling",
                                      mode, temperature=0.8, forced_tokens=forced)
    np.testing.assert_allclose(res.step_scores.cpu().numpy(), z[f"sampling.scores.{mode}"], rtol=RTOL, atol=ATOL_SCORE)


def test_sampling_scheme_philox_tokens_match_oracle_rule(engine, dev, c2):
    z, sd = c2
    N, T, seed = int(z["N"]), 6, 777
    B = z["prompts"].shape[1]
    model = H.build_model(sd, None, False).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, T, "sampling",
                                      "entropy", temperature=0.8, seed=seed, index_base=50, return_tokens=True,
                                      return_logp=True)
    items = np.repeat(np.arange(50, 50 + B), T)
    draws = np.tile(np.arange(T), B)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "sampling", "entropy", 0.0, lambda step: None,
                       temperature=0.8, sampling_u=lambda step: philox.sampling_uniform(seed, items, draws, step))
    same = (res.tokens.cpu().long() == want["tokens"]).all(dim=1)
    assert same.float().mean() > 0.9  # a CDF boundary within float noise of u may flip a token
    np.testing.assert_allclose(res.logp.cpu().numpy()[:, same.numpy()], want["logp"].numpy()[:, same.numpy()],
                               rtol=RTOL, atol=ATOL_LOGP)


def test_philox_mode_matches_oracle_with_same_counter_scheme(engine, dev, c2):
    """Free-running mode: the library's Philox masks restated in numpy (oracle/philox.py) -> same masks,
    so the run is comparable element-wise with the oracle."""
    z, sd = c2
    P, N, T, L, d, p, seed, base = int(z["P"]), int(z["N"]), 5, 2, 64, float(z["p"]), 0xC0FFEE1234, 4000
    B = z["prompts"].shape[1]
    model = H.build_model(sd, p, True).to(dev)
    res = engine.mc_score_transformer(model.native_pack(dev), torch.from_numpy(z["prompts"]).to(dev), N, T, "dropout",
                                      "entropy", seed=seed, index_base=base, return_logp=True, return_tokens=True)
    items = np.repeat(np.arange(base, base + B), T)
    draws = np.tile(np.arange(T), B)
    live = R.default_live_sites(L, True)
    want = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", "entropy", p,
                       lambda step: R.philox_mask_fn(seed, p, items, draws, step, d, live))
    same = (res.tokens.cpu().long() == want["tokens"]).all(dim=1)
    assert same.float().mean() > 0.85
    np.testing.assert_allclose(res.logp.cpu().numpy()[:, same.numpy()], want["logp"].numpy()[:, same.numpy()],
                               rtol=RTOL, atol=ATOL_LOGP)
    np.testing.assert_allclose(res.step_scores[0].cpu().numpy(), want["step_scores"][0].numpy(), rtol=RTOL, atol=ATOL_SCORE)


def test_philox_mode_is_invariant_to_chunking_and_sharding(engine, dev, c2):
    """Same seed -> bit-identical scores whether the pool is scored whole, in small chunks, or as shards
    with their global index base (what makes 1-GPU and 8-GPU runs select the same items)."""
    z, sd = c2
    N, T, p, seed = int(z["N"]), 20, float(z["p"]), 99
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    B = prompts.shape[1]
    model = H.build_model(sd, p, True).to(dev)
    pack = model.native_pack(dev)
    whole = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=seed)
    import ctypes as C
    from bayesformer_b200 import _lib
    lib = _lib.load()
    args = _lib.McArgs(n_items=B, prompt_len=prompts.shape[0], new_tokens=N, n_draws=T)
    small = lib.bfb200_mc_workspace_bytes(C.byref(pack.desc), C.byref(args), 5)
    chunked = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=seed, workspace_bytes=small)
    assert torch.equal(whole.step_scores, chunked.step_scores)
    assert torch.equal(whole.pool_scores, chunked.pool_scores)
    a, b = 7, 23
    shard = engine.mc_score_transformer(pack, prompts[:, a:b].contiguous(), N, T, "dropout", "margin", seed=seed,
                                        index_base=a)
    assert torch.equal(shard.step_scores, whole.step_scores[:, a:b])
    other = engine.mc_score_transformer(pack, prompts, N, T, "dropout", "margin", seed=seed + 1)
    assert not torch.equal(other.step_scores, whole.step_scores)


def test_philox_keep_rate(engine, dev):
    """keep-rate of the in-kernel masks: a model whose embedding is all ones exposes the mask directly."""
    from bayesformer_b200.model.transformer import MyTransformer

    torch.manual_seed(0)
    m = MyTransformer(16, 64, 8, 8, 0, bayes_dropout=0.3, bayes=True).eval()
    with torch.no_grad():
        m.embedding.weight.fill_(1.0)
        m.linear_out.weight.zero_()
        m.linear_out.weight[0].fill_(1.0)  # logit 0 = sum of kept features * scale * sqrt(d) + sum(pe)
        m.linear_out.bias.zero_()
    # with L = 0 the head sees x = keep/(1-p)*8 + pe; count kept features through logit 0 - logit 1
    m = m.to(dev)
    src = torch.zeros(1, 4096, dtype=torch.int64, device=dev)
    with torch.no_grad():
        pack = m.native_pack(dev)
        logp = engine.transformer_forward(pack, src, seed=3, draw=0)
    diff = (logp[0, :, 0] - logp[0, :, 1]).cpu().numpy()  # = logit0 - 0
    pe_sum = float(m.pos_enc.pe[0, 0].sum())
    kept = (diff - pe_sum) / (8.0 / 0.7)
    assert abs(kept.mean() / 64.0 - 0.7) < 0.01
    assert np.allclose(kept, np.round(kept), atol=1e-2)


# ----------------------------------------------------------------------------------------------
# select_samples through the reference-shaped API
# ----------------------------------------------------------------------------------------------


def test_select_samples_api(engine, dev, c2):
    from bayesformer_b200.libs import active_learning as al

    z, sd = c2
    N, p = int(z["N"]), float(z["p"])
    model = H.build_model(sd, p, True).to(dev)
    with torch.no_grad():
        picked = al.select_samples_from_tokens(model, torch.from_numpy(z["prompts"]), N, dev, nb_samples=8,
                                               compute="dropout", mode="entropy")
    assert len(picked) == 8 and len(set(picked)) == 8
    with pytest.raises(ValueError):
        al.select_samples_from_tokens(model, torch.from_numpy(z["prompts"]), N, dev, 8, compute="bogus")
    with pytest.raises(ValueError):
        al.dropout_uncertainty(model, torch.from_numpy(z["prompts"]).to(dev), N, dev, mode="bogus")
    with pytest.raises(RuntimeError):
        al.select_samples_from_tokens(model, torch.from_numpy(z["prompts"]), N, dev, nb_samples=1000, compute="dropout")
    # dropout_uncertainty returns (N, B) like the reference
    with torch.no_grad():
        u = al.dropout_uncertainty(model, torch.from_numpy(z["prompts"]).to(dev), N, dev, mode="max", nb_samples=4)
    assert tuple(u.shape) == (N, z["prompts"].shape[1]) and u.is_cuda


def test_select_samples_end_to_end_strings(engine, dev, c2):
    """The list-of-pairs entry point: tokenise -> score -> top-k; with the reference's masks unavailable in
    free-running mode, check the statistical agreement with the oracle's ranking on trained weights."""
    import random

    from bayesformer_b200.libs import active_learning as al, preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    p = float(z["p"])
    tok = ReversedPairingTokenizer()
    random.seed(123)
    pool = preprocessing.create_dataset(600, 3)
    model = H.build_model(sd, p, True).to(dev)
    picked = al.select_samples(model, tok, pool, dev, nb_samples=100, compute="dropout", mode="entropy")
    assert len(picked) == 100 and len(set(picked)) == 100
    # oracle ranking with its own (different) masks: overlap must be far above chance (100*100/600 = 16.7)
    prompts, _, P, A, _, _ = preprocessing.get_batch(pool, tok, 0, len(pool))
    g = torch.Generator().manual_seed(1)
    L, d, T = 2, 64, 20

    def step_fn(step):
        def fn(site, n_pos):
            if site not in R.default_live_sites(L, True):
                return None
            return torch.rand(n_pos, prompts.shape[1] * T, d, generator=g) >= p
        return fn

    want = R.mc_scores(sd, prompts, A + 1, T, "dropout", "entropy", p, step_fn)
    ref_pick = set(R.topk_lowest_index(want["pool_scores"].numpy(), 100).tolist())
    assert len(ref_pick & set(picked)) >= 55


def test_host_pointer_is_rejected(engine, dev):
    import ctypes as C
    from bayesformer_b200 import _lib

    lib = _lib.load()
    host = np.zeros(16, dtype=np.float32)
    out = torch.zeros(4, device=dev)
    rc = lib.bfb200_compute_uncertainty(host.ctypes.data, 4, 4, 0, out.data_ptr(), None)
    assert rc == _lib.ERR_NOT_DEVICE_PTR
    assert b"no CPU path" in lib.bfb200_last_error()


def test_generate_and_evaluate_run_natively(engine, dev, c2):
    """utils.generate / utils.evaluate (reference src/libs/utils.py:127-193, the step after acquisition): one native
    greedy-decode call; tokens equal the exact path's wherever its decisions are decisive."""
    import random

    from bayesformer_b200 import _lib
    from bayesformer_b200.libs import preprocessing, utils
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    model = H.build_model(sd, None, False).to(dev)
    prompts = torch.from_numpy(z["prompts"]).to(dev)
    N = int(z["N"])
    lib = _lib.load()
    lib.bfb200_reset_launch_count()
    with torch.no_grad():
        out = utils.generate(model, prompts, N, dev)
    assert lib.bfb200_launch_count() > 0 and tuple(out.shape) == (prompts.shape[0] + N, prompts.shape[1])
    assert torch.equal(out[: prompts.shape[0]], prompts)
    want_logp = torch.from_numpy(z["greedy.logp"])
    ok = _decisive(want_logp, 2e-2).all(dim=0)
    assert torch.equal(out[prompts.shape[0]:, ok].cpu(), want_logp.argmax(-1)[:, ok])
    tok = ReversedPairingTokenizer()
    random.seed(3)
    data = preprocessing.create_dataset(200, 3)
    acc32 = utils.evaluate(model, data, tok, 50, dev)          # deterministic decode: exact fp32 path by default
    model.native_compute = "fp16"                              # opt-in: fused tensor-core kernel
    acc = utils.evaluate(model, data, tok, 50, dev)
    assert 0.0 <= acc <= 1.0 and abs(acc - acc32) <= 0.03


def test_select_samples_accepts_a_pool_tokenised_once(engine, dev, c2):
    """TokenizedPool (tokenise once per pool instead of once per round) selects exactly what the list of pairs does."""
    import random

    from bayesformer_b200.libs import active_learning as al, preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    z, sd = c2
    tok = ReversedPairingTokenizer()
    random.seed(9)
    pool = preprocessing.create_dataset(500, 3)
    model = H.build_model(sd, None, False).to(dev)          # deterministic model: greedy scores are reproducible
    want = al.select_samples(model, tok, pool, dev, nb_samples=40, compute="greedy", mode="margin")
    tp = al.TokenizedPool(pool, tok)
    assert al.select_samples(model, tok, tp, dev, nb_samples=40, compute="greedy", mode="margin") == want
    assert len(tp) == 500 and tp[3] == pool[3]
    assert al.create_new_train_set([("1+1=", "2")], tp, want[:2]) == [("1+1=", "2"), pool[want[0]], pool[want[1]]]
