"""Pins the CPU oracle (oracle/restatement.py) to outputs of the unmodified reference (tests/golden/)."""

import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests import helpers as H

TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.fixture(scope="module")
def c2():
    z = H.load("c2_transformer.npz")
    return z, H.state_dict_from(z)


def _dropout_run(z, sd, mode):
    P, N, T, L, d = int(z["P"]), int(z["N"]), int(z["T"]), 2, 64
    pm = H.packed_masks_from_bits(H.expand_live_bits(z["dropout.keep_bits"], L), L, d)
    live = R.default_live_sites(L, True)
    return R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", mode, float(z["p"]),
                       lambda step: pm.mask_fn(step, live))


@pytest.mark.parametrize("mode", H.MODES)
def test_dropout_scores_match_reference(c2, mode):
    z, sd = c2
    out = _dropout_run(z, sd, mode)
    np.testing.assert_allclose(out["step_scores"].numpy(), z[f"dropout.scores.{mode}"], **TOL)
    T, N = int(z["T"]), int(z["N"])
    B = z["prompts"].shape[1]
    lp = out["logp"].view(N, B, T, -1).permute(2, 0, 1, 3)  # (draw, step, item, V)
    np.testing.assert_allclose(lp[:6].numpy(), z["dropout.logp_first_draws"], rtol=2e-5, atol=2e-5)
    toks = out["tokens"][:, int(z["P"]):].view(B, T, N).permute(1, 2, 0)
    assert np.array_equal(toks.numpy(), z["dropout.tokens"].astype(np.int64))
    picked = R.topk_lowest_index(out["pool_scores"].numpy(), 8)
    assert picked.tolist() == z[f"dropout.select.{mode}"].tolist()


@pytest.mark.parametrize("mode", H.MODES)
def test_greedy_scores_match_reference(c2, mode):
    z, sd = c2
    out = R.mc_scores(sd, torch.from_numpy(z["prompts"]), int(z["N"]), 1, "greedy", mode, 0.0, lambda step: None)
    np.testing.assert_allclose(out["step_scores"].numpy(), z[f"greedy.scores.{mode}"], **TOL)
    np.testing.assert_allclose(out["logp"].numpy(), z["greedy.logp"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("mode", H.MODES)
def test_sampling_scores_match_reference_with_forced_tokens(c2, mode):
    z, sd = c2
    T, N = int(z["T"]), int(z["N"])
    B = z["prompts"].shape[1]
    forced = torch.from_numpy(z["sampling.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)  # (N, item*T+draw)
    out = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "sampling", mode, 0.0, lambda step: None,
                      temperature=0.8, forced_tokens=forced)
    np.testing.assert_allclose(out["step_scores"].numpy(), z[f"sampling.scores.{mode}"], **TOL)


def test_latent_sites_forward():
    z = H.load("latent_sites.npz")
    sd = H.state_dict_from(z)
    V, d, Hh, F, L = (int(v) for v in z["dims"])
    pm = H.packed_masks_from_bits(z["bits"], L, d)
    live = set(range(R.n_sites(L)))
    out = R.transformer_forward(sd, torch.from_numpy(z["src"]), float(z["p"]), pm.mask_fn(0, live))
    np.testing.assert_allclose(out.numpy(), z["logp"], rtol=2e-5, atol=2e-5)


def test_head_in_sites_forward():
    """Head.bayes (reference transformer.py:50-53): three independent input masks per head, recorded from the reference."""
    z = H.load("head_in_sites.npz")
    sd = H.state_dict_from(z)
    V, d, Hh, F, L = (int(v) for v in z["dims"])
    pm = R.PackedMasks(1, L, z["bits"].shape[2], z["bits"].shape[3], d, Hh)
    pm.bits = z["bits"]
    live = ({R.SITE_EMBED} | {R.site_layer_out(i) for i in range(L)} | {R.site_head_out(i) for i in range(L)}
            | {R.site_head_in(i, h, c, L, Hh) for i in range(L) for h in range(Hh) for c in range(3)})
    out = R.transformer_forward(sd, torch.from_numpy(z["src"]), float(z["p"]), pm.mask_fn(0, live))
    np.testing.assert_allclose(out.numpy(), z["logp"], rtol=2e-5, atol=2e-5)
    # the site matters: without the head-input masks the outputs differ
    out2 = R.transformer_forward(sd, torch.from_numpy(z["src"]), float(z["p"]),
                                 pm.mask_fn(0, {s for s in live if s < R.n_sites(L)}))
    assert np.abs(out2.numpy() - z["logp"]).max() > 1e-2


def test_mid_model_forward():
    z = H.load("mid_model.npz")
    sd = H.state_dict_from(z)
    L, d = int(z["dims"][4]), int(z["dims"][1])
    pm = H.packed_masks_from_bits(z["bits"], L, d)
    out = R.transformer_forward(sd, torch.from_numpy(z["src"]), float(z["p"]), pm.mask_fn(0, R.default_live_sites(L, True)))
    np.testing.assert_allclose(out.numpy(), z["logp"], rtol=2e-5, atol=2e-5)


def test_compute_uncertainty_known_answers():
    # SURVEY §4 known answers
    p = torch.tensor([[0.7, 0.2, 0.1], [1 / 3, 1 / 3, 1 / 3]])
    np.testing.assert_allclose(R.compute_uncertainty(p, "max").numpy(), [0.7, 1 / 3], rtol=1e-6)
    np.testing.assert_allclose(R.compute_uncertainty(p, "margin").numpy(), [0.5, 0.0], atol=1e-7)
    np.testing.assert_allclose(R.compute_uncertainty(p, "entropy").numpy(), [0.8018185, 1.0986123], rtol=1e-6)
    with pytest.raises(ValueError):
        R.compute_uncertainty(p, "bogus")
    z = H.load("uncertainty.npz")
    for mode in H.MODES:
        np.testing.assert_allclose(R.compute_uncertainty(torch.from_numpy(z["probs"]), mode).numpy(), z[mode], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(R.compute_uncertainty(torch.from_numpy(z["probs2"]), mode).numpy(), z[mode + "2"], rtol=1e-6, atol=1e-7)


def test_classifiers_match_reference():
    z = H.load("classifiers.npz")
    x = torch.from_numpy(z["x"])
    keep = torch.from_numpy(R.unpack_bits(z["mlp.keep_bits"], 50))
    out = R.mlp_mc(x, *(torch.from_numpy(z["mlp.sd." + k]) for k in
                        ("hidden_layer.weight", "hidden_layer.bias", "output_layer.weight", "output_layer.bias")),
                   float(z["mlp.p"]), keep)
    np.testing.assert_allclose(out["probs"].numpy(), z["mlp.probs"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out["mean_p"].numpy(), z["mlp.mean"], rtol=1e-6, atol=1e-7)
    out = R.vmlp_mc(x, *(torch.from_numpy(z["vmlp.sd." + k]) for k in
                         ("hidden_layer.w_mu", "hidden_layer.w_rho", "hidden_layer.b_mu", "output_layer.w_mu",
                          "output_layer.w_rho", "output_layer.b_mu")), torch.from_numpy(z["vmlp.eps"]))
    np.testing.assert_allclose(out["probs"].numpy(), z["vmlp.probs"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out["mean_p"].numpy(), z["vmlp.mean"], rtol=1e-6, atol=1e-7)


def test_topk_rule_and_keys():
    s = np.array([1.0, 5.0, 5.0, 5.0, 2.0, -0.0, 0.0, -3.0], dtype=np.float32)
    assert R.topk_lowest_index(s, 3).tolist() == [1, 2, 3]
    keys = R.pack_key(s, np.arange(8))
    order = np.argsort(keys)[::-1]
    assert order[:5].tolist() == [1, 2, 3, 4, 0]
    with pytest.raises(RuntimeError):
        R.topk_lowest_index(s, 9)


@pytest.mark.parametrize("mode", H.MODES)
def test_char_level_c1_scores_match_reference(mode):
    """SURVEY C1 / BASELINE configs[0]: char-level model (V = 67), sequences of 32..36 tokens, the reference's own
    masks (tests/golden/c1_char_level.npz, oracle/make_golden.py char_level_c1)."""
    z = H.load("c1_char_level.npz")
    sd = H.state_dict_from(z)
    P, N, T, L, d = int(z["P"]), int(z["N"]), int(z["T"]), 2, 64
    B = z["prompts"].shape[1]
    pm = H.packed_masks_from_bits(H.expand_live_bits(z["dropout.keep_bits"], L), L, d)
    forced = torch.from_numpy(z["dropout.tokens"].astype(np.int64)).permute(1, 2, 0).reshape(N, B * T)
    out = R.mc_scores(sd, torch.from_numpy(z["prompts"]), N, T, "dropout", mode, float(z["p"]),
                      lambda step: pm.mask_fn(step, R.default_live_sites(L, True)), forced_tokens=forced)
    np.testing.assert_allclose(out["step_scores"].numpy(), z[f"dropout.scores.{mode}"], **TOL)
    lp = out["logp"].view(N, B, T, -1).permute(2, 0, 1, 3)
    np.testing.assert_allclose(lp.numpy(), z["dropout.logp"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(lp.argmax(-1).numpy(), z["dropout.tokens"].astype(np.int64))
