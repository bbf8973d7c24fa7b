"""The multi-round active-learning driver (bayesformer_b200/al_loop.py) on one GPU: a 2-round loop on a
small pool — device-side gather of the acquired rows equals `create_new_train_set` + tokenisation, the token training
path equals `preprocessing.get_batch` batch for batch, the loop runs on the native kernels and its selection equals the
public `select_samples` under the same seed."""

import copy
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from tests import helpers as H

    dev = torch.device("cuda:0")
    tok = ReversedPairingTokenizer()
    random.seed(21)
    data = dict(train=preprocessing.create_dataset(60, 3), valid=preprocessing.create_dataset(40, 3),
                test=preprocessing.create_dataset(100, 3), pool=preprocessing.create_dataset(3000, 3))
    z = H.load("c2_transformer.npz")
    model = H.build_model(H.state_dict_from(z), float(z["p"]), True).to(dev)
    return dev, tok, data, model


def test_token_dataset_batches_equal_get_batch(setup):
    from bayesformer_b200 import al_loop
    from bayesformer_b200.libs import preprocessing

    dev, tok, data, _ = setup
    mixed = data["train"] + [("7+5=", "12"), ("007+003=", "10")] + preprocessing.create_dataset(17, 2)
    ds = al_loop.TokenDataset.from_pairs(mixed, tok, dev)
    for start in range(0, len(mixed) - 1, 20):
        n = min(20, len(mixed) - start)
        X, Y, p_len, a_len, _, _ = preprocessing.get_batch(mixed, tok, start, n)
        Xd, Yd, pd, ad = ds.batch(start, n)
        assert (pd, ad) == (p_len, a_len) and torch.equal(Xd.cpu(), X) and torch.equal(Yd.cpu(), Y)


def test_two_round_loop(setup):
    from bayesformer_b200 import al_loop, engine
    from bayesformer_b200.libs import active_learning as al, preprocessing

    dev, tok, data, base = setup
    model = copy.deepcopy(base)
    pool = al_loop.ShardedTokenPool(data["pool"], tok, dev, 0, len(data["pool"]))
    with engine.profile(dev) as prof:
        recs = al_loop.active_learning_rounds(model, tok, data["train"], data["valid"], data["test"], pool, dev, rounds=2,
                                              k=50, compute="dropout", mode="entropy", epochs=2, seed=7)
    assert "mc_forward_kernel" in prof.kernels and "train_seq_kernel" in prof.kernels and "topk_select_kernel" in prof.kernels
    assert [r["train_size"] for r in recs] == [110, 160]
    assert all(len(set(r["picked"])) == 50 and 0.0 <= r["test_accuracy"] <= 1.0 for r in recs)
    # device-side gather == strings route: create_new_train_set + tokenisation (reference active_learning.py:214-233)
    Xn, Yn, pl, al_ = pool.gather_rows(recs[0]["picked"])
    new_train = al.create_new_train_set(data["train"], data["pool"], recs[0]["picked"])
    X, Y, p_len, a_len, _, _ = preprocessing.get_batch(new_train[60:], tok, 0, 50)
    assert torch.equal(Xn.cpu()[Xn.shape[0] - p_len:], X) and torch.equal(Yn.cpu()[:a_len + 1], Y)
    # round 1's selection is what the sharded public call selects for this seed on the starting weights
    from bayesformer_b200 import sharding

    first = sharding.select_samples_sharded(copy.deepcopy(base), pool.prompts, pool.A + 1, dev, nb_samples=50, compute="dropout",
                                            mode="entropy", index_base=0, n_total=len(data["pool"]), seed=7 + 1, n_draws=20)
    assert first == recs[0]["picked"]
    # fine-tuning moved the weights and the next round scored with them
    assert not torch.allclose(model.embedding.weight, base.embedding.weight)
    assert recs[1]["picked"] != recs[0]["picked"]
