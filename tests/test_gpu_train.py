"""utils.train on a CUDA device: the CUDA-graph-captured step (one graph per batch shape) must train exactly like the
eager loop it replaces (reference src/libs/utils.py:12-124) — SURVEY §8f rank 3, the step after acquisition."""

import copy
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(monkeypatch, graphs: bool, model, data, tok, dev, epochs=3):
    from bayesformer_b200.libs import utils

    monkeypatch.setenv("BFB200_GRAPH_TRAIN", "1" if graphs else "0")
    monkeypatch.setenv("BFB200_FUSED_TRAIN", "0")   # this file tests the autograd loops (the fused step: test_gpu_train_fused.py)
    m = copy.deepcopy(model)
    torch.manual_seed(5)
    hist = utils.train(m, data[0], data[1], epochs, 20, 1e-3, tok.ntokens, tok, dev)
    return m, hist


@pytest.mark.parametrize("bayes", [False, True])
def test_graphed_training_equals_eager_training(monkeypatch, bayes):
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    dev = torch.device("cuda:0")
    random.seed(3)
    torch.manual_seed(3)
    tok = ReversedPairingTokenizer()
    data = (preprocessing.create_dataset(100, 3), preprocessing.create_dataset(40, 3))
    # no random dropout => both loops are deterministic functions of the weights and the batches
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, dropout=0.0, bayes_dropout=0.0 if bayes else None, bayes=bayes).to(dev)
    eager, h_eager = _run(monkeypatch, False, model, data, tok, dev)
    graphed, h_graph = _run(monkeypatch, True, model, data, tok, dev)
    np.testing.assert_allclose(h_graph[0], h_eager[0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(h_graph[1], h_eager[1], rtol=1e-4, atol=1e-6)
    for (n, a), (_, b) in zip(eager.state_dict().items(), graphed.state_dict().items()):
        torch.testing.assert_close(b, a, rtol=1e-3, atol=1e-5, msg=lambda s, n=n: f"{n}: {s}")
    # and the trained weights moved
    assert not torch.allclose(graphed.embedding.weight, model.embedding.weight)


def test_graphed_training_with_dropout_learns(monkeypatch):
    """With the always-on MC dropout and the train-only dropout sites live, the graphed loop draws fresh masks on every
    replay (graph-safe Philox offsets): the training loss falls like the eager loop's."""
    from bayesformer_b200.libs import preprocessing, utils
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer
    from bayesformer_b200.model.transformer import MyTransformer

    dev = torch.device("cuda:0")
    random.seed(4)
    torch.manual_seed(4)
    tok = ReversedPairingTokenizer()
    data = (preprocessing.create_dataset(200, 3), preprocessing.create_dataset(40, 3))
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3, bayes=True).to(dev)
    monkeypatch.setenv("BFB200_GRAPH_TRAIN", "1")
    monkeypatch.setenv("BFB200_FUSED_TRAIN", "0")
    tr, va = utils.train(model, data[0], data[1], 12, 20, 1e-3, tok.ntokens, tok, dev)
    assert tr[-1] < 0.8 * tr[0] and np.isfinite(tr).all() and np.isfinite(va).all()
    # two replays of one graph must not reuse a mask: consecutive epochs over the same batches give different losses
    assert len(set(np.round(tr, 7))) == len(tr)
