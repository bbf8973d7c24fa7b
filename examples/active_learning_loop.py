#!/usr/bin/env python
"""Multi-round pool-based active learning with the B200-native scoring path (SURVEY config C5).

The reference's notebook (analysis/notebooks/active_learning.ipynb cells 24-42) runs ONE acquisition round per
(scheme, score); this driver repeats acquire -> extend the train set -> fine-tune -> evaluate for several rounds
using only the reference's public functions, all of which route through libbayesformer_b200.so on a CUDA device:

    python examples/active_learning_loop.py --rounds 10 --pool 100000 --k 200 --mode entropy

Fine-tuning (`utils.train`) stays eager PyTorch autograd (out of the hot path); every weight update invalidates
the packed / fp16 weight images automatically (parameter `_version`), so the next round scores with the new model.
Acquired items are kept in the pool, as in the reference.
"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from src.libs import active_learning, preprocessing, utils
from src.libs.tokenizer import ReversedPairingTokenizer
from src.model.transformer import MyTransformer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--pool", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--mode", default="entropy", choices=["max", "margin", "entropy"])
    ap.add_argument("--compute", default="dropout", choices=["greedy", "sampling", "dropout"])
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the scoring path has no CPU fallback")
    device = torch.device("cuda:0")
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    tok = ReversedPairingTokenizer()
    train_set, valid_set = preprocessing.create_dataset(500, 3), preprocessing.create_dataset(100, 3)
    test_set, pool = preprocessing.create_dataset(1000, 3), preprocessing.create_dataset(args.pool, 3)
    pool = active_learning.TokenizedPool(pool, tok)   # tokenised once, not once per round
    bayes = args.compute == "dropout"
    model = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3 if bayes else None, bayes=bayes).to(device)
    utils.train(model, train_set, valid_set, 40, 20, 1e-3, tok.ntokens, tok, device)
    print(f"round 0: test accuracy {utils.evaluate(model, test_set, tok, 250, device):.3f}")
    for r in range(1, args.rounds + 1):
        t0 = time.perf_counter()
        picked = active_learning.select_samples(model, tok, pool, device, nb_samples=args.k, compute=args.compute,
                                                mode=args.mode)
        torch.cuda.synchronize()
        t_acq = time.perf_counter() - t0
        train_set = active_learning.create_new_train_set(train_set, pool, picked)
        t0 = time.perf_counter()
        utils.train(model, train_set, valid_set, args.epochs, 20, 1e-3, tok.ntokens, tok, device)
        t_fit = time.perf_counter() - t0
        acc = utils.evaluate(model, test_set, tok, 250, device)
        print(f"round {r}: acquired {len(picked)} of {len(pool)} in {t_acq:.2f} s "
              f"({len(pool) / t_acq:,.0f} pool samples/s), fine-tune {t_fit:.1f} s, "
              f"train set {len(train_set)}, test accuracy {acc:.3f}")


if __name__ == "__main__":
    main()
