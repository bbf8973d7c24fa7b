#!/usr/bin/env python
"""BASELINE configs[4] / SURVEY C5: the active-learning loop of analysis/notebooks/active_learning.ipynb (cells 24-42),
repeated for 10 rounds, standard Transformer (greedy, sampling) against BayesFormer (MC dropout), three scores, the pool
sharded over the GPUs of one box:

    python examples/active_learning_loop.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        examples/active_learning_loop.py --pool 100000 --rounds 10             # eight GPUs

Every step runs on the native path (bayesformer_b200/al_loop.py): acquisition through libbayesformer_b200.so with one
all-gather of k keys, device-side gather of the acquired token rows, the fused fine-tune step on rank 0, parameters
re-broadcast.  Notebook sizes scaled as SURVEY 8d: train 500, valid 100, test 1000, pool 100,000, k = 200 per round, 5
fine-tune epochs per round."""
import argparse
import copy
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bayesformer_b200 import al_loop
from src.libs import preprocessing, utils
from src.libs.tokenizer import ReversedPairingTokenizer
from src.model.transformer import MyTransformer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--pool", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--init-epochs", type=int, default=40)
    ap.add_argument("--modes", default="max,margin,entropy")
    ap.add_argument("--runs", default="greedy,sampling,dropout", help="greedy / sampling: standard Transformer; dropout: BayesFormer")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the scoring path has no CPU fallback")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    tok = ReversedPairingTokenizer()
    train_set, valid_set = preprocessing.create_dataset(500, 3), preprocessing.create_dataset(100, 3)
    test_set, pool_pairs = preprocessing.create_dataset(1000, 3), preprocessing.create_dataset(args.pool, 3)
    lo, hi = rank * args.pool // world, (rank + 1) * args.pool // world
    t0 = time.perf_counter()
    pool = al_loop.ShardedTokenPool(pool_pairs[lo:hi], tok, device, lo, args.pool)
    say(f"pool of {args.pool} prompts tokenised once ({time.perf_counter() - t0:.2f} s), {world} GPU(s), shard {hi - lo} per GPU")
    bases = {}
    for bayes in (False, True):      # the notebook trains both families once, then deep-copies per experiment (cells 16-24)
        torch.manual_seed(args.seed + int(bayes))
        m = MyTransformer(tok.ntokens, 64, 8, 8, 2, bayes_dropout=0.3 if bayes else None, bayes=bayes).to(device)
        t0 = time.perf_counter()
        if rank == 0:
            utils.train(m, train_set, valid_set, args.init_epochs, 20, 1e-3, tok.ntokens, tok, device)
        al_loop._broadcast_weights(m)
        torch.cuda.synchronize()
        acc = utils.evaluate(m, test_set, tok, 250, device)
        say(f"{'BayesFormer' if bayes else 'Transformer'}: initial training {args.init_epochs} epochs in {time.perf_counter() - t0:.2f} s, "
            f"test accuracy {acc:.3f}")
        bases[bayes] = m
    summary = []
    for compute in args.runs.split(","):
        for mode in args.modes.split(","):
            model = copy.deepcopy(bases[compute == "dropout"])
            say(f"--- {'BayesFormer' if compute == 'dropout' else 'Transformer'} / {compute} / {mode}")
            t0 = time.perf_counter()
            recs = al_loop.active_learning_rounds(model, tok, train_set, valid_set, test_set, pool, device, rounds=args.rounds,
                                                  k=args.k, compute=compute, mode=mode, epochs=args.epochs, seed=args.seed, log=say)
            wall = time.perf_counter() - t0
            summary.append((compute, mode, recs[-1]["test_accuracy"], sum(r["acquire_s"] for r in recs),
                            sum(r["finetune_s"] for r in recs), wall))
    say("=== summary (compute, score, final test accuracy, total acquire s, total fine-tune s, wall s)")
    for row in summary:
        say(f"{row[0]:9s} {row[1]:8s} acc {row[2]:.3f}  acquire {row[3]:.2f} s  fine-tune {row[4]:.2f} s  wall {row[5]:.2f} s")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
