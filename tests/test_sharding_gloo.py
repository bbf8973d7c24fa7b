"""world_size-2 gloo test of the multi-GPU exchange step's host logic: per-rank candidate keys
-> all-gather -> identical final selection on every rank, equal to the single-process top-k.  The scoring,
local top-k and merge kernels are CUDA-only (tests/test_gpu_parity.py::test_topk_sharded_merge_equals_global
covers them); here the keys are produced and merged with the oracle's numpy restatement."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, k: int, out_dir: str):
    from bayesformer_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scores = np.random.default_rng(0).standard_normal(n).astype(np.float32)
        scores[::13] = 0.25  # ties across shard boundaries
        lo, hi = sharding.shard_bounds(n, world, rank)
        kl = sharding.local_candidates(k, hi - lo)
        local = R.topk_lowest_index(scores[lo:hi], kl) + lo
        keys = torch.from_numpy(R.pack_key(scores[local], local).view(np.int64))
        gathered = sharding.gather_candidate_keys(keys, k)
        assert gathered.shape == (world * k,)
        merged = np.sort(gathered.numpy().view(np.uint64))[::-1][:k]
        idx = (np.uint64(0xFFFFFFFF) - (merged & np.uint64(0xFFFFFFFF))).astype(np.int64)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), idx)
    finally:
        dist.destroy_process_group()


def test_two_rank_candidate_exchange(tmp_path):
    n, k, world = 1003, 200, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    scores = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    scores[::13] = 0.25
    want = R.topk_lowest_index(scores, k)
    for r in range(world):
        assert np.load(tmp_path / f"rank{r}.npy").tolist() == want.tolist()


def test_small_shards_pad_with_zero_keys(tmp_path):
    """k larger than a shard: every rank still contributes k slots; zero keys sort last."""
    n, k, world = 150, 100, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    scores = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    scores[::13] = 0.25
    want = R.topk_lowest_index(scores, k)
    for r in range(world):
        assert np.load(tmp_path / f"rank{r}.npy").tolist() == want.tolist()


# ----------------------------------------------------------------------------------------------
# the active-learning loop's exchange step (bayesformer_b200/al_loop.py): token rows of the acquired items gathered
# from the owning ranks with one all-reduce; CPU tensors + gloo exercise the same code the NCCL job runs
# ----------------------------------------------------------------------------------------------


def _gather_worker(rank: int, world: int, port: int, out_dir: str):
    import random

    from bayesformer_b200 import al_loop
    from bayesformer_b200.libs import preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        random.seed(3)   # every rank builds the same pool, keeps its shard
        pool = preprocessing.create_dataset(301, 3) + [("7+5=", "12"), ("12+9=", "21")]
        n = len(pool)
        lo, hi = rank * n // world, (rank + 1) * n // world
        tok = ReversedPairingTokenizer()
        shard = al_loop.ShardedTokenPool(pool[lo:hi], tok, torch.device("cpu"), lo, n)
        picked = [0, n - 1, n - 2, 150, 151, 17, 299, 3]   # both shards, the short items included
        X, Y, pl, al_ = shard.gather_rows(picked)
        torch.save({"X": X, "Y": Y, "pl": pl, "al": al_, "P": shard.P, "A": shard.A}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_of_acquired_rows(tmp_path):
    import random

    from bayesformer_b200 import al_loop
    from bayesformer_b200.libs import active_learning as al, preprocessing
    from bayesformer_b200.libs.tokenizer import ReversedPairingTokenizer

    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    random.seed(3)
    pool = preprocessing.create_dataset(301, 3) + [("7+5=", "12"), ("12+9=", "21")]
    n = len(pool)
    picked = [0, n - 1, n - 2, 150, 151, 17, 299, 3]
    tok = ReversedPairingTokenizer()
    # what the strings route produces: create_new_train_set + tokenisation of the appended items
    new_items = al.create_new_train_set([], pool, picked)
    X, Y, p_len, a_len, _, _ = preprocessing.get_batch(new_items, tok, 0, len(new_items))
    Xp, Yp, P_all, A_all, _, _ = preprocessing.get_batch(pool, tok, 0, n)
    got = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    for g in got:
        assert (g["P"], g["A"]) == (P_all, A_all)                       # global maxima, preprocessing.py:59
        assert torch.equal(g["X"], got[0]["X"]) and torch.equal(g["Y"], got[0]["Y"])   # every rank holds the same block
        assert torch.equal(g["X"][g["X"].shape[0] - p_len:], X) and torch.equal(g["Y"][:a_len + 1], Y)
        assert g["pl"].tolist() == [int((X[:, j] != tok.token_to_id["<PAD>"]).sum()) for j in range(len(picked))]
    # and a token training set extended with them serves the batches get_batch would
    base = preprocessing.create_dataset(25, 2)
    ds = al_loop.TokenDataset.from_pairs(base, tok, torch.device("cpu")).extended(got[0]["X"], got[0]["Y"], got[0]["pl"], got[0]["al"])
    joined = base + new_items
    for start in range(0, len(joined) - 1, 10):
        m = min(10, len(joined) - start)
        Xb, Yb, pb, ab, _, _ = preprocessing.get_batch(joined, tok, start, m)
        Xd, Yd, pd, ad = ds.batch(start, m)
        assert (pd, ad) == (pb, ab) and torch.equal(Xd, Xb) and torch.equal(Yd, Yb)
