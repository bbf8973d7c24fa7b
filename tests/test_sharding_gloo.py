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
