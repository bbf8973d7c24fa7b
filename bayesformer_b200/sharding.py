"""Pool sharding over the GPUs of one box: one process per GPU, contiguous index
ranges, replicated weights, and ONE exchange step — an all-gather of every rank's k packed top-k
candidate keys (8·k bytes per rank), after which every rank selects the final k itself.

torch.distributed is plumbing only: the scoring, the local top-k and the final merge run in
libbayesformer_b200.so.  Philox masks are keyed by the *global* pool index, so a sharded run scores
every item exactly as the single-GPU run does and acquires the same indices.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of pool indices owned by `rank` (sizes differ by at most one)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    lo = n_items * rank // world_size
    hi = n_items * (rank + 1) // world_size
    return lo, hi


def local_candidates(k: int, n_local: int) -> int:
    """How many candidates a shard contributes: min(k, shard size) — a shard smaller than k sends all it has."""
    return max(0, min(k, n_local))


def gather_candidate_keys(local_keys: torch.Tensor, k: int, group=None) -> torch.Tensor:
    """All-gather of per-rank candidate keys (int64 view of the packed 64-bit keys).

    Every rank contributes exactly `k` slots; a shard with fewer than k items pads with key 0, which
    sorts below every real key (real keys have the index field 0xFFFFFFFF - i > 0 or a non-zero score
    field).  Returns the (world_size * k,) tensor, identical on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    padded = torch.zeros(k, dtype=torch.int64, device=local_keys.device)
    padded[: local_keys.numel()] = local_keys.reshape(-1)[:k]
    if world == 1:
        return padded
    out = torch.empty(world * k, dtype=torch.int64, device=local_keys.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out


def select_samples_sharded(model, local_prompts: torch.Tensor, new_tokens: int, device, nb_samples: int = 200,
                           compute: str = "dropout", mode: str = "max", index_base: int = 0, n_total: int | None = None,
                           seed: int | None = None, n_draws: int = 20, temperature: float = 0.8, group=None,
                           native_compute: str | None = None) -> list[int]:
    """`select_samples` (reference src/libs/active_learning.py:158-211) over a pool sharded across ranks.

    local_prompts (P, B_local) int64 are this rank's contiguous shard starting at global index
    `index_base`; P and new_tokens must be the GLOBAL maxima (reference preprocessing.py:59).  Returns the
    same list of global indices on every rank.
    """
    from bayesformer_b200 import engine

    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("select_samples_sharded: bayesformer_b200 has no CPU path")
    if compute not in ("greedy", "sampling", "dropout"):
        raise ValueError(
            f"Unknown compute mode: {compute}. For Transformer, use 'greedy' or 'sampling'. For BayesFormer, use 'dropout'.")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1 and seed is None and compute != "greedy":
        raise ValueError("a sharded stochastic run needs an explicit `seed` shared by all ranks "
                         "(masks are keyed by (seed, global index, draw, ...))")
    with torch.no_grad():
        prompts = local_prompts.to(dev, non_blocking=True)
        n_local = prompts.shape[1]
        if n_total is None:
            n_total = n_local * world
        if nb_samples > n_total:
            raise RuntimeError("selected index k out of range")
        if native_compute is None:   # stochastic schemes: 'auto' (tensor-core path where in class); greedy: exact fp32
            native_compute = engine.default_compute(model, compute)
        pack = engine.scoring_pack(model, dev, prompts.shape[0], new_tokens, native_compute)
        res = engine.mc_score_transformer(pack, prompts, new_tokens, 1 if compute == "greedy" else n_draws, compute,
                                          mode, temperature=temperature, seed=seed, index_base=index_base)
        kl = local_candidates(nb_samples, n_local)
        if kl > 0:
            keys = engine.topk_keys(res.pool_scores, kl, index_base=index_base)
        else:
            keys = torch.empty(0, dtype=torch.int64, device=dev)
        if world == 1:
            final = keys
        else:
            final = engine.merge_topk_keys(gather_candidate_keys(keys, nb_samples, group), nb_samples)
        _, idx = engine.unpack_keys(final)
        return idx.tolist()


def repad_prompts(prompts: torch.Tensor, prompt_len: int, pad_id: int) -> torch.Tensor:
    """Left-pad a shard's `(P_local, B)` prompts with PAD to the GLOBAL prompt length (the reference pads every
    prompt of the pool to the pool-wide maximum, preprocessing.py:59-64)."""
    p_local, n = prompts.shape
    if p_local > prompt_len:
        raise ValueError("global prompt length smaller than a shard's")
    if p_local == prompt_len:
        return prompts
    pad = torch.full((prompt_len - p_local, n), pad_id, dtype=prompts.dtype, device=prompts.device)
    return torch.cat((pad, prompts), 0)


def select_samples_distributed(model, tokenizer, pool_dataset, device, nb_samples: int = 200, compute: str = "dropout",
                               mode: str = "max", seed: int | None = None, n_draws: int = 20, group=None) -> list[int]:
    """`select_samples(model, tokenizer, pool_dataset, device, ...)` (reference active_learning.py:158-211) with the
    pool sharded over the ranks of `group`: every rank passes the SAME full `pool_dataset`, tokenises only its own
    contiguous shard, agrees on the global prompt / answer lengths with one 2-element all-reduce(MAX) (set-up, not
    the data path), scores its shard and exchanges k candidate keys.  Every rank returns the same global indices."""
    from bayesformer_b200.configs import constants
    from bayesformer_b200.libs import preprocessing

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(pool_dataset), world, rank)
    model.eval()
    prompts, _, p_len, a_len, _, _ = preprocessing.get_batch(pool_dataset, tokenizer, lo, hi - lo)
    dims = torch.tensor([p_len, a_len], dtype=torch.int64, device=torch.device(device))
    if world > 1:
        dist.all_reduce(dims, op=dist.ReduceOp.MAX, group=group)
    p_len, a_len = int(dims[0]), int(dims[1])
    prompts = repad_prompts(prompts, p_len, tokenizer.token_to_id[constants.PAD_TOKEN])
    return select_samples_sharded(model, prompts, a_len + 1, device, nb_samples=nb_samples, compute=compute, mode=mode,
                                  index_base=lo, n_total=len(pool_dataset), seed=seed, n_draws=n_draws, group=group)
