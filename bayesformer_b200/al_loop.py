"""Multi-round pool-based active learning on the native path (SURVEY config C5, §8f-4).

The reference's notebook (analysis/notebooks/active_learning.ipynb cells 24-42) runs ONE acquisition round per
(scheme, score): `select_samples` -> `create_new_train_set` -> `utils.train` -> `utils.evaluate`.  This driver repeats
that round with the same public semantics (acquired items stay in the pool, as in the reference,
src/libs/active_learning.py:214-233), with the pool sharded over the ranks of one box:

  * every rank keeps its contiguous shard of the tokenised pool on its GPU (tokenised once);
  * acquisition = `sharding.select_samples_sharded` (local scoring + local top-k + ONE all-gather of k keys);
  * the acquired rows are gathered ON THE DEVICE: the owning rank index-selects its token rows, one all-reduce of the
    (P + A + 1 + 2, k) int64 block hands them to everybody (14 kB at k = 200) — no strings travel;
  * rank 0 fine-tunes (the fused step of csrc/train_fused.cu for the notebook model class) on the device-resident
    token training set and broadcasts the parameters; every rank then invalidates its packed weight images.
"""

from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from bayesformer_b200 import engine, sharding
from bayesformer_b200.configs import constants
from bayesformer_b200.libs import preprocessing, utils


class TokenDataset:
    """A training set kept as token rows on the device, laid out so that `batch(i, n)` returns exactly what
    `preprocessing.get_batch(dataset, tokenizer, i, n)` returns for the same items (prompts LEFT padded to the batch's
    longest prompt, answers + EOS RIGHT padded to the batch's longest answer; reference preprocessing.py:43-101).
    Stored globally padded: prompts (P_max, n), answers (A_max + 1, n), plus the true lengths (host)."""

    def __init__(self, prompts: torch.Tensor, answers: torch.Tensor, p_lens: np.ndarray, a_lens: np.ndarray, pad_id: int):
        self.prompts, self.answers = prompts, answers
        self.p_lens, self.a_lens = np.asarray(p_lens, dtype=np.int64), np.asarray(a_lens, dtype=np.int64)
        self.pad_id = pad_id

    @classmethod
    def from_pairs(cls, dataset, tokenizer, device) -> "TokenDataset":
        X, Y, _p, _a, _, _ = preprocessing.get_batch(dataset, tokenizer, 0, len(dataset))
        pad_id = tokenizer.token_to_id[constants.PAD_TOKEN]
        eos_id = tokenizer.token_to_id[constants.EOS_TOKEN]
        p_lens = (X != pad_id).sum(0).numpy()                  # PAD only ever appears as left padding of a prompt
        a_lens = (Y == eos_id).float().argmax(0).numpy()       # position of the EOS = number of answer tokens
        return cls(X.to(device), Y.to(device), p_lens, a_lens, pad_id)

    def __len__(self) -> int:
        return int(self.prompts.shape[1])

    def batch(self, i: int, n: int):
        """(X (P, n), Y (A + 1, n), P, A) of items [i, i + n), padded to the slice's own maxima."""
        p_len, a_len = int(self.p_lens[i:i + n].max()), int(self.a_lens[i:i + n].max())
        X = self.prompts[self.prompts.shape[0] - p_len:, i:i + n]
        Y = self.answers[:a_len + 1, i:i + n]
        return X, Y, p_len, a_len

    def extended(self, prompts: torch.Tensor, answers: torch.Tensor, p_lens, a_lens) -> "TokenDataset":
        """Copy extended by new items (`create_new_train_set`, reference active_learning.py:214-233), re-padded to the
        joint maxima."""
        P = max(self.prompts.shape[0], prompts.shape[0])
        A1 = max(self.answers.shape[0], answers.shape[0])

        def lpad(t, rows):
            return t if t.shape[0] == rows else torch.cat((torch.full((rows - t.shape[0], t.shape[1]), self.pad_id,
                                                                      dtype=t.dtype, device=t.device), t), 0)

        def rpad(t, rows):
            return t if t.shape[0] == rows else torch.cat((t, torch.full((rows - t.shape[0], t.shape[1]), self.pad_id,
                                                                         dtype=t.dtype, device=t.device)), 0)

        return TokenDataset(torch.cat((lpad(self.prompts, P), lpad(prompts, P)), 1),
                            torch.cat((rpad(self.answers, A1), rpad(answers, A1)), 1),
                            np.concatenate((self.p_lens, np.asarray(p_lens))), np.concatenate((self.a_lens, np.asarray(a_lens))),
                            self.pad_id)


class ShardedTokenPool:
    """This rank's contiguous shard [lo, hi) of a pool of `n_total` items, tokenised once and resident on the device.
    P and A are the GLOBAL maxima (reference preprocessing.py:59: the whole pool is padded as one batch)."""

    def __init__(self, shard_pairs, tokenizer, device, lo: int, n_total: int, group=None):
        X, Y, p_len, a_len, _, _ = preprocessing.get_batch(shard_pairs, tokenizer, 0, len(shard_pairs))
        pad_id = tokenizer.token_to_id[constants.PAD_TOKEN]
        eos_id = tokenizer.token_to_id[constants.EOS_TOKEN]
        dims = torch.tensor([p_len, a_len], dtype=torch.int64, device=device)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(dims, op=dist.ReduceOp.MAX, group=group)
        P, A = int(dims[0]), int(dims[1])
        X = sharding.repad_prompts(X, P, pad_id)
        if Y.shape[0] < A + 1:
            Y = torch.cat((Y, torch.full((A + 1 - Y.shape[0], Y.shape[1]), pad_id, dtype=Y.dtype)), 0)
        self.prompts, self.answers = X.to(device), Y.to(device)
        self.p_lens = (self.prompts != pad_id).sum(0)
        self.a_lens = (self.answers == eos_id).float().argmax(0)
        self.lo, self.hi, self.n_total = lo, lo + len(shard_pairs), n_total
        self.P, self.A, self.pad_id, self.group = P, A, pad_id, group

    def gather_rows(self, indices: list[int]):
        """Token rows of the acquired GLOBAL indices, assembled on the device: the owner of an index contributes its
        prompt / answer columns and lengths, everybody else zeros; one all-reduce (sum) completes the block."""
        dev = self.prompts.device
        idx = torch.tensor(indices, dtype=torch.int64, device=dev)
        mine = (idx >= self.lo) & (idx < self.hi)
        local = torch.where(mine, idx - self.lo, torch.zeros_like(idx))
        block = torch.cat((self.prompts.index_select(1, local), self.answers.index_select(1, local),
                           self.p_lens.index_select(0, local).view(1, -1), self.a_lens.index_select(0, local).view(1, -1)), 0)
        block = block * mine.view(1, -1).to(block.dtype)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(block, op=dist.ReduceOp.SUM, group=self.group)
        P, A1 = self.P, self.A + 1
        lens = block[P + A1:].cpu().numpy()
        return block[:P], block[P:P + A1], lens[0], lens[1]


def _broadcast_weights(model, group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.data, src=0, group=group)
        engine.mark_weights_changed(model)


def train_on_tokens(model, train_set: TokenDataset, valid_set: TokenDataset, nb_epochs: int, batch_size: int, lr: float,
                    vocab_size: int, device) -> tuple[list[float], list[float]]:
    """`utils.train` (reference src/libs/utils.py:12-124) over device-resident token sets: same batch walk (slices of
    `batch_size` in order, no shuffling, :60), same loss bookkeeping (:110-111), fresh AdamW state per call (:51); the
    step itself is the fused kernel when the model is in its class, else eager autograd."""
    fused = utils._fused_enabled(model, device) and max(
        train_set.prompts.shape[0] + train_set.answers.shape[0], 0) <= engine.FusedTrainer.MAX_SEQ
    trainer = utils._fused_trainer(model, lr) if fused else None
    optimizer = None if fused else torch.optim.AdamW(model.parameters(), lr=lr)
    criterion = torch.nn.CrossEntropyLoss()
    hist_t, hist_v = [], []
    try:
        for _epoch in range(nb_epochs):
            model.train()
            running = torch.zeros((), device=device)
            if fused:
                trainer.loss.zero_()
            for start in range(0, len(train_set) - 1, batch_size):
                X, Y, p_len, _a = train_set.batch(start, min(batch_size, len(train_set) - start))
                tokens = torch.cat((X, Y), 0)
                if fused:
                    trainer.step(tokens, p_len, utils._train_dropout_rate(model), float(model.bayes_dropout) if model.bayes else 0.0)
                    continue
                model.zero_grad()
                with torch.enable_grad():
                    out = model(tokens)
                    loss = criterion(out[p_len - 1:-1].reshape(-1, vocab_size), Y.reshape(-1))
                    loss.backward()
                optimizer.step()
                running += loss.detach()
            hist_t.append(float((trainer.loss if fused else running).item()) / len(train_set))
            model.eval()
            rv = torch.zeros((), device=device)
            with torch.no_grad():
                for start in range(0, len(valid_set) - 1, batch_size):
                    X, Y, p_len, _a = valid_set.batch(start, min(batch_size, len(valid_set) - start))
                    out = model(torch.cat((X, Y), 0))
                    rv += criterion(out[p_len - 1:-1].reshape(-1, vocab_size), Y.reshape(-1))
            hist_v.append(float(rv.item()) / len(valid_set))
    finally:
        engine.mark_weights_changed(model)
    return hist_t, hist_v


def evaluate_tokens(model, test_set: TokenDataset, batch_size: int, device) -> float:
    """`utils.evaluate` (reference :151-193) over a token set: exact match of the generated answer incl. EOS / PAD tail."""
    model.eval()
    hits = 0.0
    with torch.no_grad():
        for start in range(0, len(test_set) - 1, batch_size):
            X, Y, p_len, a_len = test_set.batch(start, min(batch_size, len(test_set) - start))
            produced = utils.generate(model, X, a_len + 1, device)[p_len:, :]
            hits += float(torch.all(produced == Y, dim=0).float().sum())
    return hits / len(test_set)


def active_learning_rounds(model, tokenizer, train_pairs, valid_pairs, test_pairs, pool: ShardedTokenPool, device,
                           rounds: int = 10, k: int = 200, compute: str = "dropout", mode: str = "entropy",
                           epochs: int = 5, batch_size: int = 20, lr: float = 1e-3, seed: int = 42, n_draws: int = 20,
                           group=None, log=None) -> list[dict]:
    """`rounds` x {acquire k pool items, extend the train set, fine-tune, evaluate}.  Returns one record per round with
    the acquired indices, the acquisition / gather / fine-tune times and the test accuracy (identical on every rank)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device(device)
    train = TokenDataset.from_pairs(train_pairs, tokenizer, dev)
    valid = TokenDataset.from_pairs(valid_pairs, tokenizer, dev)
    test = TokenDataset.from_pairs(test_pairs, tokenizer, dev)
    records = []
    for r in range(1, rounds + 1):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        picked = sharding.select_samples_sharded(model, pool.prompts, pool.A + 1, dev, nb_samples=k, compute=compute, mode=mode,
                                                 index_base=pool.lo, n_total=pool.n_total, seed=seed + r, n_draws=n_draws,
                                                 group=group)
        torch.cuda.synchronize(dev)
        t_acq = time.perf_counter() - t0
        t0 = time.perf_counter()
        Xn, Yn, pl, al = pool.gather_rows(picked)
        train = train.extended(Xn, Yn, pl, al)
        torch.cuda.synchronize(dev)
        t_gather = time.perf_counter() - t0
        t0 = time.perf_counter()
        if rank == 0:
            train_on_tokens(model, train, valid, epochs, batch_size, lr, tokenizer.ntokens, dev)
        _broadcast_weights(model, group)
        torch.cuda.synchronize(dev)
        t_fit = time.perf_counter() - t0
        acc = evaluate_tokens(model, test, 250, dev)
        rec = {"round": r, "picked": picked, "acquire_s": t_acq, "gather_s": t_gather, "finetune_s": t_fit,
               "train_size": len(train), "test_accuracy": acc}
        records.append(rec)
        if log is not None and rank == 0:
            log(f"round {r:2d}: acquired {len(picked)} of {pool.n_total} in {t_acq * 1e3:7.1f} ms "
                f"({pool.n_total / t_acq:,.0f} pool samples/s), gather {t_gather * 1e3:5.1f} ms, fine-tune {t_fit * 1e3:7.1f} ms "
                f"({epochs} epochs x {len(range(0, len(train) - 1, batch_size))} batches), train set {len(train)}, "
                f"test accuracy {acc:.3f}")
    return records
