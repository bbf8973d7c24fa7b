"""Pool-based active learning: uncertainty scores, stochastic scoring schemes, top-k acquisition.

Drop-in for the reference's `src/libs/active_learning.py` (same names, arguments, return types,
error behaviour).  When the model lives on a CUDA device the three `*_uncertainty` loops and
`compute_uncertainty` / `topk` run inside libbayesformer_b200.so: the `nb_samples` MC draws are
folded into the batch, masks come from a counter-based Philox stream, and the whole greedy decode
(no KV cache — fresh masks hit every position at every step, reference :145) is driven natively.
There is no CPU fallback: a model or tensor that is not on a CUDA device raises RuntimeError, and so does a
missing / stale libbayesformer_b200.so.

Semantics kept on purpose: `mode="max"` returns the max probability and
`mode="margin"` p(1) - p(2); `select_samples` then takes the LARGEST k, i.e. the most confident
items; entropy uses log(p + 1e-10); no padding mask; the step mean precedes top-k.
"""

from __future__ import annotations

import torch

from bayesformer_b200.libs import preprocessing
from bayesformer_b200.libs.tokenizer import Tokenizer
from bayesformer_b200.model.transformer import MyTransformer

_MODES = ("max", "margin", "entropy")


def _bad_mode(mode) -> ValueError:
    return ValueError(f"Unknown mode {mode}. Available modes are 'max', 'margin' and 'entropy'.")


def _no_cpu_path(what: str) -> RuntimeError:
    return RuntimeError(
        f"{what}: bayesformer_b200 runs the MC-scoring path only on a CUDA device (sm_100a) — there is no CPU "
        "fallback. Move the model and tensors to a CUDA device.")


def compute_uncertainty(probs: torch.Tensor, mode: str = "max") -> torch.Tensor:
    """Score a batch of categorical distributions `probs (B, C)` → `(B,)` (reference :10-35).

    Runs `bfb200_compute_uncertainty`.  Under autograd (training-time use, outside the hot path) the
    same formulas are applied with differentiable ATen ops on the tensor's own device."""
    if mode not in _MODES:
        raise _bad_mode(mode)
    if torch.is_grad_enabled() and probs.requires_grad:
        if mode == "max":
            return probs.max(dim=-1).values
        if mode == "margin":
            ranked = torch.sort(probs, dim=-1, descending=True).values
            return ranked[:, 0] - ranked[:, 1]
        return -(probs * torch.log(probs + 1e-10)).sum(dim=-1)
    if not probs.is_cuda:
        raise _no_cpu_path("compute_uncertainty")
    from bayesformer_b200 import engine

    return engine.compute_uncertainty(probs, mode)


def _scores(model, prompts, new_tokens, device, mode, scheme, n_draws, temperature=0.8):
    """Common body of the three `*_uncertainty` functions: one `bfb200_mc_score_transformer` call."""
    if mode not in _MODES:
        raise _bad_mode(mode)
    if not isinstance(model, MyTransformer):
        raise TypeError("model must be a bayesformer_b200 MyTransformer")
    dev = model.embedding.weight.device
    if dev.type != "cuda" or torch.device(device).type != "cuda":
        raise _no_cpu_path(f"{scheme}_uncertainty")
    if model.training and not model._train_only_dropout_inactive():
        raise RuntimeError("call model.eval() first: the native path implements inference semantics "
                           "(nn.Dropout sites off, always-on bayes sites live)")
    from bayesformer_b200 import engine

    pack = engine.scoring_pack(model, dev, prompts.shape[0], new_tokens, engine.default_compute(model, scheme))
    res = engine.mc_score_transformer(pack, prompts.to(dev), new_tokens, n_draws, scheme, mode,
                                      temperature=temperature)
    return res.step_scores


def greedy_uncertainty(model: MyTransformer, prompts: torch.Tensor, new_tokens: int, device: str,
                       mode: str = "max") -> torch.Tensor:
    """Deterministic greedy decode; per-step score of the next-token distribution → `(new_tokens, B)` (reference :38-70)."""
    return _scores(model, prompts, new_tokens, device, mode, "greedy", 1)


def sampling_uncertainty(model: MyTransformer, prompts: torch.Tensor, new_tokens: int, device: str,
                         mode: str = "max", temperature: float = 0.8, nb_samples: int = 20) -> torch.Tensor:
    """`nb_samples` temperature-sampled continuations; mean per-step score of the tempered distribution (reference :73-115)."""
    return _scores(model, prompts, new_tokens, device, mode, "sampling", nb_samples, temperature)


def dropout_uncertainty(model: MyTransformer, prompts: torch.Tensor, new_tokens: int, device: str,
                        mode: str = "max", nb_samples: int = 20) -> torch.Tensor:
    """`nb_samples` MC-dropout passes, each greedily decoding its own trajectory; mean per-step score (reference :118-155)."""
    return _scores(model, prompts, new_tokens, device, mode, "dropout", nb_samples)


_SCHEMES = {
    "greedy": greedy_uncertainty,
    "sampling": sampling_uncertainty,
    "dropout": dropout_uncertainty,
}


def select_samples_from_tokens(model: MyTransformer, prompts: torch.Tensor, new_tokens: int, device,
                               nb_samples: int = 200, compute: str = "greedy", mode: str = "max") -> list[int]:
    """`select_samples` after tokenisation: prompts `(P, B)` int64 (host or device) → k pool indices.

    This is the part of the reference's `select_samples` (:191-211) that runs on the device; the
    benchmark's end-to-end number is measured through this call with pinned host `prompts`.
    """
    if compute not in _SCHEMES:
        raise ValueError(
            f"Unknown compute mode: {compute}. For Transformer, use 'greedy' or 'sampling'. For BayesFormer, use 'dropout'."
        )
    with torch.no_grad():
        prompts = prompts.to(device, non_blocking=True)
        scores = _SCHEMES[compute](model, prompts, new_tokens, device, mode)
        from bayesformer_b200 import engine

        _, picked = engine.topk(engine.step_mean(scores), nb_samples)
        return picked.tolist()


class TokenizedPool:
    """A pool tokenised ONCE (the reference re-tokenises the whole pool in every acquisition round,
    `active_learning.py:191-194`).  Pass it to `select_samples` in place of the list of pairs; it also behaves like
    that list (`len`, indexing, iteration), so `create_new_train_set(train, pool, indices)` keeps working.  The token
    matrix stays on the device it was first scored on."""

    def __init__(self, pool_dataset: list[tuple[str, str]], tokenizer: Tokenizer):
        self.items = list(pool_dataset)
        prompts, _answers, _p_len, self.answers_length, _, _ = preprocessing.get_batch(self.items, tokenizer, 0, len(self.items))
        self.prompts = prompts
        self._on_device: dict[str, torch.Tensor] = {}

    def prompts_on(self, device) -> torch.Tensor:
        key = str(torch.device(device))
        if key not in self._on_device:
            self._on_device[key] = self.prompts.to(device)
        return self._on_device[key]

    def __len__(self) -> int:
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def __iter__(self):
        return iter(self.items)


def select_samples(model: MyTransformer, tokenizer: Tokenizer, pool_dataset, device: str,
                   nb_samples: int = 200, compute: str = "greedy", mode: str = "max") -> list[int]:
    """Indices of the `nb_samples` pool items with the largest step-averaged score (reference :158-211).
    `pool_dataset` is the reference's list of (prompt, answer) pairs, or a `TokenizedPool` built from it once."""
    model.eval()
    with torch.no_grad():
        if isinstance(pool_dataset, TokenizedPool):
            return select_samples_from_tokens(model, pool_dataset.prompts_on(device), pool_dataset.answers_length + 1, device,
                                              nb_samples, compute, mode)
        prompts, _answers, _p_len, answers_length, _, _ = preprocessing.get_batch(
            pool_dataset, tokenizer, 0, len(pool_dataset))
        return select_samples_from_tokens(model, prompts, answers_length + 1, device, nb_samples, compute, mode)


def create_new_train_set(train_dataset: list[tuple[str, str]], pool_dataset: list[tuple[str, str]],
                         indices: list[int]) -> list[tuple[str, str]]:
    """Copy of `train_dataset` extended by the selected pool items, in selection order (reference :214-233)."""
    return list(train_dataset) + [pool_dataset[i] for i in indices]
