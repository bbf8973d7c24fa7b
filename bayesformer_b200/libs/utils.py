"""Training / greedy generation / exact-match evaluation around MyTransformer.

Drop-in for the reference's `src/libs/utils.py` (`train` :12-124, `generate` :127-148,
`evaluate` :151-193).  These are the callers on either side of the scoring path:
`train` stays eager autograd; `generate` under `torch.no_grad()` on CUDA goes through the native
forward via `MyTransformer.forward`.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from bayesformer_b200.libs import preprocessing
from bayesformer_b200.libs.tokenizer import Tokenizer
from bayesformer_b200.model.transformer import MyTransformer


def _batch_loss(model, dataset, tokenizer, start, batch_size, vocab_size, device, criterion):
    prompts, answers, p_len, _a_len, _, _ = preprocessing.get_batch(dataset, tokenizer, start, batch_size)
    prompts, answers = prompts.to(device), answers.to(device)
    out = model(torch.cat((prompts, answers), 0))
    # positions P-1 .. end-1 predict the answer tokens (teacher forcing, reference :72-76)
    pred = out[p_len - 1 : -1, :, :].reshape(-1, vocab_size)
    return criterion(pred, answers.view(-1))


def train(model: MyTransformer, train_dataset, valid_dataset, nb_epochs: int, batch_size: int, lr: float,
          vocab_size: int, tokenizer: Tokenizer, device: str) -> tuple[np.ndarray, np.ndarray]:
    """AdamW + CrossEntropy on the (already log-softmaxed) outputs; returns per-epoch (train, valid) losses.

    Reporting quirk kept: epoch losses are sums of batch means divided by the dataset size (:110-111).
    """
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr)
    criterion = nn.CrossEntropyLoss()
    history_train, history_valid = [], []
    best = float("inf")
    for epoch in range(1, nb_epochs + 1):
        model.train()
        running = 0.0
        for start in range(0, len(train_dataset) - 1, batch_size):
            model.zero_grad()
            loss = _batch_loss(model, train_dataset, tokenizer, start, batch_size, vocab_size, device, criterion)
            loss.backward()
            optimizer.step()
            running += loss.item()
        model.eval()
        running_valid = 0.0
        with torch.no_grad():
            for start in range(0, len(valid_dataset) - 1, batch_size):
                running_valid += _batch_loss(model, valid_dataset, tokenizer, start, batch_size, vocab_size, device,
                                             criterion).item()
        train_loss = running / len(train_dataset)
        valid_loss = running_valid / len(valid_dataset)
        history_train.append(train_loss)
        history_valid.append(valid_loss)
        if nb_epochs < 10 or epoch % (nb_epochs // 10) == 0:
            print(f"EPOCH [{epoch} / {nb_epochs}] ----------- TRAIN LOSS : {train_loss:.4f}, VALID LOSS : {valid_loss:.4f}")
        best = min(best, valid_loss)
    print(f"Best valid loss : {best:.4f}")
    return np.array(history_train), np.array(history_valid)


def generate(model: MyTransformer, prompts: torch.Tensor, new_tokens: int, device: str) -> torch.Tensor:
    """Greedy decode with full recompute per token → `(P + new_tokens, B)` (reference :127-148).

    On a CUDA device with autograd off and a deterministic model (`bayes=False`: no always-on dropout) the whole
    decode is one `bfb200_mc_score_transformer` call (greedy scheme, one draw) that returns the trajectories; a
    bayes model keeps the per-token loop, whose `model(seq)` calls also run natively, so that every call draws
    fresh masks exactly like the reference's loop."""
    seq = prompts.to(device)
    if (isinstance(model, MyTransformer) and seq.is_cuda and not torch.is_grad_enabled() and not model.bayes
            and model._train_only_dropout_inactive() and new_tokens > 0 and seq.shape[1] > 0):
        from bayesformer_b200 import engine

        pack = engine.scoring_pack(model, seq.device, seq.shape[0], new_tokens, getattr(model, "native_compute", "auto"))
        res = engine.mc_score_transformer(pack, seq, new_tokens, 1, "greedy", "max", return_tokens=True)
        return res.tokens.t().contiguous().to(torch.int64)
    for _ in range(new_tokens):
        nxt = torch.argmax(model(seq)[-1, :, :], -1).view(1, -1)
        seq = torch.cat((seq, nxt), 0)
    return seq


def evaluate(model: MyTransformer, dataset, tokenizer: Tokenizer, batch_size: int, device: str) -> float:
    """Fraction of items whose generated answer (incl. EOS/PAD tail) matches exactly (reference :151-193)."""
    model.eval()
    hits = 0.0
    with torch.no_grad():
        for start in range(0, len(dataset) - 1, batch_size):
            prompts, answers, p_len, a_len, _, _ = preprocessing.get_batch(dataset, tokenizer, start, batch_size)
            prompts, answers = prompts.to(device), answers.to(device)
            produced = generate(model, prompts, a_len + 1, device)[p_len:, :]
            hits += torch.all(produced == answers, axis=0).float().sum()
    return (hits / len(dataset)).item()
