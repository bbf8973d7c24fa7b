"""Training / greedy generation / exact-match evaluation around MyTransformer.

Drop-in for the reference's `src/libs/utils.py` (`train` :12-124, `generate` :127-148,
`evaluate` :151-193).  These are the callers on either side of the scoring path:
`train` stays eager autograd; `generate` under `torch.no_grad()` on CUDA goes through the native
forward via `MyTransformer.forward`.
"""

from __future__ import annotations

import weakref

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


class _GraphedTrainStep:
    """One training step (forward, loss, backward, AdamW) of a fixed batch shape as ONE CUDA graph.

    The reference's step is ~600 kernels of a few microseconds each on this model (50 k parameters, 20 x 9 tokens):
    launch-bound, 4-5 ms in eager mode.  Captured once per (prompt length, answer length, batch) shape and replayed,
    the same kernels run back to back.  Capturing needs warm-up runs of the step; parameters and optimizer state are
    saved before and restored IN PLACE after them, so creating a graph never advances the training."""

    def __init__(self, model, optimizer, criterion, p_len: int, a_len: int, batch: int, vocab_size: int, device):
        self.tokens = torch.zeros(p_len + a_len + 1, batch, dtype=torch.int64, device=device)
        self.loss = None
        self.p_len, self.vocab = p_len, vocab_size
        params = [p for p in model.parameters()]
        saved_p = [p.detach().clone() for p in params]
        saved_s = {id(p): {k: v.clone() for k, v in optimizer.state.get(p, {}).items() if torch.is_tensor(v)} for p in params}
        had_state = {id(p): p in optimizer.state and len(optimizer.state[p]) > 0 for p in params}
        rng = torch.cuda.get_rng_state(device)

        def step():
            out = model(self.tokens)
            pred = out[self.p_len - 1:-1, :, :].reshape(-1, self.vocab)
            loss = criterion(pred, self.tokens[self.p_len:].reshape(-1))
            loss.backward()
            optimizer.step()
            return loss

        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):
                optimizer.zero_grad(set_to_none=True)
                step()
        torch.cuda.current_stream(device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = step()
        # undo the warm-up steps (capture itself executes nothing)
        with torch.no_grad():
            for p, sp in zip(params, saved_p):
                p.copy_(sp)
                for k, v in optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if had_state[id(p)]:
                            v.copy_(saved_s[id(p)][k])
                        else:
                            v.zero_()
        torch.cuda.set_rng_state(rng, device)

    def run(self, prompts: torch.Tensor, answers: torch.Tensor) -> torch.Tensor:
        self.tokens[:self.p_len].copy_(prompts, non_blocking=True)
        self.tokens[self.p_len:].copy_(answers, non_blocking=True)
        self.graph.replay()
        return self.loss


# model -> (capturable AdamW, {batch shape: graph}, parameter identity).  Captured steps are reused by later `train` calls
# on the same model (every active-learning round fine-tunes it again): a fresh call only resets the optimizer state in
# place (zero moments, step 0 — what a new `torch.optim.AdamW` would start from) and sets the learning rate.
_GRAPH_STATE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _graph_state(model, lr: float, criterion, device):
    ident = tuple((p.data_ptr(), tuple(p.shape)) for p in model.parameters()) + (str(torch.device(device)),)
    state = _GRAPH_STATE.get(model)
    if state is None or state[2] != ident:
        optimizer = torch.optim.AdamW(model.parameters(), lr=torch.tensor(float(lr), device=device), capturable=True)
        state = _GRAPH_STATE[model] = (optimizer, {}, ident)
    optimizer, graphs, _ = state
    with torch.no_grad():
        for group in optimizer.param_groups:
            group["lr"].fill_(float(lr))
        for st in optimizer.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
    return optimizer, graphs


_FUSED_TRAINERS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _fused_enabled(model, device) -> bool:
    """The hand-written fused step (csrc/train_fused.cu) serves the notebook model class on a CUDA device; set
    BFB200_FUSED_TRAIN=0 for the graphed / eager autograd loop."""
    import os

    if torch.device(device).type != "cuda" or os.environ.get("BFB200_FUSED_TRAIN", "1") == "0":
        return False
    from bayesformer_b200 import engine

    return isinstance(model, MyTransformer) and engine.FusedTrainer.supports(model)


def _fused_trainer(model, lr: float):
    """Per-model trainer; like the reference, every train() call starts from a FRESH AdamW state (utils.py:51)."""
    from bayesformer_b200 import engine

    tr = _FUSED_TRAINERS.get(model)
    if tr is None or not tr.valid_for(model):
        tr = _FUSED_TRAINERS[model] = engine.FusedTrainer(model, lr)
    tr.lr = float(lr)
    tr.adam_m.zero_()
    tr.adam_v.zero_()
    tr.step_count = 0
    return tr


def _train_dropout_rate(model) -> float:
    """nn.Dropout rate of the training-only sites (one rate for the whole model: MyTransformer(dropout=...))."""
    return float(model.pos_enc.dropout.p)


def _graphs_enabled(device) -> bool:
    import os

    return torch.device(device).type == "cuda" and os.environ.get("BFB200_GRAPH_TRAIN", "1") != "0"


def train(model: MyTransformer, train_dataset, valid_dataset, nb_epochs: int, batch_size: int, lr: float,
          vocab_size: int, tokenizer: Tokenizer, device: str) -> tuple[np.ndarray, np.ndarray]:
    """AdamW + CrossEntropy on the (already log-softmaxed) outputs; returns per-epoch (train, valid) losses.

    Reporting quirk kept: epoch losses are sums of batch means divided by the dataset size (:110-111).
    On a CUDA device every distinct batch shape is captured once as a CUDA graph (`_GraphedTrainStep`; set
    BFB200_GRAPH_TRAIN=0 for the plain eager loop) and the batch losses are summed on the device, read once per epoch.
    """
    use_fused = _fused_enabled(model, device)
    use_graphs = not use_fused and _graphs_enabled(device)
    criterion = nn.CrossEntropyLoss()
    trainer = None
    batches: dict[int, tuple] = {}
    if use_fused:
        # the fused step keeps a query's score row in registers: sequences of at most 16 tokens (the notebooks' are 9);
        # every batch is tokenised and copied to the device once per call
        from bayesformer_b200 import engine

        for start in range(0, len(train_dataset) - 1, batch_size):
            prompts, answers, p_len, a_len, _, _ = preprocessing.get_batch(train_dataset, tokenizer, start, batch_size)
            batches[start] = (torch.cat((prompts, answers), 0).to(device), p_len)
        if any(t.shape[0] > engine.FusedTrainer.MAX_SEQ for t, _ in batches.values()):
            use_fused, use_graphs, batches = False, _graphs_enabled(device), {}
    if use_fused:
        trainer = _fused_trainer(model, lr)
        optimizer, graphs = None, {}
    elif use_graphs:
        optimizer, graphs = _graph_state(model, lr, criterion, device)
    else:
        optimizer, graphs = torch.optim.AdamW(model.parameters(), lr=lr), {}
    history_train, history_valid = [], []
    best = float("inf")
    try:
        for epoch in range(1, nb_epochs + 1):
            model.train()
            running = 0.0
            running_dev = torch.zeros((), device=device) if use_graphs else None
            if use_fused:
                trainer.loss.zero_()
            for start in range(0, len(train_dataset) - 1, batch_size):
                if use_fused:
                    # ONE kernel per batch: forward, loss, backward, AdamW (csrc/train_fused.cu)
                    tokens, p_len = batches[start]
                    trainer.step(tokens, p_len, _train_dropout_rate(model), float(model.bayes_dropout) if model.bayes else 0.0)
                    continue
                if use_graphs:
                    # the reference walks the same slices in the same order every epoch (no shuffling, :60): each batch is
                    # tokenised and copied to the device once per call
                    if start not in batches:
                        prompts, answers, p_len, a_len, _, _ = preprocessing.get_batch(train_dataset, tokenizer, start, batch_size)
                        batches[start] = (prompts.to(device), answers.to(device), p_len, a_len)
                    prompts, answers, p_len, a_len = batches[start]
                    key = (p_len, a_len, prompts.shape[1])
                    if key not in graphs:
                        graphs[key] = _GraphedTrainStep(model, optimizer, criterion, p_len, a_len, prompts.shape[1], vocab_size,
                                                        device)
                    running_dev += graphs[key].run(prompts, answers).detach()
                    continue
                model.zero_grad()
                loss = _batch_loss(model, train_dataset, tokenizer, start, batch_size, vocab_size, device, criterion)
                loss.backward()
                optimizer.step()
                running += loss.item()
            if use_fused:
                running = float(trainer.loss.item())
            if use_graphs:
                running = float(running_dev.item())
                from bayesformer_b200 import engine

                engine.mark_weights_changed(model)   # graph replays update the weights without bumping tensor versions
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

    finally:
        if use_graphs:   # also on KeyboardInterrupt / an exception mid-epoch: the packs must never be half-fresh
            from bayesformer_b200 import engine

            engine.mark_weights_changed(model)
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

        pack = engine.scoring_pack(model, seq.device, seq.shape[0], new_tokens, engine.default_compute(model, "greedy"))
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
