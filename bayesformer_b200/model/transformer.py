"""Decoder-only transformer of the BayesFormer project behind its original class names.

Drop-in for the reference's `src/model/transformer.py`: identical constructors, parameter /
buffer names (`layers.{i}.self_attn.heads.{h}.W_q.weight`, `pos_enc.pe`, ...) and arithmetic,
so `state_dict`, `copy.deepcopy`, `.to(device)` and autograd through `forward` behave the same.

Two execution routes share these modules:

* **native** — CUDA tensors and autograd disabled (the MC-scoring situation): `MyTransformer.forward`
  hands the whole pass to libbayesformer_b200.so (`engine.transformer_forward`); dropout masks
  come from the counter-based Philox stream instead of `F.dropout`.  A missing library raises.
* **module** — anything that needs autograd (`utils.train`) or lives on the CPU runs the layer
  stack below, which restates the reference formulas with ATen ops.

Shapes are seq-first `(T, B, E)` throughout, as in the reference.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _mc_dropout(x: torch.Tensor, p: float | None) -> torch.Tensor:
    # "Bayesian" dropout of the reference: active regardless of train()/eval().
    return F.dropout(x, p=p, training=True)


class Head(nn.Module):
    """One attention head with its own bias-free Q/K/V projections (reference :9-65).

    `d_ff` is the head width (the reference's argument name).  With `bayes=True` the three
    projections each see an independently masked copy of the input (:50-53).
    """

    def __init__(self, d_model: int, d_ff: int, causal: bool = True, bayes: bool = False,
                 bayes_dropout: float | None = None):
        super().__init__()
        self.d_head = d_ff
        self.causal = causal
        self.bayes = bayes
        self.bayes_dropout = bayes_dropout
        self.W_q = nn.Linear(d_model, d_ff, bias=False)
        self.W_k = nn.Linear(d_model, d_ff, bias=False)
        self.W_v = nn.Linear(d_model, d_ff, bias=False)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        if self.bayes:
            xq, xk, xv = (_mc_dropout(x, self.bayes_dropout) for _ in range(3))
        else:
            xq = xk = xv = x
        # (T, B, dh) -> (B, T, dh)
        q = self.W_q(xq).transpose(0, 1)
        k = self.W_k(xk).transpose(0, 1)
        v = self.W_v(xv).transpose(0, 1)
        scores = (q @ k.transpose(1, 2)) / math.sqrt(self.d_head)
        if self.causal and mask is not None:
            scores = scores + mask
        weights = F.softmax(scores, dim=-1)
        return (weights @ v).transpose(0, 1)


class MultiHeadAttention(nn.Module):
    """`n_heads` independent :class:`Head` s, concatenated, then a bias-free `W_out` (reference :68-125).

    Note the reference quirk kept here: `bayes` / `bayes_dropout` are stored but not forwarded to
    the heads (:97-99), and `d_ff` is accepted and ignored.
    """

    def __init__(self, d_model: int, n_heads: int, d_ff: int, causal: bool = True,
                 bayes_dropout: float | None = None, bayes: bool = False):
        super().__init__()
        assert d_model % n_heads == 0, "d_model must be divisible by n_heads"
        self.d_model = d_model
        self.n_heads = n_heads
        self.d_head = d_model // n_heads
        self.causal = causal
        self.bayes = bayes
        self.bayes_dropout = bayes_dropout
        self.heads = nn.ModuleList(Head(d_model, self.d_head, causal) for _ in range(n_heads))
        self.W_out = nn.Linear(d_model, d_model, bias=False)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        outs = []
        for head in self.heads:
            o = head(x, mask)
            if self.bayes:
                o = _mc_dropout(o, self.bayes_dropout)
            outs.append(o)
        return self.W_out(torch.cat(outs, dim=-1))


class FeedForward(nn.Module):
    """Linear → ReLU → Linear (reference :128-170); input dropout whenever a rate is given (:167)."""

    def __init__(self, d_model: int, dim_feedforward: int, bayes_dropout: float | None = None):
        super().__init__()
        self.bayes_dropout = bayes_dropout
        self.net = nn.Sequential(
            nn.Linear(d_model, dim_feedforward),
            nn.ReLU(),
            nn.Linear(dim_feedforward, d_model),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.bayes_dropout is not None:
            x = _mc_dropout(x, self.bayes_dropout)
        return self.net(x)


class TransformerBlock(nn.Module):
    """Post-LN decoder block (reference :173-234).

    Both residual connections start from the *block input* `x` (:231-233) — the second one does
    not use the norm1 output — and that is reproduced here and in the CUDA path.
    """

    def __init__(self, d_model: int, n_heads: int, dim_feedforward: int, dropout: float = 0.1,
                 causal: bool = True, bayes_dropout: float | None = None, bayes: bool = False):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, n_heads, d_ff=dim_feedforward, causal=causal)
        self.norm1 = nn.LayerNorm(d_model)
        self.ff = FeedForward(d_model, dim_feedforward, bayes_dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.bayes = bayes
        self.bayes_dropout = bayes_dropout

    def _branch_dropout(self, t: torch.Tensor) -> torch.Tensor:
        return _mc_dropout(t, self.bayes_dropout) if self.bayes else self.dropout(t)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        attn = self.self_attn(x, mask=mask)
        hidden = self.norm1(x + self._branch_dropout(attn))
        ffn = self.ff(hidden)
        return self.norm2(x + self._branch_dropout(ffn))


class PositionalEncoding(nn.Module):
    """Sinusoidal table `pe` (max_len, 1, d) added to the input, then train-only dropout (reference :237-272)."""

    def __init__(self, d_model: int, max_len: int = 5000, dropout: float = 0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(-math.log(10000.0) * torch.arange(0, d_model, 2).float() / d_model)
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(x + self.pe[: x.shape[0], :])


def generate_subsequent_mask(sz: int) -> torch.Tensor:
    """(sz, sz) float mask, -inf strictly above the diagonal, 0 elsewhere (reference :275-287)."""
    return torch.triu(torch.full((sz, sz), float("-inf")), diagonal=1)


_MASK_CACHE: dict[tuple[int, str], torch.Tensor] = {}


def _causal_mask_on(sz: int, device: torch.device) -> torch.Tensor:
    """The same mask, built once per (size, device): the reference rebuilds it on the host and copies it over in every
    forward (:366), which also cannot be captured in a CUDA graph (utils.train's graphed step)."""
    key = (sz, str(device))
    mask = _MASK_CACHE.get(key)
    if mask is None:
        mask = _MASK_CACHE[key] = generate_subsequent_mask(sz).to(device)
    return mask


class MyTransformer(nn.Module):
    """Embedding → (+MC dropout) → ×sqrt(d) → +PE → blocks (→ MC dropout each) → vocab head → log-softmax.

    Reference :290-375.  With `bayes=True` exactly `1 + nlayers` dropout sites are live (embedding
    and every block output, :355-358 and :368-370): the blocks are always built without the bayes
    arguments (:324-335), which this class reproduces.
    """

    def __init__(self, ntoken: int, d_model: int, n_heads: int, dim_feedforward: int, nlayers: int,
                 dropout: float = 0.1, bayes_dropout: float | None = None, bayes: bool = False):
        super().__init__()
        self.ntoken = ntoken
        self.d_model = d_model
        self.embedding = nn.Embedding(ntoken, d_model)
        self.pos_enc = PositionalEncoding(d_model, dropout=dropout)
        self.layers = nn.ModuleList(
            TransformerBlock(d_model, n_heads, dim_feedforward, dropout=dropout, causal=True)
            for _ in range(nlayers)
        )
        self.linear_out = nn.Linear(d_model, ntoken)
        self._init_weights()
        self.bayes = bayes
        self.bayes_dropout = bayes_dropout

    def _init_weights(self) -> None:
        nn.init.xavier_uniform_(self.embedding.weight)
        nn.init.xavier_uniform_(self.linear_out.weight)
        nn.init.zeros_(self.linear_out.bias)

    # -- routing ---------------------------------------------------------------------------
    def _train_only_dropout_inactive(self) -> bool:
        """True when the train-only nn.Dropout sites (:231, :233, :272) cannot fire: eval mode or rate 0."""
        if not self.training:
            return True
        return self.pos_enc.dropout.p == 0 and all(l.dropout.p == 0 for l in self.layers)

    def _use_native(self, src: torch.Tensor) -> bool:
        """CUDA tensors and no autograd → the sm_100a library (inference semantics).  Anything that needs
        autograd (utils.train) runs the nn.Module stack below; that is the training path, not a fallback
        of the scoring path — `active_learning.*` never uses it."""
        if not src.is_cuda or torch.is_grad_enabled():
            return False
        return self._train_only_dropout_inactive()

    def native_pack(self, device: torch.device | None = None, compute: str = "fp32"):
        """Device-resident packed weights for the C ABI (cached; see engine.packed)."""
        from bayesformer_b200 import engine

        if device is None:
            device = self.embedding.weight.device
        return engine.packed(self, torch.device(device), compute)

    def forward(self, src: torch.Tensor) -> torch.Tensor:
        """src (T, B) token ids → (T, B, ntoken) log-probabilities."""
        if self._use_native(src):
            from bayesformer_b200 import engine

            return engine.transformer_forward(self.native_pack(src.device), src)
        return self._module_forward(src)

    def _module_forward(self, src: torch.Tensor) -> torch.Tensor:
        x = self.embedding(src)
        if self.bayes:
            x = _mc_dropout(x, self.bayes_dropout)
        x = self.pos_enc(x * math.sqrt(self.d_model))
        mask = _causal_mask_on(src.shape[0], src.device)
        for layer in self.layers:
            x = layer(x, mask=mask)
            if self.bayes:
                x = _mc_dropout(x, self.bayes_dropout)
        return F.log_softmax(self.linear_out(x), dim=-1)
