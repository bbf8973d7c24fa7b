"""Toy binary classifiers of the BayesFormer project (make_moons, 2 → hidden → 1).

Drop-in for the reference's `src/model/bayesian_classification.py`: `LinearVariational` (:10-79),
`VariationalMLP` (:82-156), `MLP` (:159-224) with the same parameters and training loops.  The
training code stays eager PyTorch (out of the hot path); what is new is `mc_predict`, the
MC loop of `src/libs/visualization.py:105-113` (`n` stochastic passes → sigmoid → mean) which on
CUDA runs as one fused kernel of libbayesformer_b200.so and also returns the acquisition scores.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import DataLoader


class LinearVariational(nn.Module):
    """Mean-field Gaussian linear layer: W ~ N(mu, softplus(rho)^2), deterministic bias (reference :10-79)."""

    def __init__(self, input_dimension: int, output_dimension: int, prior_var: float) -> None:
        super().__init__()
        self.prior_var = prior_var
        self.w_mu = nn.Parameter(torch.zeros(input_dimension, output_dimension))
        self.w_rho = nn.Parameter(torch.zeros(input_dimension, output_dimension))
        self.b_mu = nn.Parameter(torch.zeros(output_dimension))

    def sampling(self, mu: torch.Tensor, rho: torch.Tensor) -> torch.Tensor:
        """Reparameterised draw `mu + eps * log(1 + exp(rho))` (:42-44)."""
        noise = torch.randn_like(mu)
        return mu + noise * torch.log1p(torch.exp(rho))

    def kl_divergence(self) -> torch.Tensor:
        """Mean KL(q || N(0, prior_var)) over the weights, with the reference's sigma = sqrt(softplus(rho)) (:53-64)."""
        sigma_q = torch.sqrt(torch.log(1 + torch.exp(self.w_rho)))
        sigma_p = torch.full(self.w_mu.shape, float(np.sqrt(self.prior_var)))
        kl = torch.log(sigma_p / sigma_q) + (sigma_q**2 + self.w_mu**2) / (2 * sigma_p**2) - 0.5
        return torch.mean(kl)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.matmul(x, self.sampling(self.w_mu, self.w_rho)) + self.b_mu


def _fit_bce(model: nn.Module, train_loader: DataLoader, nb_epochs: int, extra_loss=None) -> None:
    # Adam(lr=0.01) + BCEWithLogits; the `lr` argument of train_model is ignored by the reference (:147, :215).
    criterion = nn.BCEWithLogitsLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=0.01)
    model.train()
    for _ in range(nb_epochs):
        for x, y in train_loader:
            optimizer.zero_grad()
            loss = criterion(model(x).squeeze(), y)
            if extra_loss is not None:
                loss = extra_loss() + loss
            loss.backward()
            optimizer.step()
    print(f"Training completed after {nb_epochs} epochs.")


class VariationalMLP(nn.Module):
    """Bayes-by-backprop MLP: every forward call samples fresh weights for both layers (reference :82-156)."""

    def __init__(self, input_dimension: int, hidden_dimension: int, output_dimension: int, prior_var: float) -> None:
        super().__init__()
        self.hidden_layer = LinearVariational(input_dimension, hidden_dimension, prior_var)
        self.output_layer = LinearVariational(hidden_dimension, output_dimension, prior_var)

    def kl_divergence(self) -> torch.Tensor:
        return self.hidden_layer.kl_divergence() + self.output_layer.kl_divergence()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.output_layer(F.relu(self.hidden_layer(x)))

    def train_model(self, train_loader: DataLoader, nb_epochs: int = 100, lr: float = 0.01) -> None:
        _fit_bce(self, train_loader, nb_epochs, extra_loss=self.kl_divergence)

    def mc_predict(self, x: torch.Tensor, nsamples: int = 100, return_scores: bool = False):
        """Predictive mean of `sigmoid(self(x))` over `nsamples` weight draws (visualization.py:105-113)."""
        if x.is_cuda and not torch.is_grad_enabled():
            from bayesformer_b200 import engine

            h, o = self.hidden_layer, self.output_layer
            n_eps = h.w_mu.numel() + o.w_mu.numel()
            # one N(0,1) tensor per layer per draw, hidden layer first — the order forward() samples in
            eps = torch.randn(nsamples, n_eps, device=x.device, dtype=torch.float32)
            res = engine.vmlp_mc_score(x, h.w_mu, h.w_rho, h.b_mu, o.w_mu, o.w_rho, o.b_mu, eps, return_probs=False)
            return res if return_scores else res.mean_p
        outs = torch.stack([torch.sigmoid(self(x)) for _ in range(nsamples)])
        return outs.mean(0).squeeze()


class MLP(nn.Module):
    """Deterministic MLP, or MC-dropout MLP when `dropout=True` (mask after the ReLU, always on; reference :159-224)."""

    def __init__(self, input_dimension: int, hidden_dimension: int, output_dimension: int, dropout: bool = False,
                 dropout_rate: float | None = None) -> None:
        super().__init__()
        self.hidden_layer = nn.Linear(input_dimension, hidden_dimension)
        self.output_layer = nn.Linear(hidden_dimension, output_dimension)
        self.dropout = dropout
        self.dropout_rate = dropout_rate

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        hidden = F.relu(self.hidden_layer(x))
        if self.dropout:
            hidden = F.dropout(hidden, p=self.dropout_rate, training=True)
        return self.output_layer(hidden)

    def train_model(self, train_loader: DataLoader, nb_epochs: int = 100, lr: float = 0.01) -> None:
        _fit_bce(self, train_loader, nb_epochs)

    def mc_predict(self, x: torch.Tensor, nsamples: int = 100, return_scores: bool = False):
        """Predictive mean of `sigmoid(self(x))` over `nsamples` dropout draws (visualization.py:105-113)."""
        if x.is_cuda and not torch.is_grad_enabled():
            from bayesformer_b200 import engine

            p = float(self.dropout_rate) if self.dropout else 0.0
            res = engine.mlp_mc_score(x, self.hidden_layer.weight, self.hidden_layer.bias, self.output_layer.weight,
                                      self.output_layer.bias, p, nsamples, return_probs=False)
            return res if return_scores else res.mean_p
        outs = torch.stack([torch.sigmoid(self(x)) for _ in range(nsamples)])
        return outs.mean(0).squeeze()
