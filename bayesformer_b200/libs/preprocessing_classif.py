"""make_moons toy data for the classifiers (mirror of reference src/libs/preprocessing_classif.py:10-46)."""

from __future__ import annotations

import torch
from sklearn.datasets import make_moons
from torch.utils.data import DataLoader, TensorDataset

from bayesformer_b200.configs import constants


def get_data(nb_samples: int = 1000, noise: float = 0.1) -> tuple[torch.Tensor, torch.Tensor]:
    """`(X (n, 2) float32, y (n,) float32)` from `make_moons(random_state=RANDOM_SEED)` (:23-28)."""
    X, y = make_moons(n_samples=nb_samples, noise=noise, random_state=constants.RANDOM_SEED)
    return torch.from_numpy(X).type(torch.float), torch.from_numpy(y).type(torch.float)


def get_dataloader(X: torch.Tensor, y: torch.Tensor, batch_size: int = 32) -> DataLoader:
    """Sequential (unshuffled) loader over `(X, y)` (:31-46)."""
    return DataLoader(TensorDataset(X, y), batch_size=batch_size)
