"""Shared fixture loading for the test-suite (oracle side only — never imported by the product)."""

from __future__ import annotations

import os

import numpy as np
import torch

from oracle import restatement as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = ("max", "margin", "entropy")


def load(name: str):
    return np.load(os.path.join(GOLDEN, name))


def sinusoid_pe(d: int, max_len: int = 5000) -> torch.Tensor:
    """pos_enc.pe buffer, rebuilt exactly as the model constructor does (it is not stored in fixtures)."""
    from bayesformer_b200.model.transformer import PositionalEncoding

    return PositionalEncoding(d, max_len=max_len).pe.clone()


def state_dict_from(npz, prefix: str = "sd.") -> dict:
    sd = {k[len(prefix):]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(prefix)}
    sd["pos_enc.pe"] = sinusoid_pe(sd["embedding.weight"].shape[1])
    return sd


def build_model(sd: dict, p, bayes: bool):
    """Our MyTransformer carrying the fixture weights."""
    from bayesformer_b200.model.transformer import MyTransformer

    dims = R.model_dims(sd)
    m = MyTransformer(dims["V"], dims["d"], dims["H"], dims["F"], dims["L"], bayes_dropout=p if bayes else None, bayes=bayes)
    m.load_state_dict(sd)
    return m.eval()


def expand_live_bits(live: np.ndarray, L: int) -> np.ndarray:
    """[step][E, LO_0..][seq][pos][w] (fixture, live sites only) -> C-ABI layout with all 1+5L sites."""
    n_steps, _, S, max_len, words = live.shape
    full = np.zeros((n_steps, R.n_sites(L), S, max_len, words), dtype=np.uint32)
    full[:, R.SITE_EMBED] = live[:, 0]
    for i in range(L):
        full[:, R.site_layer_out(i)] = live[:, 1 + i]
    return full


def packed_masks_from_bits(bits: np.ndarray, L: int, d: int) -> R.PackedMasks:
    pm = R.PackedMasks(bits.shape[0], L, bits.shape[2], bits.shape[3], d)
    pm.bits = bits
    return pm


def reference_mask_list(live: np.ndarray, T: int, P: int, N: int, d: int):
    """The reference's F.dropout call order (for draw, for step: E, LO_0, LO_1) from the fixture bits."""
    masks = []
    B = live.shape[2] // T
    for draw in range(T):
        for step in range(N):
            for site in range(live.shape[1]):
                keep = R.unpack_bits(live[step, site, draw::T, : P + step, :], d)  # (B, pos, d)
                masks.append(torch.from_numpy(np.ascontiguousarray(np.transpose(keep, (1, 0, 2)))))
    assert len(masks) and masks[0].shape[1] == B
    return masks
