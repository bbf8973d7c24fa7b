"""Record / replay shim around the UNMODIFIED reference (TEST INFRASTRUCTURE — see oracle/__init__.py).

Only usable where /root/reference exists (the build container): imports the reference's `src`
package under private module objects (our drop-in package is also called `src`) and patches
`torch.nn.functional.dropout` — the object the reference's `F.dropout` resolves to — with a shim
that is bit-identical to it and consumes the generator identically
(`x * (empty_like(x).bernoulli_(1 - p) / (1 - p))`) while recording or replaying masks.
"""

from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("BAYESFORMER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model"))


_cache = None


def load_reference() -> types.SimpleNamespace:
    """Import the reference modules without leaving them (or their name `src`) in sys.modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    # The reference's top-level `src` has no __init__.py (namespace package) and would lose against
    # our regular `src` package on sys.path, so bind the name to the reference directory explicitly.
    pkg = types.ModuleType("src")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
    sys.modules["src"] = pkg
    try:
        names = ["src.configs.constants", "src.libs.tokenizer", "src.libs.preprocessing", "src.model.transformer",
                 "src.libs.active_learning", "src.libs.utils", "src.model.bayesian_classification",
                 "src.libs.preprocessing_classif"]
        mods = {n.rsplit(".", 1)[1]: importlib.import_module(n) for n in names}
    finally:
        sys.dont_write_bytecode = dont_write
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = types.SimpleNamespace(**mods)
    return _cache


@contextlib.contextmanager
def record_dropout(log: list):
    """Every F.dropout(training=True) call appends its keep mask (bool tensor) to `log`."""
    original = F.dropout

    def shim(input, p=0.5, training=True, inplace=False):
        if not training:
            return original(input, p, training, inplace)
        mask = torch.empty_like(input).bernoulli_(1 - p)
        log.append(mask.bool())
        return input * (mask / (1 - p))

    F.dropout = shim
    try:
        yield log
    finally:
        F.dropout = original


@contextlib.contextmanager
def replay_dropout(masks: list):
    """Every F.dropout(training=True) call consumes the next recorded keep mask instead of the RNG."""
    original = F.dropout
    it = iter(masks)

    def shim(input, p=0.5, training=True, inplace=False):
        if not training:
            return original(input, p, training, inplace)
        keep = next(it)
        assert keep.shape == input.shape, (keep.shape, input.shape)
        return input * (keep.to(input.dtype) / (1 - p))

    F.dropout = shim
    try:
        yield
    finally:
        F.dropout = original
