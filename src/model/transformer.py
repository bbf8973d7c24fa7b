"""Alias of bayesformer_b200.model.transformer (reference path: src/model/transformer.py)."""
import sys

from bayesformer_b200.model import transformer as _impl

sys.modules[__name__] = _impl
