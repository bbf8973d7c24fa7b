"""Alias of bayesformer_b200.libs.tokenizer (reference path: src/libs/tokenizer.py)."""
import sys

from bayesformer_b200.libs import tokenizer as _impl

sys.modules[__name__] = _impl
