"""Alias of bayesformer_b200.libs.preprocessing (reference path: src/libs/preprocessing.py)."""
import sys

from bayesformer_b200.libs import preprocessing as _impl

sys.modules[__name__] = _impl
