"""Alias of bayesformer_b200.libs.utils (reference path: src/libs/utils.py)."""
import sys

from bayesformer_b200.libs import utils as _impl

sys.modules[__name__] = _impl
