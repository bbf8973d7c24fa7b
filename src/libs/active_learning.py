"""Alias of bayesformer_b200.libs.active_learning (reference path: src/libs/active_learning.py)."""
import sys

from bayesformer_b200.libs import active_learning as _impl

sys.modules[__name__] = _impl
