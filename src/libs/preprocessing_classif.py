"""Alias of bayesformer_b200.libs.preprocessing_classif (reference path: src/libs/preprocessing_classif.py)."""
import sys

from bayesformer_b200.libs import preprocessing_classif as _impl

sys.modules[__name__] = _impl
