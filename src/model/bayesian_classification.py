"""Alias of bayesformer_b200.model.bayesian_classification (reference path: src/model/bayesian_classification.py)."""
import sys

from bayesformer_b200.model import bayesian_classification as _impl

sys.modules[__name__] = _impl
