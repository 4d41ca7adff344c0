"""Mirrors of the reference's algo classes on the TRPO path (algos/batch_polopt.py, algos/npo.py,
algos/trpo.py) and of rllab's ConjugateGradientOptimizer (socket B3)."""
from .batch_polopt import BatchPolopt
from .npo import NPO
from .trpo import TRPO
from .conjugate_gradient_optimizer import ConjugateGradientOptimizer

__all__ = ["BatchPolopt", "NPO", "TRPO", "ConjugateGradientOptimizer"]
