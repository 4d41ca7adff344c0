"""TRPO (algos/trpo.py:7-21): NPO with rllab's ConjugateGradientOptimizer, all defaults."""
from .npo import NPO
from .conjugate_gradient_optimizer import ConjugateGradientOptimizer


class TRPO(NPO):
    def __init__(self, optimizer=None, optimizer_args=None, **kwargs):
        if optimizer is None:
            if optimizer_args is None:
                optimizer_args = dict()
            optimizer = ConjugateGradientOptimizer(**optimizer_args)
        super().__init__(optimizer=optimizer, **kwargs)
