"""TRPO (algos/trpo.py:7-21): NPO with rllab's ConjugateGradientOptimizer, all defaults.

`dist_ctx` (me_trpo_b200.parallel.DistContext, optional) row-shards the sampler over the ranks
and all-reduces the accumulators of process_samples / the optimizer."""
from .npo import NPO
from .conjugate_gradient_optimizer import ConjugateGradientOptimizer


class TRPO(NPO):
    def __init__(self, optimizer=None, optimizer_args=None, dist_ctx=None, **kwargs):
        if optimizer is None:
            if optimizer_args is None:
                optimizer_args = dict()
            optimizer = ConjugateGradientOptimizer(**optimizer_args)
        if dist_ctx is not None:
            if hasattr(optimizer, "set_dist_ctx"):
                optimizer.set_dist_ctx(dist_ctx)
            sampler_args = dict(kwargs.pop("sampler_args", None) or {})
            sampler_args.setdefault("dist_ctx", dist_ctx)
            kwargs["sampler_args"] = sampler_args
        self.dist_ctx = dist_ctx
        super().__init__(optimizer=optimizer, **kwargs)
