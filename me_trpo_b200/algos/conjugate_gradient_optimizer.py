"""ConjugateGradientOptimizer: socket B3 (algos/trpo.py:14-21, algos/npo.py:85-111).

Same constructor arguments, defaults and accept/reject rule as rllab's optimizer (SURVEY.md
Appendix A.2); `optimize(inputs)` runs the whole update -- loss_before, flat gradient, CG on
Fisher-vector products, step scaling, back-tracking line search, parameter restore -- as one
stream of CUDA kernels (metrpo_trpo_update) with no scalar returning to the host."""
import numpy as np
import torch

from ..trpo import PolicyUpdate


class ConjugateGradientOptimizer:
    def __init__(self, cg_iters=10, reg_coeff=1e-5, subsample_factor=1.0, backtrack_ratio=0.8,
                 max_backtracks=15, accept_violation=False, hvp_approach=None, num_slices=1):
        if subsample_factor != 1.0:
            raise NotImplementedError("subsample_factor != 1 is not used by the reference (algos/trpo.py:20)")
        if accept_violation:
            raise NotImplementedError("accept_violation=True is not used by the reference")
        self._cg_iters, self._reg_coeff = cg_iters, reg_coeff
        self._backtrack_ratio, self._max_backtracks = backtrack_ratio, max_backtracks
        self._target = None
        self._max_constraint_val = None
        self._constraint_name = None
        self._kernels = None
        self._dist_ctx = None
        self.last_info = None

    def set_dist_ctx(self, ctx):
        """Row-sharded samples: gradient, Fisher-vector products and (loss, kl) pairs are summed
        over the ranks of `ctx` so that every rank takes the identical natural-gradient step."""
        self._dist_ctx = ctx
        if self._kernels is not None and ctx is not None and ctx.distributed:
            self._kernels.enable_allreduce(ctx.group)

    def update_opt(self, loss, target, leq_constraint, inputs, extra_inputs=None,
                   constraint_name="constraint", *args, **kwargs):
        if loss != "surr_loss" or leq_constraint[0] != "mean_kl":
            raise NotImplementedError("the fused optimizer implements NPO's surr_loss / mean_kl pair (algos/npo.py:68-75)")
        self._target = target
        self._max_constraint_val = float(leq_constraint[1])
        self._constraint_name = constraint_name
        dims = [target.obs_dim] + list(target.hidden_sizes) + [target.action_dim]
        if self._kernels is not None:
            self._kernels.close()
        self._kernels = PolicyUpdate(dims, out_tanh=target.output_tanh, device=target.device)
        if self._dist_ctx is not None and self._dist_ctx.distributed:
            self._kernels.enable_allreduce(self._dist_ctx.group)

    # ------------------------------------------------------------------------------------------
    def _device_inputs(self, inputs):
        if self._kernels is None:
            raise RuntimeError("update_opt() must be called before optimize()")   # rllab asserts likewise
        dev = self._kernels.device
        f = lambda a: (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float32)) \
            .to(device=dev, dtype=torch.float32).contiguous()
        obs, act, adv, mean, log_std = [f(a) for a in inputs[:5]]
        N = adv.numel()
        obs, act, mean = obs.reshape(N, -1), act.reshape(N, -1), mean.reshape(N, -1)
        adv = adv.reshape(N)
        log_std = log_std.reshape(-1) if log_std.numel() == act.shape[1] else log_std.reshape(N, -1)
        valid = None
        if len(inputs) > 5 and inputs[5] is not None:
            v = inputs[5]
            v = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))
            valid = v.to(device=dev, dtype=torch.uint8).reshape(N).contiguous()
        return obs, act, adv, mean, log_std, valid

    def loss(self, inputs, extra_inputs=None):
        obs, act, adv, mean, log_std, valid = self._device_inputs(inputs)
        return self._kernels.loss_kl(self._target.flat_params(), obs, act, adv, mean, log_std, valid)[0]

    def constraint_val(self, inputs, extra_inputs=None):
        obs, act, adv, mean, log_std, valid = self._device_inputs(inputs)
        return self._kernels.loss_kl(self._target.flat_params(), obs, act, adv, mean, log_std, valid)[1]

    def optimize(self, inputs, extra_inputs=None, subsample_grouped_inputs=None):
        obs, act, adv, mean, log_std, valid = self._device_inputs(inputs)
        theta = self._target.flat_params()
        self.last_info = self._kernels.update(
            theta, obs, act, adv, mean, log_std, valid, step_size=self._max_constraint_val,
            cg_iters=self._cg_iters, reg_coeff=self._reg_coeff, backtrack_ratio=self._backtrack_ratio,
            max_backtracks=self._max_backtracks)
        self._target.set_flat_params(theta)
