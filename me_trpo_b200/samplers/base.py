"""BaseSampler.process_samples (samplers/base.py:48-182) for the non-recurrent case: baseline
prediction, TD residuals, discounted cumulative sums, advantage centring, baseline refit.

Host (NumPy) implementation operating on the list-of-paths format of socket B2; the on-device
version for flat trajectory buffers is the N1 row of SURVEY.md 8(f)."""
import numpy as np


def discount_cumsum(x, discount):
    """y_t = x_t + discount * y_{t+1} (rllab special.discount_cumsum)."""
    y = np.zeros(len(x), dtype=np.float64)
    run = 0.0
    for t in range(len(x) - 1, -1, -1):
        run = x[t] + discount * run
        y[t] = run
    return y


class BaseSampler:
    def __init__(self, algo, dist_ctx=None):
        self.algo = algo
        self.dist_ctx = dist_ctx          # me_trpo_b200.parallel.DistContext or None (single GPU)
        self._update_kernels = None

    def _kernels(self):
        if self._update_kernels is None:
            from ..trpo import PolicyUpdate
            pol = self.algo.policy
            dims = [pol.obs_dim] + list(pol.hidden_sizes) + [pol.action_dim]
            self._update_kernels = PolicyUpdate(dims, out_tanh=pol.output_tanh, device=pol.device)
            if self.dist_ctx is not None and self.dist_ctx.distributed:
                # rows are sharded over ranks: advantage moments and the baseline's normal equations
                # are summed over the group, so every rank centres / fits on the whole batch
                self._update_kernels.enable_allreduce(self.dist_ctx.group)
        return self._update_kernels

    def process_samples_flat(self, itr, flat):
        """process_samples (samplers/base.py:48-182) on the time-major DEVICE buffers of
        `obtain_samples_flat`: no list of paths, no host copy (metrpo_trpo_process +
        metrpo_trpo_fit_baseline).  Samples of paths left unfinished at the end of the buffer carry
        valids == 0 (the reference drops those paths, samplers/vectorized_sampler.py:80-105)."""
        import torch
        algo, ku, bl = self.algo, self._kernels(), self.algo.baseline
        pr = ku.process(flat["obs"], flat["rew"], flat["done"], bl.device_coeffs(ku.device),
                        discount=algo.discount, gae_lambda=algo.gae_lambda, center_adv=algo.center_adv,
                        positive_adv=getattr(algo, "positive_adv", False))
        T, B = flat["rew"].shape
        N = T * B
        log_std = torch.clamp(algo.policy.log_std, min=float(np.log(1e-6)))
        samples_data = dict(
            observations=flat["obs"].reshape(N, -1), actions=flat["act"].reshape(N, -1),
            rewards=flat["rew"].reshape(N), returns=pr["ret"].reshape(N), advantages=pr["adv"].reshape(N),
            valids=pr["valid"].reshape(N), env_infos={},
            agent_infos=dict(mean=flat["mean"].reshape(N, -1), log_std=log_std), stats=pr["stats"])
        # after the advantages: iteration j uses the fit of j-1 (:167)
        bl.set_device_coeffs(ku.fit_baseline(flat["obs"], pr["ret"], pr["valid"], flat["done"],
                                             reg_coeff=bl._reg_coeff))
        return samples_data

    def process_samples(self, itr, paths):
        algo = self.algo
        gamma, lam = algo.discount, algo.gae_lambda
        path_baselines = [algo.baseline.predict(p) for p in paths]
        for p, b in zip(paths, path_baselines):
            b1 = np.append(b, 0.0)
            deltas = p["rewards"] + gamma * b1[1:] - b1[:-1]
            p["advantages"] = discount_cumsum(deltas, gamma * lam)
            p["returns"] = discount_cumsum(p["rewards"], gamma)
        cat = lambda key: np.concatenate([p[key] for p in paths])
        advantages = cat("advantages")
        if algo.center_adv:
            advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
        if getattr(algo, "positive_adv", False):
            advantages = advantages - advantages.min() + 1e-8
        samples_data = dict(
            observations=cat("observations"), actions=cat("actions"), rewards=cat("rewards"),
            returns=cat("returns"), advantages=advantages, env_infos={},
            agent_infos={k: np.concatenate([p["agent_infos"][k] for p in paths])
                         for k in paths[0]["agent_infos"]},
            paths=paths)
        algo.baseline.fit(paths)   # after the advantages: iteration j uses the fit of j-1
        return samples_data
