"""BatchPolopt (algos/batch_polopt.py:17-108): owns env / policy / baseline / sampler and the
sampling hyper-parameters the sampler reads (`batch_size`, `max_path_length`, `discount`,
`gae_lambda`, `center_adv`, `positive_adv`)."""
from ..samplers import VectorizedSampler


class BatchPolopt:
    def __init__(self, env, policy, baseline, scope=None, n_itr=500, start_itr=0, batch_size=5000,
                 max_path_length=500, discount=0.99, gae_lambda=1, plot=False, pause_for_plot=False,
                 center_adv=True, positive_adv=False, store_paths=False, whole_paths=True,
                 fixed_horizon=False, sampler_cls=None, sampler_args=None, force_batch_sampler=False,
                 **kwargs):
        self.env, self.policy, self.baseline, self.scope = env, policy, baseline, scope
        self.n_itr, self.start_itr = n_itr, start_itr
        self.batch_size, self.max_path_length = batch_size, max_path_length
        self.discount, self.gae_lambda = discount, gae_lambda
        self.plot, self.pause_for_plot = plot, pause_for_plot
        self.center_adv, self.positive_adv = center_adv, positive_adv
        self.store_paths, self.whole_paths, self.fixed_horizon = store_paths, whole_paths, fixed_horizon
        self.kwargs = kwargs
        if sampler_cls is None:
            if not getattr(self.policy, "vectorized", False) or force_batch_sampler:
                # the reference falls back to the single-env BatchSampler (:88-91); the imaginary
                # rollout path always uses the vectorized sampler (model_based_rl.py:375-380)
                raise NotImplementedError("BatchSampler is outside the imaginary-rollout hot path")
            sampler_cls = VectorizedSampler
        if sampler_args is None:
            sampler_args = dict()
        self.sampler = sampler_cls(self, **sampler_args)
        self.init_opt()

    def start_worker(self):
        self.sampler.start_worker()

    def shutdown_worker(self):
        self.sampler.shutdown_worker()

    def obtain_samples(self, itr, determ=False):
        return self.sampler.obtain_samples(itr, determ)

    def process_samples(self, itr, paths):
        return self.sampler.process_samples(itr, paths)

    # -- device-resident variants (no list of paths, no host copy) -------------------------------
    def obtain_samples_flat(self, itr, determ=False):
        return self.sampler.obtain_samples_flat(itr, determ)

    def process_samples_flat(self, itr, flat):
        return self.sampler.process_samples_flat(itr, flat)

    def init_opt(self):
        raise NotImplementedError

    def optimize_policy(self, itr, samples_data):
        raise NotImplementedError
