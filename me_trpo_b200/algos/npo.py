"""NPO (algos/npo.py:11-115): natural policy optimisation.  The reference builds the TF graph
surr_loss = -mean(lr * adv), mean_kl = mean(kl_sym(old, new)) (:68-75) and hands it to the
optimizer (:85-91); here the optimizer owns fused CUDA kernels for exactly that loss/constraint
pair, so `init_opt` passes the names instead of symbolic tensors."""
from .batch_polopt import BatchPolopt


class NPO(BatchPolopt):
    def __init__(self, optimizer=None, optimizer_args=None, step_size=0.01, **kwargs):
        if optimizer is None:
            # reference default: PenaltyLbfgsOptimizer (:24-27) -- never used by ME-TRPO, which
            # always goes through TRPO (algos/trpo.py:17-20)
            raise NotImplementedError("NPO needs an optimizer; use TRPO for the reference's path")
        self.optimizer = optimizer
        self.step_size = step_size
        super().__init__(**kwargs)

    def init_opt(self):
        self.optimizer.update_opt(
            loss="surr_loss", target=self.policy, leq_constraint=("mean_kl", self.step_size),
            inputs=["obs", "action", "advantage", "old_mean", "old_log_std"], constraint_name="mean_kl")
        return dict()

    def optimize_policy(self, itr, samples_data):
        agent_infos = samples_data["agent_infos"]
        all_input_values = (samples_data["observations"], samples_data["actions"],
                            samples_data["advantages"], agent_infos["mean"], agent_infos["log_std"])
        if samples_data.get("valids") is not None:     # flat device buffers carry a validity mask
            all_input_values += (samples_data["valids"],)
        self.optimizer.optimize(all_input_values)
        return dict()
