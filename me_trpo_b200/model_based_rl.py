"""optimize_policy: the TRPO branch of the reference's policy-improvement controller
(model_based_rl.py:1084-1337).

Per iteration j (model_based_rl.py:1171-1180):
    algo.start_worker(); paths = algo.obtain_samples(j); samples_data = algo.process_samples(j,
    paths); algo.optimize_policy(j, samples_data)
every `log_every` iterations (:1209-1248) the K per-model validation costs of the deterministic
policy are evaluated on the fixed validation initial states (build_policy_graph, :122-142) -- here
ONE launch of the persistent rollout kernel in per-model mode (metrpo_rollout_model_costs) instead
of a T-times-unrolled TF graph -- and fed to is_done (:1339-1371); the best policy so far plays the
role of the `policy.ckpt` checkpoint and is restored at the end (log_and_restore, :1374-1400).

What is NOT here (out of scope): the real-simulator validation cost
(evaluate_fixed_init_trajectories needs MuJoCo) -- `real_cost_fn` may be supplied by the caller,
default 0.0 -- and the bptt / l-bfgs / svg branches.
"""
import logging
from collections import namedtuple

import numpy as np

from .rollout import EnsembleRollout
from .utils import stop_critereon

Policy_opt_params = namedtuple(
    "Policy_opt_params",
    "mode whole T gamma log_every num_iters_threshold max_iters stop_critereon batch_size")


def policy_opt_params_from_json(d):
    """params/params-<env>.json 'policy_opt_params' (namedtuples.py / training.py:300-318)."""
    sc = d["stop_critereon"]
    return Policy_opt_params(
        mode=d.get("mode", "estimated"), whole=bool(d.get("whole", False)), T=int(d["T"]),
        gamma=float(d["gamma"]), log_every=int(d["log_every"]),
        num_iters_threshold=int(d["num_iters_threshold"]), max_iters=int(d["max_iters"]),
        stop_critereon=stop_critereon(sc["threshold"], sc["offset"],
                                      sc.get("percent_models_threshold", 0.5)),
        batch_size=int(d.get("batch_size", 500)))


def is_done(policy_opt_params, min_validation_costs, candidates, logger=None):
    """model_based_rl.py:1339-1371."""
    mode = policy_opt_params.mode
    if mode == "real":
        return min_validation_costs["real"] < candidates["real"]
    if mode == "trpo_mean":
        assert "trpo_mean" in min_validation_costs.keys()
        return min_validation_costs["trpo_mean"] < candidates["trpo_mean"]
    if mode == "one_model":
        return min_validation_costs["estimated"][0] < candidates["estimated"][0]
    if mode == "no_early":
        return False
    assert "estimated" in mode
    for _mode in min_validation_costs.keys():
        if "estimated" in _mode and policy_opt_params.stop_critereon(
                min_validation_costs[_mode], candidates[_mode], mode="vector"):
            if logger:
                logger.info("\t### %s tells us to stop." % _mode)
            return True
    return False


def update_stats(min_validation_costs, candidates, whole=False):
    """model_based_rl.py:1403-1419."""
    for _mode in min_validation_costs.keys():
        costs = min_validation_costs[_mode]
        if hasattr(costs, "__iter__") and len(costs) != 1:
            if whole:
                min_validation_costs[_mode][:] = candidates[_mode][:]
            else:
                to_update = costs > candidates[_mode]
                min_validation_costs[_mode][to_update] = candidates[_mode][to_update]
        elif whole or costs > candidates[_mode]:
            min_validation_costs[_mode] = candidates[_mode]


class PolicyCostEvaluator:
    """`policy_costs[scope]` of the reference: K per-model discounted costs of the deterministic
    policy from fixed initial states (model_based_rl.py:122-142, evaluated at :1237-1248)."""

    def __init__(self, env, policy, n_rows, T, gamma):
        self.env, self.policy, self.T, self.gamma = env, policy, int(T), float(gamma)
        self.rollout = EnsembleRollout(env.env_name, env.n_models, int(n_rows), self.T,
                                       hidden=env.hidden, policy_hidden=policy.hidden_sizes,
                                       sam_mode=env.sam_mode, policy_out_tanh=policy.output_tanh,
                                       device=env.device)
        self.refresh_models()

    def refresh_models(self):
        self.rollout.set_dynamics_ensemble(self.env.models)
        self.rollout.set_normalization(**self.env.norm)

    def __call__(self, init_states):
        pol = self.policy
        self.rollout.set_policy(pol.W, pol.b, pol.log_std)
        costs = self.rollout.model_costs(self.T, init_states, self.gamma)
        self.rollout.synchronize()
        return costs.cpu().numpy().astype(np.float32)

    def close(self):
        self.rollout.close()


def optimize_policy(algo, policy_opt_params, policy_validation_init, logger=None, real_cost_fn=None,
                    flat=True):
    """TRPO branch of model_based_rl.py:1084-1337.  `algo` is the TRPO object (kwargs['rllab_algo']),
    `policy_validation_init` [n,S] the fixed validation start states (:444-487).  flat=True keeps
    the samples on the device (obtain_samples_flat / process_samples_flat); flat=False walks the
    reference's list-of-paths route.  Returns the reference's result dict (:1329-1337)."""
    logger = logger or logging.getLogger("me_trpo_b200")
    pop = policy_opt_params
    policy, env = algo.policy, algo.env
    evaluator = PolicyCostEvaluator(env, policy, len(policy_validation_init), pop.T, pop.gamma)
    real = (lambda: float(real_cost_fn(policy))) if real_cost_fn is not None else (lambda: 0.0)
    mode_order = ["real", "estimated"]
    trpo_mean_costs, training_costs, real_validation_costs = [], [], []
    estimated_validation_costs = {}

    # iteration 0 (:1143-1168): costs of the incoming policy are the first "best"
    min_validation_costs = {"real": real(), "estimated": evaluator(policy_validation_init)}
    if pop.mode == "trpo_mean":
        min_validation_costs["trpo_mean"] = np.inf
    best_index = 0
    best_params = policy.flat_params().clone()          # plays policy.ckpt (:1287-1289)
    real_current_validation_cost = min_validation_costs["real"]
    candidates = {}
    j = 0
    for j in range(1, pop.max_iters + 1):
        algo.start_worker()                                              # :1175
        if flat:
            samples_data = algo.process_samples_flat(j, algo.obtain_samples_flat(j))
        else:
            samples_data = algo.process_samples(j, algo.obtain_samples(j))   # :1177-1178
        algo.optimize_policy(j, samples_data)                            # :1179
        training_cost = 0
        if j % pop.log_every == 0:                                       # :1209
            if pop.mode == "trpo_mean":                                  # :1218-1227
                determ_paths = algo.obtain_samples(j, determ=True)
                candidates["trpo_mean"] = float(np.mean([-np.sum(p["rewards"]) for p in determ_paths]))
                if "trpo_mean" != mode_order[1]:
                    mode_order.insert(1, "trpo_mean")
            else:
                candidates["trpo_mean"] = 0.0
            trpo_mean_costs.append(candidates["trpo_mean"])
            training_costs.append(training_cost)
            est = evaluator(policy_validation_init)                      # :1237-1248
            estimated_validation_costs.setdefault("estimated", []).append(float(np.mean(est)))
            candidates["estimated"] = est
            candidates["real"] = real()                                  # :1251-1262
            real_validation_costs.append(candidates["real"])
            logger.info("iter %d" % j)
            if not is_done(pop, min_validation_costs, candidates, logger):   # :1283-1293
                best_index = j
                real_current_validation_cost = candidates["real"]
                best_params = policy.flat_params().clone()
                update_stats(min_validation_costs, candidates, pop.whole)
            if j - best_index >= pop.num_iters_threshold:                # :1296-1298
                break
    logger.info("Stop at iter %d. Recover to iter %d." % (j, best_index))    # log_and_restore
    policy.set_flat_params(best_params)
    evaluator.close()
    if pop.mode in ("one_model", "no_early"):
        min_val_cost = min_validation_costs["estimated"][0]
    else:
        min_val_cost = np.mean(min_validation_costs[pop.mode])
    return {"real_validation_costs": real_validation_costs, "training_costs": training_costs,
            "estimated_validation_costs": estimated_validation_costs, "best_index": best_index,
            "best_cost": min_val_cost, "trpo_mean_costs": trpo_mean_costs,
            "real_current_validation_cost": real_current_validation_cost,
            "min_validation_costs": min_validation_costs}
