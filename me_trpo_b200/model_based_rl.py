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
                                       device=env.device, precision=getattr(env, "precision", "bf16"))
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


# ================================================================================================
# Outer loop: collect -> fit the ensemble -> improve the policy  (model_based_rl.py:231-755)
# ================================================================================================
def prepare_policy(W, b, param_noise, diff_weights, initial_param_std, rng):
    """env_helpers.py:50-59: per-episode parameter-space exploration.  Returns perturbed COPIES of
    the mean-network weights -- flat_weight_update = param_noise * diff_weights * randn(n), applied
    to the biases then the weight matrices (the reference's flat order, model_based_rl.py:419-421;
    log_std is not perturbed) -- and mean(|update|).  `diff_weights` maps like
    policy.get_param_values(): (W, b) per layer, then log_std (ignored)."""
    if diff_weights is None:
        assert initial_param_std == 0.0
        return W, b, 0.0
    dW, db, o = [], [], 0
    for w, v in zip(W, b):
        dW.append(np.asarray(diff_weights[o:o + w.size]).reshape(w.shape)); o += w.size
        db.append(np.asarray(diff_weights[o:o + v.size])); o += v.size
    n_vars = sum(v.size for v in b) + sum(w.size for w in W)
    z = rng.randn(n_vars)
    W2, b2, o, total = [], [], 0, 0.0
    for v, d in zip(b, db):
        upd = param_noise * d * z[o:o + v.size]; o += v.size
        b2.append(v + upd); total += np.abs(upd).sum()
    for w, d in zip(W, dW):
        upd = param_noise * d * z[o:o + w.size].reshape(w.shape); o += w.size
        W2.append(w + upd); total += np.abs(upd).sum()
    return W2, b2, total / n_vars


def sample_trajectories(real_env, policy, exploration, batch_size, max_timestep, rng, logger=None,
                        diff_weights=None):
    """env_helpers.py:352-460.  Per episode the policy's mean network is perturbed in parameter
    space (prepare_policy; the reference saves / restores a checkpoint around it, here the episode
    simply runs on a perturbed copy); per step get_action (env_helpers.py:37-48) adds
    action_noise * randn(1) -- ONE scalar draw broadcast over all action dimensions, because the
    reference takes `len(action)` of a [1, A] array -- and clips to the action bounds.
    Returns (Os, As, Rs, info)."""
    import torch
    Os, As, Rs = [], [], []
    counter = 1
    with torch.no_grad():
        W0 = [w.cpu().numpy().astype(np.float64) for w in policy.W]
        b0 = [v.cpu().numpy().astype(np.float64) for v in policy.b]
    n = len(W0)

    def mean_action(o, W, b):
        h = o
        for i in range(n):
            h = h @ W[i] + b[i]
            if i < n - 1 or policy.output_tanh:
                h = np.tanh(h)
        return h

    changes = []
    while counter <= batch_size:
        o, a, r = [real_env.reset()], [], []
        W, b, change = prepare_policy(W0, b0, exploration.get("param_noise", 0.0), diff_weights,
                                      exploration.get("initial_param_std", 0.0), rng)
        changes.append(change)
        for t in range(max_timestep):
            noise = exploration["action_noise"] * (rng.uniform() if exploration.get("vary_trajectory_noise") else 1.0)
            act = np.clip(mean_action(o[-1], W, b) + noise * rng.randn(1), -1.0, 1.0)   # get_action
            obs, rew, done, _ = real_env.step(act)
            o.append(obs); a.append(act); r.append(rew)
            counter += 1
            if done:
                break
        Os.append(o); As.append(a); Rs.append(r)
    info = dict(EpisodesCollected=len(Os), TimeStepsCollected=counter - 1,
                avg_eps_reward=float(np.mean([np.sum(x) for x in Rs])),
                avg_weight_change=float(np.mean(changes)))
    return Os, As, Rs, info


def collect_data(real_env, policy, sample_size, dynamics_data, dynamics_validation, input_rms, output_rms,
                 rollout_params, rng, logger=None, diff_weights=None, dist_ctx=None):
    """model_based_rl.py:758-857 (use_same_dataset / trajectory split).  With several ranks only
    rank 0 steps the real environment; the new transitions are broadcast."""
    from .dynamics import add_rollout_data
    if sample_size == 0:
        return {}
    x_all = y_all = None
    info = {}
    if dist_ctx is None or dist_ctx.rank == 0:
        Os, As, Rs, info = sample_trajectories(real_env, policy, rollout_params["exploration"], sample_size,
                                               rollout_params["max_timestep"], rng, logger, diff_weights)
        x_all, y_all = [], []
        for o, a in zip(Os, As):
            for t in range(len(o) - 1):
                x_all.append(np.concatenate([o[t], a[t]]))
                y_all.append(o[t + 1])
        x_all, y_all = np.asarray(x_all, np.float32), np.asarray(y_all, np.float32)
    if dist_ctx is not None and dist_ctx.distributed:
        stats = np.asarray([info.get("EpisodesCollected", 0), info.get("TimeStepsCollected", 0),
                            info.get("avg_eps_reward", 0.0), info.get("avg_weight_change", 0.0)], np.float64) \
            if dist_ctx.rank == 0 else None
        x_all, y_all, stats = dist_ctx.broadcast_arrays([x_all, y_all, stats], dynamics_data.device)
        info = dict(EpisodesCollected=int(stats[0]), TimeStepsCollected=int(stats[1]),
                    avg_eps_reward=float(stats[2]), avg_weight_change=float(stats[3]))
    assert len(x_all) >= sample_size
    add_rollout_data(x_all, y_all, dynamics_data, dynamics_validation, input_rms, output_rms,
                     rollout_params["split_ratio"])
    return info


def train_models(real_env, nn_env, algo, fit, params, snapshot_dir=None, seed=0, logger=None,
                 sweep_iters=None, policy_validation_init=None, dist_ctx=None):
    """The sweep loop of train_models (model_based_rl.py:546-755) for algo 'trpo': every sweep
    collects `sample_size` real transitions, refits the K dynamics models on the device
    (optimize_models), pushes the new weights + normalisers into the imaginary env, resets
    log_std (training.py:368-370) and runs optimize_policy.  One progress.csv row per sweep with
    the reference's column names."""
    import csv
    import os
    import time
    import torch
    from .dynamics import RunningMeanStd, data_collection, optimize_models
    from .parallel import DistContext
    ctx = dist_ctx if dist_ctx is not None else DistContext()
    K_total = int(params["n_models"])
    logger = logger or logging.getLogger("me_trpo_b200")
    rng = np.random.RandomState(seed)
    rp, dop, pop_json = params["rollout_params"], params["dynamics_opt_params"], params["policy_opt_params"]
    pop = policy_opt_params_from_json(pop_json)
    policy = algo.policy
    dev = fit.device
    dynamics_data = data_collection(rp["training_data_size"], dev)
    dynamics_validation = data_collection(rp["validation_data_size"], dev)
    input_rms = RunningMeanStd(shape=(fit.S + fit.A,), device=dev)
    diff_rms = RunningMeanStd(shape=(fit.S,), device=dev)
    if policy_validation_init is None:                                      # :444-487
        policy_validation_init = (np.asarray([real_env.reset() for _ in range(pop.batch_size)], np.float32)
                                  if ctx.rank == 0 else None)
        (policy_validation_init,) = ctx.broadcast_arrays([policy_validation_init], dev)
    sweep_iters = int(sweep_iters or params["sweep_iters"])
    rows, start_time = [], time.time()
    diff_weights = None
    for count in range(1, sweep_iters + 1):
        t0 = time.time()
        reinit_every = int(dop["reinitialize"])
        reinitialize = (count == 1) or not (reinit_every <= 0 or count % reinit_every != 1)   # :550-556
        info = collect_data(real_env, policy, params["sample_size"], dynamics_data, dynamics_validation,
                            input_rms, diff_rms, rp, rng, logger, diff_weights, dist_ctx=ctx)   # :573-588
        t1 = time.time()
        norm = dict(in_mean=input_rms.mean, in_std=input_rms.std, diff_mean=diff_rms.mean, diff_std=diff_rms.std)
        fit.set_normalization(**norm)
        dlog = optimize_models(fit, dynamics_data, dynamics_validation, batch_size=min(dop["batch_size"], fit.max_rows),
                               learning_rate=dop["learning_rate"], log_every=dop["log_every"],
                               num_passes_threshold=dop["num_passes_threshold"], max_passes=dop["max_passes"],
                               reinitialize=reinitialize, rng=None, seed=seed + count + 7919 * ctx.rank, logger=None)
        torch.cuda.synchronize()
        t2 = time.time()
        # new weights -> imaginary env (with G ranks: every owner broadcasts the models it fitted)
        nn_env.models = ctx.gather_models(fit.get_ensemble(), K_total)
        if ctx.distributed:      # per-model diagnostics of the other ranks' models
            stats = torch.tensor([float(dlog["n_updates"]), float(dlog["min_sum_validation_loss"])],
                                 dtype=torch.float64, device=dev)
            import torch.distributed as dist
            mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=ctx.group)
            dist.all_reduce(stats, group=ctx.group)
            dlog["n_updates"], dlog["min_sum_validation_loss"] = int(mx[0].item()), float(stats[1].item())
        nn_env.norm = {k: v.clone() for k, v in norm.items()}
        if pop_json["trpo"].get("reset", False):                             # training.py:368-370
            policy.log_std.fill_(float(np.log(pop_json["trpo"]["init_std"])))
        old_w = policy.get_param_values()
        plog = optimize_policy(algo, pop, policy_validation_init, logger=logger)
        new_w = policy.get_param_values()
        torch.cuda.synchronize()
        t3 = time.time()
        if (np.abs(new_w - old_w) > 0).any():
            diff_weights = np.abs(new_w - old_w)
        row = {"collect_data_time": t1 - t0, "model_opt_time": t2 - t1, "policy_opt_time": t3 - t2,
               "Time": time.time() - start_time, "ItrTime": time.time() - t0,
               "MaxPolicyWeightDiff": float(np.amax(diff_weights)) if diff_weights is not None else 0,
               "MinPolicyWeightDiff": float(np.amin(diff_weights)) if diff_weights is not None else 0,
               "AvgPolicyWeightDiff": float(np.mean(diff_weights)) if diff_weights is not None else 0,
               "EpisodesCollected": info.get("EpisodesCollected", 0),
               "TimeStepsCollected": info.get("TimeStepsCollected", 0),
               "# model updates": dlog["n_updates"],
               "training_dynamics_min_sum_validation_loss": dlog["min_sum_validation_loss"],
               "estimated_policy_mean_min_validation_cost": float(np.mean(plog["min_validation_costs"]["estimated"])),
               "real_policy_mean_min_validation_cost": float(np.mean(plog["min_validation_costs"]["real"])),
               "real_current_validation_cost": plog["real_current_validation_cost"],
               "# policy updates": plog["best_index"], "avg_eps_reward": info.get("avg_eps_reward", 0.0)}
        rows.append(row)
        logger.info("sweep %d: %s" % (count, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}))
        if snapshot_dir:
            os.makedirs(snapshot_dir, exist_ok=True)
            with open(os.path.join(snapshot_dir, "progress.csv"), "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
                w.writeheader()
                w.writerows(rows)
    return rows
