"""train(variant): the reference's experiment body (training.py:17-403) for algo 'trpo' on the
B200-native components: builds the real-env adapter, the K-model ensemble (EnsembleFit), the
imaginary environment (NeuralNetEnv), the Gaussian MLP policy, LinearFeatureBaseline and TRPO with
the fused sampler, then runs train_models."""
import json
import logging
import os

import numpy as np
import torch

from .algos import TRPO
from .baselines import LinearFeatureBaseline
from .dynamics import EnsembleFit
from .env_helpers import NeuralNetEnv
from .envs import ENV_SPECS, canonical_env_name, drop_cols_from_params
from .model_based_rl import train_models
from .policies import GaussianMLPPolicy
from .real_env import make_real_env
from .synthetic import default_norm, init_dynamics


def set_global_seeds(seed):                     # utils.py:34-37
    np.random.seed(seed)
    torch.manual_seed(seed)


def train(variant, snapshot_dir=None, sampler_n_envs=None, sweep_iters=None, device="cuda", dist_ctx=None):
    """`dist_ctx` (parallel.DistContext; default: from the torchrun environment): with G > 1 ranks
    the imaginary rollouts are row-sharded, the TRPO accumulators all-reduced, the K models fitted
    k = rank (mod G) per rank and re-broadcast, and only rank 0 touches the real environment and
    writes the snapshot directory."""
    from .parallel import DistContext
    if dist_ctx is None and int(os.environ.get("WORLD_SIZE", "1")) > 1 and device == "cuda":
        device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))           # one process per GPU
        torch.cuda.set_device(device)
    ctx = dist_ctx if dist_ctx is not None else DistContext.from_env(device)
    if ctx.distributed and not ctx.rank == 0:
        snapshot_dir = None
    params = variant["params"]
    seed = int(variant.get("seed", 0))
    set_global_seeds(seed)                                                     # training.py:18
    if params["algo"] != "trpo":
        raise NotImplementedError("algo %r: only the TRPO path is built (SURVEY.md section 8)" % params["algo"])
    logger = logging.getLogger("me_trpo_b200")
    name = canonical_env_name(params["env"])
    spec = ENV_SPECS[name]
    S, A = spec["S"], spec["A"]
    if snapshot_dir:
        os.makedirs(snapshot_dir, exist_ok=True)
        with open(os.path.join(snapshot_dir, "params.json"), "w") as f:       # training.py:43-45
            json.dump(params, f, indent=1)
    real_env = make_real_env(name, seed)
    dm = params["dynamics_model"]
    hidden = dm["hidden_layers"]
    assert len(hidden) == 2 and hidden[0] == hidden[1], "the kernels cover the shipped 2 x H dynamics MLPs"
    assert dm.get("nonlinearity", ["tf.nn.relu"] * 2) == ["tf.nn.relu", "tf.nn.relu"]
    drop = drop_cols_from_params(dm)
    K = int(params["n_models"])
    pop = params["policy_opt_params"]
    rng = np.random.RandomState(seed)
    models = init_dynamics(rng, S, A, drop, hidden[0], K, out_scale=1.0)       # identical on every rank
    mine = ctx.my_models(K)                                                     # models this rank fits
    if not mine:
        raise RuntimeError("more ranks (%d) than dynamics models (%d): nothing to fit on rank %d"
                           % (ctx.world_size, K, ctx.rank))
    fit = EnsembleFit(S, A, drop, hidden[0], len(mine), max_rows=max(4096, params["dynamics_opt_params"]["batch_size"]),
                      device=device)
    fit.set_ensemble([models[k] for k in mine])
    nn_env = NeuralNetEnv(name, models, default_norm(S, A), sam_mode=pop.get("sam_mode", "step_rand"),
                          reset_sampler=lambda n: np.asarray([real_env.reset() for _ in range(n)], np.float32),
                          hidden=hidden[0], device=device, policy_hidden=tuple(params["policy"]["hidden_layers"]),
                          precision=dm.get("rollout_precision", "bf16"))      # extra key: "fp32" = reference arithmetic
    policy = GaussianMLPPolicy(S, A, tuple(params["policy"]["hidden_layers"]), init_std=pop["trpo"]["init_std"],
                               output_tanh=(params["policy"].get("output_nonlinearity") == "tf.tanh"),
                               device=device, seed=seed)
    algo = TRPO(env=nn_env, policy=policy, baseline=LinearFeatureBaseline(env_spec=nn_env.spec),
                batch_size=pop["trpo"]["batch_size"], max_path_length=pop["T"], discount=pop["trpo"]["discount"],
                step_size=pop["trpo"]["step_size"],
                sampler_args=dict(n_envs=sampler_n_envs, seed=seed), dist_ctx=ctx)   # training.py:355-366
    ctx.broadcast_policy(policy)
    rows = train_models(real_env, nn_env, algo, fit, params, snapshot_dir=snapshot_dir, seed=seed, logger=logger,
                        sweep_iters=sweep_iters, dist_ctx=ctx)
    algo.shutdown_worker()
    fit.close()
    return dict(progress=rows, policy=policy)
