"""Synthetic problem generator for benchmarks and demos (MuJoCo / trained checkpoints are not
available offline): random-init networks of the reference's architecture and initialisers.

  dynamics  tf.contrib.layers.xavier_initializer() on W *and* b (training.py:179,187-194), with the
            output layer scaled by `out_scale` (0.1) so a 1000-step open-loop rollout of a random
            net stays finite (SURVEY.md 8d records the factor)
  policy    rllab GaussianMLPPolicy init: Xavier-uniform W, zero b, log_std = log(init_std)
  norm      mu_in = 0, sigma_in = 1, mu_delta = 0, sigma_delta = 0.1 (the RunningMeanStd floor,
            running_mean_std.py:25)

The draw order from the RandomState is part of the contract: tests regenerate the same problem
on the oracle side from the same seed.
"""
import numpy as np

from .envs import ENV_SPECS, canonical_env_name


def xavier_uniform(rng, shape):
    fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_dynamics(rng, S, A, drop, hidden, K, out_scale=0.1):
    din = S + A - drop
    models = []
    for _ in range(K):
        models.append(dict(
            W0=xavier_uniform(rng, (din, hidden)), b0=xavier_uniform(rng, (hidden,)),
            W1=xavier_uniform(rng, (hidden, hidden)), b1=xavier_uniform(rng, (hidden,)),
            W2=xavier_uniform(rng, (hidden, S)) * np.float32(out_scale),
            b2=xavier_uniform(rng, (S,)) * np.float32(out_scale)))
    return models


def init_policy(rng, S, hidden, A, init_std=1.0):
    dims = [S] + list(hidden) + [A]
    W = [xavier_uniform(rng, (dims[i], dims[i + 1])) for i in range(len(dims) - 1)]
    b = [np.zeros(dims[i + 1], np.float32) for i in range(len(dims) - 1)]
    return dict(W=W, b=b, log_std=np.full(A, np.log(init_std), np.float32))


def default_norm(S, A):
    return dict(in_mean=np.zeros(S + A, np.float32), in_std=np.ones(S + A, np.float32),
                diff_mean=np.zeros(S, np.float32), diff_std=np.full(S, 0.1, np.float32))


def make_problem(env, n_models, n_rows, hidden=None, seed=0, pool_rows=None):
    """(spec, models, policy, norm, init_states[B,S], reset_pool[R,S]) for `env`."""
    name = canonical_env_name(env)
    spec = ENV_SPECS[name]
    hidden = int(hidden or spec["hidden"])
    rng = np.random.RandomState(seed)
    models = init_dynamics(rng, spec["S"], spec["A"], spec["drop"], hidden, n_models)
    pol = init_policy(rng, spec["S"], spec["policy_hidden"], spec["A"])
    norm = default_norm(spec["S"], spec["A"])
    init = rng.normal(0, 0.1, (n_rows, spec["S"])).astype(np.float32)
    pool = rng.normal(0, 0.1, (pool_rows or n_rows, spec["S"])).astype(np.float32)
    return spec, models, pol, norm, init, pool
