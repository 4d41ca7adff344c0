"""ORACLE (test infrastructure, not product code): analytic cost / done functions and the
dimensions of the reference's imaginary environments.

Restates, per env, `cost_np_vec` (per-row cost; reward = -cost, env_helpers.py:601) and `is_done`
from the reference's env classes.  PINNED: tests/test_ref_fixtures.py compares every function
here with the reference's own env classes (envs/com_*_env.py executed under shims) on inputs that
exercise the clip / penalty / NaN branches.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

Reference lines:
  swimmer      envs/com_swimmer_env.py:112-114        ctrl_cost_coeff 1e-2 (:43)
  half-cheetah envs/com_half_cheetah_env.py:72-75     ctrl_cost_coeff 1e-1 (:21)
  hopper       envs/com_hopper_env.py:94-104          ctrl_cost_coeff 1e-2 (:30)
  ant          envs/com_ant_env.py:77-83, is_done 88-101
  humanoid     envs/com_simple_humanoid_env.py:105-109 ctrl_cost_coeff 1e-3 (:27)
  snake        envs/com_snake_env.py:81-84            idx 7 (:12), ctrl_cost_coeff 1e-2 (:21)
State/action dims: SURVEY.md section 6 (derived from the obs builders + MuJoCo XML).
"""
import numpy as np

# name -> (S, A, drop_cols, dynamics hidden, policy hidden)   (params/params-<env>.json)
ENV_SPECS = {
    "swimmer": dict(env_id=0, S=10, A=2, drop=2, hidden=512, policy_hidden=(32, 32)),
    "half-cheetah": dict(env_id=1, S=18, A=6, drop=1, hidden=1024, policy_hidden=(32, 32)),
    "hopper": dict(env_id=2, S=11, A=3, drop=0, hidden=1024, policy_hidden=(32, 32)),
    "ant": dict(env_id=3, S=29, A=8, drop=2, hidden=1024, policy_hidden=(32, 32)),
    "humanoid": dict(env_id=4, S=55, A=21, drop=0, hidden=1024, policy_hidden=(100, 50, 25)),
    "snake": dict(env_id=5, S=14, A=4, drop=2, hidden=1024, policy_hidden=(32, 32)),
}
ENV_SPECS["half_cheetah"] = ENV_SPECS["half-cheetah"]  # get_env spelling (env_helpers.py:18)
ENV_NAMES = ["swimmer", "half-cheetah", "hopper", "ant", "humanoid", "snake"]


def env_name(env):
    if isinstance(env, (int, np.integer)):
        return ENV_NAMES[int(env)]
    return "half-cheetah" if env == "half_cheetah" else env


def cost_np_vec(env, x, u, x_next):
    """Per-row cost [B]; `u` must already be clipped to [-1, 1] (the reference asserts it)."""
    env = env_name(env)
    assert u.size == 0 or np.amax(np.abs(u)) <= 1.0
    if env == "swimmer":
        return -(x_next[:, 5] - 1e-2 * np.mean(np.square(u), axis=1))
    if env == "half-cheetah":
        return -np.clip(x_next[:, 9] - 1e-1 * 0.5 * np.sum(np.square(u), axis=1), -10, 10)
    if env == "hopper":
        vel, height, ang = x_next[:, 5], x_next[:, 0], x_next[:, 1]
        return -(vel - 1e-2 * 0.5 * np.sum(np.square(u), axis=1)
                 - 10 * np.maximum(0.45 - height, 0)
                 - 10 * np.maximum(np.abs(ang) - .2, 0)
                 - np.sum(np.maximum(np.abs(x_next[:, 2:]) - 100, 0), axis=1))
    if env == "ant":
        return -(x_next[:, 15] - 1e-2 * 0.5 * np.sum(np.square(u), axis=1) + 0.05)
    if env == "humanoid":
        return (x_next[:, -1] - 1.5) ** 2 + 1e-2 * 1e-3 * np.sum(np.square(u), axis=1)
    if env == "snake":
        return -(x_next[:, 7] - 1e-2 * 0.5 * np.sum(np.square(u), axis=1))
    raise ValueError(env)


def is_done(env, x, x_next):
    """bool [B].  Only Ant defines is_done; the others use NeuralNetEnv's all-False default
    (env_helpers.py:537)."""
    env = env_name(env)
    if env == "ant":
        notdone = np.logical_and(np.logical_and(x_next[:, 2] >= 0.2, x_next[:, 2] <= 1.0),
                                 np.amin(np.isfinite(x_next), axis=1))
        return np.invert(notdone)
    return np.zeros(len(x_next), dtype=bool)
