"""cost_np_vec / is_done of the reference's env classes (envs/com_*_env.py) for HOST-side use
(data collection and real-env validation; the imaginary rollout fuses the same formulas into the
CUDA kernel, csrc/rollout_kernel.cuh env_cost / env_is_done).

  swimmer com_swimmer_env.py:112-114 | half-cheetah com_half_cheetah_env.py:72-75 | hopper
  com_hopper_env.py:94-104 | ant com_ant_env.py:77-101 | humanoid com_simple_humanoid_env.py:105-109
  | snake com_snake_env.py:81-84"""
import numpy as np

from .envs import canonical_env_name


def cost_np_vec(env, x, u, x_next):
    env = canonical_env_name(env)
    assert u.size == 0 or np.amax(np.abs(u)) <= 1.0
    su2 = np.sum(np.square(u), axis=1)
    if env == "swimmer":
        return -(x_next[:, 5] - 1e-2 * np.mean(np.square(u), axis=1))
    if env == "half-cheetah":
        return -np.clip(x_next[:, 9] - 1e-1 * 0.5 * su2, -10, 10)
    if env == "hopper":
        return -(x_next[:, 5] - 1e-2 * 0.5 * su2 - 10 * np.maximum(0.45 - x_next[:, 0], 0)
                 - 10 * np.maximum(np.abs(x_next[:, 1]) - .2, 0)
                 - np.sum(np.maximum(np.abs(x_next[:, 2:]) - 100, 0), axis=1))
    if env == "ant":
        return -(x_next[:, 15] - 1e-2 * 0.5 * su2 + 0.05)
    if env == "humanoid":
        return (x_next[:, -1] - 1.5) ** 2 + 1e-2 * 1e-3 * su2
    if env == "snake":
        return -(x_next[:, 7] - 1e-2 * 0.5 * su2)
    raise ValueError(env)


def is_done(env, x, x_next):
    if canonical_env_name(env) == "ant":
        notdone = np.logical_and(np.logical_and(x_next[:, 2] >= 0.2, x_next[:, 2] <= 1.0),
                                 np.amin(np.isfinite(x_next), axis=1))
        return np.invert(notdone)
    return np.zeros(len(x_next), dtype=bool)
