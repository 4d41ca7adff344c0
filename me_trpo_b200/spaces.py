"""Minimal observation / action space objects (the reference gets these from rllab; only
`shape`, `flat_dim`, `bounds` and `flatten_n` are touched on the rollout path:
env_helpers.py:579,599 and samplers/vectorized_sampler.py:94-95)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape):
        self.low = np.full(shape, low, dtype=np.float32)
        self.high = np.full(shape, high, dtype=np.float32)
        self.shape = tuple(shape)

    @property
    def flat_dim(self):
        return int(np.prod(self.shape))

    @property
    def bounds(self):
        return self.low, self.high

    def flatten_n(self, xs):
        return np.asarray(xs).reshape(len(xs), -1)


class EnvSpec:
    def __init__(self, observation_space, action_space):
        self.observation_space = observation_space
        self.action_space = action_space
