"""Real-environment adapter (the part of the reference that needs MuJoCo, out of the hot path).

ME-TRPO touches the real simulator in three places: collecting transitions for the dynamics fit
(sample_trajectories, env_helpers.py:352-460), supplying reset states to the imaginary vec-env
(env_helpers.py:552-555,590-593) and the real validation cost (evaluate_fixed_init_trajectories).
All three go through this interface:

    class RealEnv:                      # what get_env(...) returns in the reference, normalised
        S, A                            # observation / action dims
        reset() -> obs[S]
        step(action[A] in [-1,1]) -> (obs[S], reward, done, info)

MuJoCo is not part of this repository; `make_real_env(name)` returns a user-registered
environment (REGISTRY) or the synthetic stand-in below, so that the whole loop runs offline."""
import numpy as np

from .envs import ENV_SPECS, canonical_env_name

REGISTRY = {}   # name -> callable() -> RealEnv ; register MuJoCo-backed envs here


def register(name, factory):
    REGISTRY[canonical_env_name(name)] = factory


class SyntheticEnv:
    """Stand-in simulator with the env's dimensions: smooth random nonlinear dynamics
    x' = x + dt * tanh([x,u] W1) W2 (fixed seed), the env's analytic reward and is_done, reset
    states ~ N(0, 0.1^2) (MujocoEnv.reset draws init_qpos/qvel + noise).  It exists so
    that collect -> fit -> optimise runs end to end without MuJoCo; it is NOT a physics model."""

    def __init__(self, name, seed=0, dt=0.05, hidden=64):
        from . import env_costs
        self.name = canonical_env_name(name)
        spec = ENV_SPECS[self.name]
        self.S, self.A = spec["S"], spec["A"]
        rng = np.random.RandomState(1000 + seed)
        self._W1 = rng.normal(0, 1.0 / np.sqrt(self.S + self.A), (self.S + self.A, hidden))
        self._W2 = rng.normal(0, 1.0 / np.sqrt(hidden), (hidden, self.S))
        self._dt = dt
        self._rng = np.random.RandomState(seed)
        self._cost, self._done = env_costs.cost_np_vec, env_costs.is_done
        self._x = None

    def reset(self):
        self._x = self._rng.normal(0, 0.1, self.S)
        if self.name == "ant":
            self._x[2] = 0.6          # inside the healthy band of is_done
        return self._x.copy()

    def step(self, action):
        u = np.clip(np.asarray(action, np.float64).reshape(-1), -1.0, 1.0)
        x = self._x
        xn = x + self._dt * np.tanh(np.concatenate([x, u]) @ self._W1) @ self._W2
        r = -float(self._cost(self.name, x[None], u[None], xn[None])[0])
        d = bool(self._done(self.name, x[None], xn[None])[0])
        self._x = xn
        return xn.copy(), r, d, {}


def make_real_env(name, seed=0):
    name = canonical_env_name(name)
    if name in REGISTRY:
        return REGISTRY[name]()
    return SyntheticEnv(name, seed)
