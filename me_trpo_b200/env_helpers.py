"""NeuralNetEnv / VecSimpleEnv: drop-in mirrors of the reference's imaginary environment
(env_helpers.py:532-635) whose step runs on the B200 through libmetrpo.so.

Differences forced by the environment, all explicit:
  * the TF dynamics graph (`dynamics_in`, `dynamics_outs`) is replaced by the ensemble weights
    (`models`: list of dicts W0,b0,W1,b1,W2,b2 in TF layout) + RunningMeanStd constants (`norm`);
  * the real simulator's reset() (env_helpers.py:552-555,592) is replaced by `reset_sampler(n)`,
    a callable returning n start states (MuJoCo is not part of this repo);
  * `cost_np` / `is_done` are fused in the kernel and selected by the env name.
"""
import numpy as np
import torch

from .envs import ENV_SPECS, canonical_env_name
from .rollout import EnsembleRollout
from .spaces import Box, EnvSpec


class NeuralNetEnv:
    """rllab-Env-like object; `vectorized = True` makes VectorizedSampler ask it for a
    vec_env_executor (samplers/vectorized_sampler.py:29-33)."""

    def __init__(self, env, models, norm, sam_mode="step_rand", reset_sampler=None, hidden=None,
                 device=None, policy_hidden=None, precision="bf16"):
        self.vectorized = True
        self.env_name = canonical_env_name(env)
        spec = ENV_SPECS[self.env_name]
        self.S, self.A = spec["S"], spec["A"]
        self.models, self.norm = list(models), dict(norm)
        self.n_models = len(self.models)
        self.sam_mode = sam_mode
        self.hidden = int(hidden or self.models[0]["W1"].shape[0])
        self.policy_hidden = policy_hidden
        self.device = device
        self.precision = precision     # "bf16" (tensor cores) | "fp32" (reference arithmetic, slow)
        self.reset_sampler = reset_sampler or (
            lambda n: np.random.normal(0.0, 0.1, size=(n, self.S)).astype(np.float32))
        self._single = None
        self._state = None

    @property
    def observation_space(self):
        return Box(-np.inf, np.inf, (self.S,))

    @property
    def action_space(self):
        return Box(-1.0, 1.0, (self.A,))   # rllab normalize(env): bounds are [-1, 1]

    @property
    def spec(self):
        return EnvSpec(self.observation_space, self.action_space)

    def vec_env_executor(self, n_envs, max_path_length):
        return VecSimpleEnv(env=self, n_envs=n_envs, max_path_length=max_path_length)

    # single-env protocol (env_helpers.py:552-566) through a 1-row executor
    def reset(self):
        if self._single is None:
            self._single = VecSimpleEnv(self, 1, 1 << 30)
        self._state = self._single.reset()[0]
        return np.copy(self._state)

    def step(self, action):
        obs, rew, done, _ = self._single.step(np.asarray(action, np.float32)[None])
        self._state = obs[0]
        return self._state, float(rew[0]), bool(done[0]), {}

    def terminate(self):
        pass


class VecSimpleEnv:
    """env_helpers.py:575-635 with the state resident on the GPU."""

    def __init__(self, env, n_envs, max_path_length, rng=None):
        self.env = env
        self.n_envs = self.num_envs = int(n_envs)
        self.max_path_length = int(max_path_length)
        self.rng = np.random if rng is None else rng
        self.rollout = EnsembleRollout(env.env_name, env.n_models, self.n_envs, self.max_path_length,
                                       hidden=env.hidden, sam_mode=env.sam_mode, device=env.device,
                                       policy_hidden=env.policy_hidden,
                                       precision=getattr(env, "precision", "bf16"))
        self.rollout.set_dynamics_ensemble(env.models)
        self.rollout.set_normalization(**env.norm)
        self.device = self.rollout.device
        self.cur_model_idx = self.rng.randint(env.n_models, size=(self.n_envs,))   # :583
        self._obs = torch.zeros(self.n_envs, env.S, device=self.device)           # :580
        self._needs_reset = True

    def reset(self, dones=None):
        """No argument: reset every row (samplers/vectorized_sampler.py:49)."""
        if dones is None:
            dones = np.ones(self.n_envs, dtype=bool)
        dones = np.asarray(dones, dtype=bool)
        n = int(dones.sum())
        if n:
            rows = np.nonzero(dones)[0]
            fresh_np = np.empty((n, self.env.S), np.float32)
            for j, i in enumerate(rows):      # row order, simulator reset then model draw (:590-593)
                fresh_np[j] = np.asarray(self.env.reset_sampler(1), np.float32)[0]
                self.cur_model_idx[i] = self.rng.randint(self.env.n_models)
            fresh = torch.as_tensor(fresh_np, device=self.device)
            self._obs[torch.as_tensor(dones, device=self.device)] = fresh
            if dones.all():
                self.rollout.reset(self._obs)      # ts = 0 for every row
                self._needs_reset = False
            else:
                raise NotImplementedError(
                    "partial reset() from the host: done rows are already replaced inside step()")
        return self._obs[torch.as_tensor(dones, device=self.device)].cpu().numpy()

    def step(self, actions):
        """(obs, rewards, dones, env_infos) with obs the post-reset states (:605-607)."""
        if self._needs_reset:
            raise RuntimeError("VecSimpleEnv.step() before reset()")
        B, K = self.n_envs, self.env.n_models
        mode = self.env.sam_mode
        model_idx = None
        if mode == "step_rand":
            model_idx = self.rng.randint(K, size=B)                    # :619
        elif mode == "eps_rand":
            model_idx = self.cur_model_idx                              # :621-622
        std_noise = self.rng.normal(size=(B, self.env.S)).astype(np.float32) if mode == "model_mean_std" else None
        # the kernel needs a reset state per row that may finish; the REAL resets are drawn below
        # for the done rows only, in row order, like the reference's reset(dones) loop (:590-593)
        obs, rew, done = self.rollout.step(actions, self._obs, model_idx=model_idx, std_noise=std_noise)
        self.rollout.synchronize()
        done_np = done.cpu().numpy().astype(bool)
        if done_np.any():
            rows = np.nonzero(done_np)[0]
            fresh = np.empty((len(rows), self.env.S), np.float32)
            for j, i in enumerate(rows):                                # :590-593
                fresh[j] = np.asarray(self.env.reset_sampler(1), np.float32)[0]
                self.cur_model_idx[i] = self.rng.randint(K)
            self.rollout.set_rows(rows, fresh)
            obs[torch.as_tensor(rows, device=obs.device, dtype=torch.long)] = torch.as_tensor(fresh, device=obs.device)
        self._obs = obs
        return obs.cpu().numpy(), rew.cpu().numpy(), done_np, dict()

    def terminate(self):
        self.rollout.close()
