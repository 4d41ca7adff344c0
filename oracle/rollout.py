"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's imaginary
rollout -- VecSimpleEnv (env_helpers.py:575-635) driven by VectorizedSampler.obtain_samples
(samplers/vectorized_sampler.py:45-116) -- with every source of randomness made an explicit input.

PARITY PINNED (round 2): tests/test_ref_fixtures.py compares this module with
tests/golden/ref_fixtures.npz, outputs of the REFERENCE'S OWN VecSimpleEnv / NeuralNetEnv /
VectorizedSampler / build_policy_graph code executed in the build container under TF / rllab shims
(tests/golden/make_ref_fixtures.py): all six sam_modes, timeouts, Ant's early termination with
row-ordered reset consumption, the f64/f32 dtype drift.  What stays restated-from-memory is the
part of rllab the reference does not vendor (GaussianMLPPolicy.get_actions).

Randomness: the reference draws policy noise, per-step model indices and
real-simulator reset states from one interleaved NumPy MT19937 stream; a device kernel cannot
reproduce that, so here -- as in the C ABI (include/metrpo.h metrpo_rollout_run) -- they are inputs:
  eps[t, b, :]        N(0,1) policy noise           (rllab get_actions: rnd * exp(log_std) + mean)
  model_idx[t, b]     step_rand / eps_rand index    (env_helpers.py:619 / :583,593)
  std_noise[t, b, :]  model_mean_std noise          (env_helpers.py:626)
  reset_pool[r, :]    pre-sampled real-env reset states (env_helpers.py:592 calls MuJoCo reset)
or they are all derived from the counter-based Philox4x32-10 generator below, which the CUDA kernel
implements bit-identically (integer part) in csrc/philox.cuh.

Reset-pool consumption.  The reference overwrites done rows in row order with fresh simulator
resets (env_helpers.py:590-593).  `reset_mode="per_row"` (what the device kernel does) gives the
n-th reset of row i the pool entry (n*B + i) % R; `reset_mode="ordered"` consumes the pool
sequentially over done rows in row order exactly like the reference loop.  The two coincide
whenever all rows finish together (every env but Ant: only the timeout sets done).
"""
import numpy as np

from . import envs as _envs
from . import models as _models

# ------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,c2,c3), key = (k0,k1)
# ------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
STREAM_EPS = 0          # + action block index (4 normals per block)
STREAM_IDX = 0x10000    # step_rand model index
STREAM_EIDX = 0x10001   # eps_rand model index (counter c0 = episode number of the row)
STREAM_STD = 0x20000    # + state block index
PHILOX_C3 = 0x4D455452  # 'METR'


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def _u01_open_closed(x):   # (0, 1]
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)


def _u01_closed_open(x):   # [0, 1)
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _box_muller(xa, xb):
    """Two N(0,1) fp32 values from two uint32 (same formula as the kernel; evaluated in float64 and
    rounded, the kernel evaluates logf/sincosf in fp32 -> agreement to ~1e-6)."""
    u1 = _u01_open_closed(xa).astype(np.float64)
    u2 = _u01_closed_open(xb).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def philox_normal(seed, step, rows, n, stream_base):
    """[len(rows), n] N(0,1) fp32 for the 64-bit global step index `step` and global row ids
    (c0 = low word of the step, c3 = 'METR' ^ high word; csrc/philox.cuh)."""
    rows = np.asarray(rows, dtype=np.uint32)
    out = np.empty((len(rows), n), np.float32)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    step = int(step)
    for blk in range((n + 3) // 4):
        x0, x1, x2, x3 = philox4x32(np.uint32(step & 0xFFFFFFFF), rows, np.uint32(stream_base + blk),
                                    np.uint32(PHILOX_C3 ^ ((step >> 32) & 0xFFFFFFFF)), k0, k1)
        n0, n1 = _box_muller(x0, x1)
        n2, n3 = _box_muller(x2, x3)
        for j, v in enumerate((n0, n1, n2, n3)):
            if 4 * blk + j < n:
                out[:, 4 * blk + j] = v
    return out


def philox_index(seed, counter, rows, K, stream):
    """[len(rows)] int32 in [0, K): mulhi(x0, K).  `counter`: 64-bit scalar step or a uint32 array
    of per-row episode numbers."""
    rows = np.asarray(rows, dtype=np.uint32)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    if np.ndim(counter) == 0:
        c = int(counter)
        lo, hi = np.uint32(c & 0xFFFFFFFF), (c >> 32) & 0xFFFFFFFF
    else:
        lo, hi = np.asarray(counter, dtype=np.uint32), 0
    x0, _, _, _ = philox4x32(lo, rows, np.uint32(stream), np.uint32(PHILOX_C3 ^ hi), k0, k1)
    return ((x0.astype(np.uint64) * np.uint64(K)) >> np.uint64(32)).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# noise sources
# ------------------------------------------------------------------------------------------------
class ExplicitNoise:
    def __init__(self, eps=None, model_idx=None, std_noise=None):
        self.eps, self.model_idx, self.std_noise = eps, model_idx, std_noise

    def get_eps(self, t, B, A):
        return np.asarray(self.eps[t], np.float32)

    def get_model_idx(self, t, B, K, episode):
        return np.asarray(self.model_idx[t], np.int64)

    def get_std_noise(self, t, B, S):
        return np.asarray(self.std_noise[t], np.float32)


class PhiloxNoise:
    """Counter-based noise identical to the kernel's eps == NULL / model_idx == NULL path."""

    def __init__(self, seed, offset=0, row0=0, sam_mode="step_rand"):
        self.seed, self.offset, self.row0, self.sam_mode = int(seed), int(offset), int(row0), sam_mode

    def _rows(self, B):
        return np.arange(self.row0, self.row0 + B, dtype=np.uint32)

    def get_eps(self, t, B, A):
        return philox_normal(self.seed, self.offset + t, self._rows(B), A, STREAM_EPS)

    def get_model_idx(self, t, B, K, episode):
        if self.sam_mode == "eps_rand":
            return philox_index(self.seed, episode.astype(np.uint32), self._rows(B), K,
                                STREAM_EIDX).astype(np.int64)
        return philox_index(self.seed, self.offset + t, self._rows(B), K, STREAM_IDX).astype(np.int64)

    def get_std_noise(self, t, B, S):
        return philox_normal(self.seed, self.offset + t, self._rows(B), S, STREAM_STD)


# ------------------------------------------------------------------------------------------------
# VecSimpleEnv restatement
# ------------------------------------------------------------------------------------------------
class VecSimpleEnvOracle:
    """env_helpers.py:575-635.  `models/norm` play the role of the TF dynamics graph
    (dynamics_in -> dynamics_outs), `reset_pool` the real simulator's reset()."""

    def __init__(self, env, models, norm, n_envs, max_path_length, sam_mode, noise, reset_pool,
                 S, A, drop, dtype=np.float32, mma="fp32", reset_mode="per_row"):
        self.env, self.models, self.norm = _envs.env_name(env), models, norm
        self.n_envs = self.num_envs = n_envs
        self.max_path_length = max_path_length
        self.sam_mode, self.noise = sam_mode, noise
        self.reset_pool = np.asarray(reset_pool, np.float32)
        self.S, self.A, self.drop, self.dtype, self.mma = S, A, drop, dtype, mma
        self.reset_mode = reset_mode
        self.states = np.zeros((n_envs, S), dtype)          # :580
        self.ts = np.zeros((n_envs,))                        # :581
        self.n_resets = np.zeros(n_envs, np.int64)           # resets consumed per row (after init)
        self._pool_cursor = 0
        self.t = 0                                           # global step counter (noise index)

    def set_states(self, states):
        """The initial vec_env.reset() (samplers/vectorized_sampler.py:49) with given states."""
        self.states = np.array(states, self.dtype)
        self.ts[:] = 0
        return self.states.copy()

    def reset(self, dones):                                  # :585-595
        dones = np.asarray(dones, bool)
        R = len(self.reset_pool)
        for i, done in enumerate(dones):                     # row order, like the reference loop
            if done:
                if self.reset_mode == "ordered":
                    r = self._pool_cursor % R
                    self._pool_cursor += 1
                else:
                    r = (self.n_resets[i] * self.n_envs + i) % R
                self.states[i] = self.reset_pool[r]
                self.n_resets[i] += 1
        self.ts[dones] = 0
        return self.states[dones]

    def step(self, actions):                                 # :597-607
        self.ts += 1
        actions = np.clip(actions, -1.0, 1.0).astype(self.dtype)          # :599 (normalize(env))
        next_observations = self.get_next_observation(actions)
        rewards = -_envs.cost_np_vec(self.env, self.states, actions, next_observations)  # :601
        self.states = next_observations                                    # :602
        dones = _envs.is_done(self.env, self.states, next_observations)    # :603
        dones[self.ts >= self.max_path_length] = True                      # :604
        if np.any(dones):
            self.reset(dones)                                              # :605-606
        self.t += 1
        return self.states.copy(), rewards, dones, dict()

    def get_next_observation(self, actions):                 # :609-635
        B, K = self.n_envs, len(self.models)
        xu = np.concatenate([self.states, actions], axis=1)
        cand = _models.ensemble_forward(self.models, self.norm, xu, self.S, self.drop,
                                        self.dtype, self.mma)               # [K,B,S]
        m = self.sam_mode
        if m in ("step_rand", "eps_rand"):
            idx = self.noise.get_model_idx(self.t, B, K, self.n_resets)
            return cand[idx, np.arange(B)]
        if m == "model_mean_std":
            std = np.std(cand, axis=0)
            return (np.mean(cand, axis=0) + self.noise.get_std_noise(self.t, B, self.S) * std
                    ).astype(self.dtype)
        if m == "model_mean":
            return np.mean(cand, axis=0)
        if m == "model_med":
            return np.median(cand, axis=0)
        if m == "one_model":
            return cand[0]
        raise AssertionError("sam mode %s is not defined." % m)


# ------------------------------------------------------------------------------------------------
# VectorizedSampler.obtain_samples restatement
# ------------------------------------------------------------------------------------------------
def get_actions(pol, obses, eps, dtype=np.float32, out_tanh=False):
    """rllab GaussianMLPPolicy.get_actions: actions = rnd*exp(log_std) + mean."""
    mean = _models.policy_forward(pol, obses, dtype, out_tanh)
    log_std = np.maximum(pol["log_std"].astype(dtype), dtype(np.log(1e-6)))   # min_std clamp
    log_std = np.broadcast_to(log_std, mean.shape)
    actions = eps.astype(dtype) * np.exp(log_std) + mean
    return actions, dict(mean=mean, log_std=log_std)


def obtain_samples(vec_env, pol, init_states, batch_size, determ=False, out_tanh=False,
                   max_steps=None):
    """samplers/vectorized_sampler.py:45-116, including the per-env Python bookkeeping.
    Returns the list of COMPLETED path dicts (whole paths only)."""
    paths = []
    n_samples = 0
    obses = vec_env.set_states(init_states)                               # :49
    running_paths = [None] * vec_env.num_envs
    A, dtype = vec_env.A, vec_env.dtype
    steps = 0
    while n_samples < batch_size and (max_steps is None or steps < max_steps):   # :60
        eps = vec_env.noise.get_eps(vec_env.t, vec_env.num_envs, A)
        actions, agent_infos = get_actions(pol, obses, eps, dtype, out_tanh)      # :63
        if determ:
            actions = agent_infos["mean"]                                         # :64-65
        next_obses, rewards, dones, _ = vec_env.step(actions)                     # :69
        for idx in range(vec_env.num_envs):                                       # :80-105
            if running_paths[idx] is None:
                running_paths[idx] = dict(observations=[], actions=[], rewards=[], mean=[],
                                          log_std=[])
            rp = running_paths[idx]
            rp["observations"].append(obses[idx])
            rp["actions"].append(actions[idx])          # UNCLIPPED action (:92)
            rp["rewards"].append(rewards[idx])
            rp["mean"].append(agent_infos["mean"][idx])
            rp["log_std"].append(agent_infos["log_std"][idx])
            if dones[idx]:
                paths.append(dict(
                    observations=np.asarray(rp["observations"]),
                    actions=np.asarray(rp["actions"]),
                    rewards=np.asarray(rp["rewards"]),
                    env_infos=dict(),
                    agent_infos=dict(mean=np.asarray(rp["mean"]), log_std=np.asarray(rp["log_std"])),
                ))
                n_samples += len(rp["rewards"])
                running_paths[idx] = None
        obses = next_obses
        steps += 1
    return paths


def rollout_flat(env, pol, models, norm, init_states, reset_pool, noise, n_steps, max_path_length,
                 sam_mode="step_rand", determ=False, out_tanh=False, dtype=np.float32, mma="fp32",
                 reset_mode="per_row", teacher_states=None):
    """Same semantics as obtain_samples for exactly n_steps steps, returning the time-major flat
    buffers of the C ABI: obs[T,B,S] (pre-step), act[T,B,A] (unclipped), mean[T,B,A], rew[T,B],
    done[T,B], final_states[B,S].

    teacher_states[T,B,S] (optional): teacher forcing -- the pre-step observation of step t is
    replaced by teacher_states[t] (per-step parity check without open-loop error compounding)."""
    spec = _envs.ENV_SPECS[_envs.env_name(env)]
    S, A, drop = spec["S"], spec["A"], spec["drop"]
    B = len(init_states)
    ve = VecSimpleEnvOracle(env, models, norm, B, max_path_length, sam_mode, noise, reset_pool,
                            S, A, drop, dtype, mma, reset_mode)
    obses = ve.set_states(init_states)
    out = dict(obs=np.zeros((n_steps, B, S), np.float32), act=np.zeros((n_steps, B, A), np.float32),
               mean=np.zeros((n_steps, B, A), np.float32), rew=np.zeros((n_steps, B), np.float32),
               done=np.zeros((n_steps, B), np.uint8))
    for t in range(n_steps):
        if teacher_states is not None:
            ve.states = np.array(teacher_states[t], dtype)
            obses = ve.states.copy()
        eps = noise.get_eps(ve.t, B, A)
        actions, infos = get_actions(pol, obses, eps, dtype, out_tanh)
        if determ:
            actions = infos["mean"]
        next_obses, rewards, dones, _ = ve.step(actions)
        out["obs"][t], out["act"][t], out["mean"][t] = obses, actions, infos["mean"]
        out["rew"][t], out["done"][t] = rewards, dones
        obses = next_obses
    out["final_states"] = np.asarray(obses, np.float32)
    return out


def paths_from_flat(flat, log_std):
    """Split flat time-major buffers into the reference's list of completed path dicts, in the
    order obtain_samples appends them (step-major, then row)."""
    T, B = flat["rew"].shape
    start = np.zeros(B, np.int64)
    paths = []
    for t in range(T):
        for b in np.nonzero(flat["done"][t])[0]:
            sl = slice(start[b], t + 1)
            L = t + 1 - start[b]
            paths.append(dict(
                observations=flat["obs"][sl, b], actions=flat["act"][sl, b],
                rewards=flat["rew"][sl, b], env_infos=dict(),
                agent_infos=dict(mean=flat["mean"][sl, b],
                                 log_std=np.broadcast_to(log_std, (L, len(log_std))).copy())))
            start[b] = t + 1
    return paths


# ------------------------------------------------------------------------------------------------
# build_policy_graph restatement (per-model validation cost; R12)
# ------------------------------------------------------------------------------------------------
def cost_tf(env, x, u, x_next, dones=None):
    """envs/com_*_env.py cost_tf: batch MEAN of the per-row cost; Ant's takes the running `dones`
    mask and zeroes the rows that already terminated (envs/com_ant_env.py:70-75)."""
    c = _envs.cost_np_vec(env, x, u, x_next)
    if dones is not None:
        c = c * (1 - dones)
    return np.mean(c, dtype=c.dtype)


def model_costs(env, pol, models, norm, init_states, n_steps, gamma=1.0, out_tanh=False,
                dtype=np.float32, mma="fp32", return_rows=False):
    """model_based_rl.py:122-142: for each model i, x <- init; for t < T: u = clip(policy(x)) with
    stochastic = 0 (:130), x' = dynamics_model_i([x,u]) (:132-134), cost += gamma**t *
    cost_tf(x,u,x'[,dones]) (:136-141), dones = max(dones, is_done_tf(x,x')) after the cost (Ant
    only, :137), x <- x'.  Returns [K] (and the per-row discounted sums [K,B])."""
    name = _envs.env_name(env)
    spec = _envs.ENV_SPECS[name]
    S, drop = spec["S"], spec["drop"]
    K, B = len(models), len(init_states)
    costs = np.zeros(K, dtype)
    rows = np.zeros((K, B), dtype)
    for i, m in enumerate(models):
        x = np.array(init_states, dtype)
        dones = np.zeros(B, dtype) if name == "ant" else None
        for t in range(n_steps):
            u = np.clip(_models.policy_forward(pol, x, dtype, out_tanh), -1.0, 1.0).astype(dtype)
            x_next = _models.dynamics_forward(m, norm, np.concatenate([x, u], 1), S, drop, dtype, mma)
            g = dtype(gamma ** t)
            c_rows = _envs.cost_np_vec(name, x, u, x_next).astype(dtype)
            if dones is not None:
                c_rows = c_rows * (1 - dones)
            costs[i] += g * cost_tf(name, x, u, x_next, dones)
            rows[i] += g * c_rows
            if dones is not None:
                dones = np.maximum(dones, _envs.is_done(name, x, x_next).astype(dtype))
            x = x_next
    return (costs, rows) if return_rows else costs
