"""GPU tests of the reference-facing sockets: B1 vec-env (NeuralNetEnv /
VecSimpleEnv.reset/step) and B2 sampler (VectorizedSampler.obtain_samples -> list of path dicts)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from oracle import rollout as orl, envs as oe  # noqa: E402

pytestmark = pytest.mark.gpu


class _PoolSampler:
    """Stands in for the real simulator's reset(): hands out rows of a fixed pool in order."""

    def __init__(self, pool):
        self.pool, self.i = pool, 0

    def __call__(self, n):
        idx = (self.i + np.arange(n)) % len(self.pool)
        self.i += n
        return self.pool[idx]


def test_vec_env_socket_matches_oracle_step_by_step():
    from me_trpo_b200.env_helpers import NeuralNetEnv
    env, K, B, T, T_max, hidden = "half-cheetah", 3, 140, 6, 100, 256
    inp = mg.make_inputs(env, K, B, T, hidden)
    rng = np.random.RandomState(5)
    nn_env = NeuralNetEnv(env, inp["models"], inp["norm"], sam_mode="step_rand",
                          reset_sampler=_PoolSampler(inp["init"]))
    assert nn_env.vectorized and nn_env.n_models == K
    assert nn_env.action_space.bounds[0].min() == -1 and nn_env.observation_space.shape == (18,)
    ve = nn_env.vec_env_executor(n_envs=B, max_path_length=T_max)
    ve.rng = np.random.RandomState(11)
    obs = ve.reset()
    assert obs.shape == (B, 18) and ve.num_envs == B
    np.testing.assert_array_equal(obs, inp["init"])
    spec = oe.ENV_SPECS[env]
    ref_rng = np.random.RandomState(11)
    ref_rng.randint(K, size=B)                                       # the reset() draw (:593)
    states = inp["init"].copy()
    for t in range(T):
        actions = rng.normal(size=(B, 6)).astype(np.float32) * 1.5   # unclipped
        o, r, d, info = ve.step(actions)
        idx = ref_rng.randint(K, size=B)
        oracle = orl.VecSimpleEnvOracle(env, inp["models"], inp["norm"], B, T_max, "step_rand",
                                        orl.ExplicitNoise(None, idx[None]), inp["pool"], spec["S"], spec["A"],
                                        spec["drop"], mma="bf16")
        oracle.set_states(states)
        o_ref, r_ref, d_ref, _ = oracle.step(actions)
        assert info == {} and o.shape == (B, 18) and d.dtype == bool
        assert np.max(np.abs(o - o_ref)) < 5e-5 and np.max(np.abs(r - r_ref)) < 5e-5
        assert not d.any()
        states = o
    ve.terminate()


def test_vec_env_socket_timeout_and_reset():
    from me_trpo_b200.env_helpers import NeuralNetEnv
    env, K, B, T_max, hidden = "swimmer", 2, 100, 3, 512
    inp = mg.make_inputs(env, K, B, 4, hidden)
    sampler = _PoolSampler(inp["pool"])
    nn_env = NeuralNetEnv(env, inp["models"], inp["norm"], sam_mode="eps_rand", reset_sampler=sampler)
    ve = nn_env.vec_env_executor(B, T_max)
    ve.reset()
    zeros = np.zeros((B, 2), np.float32)
    for t in range(T_max):
        before = sampler.i
        o, r, d, _ = ve.step(zeros)
        if t < T_max - 1:
            assert not d.any()
        else:
            assert d.all()                                           # ts >= max_path_length (:604)
            fresh = inp["pool"][(before + np.arange(B)) % len(inp["pool"])]
            np.testing.assert_array_equal(o, fresh)                  # post-reset observations (:605-607)
    with pytest.raises(RuntimeError):
        nn_env.vec_env_executor(B, T_max).step(zeros)                # step before reset
    ve.terminate()


class _Algo:
    discount, gae_lambda, center_adv, positive_adv = 1.0, 1.0, True, False


def test_sampler_socket_returns_reference_style_paths():
    from me_trpo_b200.baselines import LinearFeatureBaseline
    from me_trpo_b200.env_helpers import NeuralNetEnv
    from me_trpo_b200.policies import GaussianMLPPolicy
    from me_trpo_b200.samplers import VectorizedSampler
    env, K, B, T_max, hidden = "half-cheetah", 5, 256, 20, 256
    inp = mg.make_inputs(env, K, B, 1, hidden)
    algo = _Algo()
    algo.env = NeuralNetEnv(env, inp["models"], inp["norm"], reset_sampler=_PoolSampler(inp["pool"]))
    algo.policy = GaussianMLPPolicy(18, 6, (32, 32))
    algo.baseline = LinearFeatureBaseline()
    algo.batch_size, algo.max_path_length = 2 * B * T_max, T_max
    smp = VectorizedSampler(algo, n_envs=B)
    smp.start_worker()
    paths = smp.obtain_samples(0)
    assert len(paths) == 2 * B                                        # whole paths only (:94-103)
    assert sum(len(p["rewards"]) for p in paths) >= algo.batch_size
    p = paths[0]
    assert p["observations"].shape == (T_max, 18) and p["actions"].shape == (T_max, 6)
    assert p["agent_infos"]["mean"].shape == (T_max, 6) and p["agent_infos"]["log_std"].shape == (T_max, 6)
    assert max(np.abs(q["actions"]).max() for q in paths) > 1.0      # unclipped actions are stored (:92)
    # the recorded mean is the policy's mean at the recorded observation
    mean, _ = algo.policy.mean_and_log_std(p["observations"])
    assert np.max(np.abs(mean.cpu().numpy() - p["agent_infos"]["mean"])) < 1e-5
    det = smp.obtain_samples(1, determ=True)
    np.testing.assert_array_equal(det[0]["actions"], det[0]["agent_infos"]["mean"])
    sd = smp.process_samples(0, paths)
    assert sd["observations"].shape[0] == 2 * B * T_max and abs(sd["advantages"].mean()) < 1e-6
    # default n_envs rule of the reference (:26-27)
    smp2 = VectorizedSampler(algo)
    smp2.start_worker()
    assert smp2._n_envs == 100
    smp2.shutdown_worker()
    smp.shutdown_worker()


def test_ant_sampler_follows_reference_stop_rule():
    """Early-terminating paths (Ant): obtain_samples stops after the first step at which the COMPLETED
    paths hold batch_size samples (samplers/vectorized_sampler.py:60,96-105) -- not after a fixed
    number of horizons.  fp32 mode, so that done thresholds match the fp32 oracle exactly."""
    from me_trpo_b200.env_helpers import NeuralNetEnv
    from me_trpo_b200.policies import GaussianMLPPolicy
    from me_trpo_b200.samplers.vectorized_sampler import VectorizedSampler
    env, K, B, T_max, hidden, batch = "ant", 3, 40, 12, 256, 1000
    inp = mg.make_inputs(env, K, B, 4, hidden)
    rs = np.random.RandomState(3)
    big_pool = rs.normal(0, 0.1, (4000, 29)).astype(np.float32)
    big_pool[:, 2] = rs.uniform(0.22, 0.98, size=4000)            # start inside the healthy band, some near the edge
    nn_env = NeuralNetEnv(env, inp["models"], inp["norm"], reset_sampler=_PoolSampler(big_pool), hidden=hidden,
                          precision="fp32")
    pol = GaussianMLPPolicy(29, 8, (32, 32), device="cuda", seed=2)
    algo = _Algo()
    algo.env, algo.policy, algo.batch_size, algo.max_path_length = nn_env, pol, batch, T_max
    smp = VectorizedSampler(algo, n_envs=B, seed=11)
    smp.start_worker()
    paths = smp.obtain_samples(0)
    # oracle: the reference loop with the same Philox streams, per-row pool rule, fp32 arithmetic
    W = [w.cpu().numpy() for w in pol.W]; b = [v.cpu().numpy() for v in pol.b]
    opol = dict(W=W, b=b, log_std=pol.log_std.cpu().numpy())
    T_fixed = -(-batch // (B * T_max)) * T_max
    cap = T_fixed + 4 * T_max
    n_res = -(-cap // T_max) * 4
    init, pool = big_pool[:B], big_pool[B:B + B * n_res]
    spec = oe.ENV_SPECS[env]
    ve = orl.VecSimpleEnvOracle(env, inp["models"], inp["norm"], B, T_max, "step_rand", orl.PhiloxNoise(11, 0, 0, "step_rand"),
                                pool, spec["S"], spec["A"], spec["drop"], reset_mode="per_row")
    with np.errstate(invalid="ignore"):
        ref = orl.obtain_samples(ve, opol, init, batch)
    assert ve.t % T_max != 0 or len(ref) != B * (ve.t // T_max), "case must contain early terminations"
    assert len(paths) == len(ref)
    assert sum(len(p["rewards"]) for p in paths) == sum(len(p["rewards"]) for p in ref) >= batch
    for a, r in zip(paths, ref):
        assert len(a["rewards"]) == len(r["rewards"])
        np.testing.assert_allclose(a["observations"], r["observations"], atol=5e-5)
        np.testing.assert_allclose(a["rewards"], r["rewards"], atol=5e-5)
    smp.shutdown_worker()
