"""CPU tests of the host-side mirrors of the reference sockets (no CUDA needed): path cutting,
process_samples (GAE / returns / centring) against the oracle-style recomputation, the linear
baseline, the sampler's n_envs rule, env-name handling and row sharding (incl. a 2-process gloo run)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from oracle import rollout as orl  # noqa: E402


def test_env_names_and_drop_cols():
    from me_trpo_b200.envs import ENV_SPECS, canonical_env_name, drop_cols_from_params
    assert canonical_env_name("half_cheetah") == canonical_env_name("half-cheetah") == "half-cheetah"
    with pytest.raises(AssertionError):
        canonical_env_name("walker")
    assert drop_cols_from_params({"ignore_xy_input": True}) == 2
    assert drop_cols_from_params({"ignore_x_input": True}) == 1
    assert drop_cols_from_params({"ignore_x_input": False}) == 0
    from oracle.envs import ENV_SPECS as O
    for k, v in ENV_SPECS.items():     # product table == oracle table
        assert (v["S"], v["A"], v["drop"], v["hidden"]) == (O[k]["S"], O[k]["A"], O[k]["drop"], O[k]["hidden"])


def test_paths_from_flat_matches_oracle_obtain_samples():
    from me_trpo_b200.samplers.vectorized_sampler import paths_from_flat
    name, env, K, B, T, T_max, hidden, sam_mode, _ = mg.CASES[0]
    inp = mg.make_inputs(env, K, B, T, hidden)
    noise = orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"])
    flat = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T, T_max)
    mine = paths_from_flat(flat, inp["pol"]["log_std"])
    ref = orl.paths_from_flat(flat, inp["pol"]["log_std"])
    assert len(mine) == len(ref) == B
    for a, b in zip(mine, ref):
        for k in ("observations", "actions", "rewards"):
            np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_array_equal(a["agent_infos"]["mean"], b["agent_infos"]["mean"])
        assert a["agent_infos"]["log_std"].shape == a["actions"].shape


class _Algo:
    discount, gae_lambda, center_adv, positive_adv = 0.99, 0.95, True, False


def test_process_samples_gae_and_returns():
    from me_trpo_b200.baselines import LinearFeatureBaseline
    from me_trpo_b200.samplers.base import BaseSampler, discount_cumsum
    rng = np.random.RandomState(0)
    paths = [dict(observations=rng.randn(L, 5), actions=rng.randn(L, 2), rewards=rng.randn(L),
                  agent_infos=dict(mean=rng.randn(L, 2), log_std=np.zeros((L, 2)))) for L in (7, 3, 11)]
    algo = _Algo()
    algo.baseline = LinearFeatureBaseline()
    # discount_cumsum == lfilter([1], [1, -d], x[::-1])[::-1]
    import scipy.signal
    x = rng.randn(9)
    np.testing.assert_allclose(discount_cumsum(x, 0.9), scipy.signal.lfilter([1], [1, -0.9], x[::-1])[::-1], atol=1e-12)
    sd = BaseSampler(algo).process_samples(0, paths)
    # first call: baseline predicts zeros -> advantages are lambda-discounted reward sums, centred
    adv = np.concatenate([discount_cumsum(p["rewards"], 0.99 * 0.95) for p in paths])
    np.testing.assert_allclose(sd["advantages"], (adv - adv.mean()) / (adv.std() + 1e-8), atol=1e-12)
    np.testing.assert_allclose(sd["returns"], np.concatenate([discount_cumsum(p["rewards"], 0.99) for p in paths]))
    assert sd["observations"].shape == (21, 5) and sd["agent_infos"]["mean"].shape == (21, 2)
    # the baseline was fitted AFTER the advantages were computed (samplers/base.py:167)
    assert algo.baseline._coeffs is not None and algo.baseline._coeffs.shape == (2 * 5 + 4,)
    sd2 = BaseSampler(algo).process_samples(1, paths)
    assert not np.allclose(sd2["advantages"], sd["advantages"])


def test_linear_baseline_recovers_linear_returns():
    from me_trpo_b200.baselines import LinearFeatureBaseline
    rng = np.random.RandomState(1)
    w = rng.randn(3)
    paths = []
    for L in (20, 30):
        o = rng.randn(L, 3)
        paths.append(dict(observations=o, rewards=np.zeros(L), returns=o.dot(w) + 0.5))
    bl = LinearFeatureBaseline(reg_coeff=1e-8)
    bl.fit(paths)
    np.testing.assert_allclose(bl.predict(paths[0]), paths[0]["returns"], atol=1e-4)


def test_shard_rows_partitions():
    from me_trpo_b200.parallel import shard_rows
    for n, w in [(4096, 8), (4097, 8), (5, 8), (100, 3)]:
        blocks = [shard_rows(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from me_trpo_b200.parallel import shard_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name, env, K, B, T, T_max, hidden, sam_mode, _ = mg.CASES[1]      # Philox noise case
    inp = mg.make_inputs(env, K, B, T, hidden)
    lo, hi = shard_rows(B, rank, world)
    # rows are independent: a rank rolls out its block with noise keyed by GLOBAL row (row0 = lo).
    # Each rank needs its own reset pool slice consistent with the per-row rule (n*B + i) % R:
    pool = inp["pool"]
    R = len(pool)
    n_res = -(-T // T_max)
    local_pool = np.stack([pool[(n * B + i) % R] for n in range(n_res) for i in range(lo, hi)])
    noise = orl.PhiloxNoise(1234, 7, lo, sam_mode)
    part = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"][lo:hi], local_pool, noise, T, T_max)
    t = torch.tensor(part["obs"])
    gathered = [torch.zeros(T, shard_rows(B, r, world)[1] - shard_rows(B, r, world)[0], t.shape[2]) for r in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        np.save(os.path.join(tmp, "sharded_obs.npy"), torch.cat(gathered, dim=1).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_rollout_equals_unsharded_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 path's host logic: block sharding + global-row noise keys make
    the concatenation of per-rank rollouts identical to the single-rank rollout (no collective is
    needed during the rollout itself; the gather here only checks the result)."""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded_obs.npy"))
    name, env, K, B, T, T_max, hidden, sam_mode, _ = mg.CASES[1]
    inp = mg.make_inputs(env, K, B, T, hidden)
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"],
                           orl.PhiloxNoise(1234, 7, 0, sam_mode), T, T_max)
    np.testing.assert_array_equal(got, ref["obs"])


def test_sample_trajectories_param_noise_and_scalar_action_noise():
    """env_helpers.py:37-59,352-460: per-episode parameter-space perturbation
    param_noise * diff_weights * randn (biases then matrices; restored after the episode) and ONE
    scalar action-noise draw per step broadcast over the action dimensions."""
    torch = pytest.importorskip("torch")
    from me_trpo_b200.model_based_rl import prepare_policy, sample_trajectories
    from me_trpo_b200.policies import GaussianMLPPolicy
    from me_trpo_b200.real_env import SyntheticEnv
    pol = GaussianMLPPolicy(10, 2, (32, 32), device="cpu", seed=1)
    W = [w.numpy().astype(np.float64) for w in pol.W]
    b = [v.numpy().astype(np.float64) for v in pol.b]
    n_wb = sum(w.size for w in W) + sum(v.size for v in b)
    diff = np.full(n_wb + 2, 0.01)
    # no diff_weights yet (first sweep): nothing is perturbed, initial_param_std must be 0
    W1, b1, ch = prepare_policy(W, b, 3.0, None, 0.0, np.random.RandomState(0))
    assert ch == 0.0 and W1 is W
    with pytest.raises(AssertionError):
        prepare_policy(W, b, 3.0, None, 0.5, np.random.RandomState(0))
    # the reference's draw order: randn(n) over [b0, b1, b2, W0, W1, W2]
    W2, b2, ch = prepare_policy(W, b, 3.0, diff, 0.0, np.random.RandomState(0))
    z = np.random.RandomState(0).randn(n_wb)
    np.testing.assert_allclose(b2[0] - b[0], 3.0 * 0.01 * z[:32])
    o = sum(v.size for v in b)
    np.testing.assert_allclose(W2[0] - W[0], (3.0 * 0.01 * z[o:o + W[0].size]).reshape(W[0].shape))
    assert ch == pytest.approx(np.mean(np.abs(3.0 * 0.01 * z)))
    assert np.array_equal(W[0], pol.W[0].numpy().astype(np.float64))     # originals untouched
    # scalar action noise: with zero weights the action is noise * randn(1) in every dimension
    for w in pol.W:
        w.zero_()
    env = SyntheticEnv("swimmer", seed=0)
    expl = dict(action_noise=0.3, vary_trajectory_noise=False, param_noise=0.0, initial_param_std=0.0)
    Os, As, Rs, info = sample_trajectories(env, pol, expl, 20, 10, np.random.RandomState(3))
    acts = np.concatenate([np.asarray(a) for a in As])
    assert acts.shape[1] == 2 and np.array_equal(acts[:, 0], acts[:, 1]) and np.abs(acts).max() > 0
    assert info["TimeStepsCollected"] >= 20 and info["avg_weight_change"] == 0.0
