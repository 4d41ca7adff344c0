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


# ---------------------------------------------------------------------------------------------
# multi-GPU host logic of the PRODUCT (me_trpo_b200.parallel.DistContext, the sampler's row
# sharding, model ownership) over gloo, world size 2
# ---------------------------------------------------------------------------------------------
class _FakeRollout:
    """Stands where EnsembleRollout does: records what the sampler hands to the device."""
    log = []

    def __init__(self, env, n_models, n_envs, max_path_length, row_offset=0, **kw):
        self.B, self.row_offset, self.device = n_envs, row_offset, "cpu"

    def set_dynamics_ensemble(self, m): pass
    def set_normalization(self, **k): pass
    def set_policy(self, *a): pass
    def synchronize(self): pass
    def close(self): pass

    def run(self, T, init, pool, seed=0, offset=0, determ=False):
        _FakeRollout.log.append(dict(T=T, init=np.array(init), pool=np.array(pool), rows=self.B,
                                     row_offset=self.row_offset, seed=seed, offset=offset))
        return {}


def _dist_host_worker(rank, world, port, tmp):
    import types
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from me_trpo_b200.parallel import DistContext, local_reset_pool
    from me_trpo_b200.policies import GaussianMLPPolicy
    from me_trpo_b200.samplers import vectorized_sampler as vs
    ctx = DistContext(rank, world)
    out = {}
    # 1. policy broadcast: ranks start from different parameters, end on rank 0's
    pol = GaussianMLPPolicy(18, 6, (32, 32), device="cpu", seed=100 + rank)
    ctx.broadcast_policy(pol)
    out["theta"] = pol.get_param_values()
    # 2. model ownership + weight exchange: model k is fitted by rank k % G
    K = 5
    mine = ctx.my_models(K)
    local = [dict(W0=torch.full((3, 4), float(k)), b0=torch.full((4,), 10.0 + k)) for k in mine]
    allm = ctx.gather_models(local, K)
    out["mine"] = np.asarray(mine)
    out["gathered"] = np.asarray([[float(m["W0"][0, 0]), float(m["b0"][0])] for m in allm])
    # 3. rank 0's arrays everywhere
    arrs = [np.arange(12, dtype=np.float32).reshape(3, 4) + 1, np.asarray([7, 8], np.int64)] if rank == 0 else [None, None]
    a, b = ctx.broadcast_arrays(arrs, "cpu")
    out["bc_a"], out["bc_b"] = a, b
    # 4. the sampler shards rows, keys noise by global row and slices the reset pool; only rank 0
    #    draws from the simulator
    draws = []
    rs = np.random.RandomState(5)

    def reset_sampler(n):
        draws.append(n)
        return rs.normal(size=(n, 18)).astype(np.float32)
    env = types.SimpleNamespace(vectorized=True, env_name="half-cheetah", n_models=K, hidden=256, sam_mode="step_rand",
                                device="cpu", models=[], norm={}, spec=None, reset_sampler=reset_sampler)
    algo = types.SimpleNamespace(env=env, policy=types.SimpleNamespace(hidden_sizes=(32, 32), output_tanh=False, W=[], b=[],
                                                                        log_std=None),
                                 batch_size=1000, max_path_length=10, discount=1.0, gae_lambda=1.0)
    vs.EnsembleRollout = _FakeRollout
    smp = vs.VectorizedSampler(algo, n_envs=37, seed=3, dist_ctx=ctx)
    smp.start_worker()
    smp.obtain_samples_flat(0)
    smp.obtain_samples_flat(1)
    rec = _FakeRollout.log
    out["rows"] = np.asarray([r["rows"] for r in rec]); out["row_offset"] = np.asarray([r["row_offset"] for r in rec])
    out["T"] = np.asarray([r["T"] for r in rec]); out["offsets"] = np.asarray([r["offset"] for r in rec])
    out["init0"], out["pool0"] = rec[0]["init"], rec[0]["pool"]
    out["n_draws"] = np.asarray(len(draws))
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_dist_context_and_sharded_sampler_gloo(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from me_trpo_b200.parallel import local_reset_pool, shard_rows
    from me_trpo_b200.policies import GaussianMLPPolicy
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_dist_host_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(2)]
    ref = GaussianMLPPolicy(18, 6, (32, 32), device="cpu", seed=100).get_param_values()
    np.testing.assert_array_equal(r[0]["theta"], ref)
    np.testing.assert_array_equal(r[1]["theta"], ref)                 # rank 1 now holds rank 0's policy
    assert r[0]["mine"].tolist() == [0, 2, 4] and r[1]["mine"].tolist() == [1, 3]
    want = np.asarray([[k, 10.0 + k] for k in range(5)])
    np.testing.assert_array_equal(r[0]["gathered"], want)
    np.testing.assert_array_equal(r[1]["gathered"], want)
    np.testing.assert_array_equal(r[1]["bc_a"], np.arange(12, dtype=np.float32).reshape(3, 4) + 1)
    assert r[1]["bc_b"].tolist() == [7, 8]
    # sampler: 37 rows -> blocks [0,19) and [19,37); T covers batch_size with whole horizons
    B, T_max = 37, 10
    T = -(-1000 // (B * T_max)) * T_max
    for i in range(2):
        lo, hi = shard_rows(B, i, 2)
        assert r[i]["rows"].tolist() == [hi - lo] * 2 and r[i]["row_offset"].tolist() == [lo] * 2
        assert r[i]["T"].tolist() == [T, T] and r[i]["offsets"].tolist() == [0, 1 << 20]
    assert int(r[0]["n_draws"]) == 4 and int(r[1]["n_draws"]) == 0       # only rank 0 touched the simulator
    rs = np.random.RandomState(5)
    n_res = -(-T // T_max)
    init = rs.normal(size=(B, 18)).astype(np.float32)
    pool = rs.normal(size=(B * n_res, 18)).astype(np.float32)
    np.testing.assert_array_equal(np.concatenate([r[0]["init0"], r[1]["init0"]]), init)
    for i in range(2):
        lo, hi = shard_rows(B, i, 2)
        np.testing.assert_array_equal(r[i]["pool0"], local_reset_pool(pool, B, lo, hi, n_res))
        # the kernel's per-row rule with LOCAL sizes picks the entry the unsharded run would
        nb = hi - lo
        for n in range(n_res):
            for j in (0, nb - 1):
                np.testing.assert_array_equal(r[i]["pool0"][(n * nb + j) % len(r[i]["pool0"])], pool[(n * B + lo + j) % len(pool)])


def test_completed_samples_per_step_counts_like_the_reference_loop():
    """n_samples of samplers/vectorized_sampler.py:96-104 (samples of COMPLETED paths after each
    step) from time-major done flags, against a brute-force replay of the per-env bookkeeping."""
    torch = pytest.importorskip("torch")
    from me_trpo_b200.samplers.vectorized_sampler import VectorizedSampler
    rs = np.random.RandomState(0)
    done = (rs.rand(57, 23) < 0.07)
    done[19] |= rs.rand(23) < 0.5
    cum = VectorizedSampler.completed_samples_per_step(torch.tensor(done.astype(np.uint8))).numpy()
    n, start, ref = 0, np.zeros(23, int), []
    for t in range(57):
        for b in range(23):
            if done[t, b]:
                n += t + 1 - start[b]
                start[b] = t + 1
        ref.append(n)
    np.testing.assert_array_equal(cum, ref)
