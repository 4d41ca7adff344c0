"""CPU tests of the oracle (test infrastructure) against committed golden vectors, the published
Philox4x32-10 known-answer vectors and the properties the reference itself states
(SURVEY.md section 4): test_runningmeanstd (running_mean_std.py:44-60), the test_policy_cost
equivalence "fused T-step rollout == T single steps" (env_helpers.py:271-305), |u| <= 1."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from oracle import envs as oe, models as om, rollout as orl  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "rollout_golden.npz"))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        got = orl.philox4x32(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_philox_normal_moments_and_index_range():
    z = orl.philox_normal(7, 3, np.arange(20000), 6, orl.STREAM_EPS)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    idx = orl.philox_index(7, np.uint32(3), np.arange(20000), 5, orl.STREAM_IDX)
    assert idx.min() == 0 and idx.max() == 4
    assert np.all(np.abs(np.bincount(idx) / 20000.0 - 0.2) < 0.02)


def test_philox_step_counter_is_64_bit():
    """Offsets that differ by 2**32 must not replay the stream (VectorizedSampler spaces its calls
    by 2**20 steps for the whole run: call n and call n + 4096 used to collide)."""
    rows = np.arange(64)
    a = orl.philox_normal(7, 5, rows, 6, orl.STREAM_EPS)
    b = orl.philox_normal(7, 5 + (1 << 32), rows, 6, orl.STREAM_EPS)
    assert not np.array_equal(a, b) and abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 0.2
    ia = orl.philox_index(7, 5, rows, 5, orl.STREAM_IDX)
    ib = orl.philox_index(7, 5 + (1 << 32), rows, 5, orl.STREAM_IDX)
    assert not np.array_equal(ia, ib)
    n1 = orl.PhiloxNoise(7, offset=(1 << 32) - 2)       # the carry into the high word
    np.testing.assert_array_equal(n1.get_eps(2, 64, 6), orl.philox_normal(7, 1 << 32, rows, 6, orl.STREAM_EPS))
    np.testing.assert_array_equal(n1.get_eps(1, 64, 6), orl.philox_normal(7, (1 << 32) - 1, rows, 6, orl.STREAM_EPS))


@pytest.mark.parametrize("case", mg.CASES, ids=[c[0] for c in mg.CASES])
@pytest.mark.parametrize("mma", ["fp32", "bf16"])
def test_oracle_matches_golden(case, mma):
    res = mg.run_case(case, mma)
    for k, v in res.items():
        ref = GOLD["%s/%s/%s" % (case[0], mma, k)]
        if k == "done":
            assert np.array_equal(v, ref)
        else:
            np.testing.assert_allclose(v, ref, rtol=0, atol=2e-6)


def test_bf16_rounding():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.14159, 1e-40, np.inf], np.float32)
    r = om.bf16_round(x)
    assert r[0] == 1.0 and r[1] == 1.0            # tie -> even
    assert r[2] == np.float32(1.0078125)
    assert abs(r[3] + 3.140625) < 1e-6
    assert np.isinf(r[5])


def test_runningmeanstd_property():
    """running_mean_std.py:44-60: after updates, mean/std equal those of the concatenated data."""
    rng = np.random.RandomState(0)
    for x1, x2 in [(rng.randn(3), rng.randn(4)), (rng.randn(3, 2), rng.randn(4, 2))]:
        # with the reference's tiny-count initialisation (epsilon = 0) the identity is exact
        rms = om.RunningMeanStd(epsilon=0.0, shape=x1.shape[1:])
        rms.update(x1)
        rms.update(x2)
        x = np.concatenate([x1, x2], axis=0)
        np.testing.assert_allclose(rms.mean, x.mean(axis=0), atol=1e-6)
        np.testing.assert_allclose(rms.std, np.sqrt(np.maximum(x.var(axis=0), 1e-2)), atol=1e-6)


def test_runningmeanstd_std_floor():
    rms = om.RunningMeanStd(shape=(2,))
    rms.update(np.full((100, 2), 3.0, np.float32))
    assert np.allclose(rms.std, 0.1, atol=1e-6)   # sqrt(max(var, 1e-2))  (:22-27)


@pytest.mark.parametrize("env", ["swimmer", "half-cheetah", "hopper", "ant", "humanoid", "snake"])
def test_costs_against_hand_formulas(env):
    spec = oe.ENV_SPECS[env]
    rng = np.random.RandomState(1)
    x = rng.randn(5, spec["S"]); xn = rng.randn(5, spec["S"]) * 3
    u = np.clip(rng.randn(5, spec["A"]), -1, 1)
    c = oe.cost_np_vec(env, x, u, xn)
    su2 = (u ** 2).sum(1)
    want = {
        "swimmer": lambda: -(xn[:, 5] - 0.01 * (u ** 2).mean(1)),
        "half-cheetah": lambda: -np.clip(xn[:, 9] - 0.05 * su2, -10, 10),
        "hopper": lambda: -(xn[:, 5] - 0.005 * su2 - 10 * np.maximum(0.45 - xn[:, 0], 0)
                            - 10 * np.maximum(np.abs(xn[:, 1]) - 0.2, 0)
                            - np.maximum(np.abs(xn[:, 2:]) - 100, 0).sum(1)),
        "ant": lambda: -(xn[:, 15] - 0.005 * su2 + 0.05),
        "humanoid": lambda: (xn[:, -1] - 1.5) ** 2 + 1e-5 * su2,
        "snake": lambda: -(xn[:, 7] - 0.005 * su2),
    }[env]()
    np.testing.assert_allclose(c, want, rtol=1e-12)
    with pytest.raises(AssertionError):           # the reference asserts |u| <= 1
        oe.cost_np_vec(env, x, u * 5, xn)


def test_ant_is_done():
    xn = np.zeros((4, 29)); xn[:, 2] = [0.5, 0.1, 1.2, 0.5]; xn[3, 7] = np.nan
    assert oe.is_done("ant", xn, xn).tolist() == [False, True, True, True]
    assert not oe.is_done("hopper", xn[:, :11], xn[:, :11]).any()


def test_fused_rollout_equals_single_steps():
    """test_policy_cost equivalence (env_helpers.py:271-305): the fused T-step rollout equals
    T applications of VecSimpleEnv.step with the same noise."""
    case = mg.CASES[0]
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    inp = mg.make_inputs(env, K, B, T, hidden)
    noise = orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"])
    flat = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T, T_max)
    spec = inp["spec"]
    ve = orl.VecSimpleEnvOracle(env, inp["models"], inp["norm"], B, T_max, sam_mode, noise, inp["pool"],
                                spec["S"], spec["A"], spec["drop"])
    obs = ve.set_states(inp["init"])
    for t in range(T):
        a, info = orl.get_actions(inp["pol"], obs, inp["eps"][t])
        assert np.array_equal(obs, flat["obs"][t])
        obs, r, d, _ = ve.step(a)
        np.testing.assert_array_equal(r.astype(np.float32), flat["rew"][t])
        assert np.array_equal(d.astype(np.uint8), flat["done"][t])


def test_obtain_samples_paths_match_flat_buffers():
    case = mg.CASES[0]
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    inp = mg.make_inputs(env, K, B, 8, hidden)
    noise = orl.ExplicitNoise(inp["eps"].repeat(2, 0)[:8], inp["mi"].repeat(2, 0)[:8])
    spec = inp["spec"]
    ve = orl.VecSimpleEnvOracle(env, inp["models"], inp["norm"], B, T_max, sam_mode, noise, inp["pool"],
                                spec["S"], spec["A"], spec["drop"])
    paths = orl.obtain_samples(ve, inp["pol"], inp["init"], batch_size=B * 8)
    flat = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, 8, T_max)
    p2 = orl.paths_from_flat({k: v for k, v in flat.items() if k != "final_states"}, inp["pol"]["log_std"])
    assert len(paths) == len(p2) == 2 * B         # every row completes 2 paths of length T_max = 4
    for a, b in zip(paths, p2):
        assert a["observations"].shape == (T_max, spec["S"])
        np.testing.assert_array_equal(a["observations"], b["observations"])
        np.testing.assert_array_equal(a["actions"], b["actions"])       # UNCLIPPED actions (:92)
        np.testing.assert_array_equal(a["rewards"].astype(np.float32), b["rewards"])
    assert max(np.abs(p["actions"]).max() for p in paths) > 1.0         # clipping is not recorded


def test_reset_modes_coincide_when_rows_finish_together():
    name, env, K, B, T, T_max, hidden, sam_mode, _ = mg.CASES[0]
    inp = mg.make_inputs(env, K, B, T, hidden)
    noise = orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"])
    a = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T, T_max,
                         reset_mode="per_row")
    b = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T, T_max,
                         reset_mode="ordered")
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_bf16_mode_close_to_fp32_reference():
    """Stated tolerance of the tensor-core arithmetic w.r.t. the reference's fp32 path, one step,
    teacher-forced: |next_state error| <= 5e-4 for unit-scale inputs."""
    name, env, K, B, T, T_max, hidden, sam_mode, _ = mg.CASES[0]
    a = GOLD["%s/fp32/obs" % name][1]   # state after one step
    b = GOLD["%s/bf16/obs" % name][1]
    assert np.max(np.abs(a - b)) < 5e-4
