"""R12: per-model validation cost (build_policy_graph, model_based_rl.py:122-142) and the
early-stopping logic that consumes it (utils.py:285-296, model_based_rl.py:1339-1419).

Tolerances: device vs oracle in the kernel's arithmetic (mma="bf16") 2e-4 * T on the per-row
discounted sums (T steps of open-loop rollout, errors add), per-model means 1e-4 * T; vs the
fp32-arithmetic oracle 2e-3 * T."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402
from oracle import rollout as orl  # noqa: E402
from oracle import envs as oe  # noqa: E402


# ---------------------------------------------------------------------------------------------
# CPU: oracle properties + host logic
# ---------------------------------------------------------------------------------------------
def test_oracle_model_costs_equals_summed_cost_of_one_model_rollout():
    """With K = 1, sam_mode one_model, determ policy and no timeout inside the horizon, the
    sampler path visits the same states; -sum(rewards) averaged over rows must equal the
    per-model cost (the reference's own `test_policy_cost` equivalence, env_helpers.py:271-305)."""
    env, B, T, hidden = "half-cheetah", 16, 6, 64
    inp = mg.make_inputs(env, 1, B, T, hidden)
    costs, rows = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], inp["init"], T, gamma=1.0,
                                  return_rows=True)
    flat = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"],
                            orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"]), T, 1000, "one_model",
                            determ=True)
    np.testing.assert_allclose(rows[0], -flat["rew"].sum(0), rtol=0, atol=1e-5)
    np.testing.assert_allclose(costs[0], np.mean(-flat["rew"].sum(0)), rtol=0, atol=1e-5)


def test_oracle_model_costs_discount_and_ant_mask():
    env, B, T, hidden = "ant", 32, 5, 64
    inp = mg.make_inputs(env, 2, B, T, hidden)
    init = inp["init"].copy()
    init[:, 2] = 0.5              # healthy height ...
    init[:4, 2] = 5.0             # ... except rows 0-3, which leave the band from step 0 on
    c1, r1 = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], init, T, gamma=0.9, return_rows=True)
    c0, r0 = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], init, 1, gamma=0.9, return_rows=True)
    # the first step is never masked (dones updates AFTER the cost, model_based_rl.py:136-137)
    assert np.all(r0 != 0)
    spec = oe.ENV_SPECS["ant"]
    assert r1.shape == (2, B) and c1.shape == (2,)
    # a row that is done after step 0 stops accumulating: its T-step sum equals its 1-step sum
    x1 = orl.model_costs  # noqa: F841
    from oracle import models as om
    u = np.clip(om.policy_forward(inp["pol"], init), -1, 1).astype(np.float32)
    xn = om.dynamics_forward(inp["models"][0], inp["norm"], np.concatenate([init, u], 1), spec["S"], spec["drop"])
    dead = oe.is_done("ant", init, xn)
    assert dead[:4].all()
    np.testing.assert_allclose(r1[0][dead], r0[0][dead], rtol=0, atol=0)


def test_stop_critereon_and_is_done_logic():
    from me_trpo_b200.utils import stop_critereon
    from me_trpo_b200 import model_based_rl as mb
    sc = stop_critereon(threshold=0.1, offset=1e-5, percent_models_threshold=0.30)
    old = np.array([1.0, 1.0, 1.0, 1.0, 1.0])
    assert not sc(old, np.array([0.9, 0.9, 0.9, 0.9, 1.1]), mode="vector")    # 20 % worse
    assert sc(old, np.array([0.9, 0.9, 0.9, 1.1, 1.1]), mode="vector")        # 40 % worse
    assert sc(1.0, 1.2) and not sc(1.0, 1.05)                                  # scalar mode
    with pytest.raises(AssertionError):
        sc(old, [1.0] * 5, mode="vector")                                      # must be ndarray
    pop = mb.policy_opt_params_from_json(dict(
        mode="estimated", whole=True, T=100, gamma=1.0, log_every=5, num_iters_threshold=25,
        max_iters=400, stop_critereon=dict(offset=1e-5, threshold=0.1, percent_models_threshold=0.3)))
    mins = {"real": 0.0, "estimated": old.copy()}
    assert mb.is_done(pop, mins, {"real": 0.0, "estimated": np.array([2.0, 2.0, 0.5, 0.5, 0.5])})
    assert not mb.is_done(pop, mins, {"real": 5.0, "estimated": np.array([2.0, 0.5, 0.5, 0.5, 0.5])})
    # update_stats: whole=True copies the candidate set; whole=False keeps per-entry minima
    cand = {"real": -1.0, "estimated": np.array([2.0, 0.5, 0.5, 0.5, 0.5])}
    m1 = {"real": 0.0, "estimated": old.copy()}
    mb.update_stats(m1, cand, whole=True)
    assert m1["real"] == -1.0 and np.array_equal(m1["estimated"], cand["estimated"])
    m2 = {"real": -2.0, "estimated": old.copy()}
    mb.update_stats(m2, cand, whole=False)
    assert m2["real"] == -2.0 and np.array_equal(m2["estimated"], [1.0, 0.5, 0.5, 0.5, 0.5])
    for mode, exp in (("real", True), ("no_early", False), ("one_model", True)):
        p2 = pop._replace(mode=mode)
        assert mb.is_done(p2, {"real": 0.0, "estimated": old}, {"real": 1.0, "estimated": old + 1}) == exp


# ---------------------------------------------------------------------------------------------
# GPU: kernel (per-model mode) vs oracle
# ---------------------------------------------------------------------------------------------
PM_CASES = [
    # name, env, K, n_rows, n_envs (handle), T, hidden, gamma
    ("hc_500rows", "half-cheetah", 5, 500, 500, 8, 256, 1.0),
    ("hc_rows_lt_envs", "half-cheetah", 5, 130, 1024, 6, 256, 0.97),
    ("ant_mask", "ant", 4, 256, 256, 6, 256, 0.99),
    ("hopper_k1", "hopper", 1, 64, 64, 5, 256, 1.0),
    ("humanoid", "humanoid", 3, 200, 256, 4, 128, 0.99),
    ("swimmer_many_tiles", "swimmer", 5, 4096, 4096, 3, 256, 1.0),   # 32 tiles > 29 slots: 2 tiles on some slots
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", PM_CASES, ids=[c[0] for c in PM_CASES])
def test_model_costs_match_oracle(case):
    from me_trpo_b200.rollout import EnsembleRollout
    name, env, K, n, B, T, hidden, gamma = case
    inp = mg.make_inputs(env, K, n, 1, hidden)
    init = inp["init"].copy()
    if env == "ant":
        init[:, 2] = 0.5
        init[::7, 2] = 3.0        # some rows terminate at step 0 -> masked afterwards
    ro = EnsembleRollout(env, K, B, 1000, hidden=hidden)
    ro.set_dynamics_ensemble(inp["models"])
    ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    costs, rows = ro.model_costs(T, init, gamma, return_rows=True)
    ro.synchronize()
    assert ro.last_launches() == 2
    costs, rows = costs.cpu().numpy(), rows.cpu().numpy()
    for mma, tol in (("bf16", 2e-4 * T), ("fp32", 2e-3 * T)):
        c_ref, r_ref = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], init, T, gamma,
                                       mma=mma, return_rows=True)
        assert np.max(np.abs(rows - r_ref)) <= tol, (name, mma, np.max(np.abs(rows - r_ref)))
        assert np.max(np.abs(costs - c_ref)) <= tol / 2, (name, mma)
    np.testing.assert_allclose(costs, rows.astype(np.float64).mean(1), rtol=1e-6, atol=1e-6)
    # a sampler run afterwards on the same handle still works (mode flag does not leak)
    out = ro.run(2, inp["init"] if n == B else np.zeros((B, inp["init"].shape[1]), np.float32),
                 np.zeros((B, inp["init"].shape[1]), np.float32))
    ro.synchronize()
    assert np.isfinite(out["rew"].cpu().numpy()).all()
    ro.close()


@pytest.mark.gpu
def test_optimize_policy_controller_runs_and_restores_best_policy():
    """The TRPO branch of optimize_policy (model_based_rl.py:1171-1299) end to end on a small
    problem: evaluates per-model costs every log_every iterations, stops on the threshold and
    leaves the best-so-far policy in place."""
    import test_algo_gpu as tag
    from me_trpo_b200 import model_based_rl as mb
    algo = tag._make_algo(B=128, T_max=20, hidden=256, K=3)
    pop = mb.policy_opt_params_from_json(dict(
        mode="estimated", whole=True, T=20, gamma=1.0, log_every=2, num_iters_threshold=4, max_iters=6,
        stop_critereon=dict(offset=1e-5, threshold=0.1, percent_models_threshold=0.3)))
    val_init = np.random.RandomState(3).normal(0, 0.1, (100, 18)).astype(np.float32)
    res = mb.optimize_policy(algo, pop, val_init)
    assert res["best_index"] % 2 == 0 and 0 <= res["best_index"] <= 6
    assert len(res["estimated_validation_costs"]["estimated"]) >= 2
    assert res["min_validation_costs"]["estimated"].shape == (3,)
    # the policy left in place is the best one: its costs are the recorded minima (whole=True)
    ev = mb.PolicyCostEvaluator(algo.env, algo.policy, 100, 20, 1.0)
    np.testing.assert_allclose(ev(val_init), res["min_validation_costs"]["estimated"], rtol=1e-5, atol=1e-5)
    ev.close()
    algo.shutdown_worker()
