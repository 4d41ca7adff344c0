"""CPU tests of the TRPO oracle (oracle/trpo.py) and of the host-side mirrors of the reference's
algo classes.  The oracle is pinned by properties (the reference ships no vectors for this path):
scipy's lfilter form of discount_cumsum, finite differences of the float64 graph, CG on a known
SPD system, and equivalence of the flat [T,B] processing with the reference-style path lists."""
import numpy as np
import pytest

from oracle import trpo as ot, rollout as orl, models as om


def _flat_case(T=30, B=16, S=18, seed=4, p_done=0.04, T_max=15):
    rng = np.random.RandomState(seed)
    obs = rng.normal(0, 3.0, (T, B, S)).astype(np.float32)
    rew = rng.normal(-1, 1, (T, B)).astype(np.float32)
    done = rng.rand(T, B) < p_done
    ts = np.zeros(B, int)
    for t in range(T):
        ts += 1
        done[t] |= ts >= T_max
        ts[done[t]] = 0
    return dict(obs=obs, rew=rew, done=done.astype(np.uint8))


def test_discount_cumsum_is_the_lfilter_form():
    import scipy.signal
    rng = np.random.RandomState(0)
    x = rng.normal(size=57)
    for d in (1.0, 0.99, 0.5):
        ref = scipy.signal.lfilter([1], [1, -d], x[::-1], axis=0)[::-1]   # rllab special.discount_cumsum
        assert np.allclose(ot.discount_cumsum(x, d), ref, atol=1e-12)


def test_flat_processing_equals_path_list_processing():
    fl = _flat_case()
    A = 6
    flat = dict(obs=fl["obs"], rew=fl["rew"], done=fl["done"], act=np.zeros((30, 16, A), np.float32),
                mean=np.zeros((30, 16, A), np.float32))
    paths = orl.paths_from_flat(flat, np.zeros(A, np.float32))
    bl = ot.LinearFeatureBaselineOracle()
    rng = np.random.RandomState(1)
    bl.coeffs = rng.normal(0, 0.1, 2 * 18 + 4)
    data = ot.process_samples(paths, bl, 0.99, 0.97)
    ref = ot.process_flat(fl, rng.__class__(1).normal(0, 0.1, 2 * 18 + 4), 0.99, 0.97)
    v = ref["valid"]
    assert v.sum() == len(data["advantages"]) and v.sum() < v.size
    assert np.allclose(np.sort(ot.center_advantages(ref["adv_raw"][v])), np.sort(data["advantages"]), atol=1e-10)
    assert np.allclose(np.sort(ref["ret"][v]), np.sort(data["returns"]), atol=1e-10)


def test_baseline_fit_recovers_feature_linear_returns():
    fl = _flat_case(T=40, B=32, S=5, seed=2)
    flat = dict(obs=fl["obs"], rew=fl["rew"], done=fl["done"], act=np.zeros((40, 32, 2), np.float32),
                mean=np.zeros((40, 32, 2), np.float32))
    paths = orl.paths_from_flat(flat, np.zeros(2, np.float32))
    c = np.random.RandomState(3).normal(size=14)
    for p in paths:
        p["returns"] = ot.baseline_features(p["observations"], len(p["rewards"])).dot(c)
    bl = ot.LinearFeatureBaselineOracle()
    bl.fit(paths)
    pred = np.concatenate([bl.predict(p) for p in paths])
    assert np.allclose(pred, np.concatenate([p["returns"] for p in paths]), atol=1e-3)


def _problem(N=200, seed=0):
    rng = np.random.RandomState(seed)
    S, A, hidden = 7, 3, (8, 8)
    pol = om.init_policy(rng, S, hidden, A)
    pol["log_std"] = rng.uniform(-0.5, 0.1, size=A).astype(np.float32)
    obs = rng.normal(size=(N, S))
    mean = om.policy_forward(pol, obs.astype(np.float32), np.float32).astype(np.float64)
    act = mean + rng.normal(size=(N, A)) * np.exp(pol["log_std"])
    adv = rng.normal(size=N)
    dims = [S] + list(hidden) + [A]
    return pol, dims, (obs, act, adv, mean, np.tile(pol["log_std"], (N, 1)))


def test_gradient_and_hvp_against_finite_differences():
    pol, dims, inp = _problem()
    orc = ot.TRPOOracle(dims)
    th = ot.flatten_params(pol) + np.random.RandomState(1).normal(0, 0.05, ot.flatten_params(pol).shape)
    g = orc.grad(th, inp)
    rng = np.random.RandomState(2)
    for _ in range(3):
        d = rng.normal(size=th.shape); d /= np.linalg.norm(d)
        h = 1e-5
        fd = (orc.loss_kl(th + h * d, inp)[0] - orc.loss_kl(th - h * d, inp)[0]) / (2 * h)
        assert abs(fd - g.dot(d)) < 1e-7
    # Hessian of mean_kl at old == new: v.H.v == second difference of the KL along v
    th0 = ot.flatten_params(pol)
    v = rng.normal(size=th0.shape); v /= np.linalg.norm(v)
    hv = orc.hvp(th0, inp, v) - orc.reg_coeff * v
    h = 1e-4
    k = lambda t: orc.loss_kl(t, inp)[1]
    second = (k(th0 + h * v) - 2 * k(th0) + k(th0 - h * v)) / h ** 2
    assert abs(second - v.dot(hv)) < 1e-5 * max(1.0, abs(second))
    assert abs(k(th0)) < 1e-12          # KL(old || old) = 0


def test_cg_solves_spd_system():
    rng = np.random.RandomState(0)
    M = rng.normal(size=(12, 12)); Aspd = M.dot(M.T) + 12 * np.eye(12)
    b = rng.normal(size=12)
    orc = ot.TRPOOracle([2, 2, 2], cg_iters=12)
    x = orc.cg(lambda p: Aspd.dot(p), b)
    assert np.allclose(Aspd.dot(x), b, atol=1e-6)


def test_optimize_improves_surrogate_within_trust_region():
    pol, dims, inp = _problem(N=500, seed=3)
    orc = ot.TRPOOracle(dims)
    th = ot.flatten_params(pol)
    new, info = orc.optimize(th, inp)
    assert info["accepted"] and info["loss_after"] < info["loss_before"] and info["kl"] <= 0.01
    assert not np.allclose(new, th)
    # zero advantages: nothing to gain -> step rejected, parameters restored
    inp0 = (inp[0], inp[1], np.zeros_like(inp[2]), inp[3], inp[4])
    new0, info0 = orc.optimize(th, inp0)
    assert not info0["accepted"] and np.array_equal(new0, th)
