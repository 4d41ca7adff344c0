"""The oracle (and the product's host-side mirrors) against tests/golden/ref_fixtures.npz -- outputs
of the REFERENCE'S OWN code (tests/golden/make_ref_fixtures.py executes /root/reference's files
under TF/rllab shims in the build container).  This is what pins the oracle: every comparison
below is reference output vs oracle output on the same inputs.

Tolerances: exact (or 1e-12 relative) where the reference computes in float64 NumPy; 2e-5 absolute
/ relative where an fp32 matrix product is re-associated (NumPy BLAS blocking vs the row order of
the restatement)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_inputs as RI  # noqa: E402
from oracle import envs as oe, models as om, rollout as orl, trpo as otr, fit as ofit  # noqa: E402

FIX = np.load(os.path.join(HERE, "golden", "ref_fixtures.npz"))
F32 = dict(rtol=2e-5, atol=2e-5)


def fx(prefix):
    return {k[len(prefix):]: FIX[k] for k in FIX.files if k.startswith(prefix)}


def world_inputs(env, hidden, K, seed, out_scale=1.0, tag=None):
    """The same weights make_ref_fixtures.World loaded into the reference graph; normaliser
    constants are the REFERENCE's RunningMeanStd outputs (checked against the oracle's below)."""
    _, _, S, A, _ = RI.ENVS[env]
    models = RI.dynamics_weights(seed, S, A, RI.DROP[env], hidden, K, out_scale)
    pol = RI.policy_weights(seed + 1, S, (100, 50, 25) if env == "humanoid" else (32, 32), A)
    xu, diff = RI.rms_data(seed + 2, S, A)
    rin, rdf = om.RunningMeanStd(0.0, (S + A,)), om.RunningMeanStd(0.0, (S,))
    with np.errstate(invalid="ignore", divide="ignore"):
        rin.update(xu[:150]); rdf.update(diff[:150])
        first = dict(in_mean=rin.mean, in_std=rin.std, diff_mean=rdf.mean, diff_std=rdf.std)
        rin.update(xu[150:]); rdf.update(diff[150:])
    norm = dict(in_mean=rin.mean, in_std=rin.std, diff_mean=rdf.mean, diff_std=rdf.std)
    return S, A, models, pol, norm, first


# ---------------------------------------------------------------------------------------------
# A. cost / done (R5)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env", list(RI.ENVS))
def test_cost_and_done_match_reference_env_classes(env):
    f = fx("A_costs__%s__" % env)
    x, u, xn = f["x"], f["u"], f["x_next"]
    with np.errstate(invalid="ignore"):
        c = oe.cost_np_vec(env, x, u, xn)
    np.testing.assert_allclose(c, f["cost_np_vec"], rtol=1e-12, atol=0, equal_nan=True)
    assert c.dtype == f["cost_np_vec"].dtype == np.float64          # u is f64 in the sampler path
    np.testing.assert_allclose(np.mean(c), f["cost_np"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(oe.is_done(env, x, xn), f["is_done"])
    if env == "ant":
        assert f["is_done"][:5].tolist() == [False, False, True, True, True]   # bounds, nan, inf, nan z
        np.testing.assert_array_equal(f["is_done_tf"], f["is_done"].astype(np.float32))
        rows = f["cost_tf_rows"]
        got = orl.cost_tf(env, x[rows], u[rows].astype(np.float32), xn[rows], f["dones_in"][rows])
    else:
        got = orl.cost_tf(env, x, u.astype(np.float32), xn)
    np.testing.assert_allclose(got, f["cost_tf"], **F32)


@pytest.mark.parametrize("env", list(RI.ENVS))
def test_product_host_cost_mirror_matches_reference(env):
    from me_trpo_b200 import env_costs
    f = fx("A_costs__%s__" % env)
    with np.errstate(invalid="ignore"):
        c = env_costs.cost_np_vec(env, f["x"], f["u"], f["x_next"])
    np.testing.assert_allclose(c, f["cost_np_vec"], rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_array_equal(env_costs.is_done(env, f["x"], f["x_next"]), f["is_done"])


# ---------------------------------------------------------------------------------------------
# B. RunningMeanStd (R9)
# ---------------------------------------------------------------------------------------------
def test_running_mean_std_matches_reference_class():
    f = fx("B_rms__")
    r = om.RunningMeanStd(epsilon=1e-2, shape=(5,))
    np.testing.assert_allclose(r.mean, f["default_empty_mean"], rtol=1e-6)
    np.testing.assert_allclose(r.std, f["default_empty_std"], rtol=1e-6)
    r.update(f["x1"])
    np.testing.assert_allclose(r.mean, f["mean1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r.std, f["std1"], rtol=2e-4, atol=1e-6)   # E[x^2]-E[x]^2 cancellation in fp32
    r.update(f["x2"])
    np.testing.assert_allclose(r.mean, f["mean2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r.std, f["std2"], rtol=2e-4, atol=1e-6)
    assert f["std2"][1] == pytest.approx(0.1, rel=1e-6)                  # the sqrt(1e-2) floor (:25)


# ---------------------------------------------------------------------------------------------
# C. dynamics_model / policy_model / fit loss (R7, R8, N3)
# ---------------------------------------------------------------------------------------------
MODEL_CASES = [("swimmer", "swimmer", (256, 256), 3, 21, 32), ("half-cheetah", "half-cheetah", (256, 256), 5, 22, 32),
               ("hopper", "hopper", (256, 256), 3, 23, 32), ("ant", "ant", (256, 256), 3, 24, 32),
               ("humanoid", "humanoid", (256, 256), 2, 25, 32), ("snake", "snake", (256, 256), 2, 26, 32),
               ("half-cheetah-h1024", "half-cheetah", (1024, 1024), 2, 27, 16)]


@pytest.mark.parametrize("tag,env,hidden,K,seed,B", MODEL_CASES)
def test_models_match_reference_graph(tag, env, hidden, K, seed, B):
    f = fx("C_models__%s__" % tag)
    S, A, models, pol, norm, first = world_inputs(env, hidden, K, seed)
    for k in norm:
        np.testing.assert_allclose(first[k], f[k + "_first"], rtol=3e-4, atol=1e-6)
        np.testing.assert_allclose(norm[k], f[k], rtol=3e-4, atol=1e-6)
    ref_norm = {k: f[k] for k in norm}               # use the reference's constants from here on
    s = RI.states(seed + 3, B, S)
    a = np.clip(RI.actions(seed + 4, (B, A)), -1, 1)
    xu = np.concatenate([s, a], axis=1).astype(np.float32)            # the feed cast (f64 -> f32)
    cand = om.ensemble_forward(models, ref_norm, xu, S, RI.DROP[env])
    assert cand.shape == f["dyn_out"].shape == (K, B, S)
    np.testing.assert_allclose(cand, f["dyn_out"], **F32)
    np.testing.assert_allclose(om.policy_forward(pol, s), f["policy_mean"], **F32)
    # rllab get_actions (restated in the shim): actions = rnd * exp(log_std) + mean, float64
    acts, infos = orl.get_actions(pol, s, f["get_actions_rnd"], dtype=np.float32)
    np.testing.assert_allclose(acts, f["get_actions"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(infos["log_std"], f["get_actions_log_std"], rtol=1e-6)
    # fit loss: build_dynamics_graph's per-model prediction loss on column blocks of the full batch
    xf, yf = f["fit_x_full"], f["fit_y_full"]
    losses = [ofit.prediction_loss(m, ref_norm, ofit.get_ith_tensor(xf, i, S + A),
                                   ofit.get_ith_tensor(yf, i, S), S, RI.DROP[env])
              for i, m in enumerate(models)]
    np.testing.assert_allclose(losses, f["fit_losses"], rtol=5e-5)   # regulariser constant is 0.0
    np.testing.assert_allclose(np.sum(losses), f["fit_prediction_loss"], rtol=5e-5)


# ---------------------------------------------------------------------------------------------
# D. VecSimpleEnv (R3, R4, R6)
# ---------------------------------------------------------------------------------------------
VEC_CASES = [("half-cheetah", 5, 31), ("ant", 4, 32), ("hopper", 3, 33)]


def run_oracle_vec(env, sam_mode, f, models, norm, S, A, dtype=np.float32, mma="fp32"):
    K, B, T, mpl, _ = [int(v) for v in f["cfg"]]
    noise = orl.ExplicitNoise(model_idx=f.get("model_idx"), std_noise=f.get("std_noise"))
    pool = f["pool"]
    ve = orl.VecSimpleEnvOracle(env, models, norm, B, mpl, sam_mode, noise, pool[B:], S, A, RI.DROP[env],
                                dtype=dtype, mma=mma, reset_mode="ordered")
    obs0 = ve.set_states(pool[:B])                 # the initial reset() consumed the first B entries
    out = dict(obs0=obs0, states=[], rewards=[], dones=[])
    for t in range(T):
        with np.errstate(invalid="ignore"):
            s, r, d, _ = ve.step(f["actions"][t])
        out["states"].append(s); out["rewards"].append(r); out["dones"].append(d)
    return {k: (np.stack(v) if isinstance(v, list) else v) for k, v in out.items()}, ve


@pytest.mark.parametrize("env,K,seed", VEC_CASES)
@pytest.mark.parametrize("sam_mode", RI.SAM_MODES)
def test_vec_simple_env_matches_reference_class(env, K, seed, sam_mode):
    f = fx("D_vec__%s__%s__" % (env, sam_mode))
    S, A, models, pol, norm, _ = world_inputs(env, (256, 256), K, seed)
    out, ve = run_oracle_vec(env, sam_mode, f, models, norm, S, A)
    np.testing.assert_array_equal(out["obs0"], f["obs0"].astype(np.float32))
    np.testing.assert_array_equal(out["dones"], f["dones"])
    tol = dict(rtol=1e-4, atol=1e-4) if sam_mode == "model_mean_std" else F32
    np.testing.assert_allclose(out["states"], f["states"], **tol)
    np.testing.assert_allclose(out["rewards"], f["rewards"], **tol)
    # the reference consumed exactly one simulator reset per done row, in row order
    assert ve._pool_cursor + ve.n_envs == int(f["n_reset_calls"][-1]) == len(f["pool"])
    # dtype drift of the reference: states come back f32 from the TF run except in
    # model_mean_std, where the f64 np.random.normal promotes them; rewards are f64 (u is f64)
    assert str(f["states_dtype"]) == ("float64" if sam_mode == "model_mean_std" else "float32")
    assert str(f["rewards_dtype"]) == "float64"
    if env == "ant":
        assert 0.05 < f["dones"].mean() < 0.95       # early terminations, ragged resets


# ---------------------------------------------------------------------------------------------
# E. TRPO inner iteration: sampler + process_samples + surrogate / KL (R1, R2, R10, R11)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env,K,seed", [("half-cheetah", 5, 41), ("ant", 3, 42)])
def test_trpo_iteration_matches_reference_classes(env, K, seed):
    f = fx("E_iter__%s__" % env)
    K_, batch_size, T, n_iters, _ = [int(v) for v in f["cfg"]]
    S, A, models, pol, norm, _ = world_inputs(env, (256, 256), K, seed)
    discount = float(f["discount"])
    baseline = otr.LinearFeatureBaselineOracle()
    pool = f["pool"]
    n_used = 0
    for j in range(n_iters):
        g = fx("E_iter__%s__it%d__" % (env, j))
        n_envs = max(1, min(int(batch_size / T), 100))          # start_worker (:26-27)
        assert n_envs == int(g["n_envs"])
        lo, hi = [int(v) for v in g["reset_calls"]]
        assert lo == n_used
        my_pool = pool[lo:hi]
        noise = orl.ExplicitNoise(eps=g["eps"], model_idx=g["model_idx"])   # the reference's draws
        ve = orl.VecSimpleEnvOracle(env, models, norm, n_envs, T, "step_rand", noise, my_pool[n_envs:],
                                    S, A, RI.DROP[env], reset_mode="ordered")
        with np.errstate(invalid="ignore"):
            paths = orl.obtain_samples(ve, pol, my_pool[:n_envs], batch_size)
        assert len(paths) == int(g["n_paths"])
        np.testing.assert_array_equal([len(p["rewards"]) for p in paths], g["path_len"])
        assert ve._pool_cursor + n_envs == hi - lo           # same number of simulator resets
        n_used = hi
        for p in paths:
            p["agent_infos"]["log_std"] = np.asarray(p["agent_infos"]["log_std"])
        data = otr.process_samples(paths, baseline, discount)
        np.testing.assert_allclose(data["observations"], g["observations"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(data["actions"], g["actions"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(data["rewards"], g["rewards"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(data["returns"], g["returns"], rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(data["agent_infos"]["mean"], g["mean"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(data["advantages"], g["advantages"], rtol=2e-3, atol=2e-3)
        # exact part of process_samples: feed the reference's own paths
        base2 = otr.LinearFeatureBaselineOracle()
        base2.coeffs = None if j == 0 else prev_coeffs
        exact = otr.process_samples(_split_paths(g), base2, discount)
        np.testing.assert_allclose(exact["advantages"], g["advantages"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(exact["returns"], g["returns"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(base2.coeffs, g["baseline_coeffs"], rtol=1e-4, atol=1e-6)   # lstsq, cond ~1e9
        prev_coeffs = g["baseline_coeffs"]
        # surrogate loss / mean KL of NPO.init_opt at the sampling policy and at a moved policy
        dims = [S] + [32, 32] + [A]
        tr = otr.TRPOOracle(dims, step_size=float(f["max_constraint_val"]))
        inputs = (g["observations"], g["actions"], g["advantages"], g["mean"], g["log_std"])
        theta = otr.flatten_params(pol)
        loss, kl = tr.loss_kl(theta, inputs)
        assert abs(loss - float(g["surr_loss"])) < 2e-6 and abs(kl - float(g["mean_kl"])) < 1e-6
        loss, kl = tr.loss_kl(theta + g["param_step"].astype(np.float64), inputs)
        assert abs(loss - float(g["surr_loss_moved"])) < 2e-5 * max(1, abs(loss))
        assert abs(kl - float(g["mean_kl_moved"])) < 2e-5
    assert float(f["max_constraint_val"]) == 0.01             # params trpo.step_size


def _split_paths(g):
    """The reference's concatenated samples_data cut back into its paths (path_len order)."""
    paths, o = [], 0
    for L in g["path_len"]:
        L = int(L)
        paths.append(dict(observations=g["observations"][o:o + L], actions=g["actions"][o:o + L],
                          rewards=g["rewards"][o:o + L], env_infos={},
                          agent_infos=dict(mean=g["mean"][o:o + L], log_std=g["log_std"][o:o + L])))
        o += L
    return paths


# ---------------------------------------------------------------------------------------------
# F. per-model validation cost + stop logic (R12)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env,K,seed", [("half-cheetah", 5, 51), ("ant", 4, 52), ("hopper", 3, 53),
                                        ("humanoid", 2, 54)])
def test_model_costs_match_reference_policy_graph(env, K, seed):
    f = fx("F_costs__%s__" % env)
    K_, B, T, _ = [int(v) for v in f["cfg"]]
    S, A, models, pol, norm, _ = world_inputs(env, (256, 256), K, seed)
    with np.errstate(invalid="ignore"):
        costs = orl.model_costs(env, pol, models, norm, f["init"], T, gamma=float(f["gamma"]))
    np.testing.assert_allclose(costs, f["policy_costs"], rtol=2e-4, atol=2e-4)


def test_stop_logic_matches_reference_functions():
    from me_trpo_b200 import utils as putils, model_based_rl as pm
    f = fx("F_stop__")
    crit = putils.stop_critereon(threshold=0.1, offset=1e-5, percent_models_threshold=0.3)
    got_v = [crit(o, n, mode="vector") for o, n in zip(f["old"], f["new"])]
    got_s = [crit(float(o[0]), float(n[0])) for o, n in zip(f["old"], f["new"])]
    np.testing.assert_array_equal(got_v, f["vector"])
    np.testing.assert_array_equal(got_s, f["scalar"])
    assert 0 < f["vector"].mean() < 1
    if hasattr(pm, "is_done"):
        import types
        pop = types.SimpleNamespace(mode="estimated", stop_critereon=crit)
        got = [pm.is_done(pop, {"real": 0.0, "estimated": o.copy()}, {"real": 1.0, "estimated": n.copy()})
               for o, n in zip(f["old"], f["new"])]
        np.testing.assert_array_equal(got, f["is_done_estimated"])
        pop = types.SimpleNamespace(mode="real", stop_critereon=crit)
        got = [pm.is_done(pop, {"real": float(o[0])}, {"real": float(n[0])}) for o, n in zip(f["old"], f["new"])]
        np.testing.assert_array_equal(got, f["is_done_real"])
    if hasattr(pm, "update_stats"):
        for whole, key in ((True, "update_whole"), (False, "update_part")):
            for o, n, want in zip(f["old"], f["new"], f[key]):
                m = {"real": float(o[0]), "estimated": o.copy()}
                pm.update_stats(m, {"real": float(n[0]), "estimated": n.copy()}, whole=whole)
                np.testing.assert_array_equal(np.append(m["estimated"], m["real"]), want)


# ---------------------------------------------------------------------------------------------
# G. data_collection
# ---------------------------------------------------------------------------------------------
def test_get_ith_tensor_matches_reference():
    f = fx("G_ith__")
    got = np.stack([ofit.get_ith_tensor(f["t"], i, 4) for i in range(3)])
    np.testing.assert_array_equal(got, f["out"])


class _FixedUniform:
    def __init__(self, u):
        self.u = u

    def uniform(self, lo, hi, size):
        assert len(self.u) == size
        return self.u


def _drive_data_collection(dc, f, to_np):
    log = []
    for step in range(4):
        dc.add_data(f["add%d_x" % step], f["add%d_y" % step])
        np.testing.assert_array_equal(to_np(dc.x), f["after%d_x" % step])
        np.testing.assert_array_equal(to_np(dc.y), f["after%d_y" % step])
        log.append([dc.get_num_data(), dc.cur_idx])
        bx, _ = dc.get_next_batch(16)
        np.testing.assert_array_equal(to_np(bx), f["next%d_x" % step])
        log.append([dc.get_num_data(), dc.cur_idx])
        sx, sy = dc.sample(9, _FixedUniform(f["sample%d_u" % step]))
        np.testing.assert_array_equal(to_np(sx), f["sample%d_x" % step])
        np.testing.assert_array_equal(to_np(sy), f["sample%d_y" % step])
    np.testing.assert_array_equal(log, f["log"])


def test_oracle_data_collection_matches_reference_class():
    _drive_data_collection(ofit.DataCollection(max_size=50), fx("G_dc__"), np.asarray)


def test_product_data_collection_matches_reference_class():
    """Host logic of me_trpo_b200.dynamics.data_collection (FIFO cap, cursor, index sampling); the
    buffers are torch tensors, placed on the CPU here."""
    pytest.importorskip("torch")
    from me_trpo_b200.dynamics import data_collection
    _drive_data_collection(data_collection(max_size=50, device="cpu"), fx("G_dc__"), lambda t: t.numpy())
