"""The CUDA path (through the C ABI) against tests/golden/ref_fixtures.npz -- outputs of the
REFERENCE'S OWN code (see tests/golden/make_ref_fixtures.py).  Inputs come from
tests/golden/ref_inputs.py + the fixtures; the oracle is used only to locate rows that sit on a
done threshold (where bf16 and fp32 may legitimately disagree).

The fixtures' networks are UNSCALED Xavier-initialised nets (out_scale = 1, states of magnitude
1..10, an expansive map) -- a hard case for reduced precision.  Tolerances (reference = fp32
tf.matmul semantics, kernel = bf16 operands / fp32 accumulate; measured maxima in brackets,
B200, round 2):
  one teacher-forced step   next state <= 5e-3 abs [4.3e-3 at |x| ~ 10]; reward <= 5e-3 * gain,
                            gain = 1 + the env's largest penalty slope (hopper 10x) [2.1e-2 hopper,
                            8e-4 half-cheetah]                                   (TOL_STEP)
  per-model T-step cost     <= 3e-3 * T + 5e-3 * |cost|  [half-cheetah T=12: 2.8e-2 on 13.2;
                            hopper T=8: 0.28 on 81; humanoid T=6: 2.5e-2 on 9.6]   (R12)
  TRPO half (fp32 kernels)  advantages <= 2e-4, loss / KL <= 2e-5"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_inputs as RI  # noqa: E402

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(HERE, "golden", "ref_fixtures.npz"))
TOL_STEP = 5e-3
REWARD_GAIN = {"hopper": 11.0}      # 10 * max(0.45 - height, 0) etc. (com_hopper_env.py:94-104)
HID = (256, 256)


def fx(prefix):
    return {k[len(prefix):]: FIX[k] for k in FIX.files if k.startswith(prefix)}


def _norm(f):
    return {k: f[k] for k in ("in_mean", "in_std", "diff_mean", "diff_std")}


def _policy_hidden(env):
    return (100, 50, 25) if env == "humanoid" else (32, 32)


def _rollout(env, K, B, mpl, sam_mode, models, norm, pol=None):
    from me_trpo_b200.rollout import EnsembleRollout
    ro = EnsembleRollout(env, K, B, mpl, hidden=HID[0], sam_mode=sam_mode, device="cuda:0")
    ro.set_dynamics_ensemble(models)
    ro.set_normalization(**norm)
    if pol is not None:
        ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    return ro


def _replay_ts(dones, mpl):
    """timeout[t,row]: the reference's ts reached max_path_length at step t (env_helpers.py:604)."""
    T, B = dones.shape
    ts = np.zeros(B, int)
    timeout = np.zeros((T, B), bool)
    for t in range(T):
        ts += 1
        timeout[t] = ts >= mpl
        ts[dones[t].astype(bool)] = 0
    return timeout


VEC_CASES = [("half-cheetah", 5, 31), ("ant", 4, 32), ("hopper", 3, 33)]


@pytest.mark.parametrize("env,K,seed", VEC_CASES)
@pytest.mark.parametrize("sam_mode", RI.SAM_MODES)
def test_cuda_step_matches_reference_vec_env_teacher_forced(env, K, seed, sam_mode):
    """metrpo_rollout_step against the reference's VecSimpleEnv.step, one step at a time from the
    reference's own states (no open-loop compounding), all six sam_modes."""
    from oracle import models as om
    f = fx("D_vec__%s__%s__" % (env, sam_mode))
    K_, B, T, mpl, _ = [int(v) for v in f["cfg"]]
    _, _, S, A, _ = RI.ENVS[env]
    models = RI.dynamics_weights(seed, S, A, RI.DROP[env], HID, K)
    norm = _norm(f)
    ro = _rollout(env, K, B, 1 << 20, sam_mode, models, norm)
    timeout = _replay_ts(f["dones"], mpl)
    worst_s = worst_r = 0.0
    n_compared = 0
    for t in range(T):
        pre = f["obs0"] if t == 0 else f["states"][t - 1]
        ro.reset(pre.astype(np.float32))
        mi = f["model_idx"][t] if "model_idx" in f else None
        sn = f["std_noise"][t].astype(np.float32) if "std_noise" in f else None
        obs, rew, done = ro.step(f["actions"][t].astype(np.float32), f["states"][t].astype(np.float32),
                                 model_idx=mi, std_noise=sn)
        ro.synchronize()
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool)
        ok = ~timeout[t]
        if env == "ant":     # rows whose fp32 candidate sits on a done threshold may flip in bf16
            a = np.clip(f["actions"][t], -1, 1).astype(np.float32)
            cand = om.ensemble_forward(models, norm, np.concatenate([pre.astype(np.float32), a], 1), S, RI.DROP[env])
            if sam_mode in ("step_rand", "eps_rand"):
                z = cand[mi, np.arange(B), 2]
            elif sam_mode == "one_model":
                z = cand[0, :, 2]
            elif sam_mode == "model_med":
                z = np.median(cand[:, :, 2], axis=0)
            else:
                z = cand[:, :, 2].mean(axis=0) + (0 if sn is None else sn[:, 2] * cand[:, :, 2].std(axis=0))
            ok &= np.minimum(np.abs(z - 0.2), np.abs(z - 1.0)) > 0.02
            n_compared += int(ok.sum())
        dr = np.abs(rew - f["rewards"][t])[np.isfinite(f["rewards"][t])].max()
        worst_r = max(worst_r, dr)
        if not ok.any():          # a step where every row timed out in the reference
            continue
        np.testing.assert_array_equal(done[ok], f["dones"][t][ok])
        # rows that are done in both got the reference's post-reset state (handed in as reset_states)
        ds = np.abs(obs[ok] - f["states"][t][ok]).max()
        worst_s = max(worst_s, ds)
    print("teacher-forced %s/%s: max |ds| %.2e  max |dr| %.2e" % (env, sam_mode, worst_s, worst_r))
    assert worst_s <= TOL_STEP and worst_r <= TOL_STEP * REWARD_GAIN.get(env, 1.0)
    assert env != "ant" or n_compared >= 40         # enough rows away from the done thresholds
    ro.close()


@pytest.mark.parametrize("env,K,seed", [("half-cheetah", 5, 31), ("hopper", 3, 33)])
def test_cuda_step_open_loop_timeouts_match_reference(env, K, seed):
    """Open loop with the device's own timeout counter: dones bit-exact, post-reset rows are the
    pool entries the reference consumed, states within the compounded bf16 tolerance."""
    f = fx("D_vec__%s__step_rand__" % env)
    K_, B, T, mpl, _ = [int(v) for v in f["cfg"]]
    _, _, S, A, _ = RI.ENVS[env]
    models = RI.dynamics_weights(seed, S, A, RI.DROP[env], HID, K)
    ro = _rollout(env, K, B, mpl, "step_rand", models, _norm(f))
    ro.reset(f["obs0"].astype(np.float32))
    for t in range(T):
        obs, rew, done = ro.step(f["actions"][t].astype(np.float32), f["states"][t].astype(np.float32),
                                 model_idx=f["model_idx"][t])
        ro.synchronize()
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), f["dones"][t])
        d = f["dones"][t].astype(bool)
        np.testing.assert_array_equal(obs.cpu().numpy()[d], f["states"][t][d])
        assert np.abs(obs.cpu().numpy() - f["states"][t]).max() <= TOL_STEP * mpl
        assert np.abs(rew.cpu().numpy() - f["rewards"][t]).max() <= TOL_STEP * mpl * REWARD_GAIN.get(env, 1.0)
    ro.close()


@pytest.mark.parametrize("env,K,seed", [("half-cheetah", 5, 51), ("ant", 4, 52), ("hopper", 3, 53),
                                        ("humanoid", 2, 54)])
def test_cuda_model_costs_match_reference_policy_graph(env, K, seed):
    """metrpo_rollout_model_costs against the reference's build_policy_graph (unrolled T-step
    graph per model, Ant (1 - dones) mask, gamma**t)."""
    f = fx("F_costs__%s__" % env)
    K_, B, T, _ = [int(v) for v in f["cfg"]]
    _, _, S, A, _ = RI.ENVS[env]
    models = RI.dynamics_weights(seed, S, A, RI.DROP[env], HID, K)
    pol = RI.policy_weights(seed + 1, S, _policy_hidden(env), A)
    ro = _rollout(env, K, B, T, "step_rand", models, _norm(f), pol)
    costs = ro.model_costs(T, f["init"].astype(np.float32), gamma=float(f["gamma"]))
    ro.synchronize()
    got, ref = costs.cpu().numpy(), f["policy_costs"]
    tol = 3e-3 * T + 5e-3 * np.abs(ref)
    print("model costs %s: dev %s ref %s" % (env, got, ref))
    assert np.all(np.abs(got - ref) <= tol)
    ro.close()


def _grid_from_reference(g, S, A):
    """Lay the reference's concatenated samples back onto the [step, env] grid the sampler ran on
    (path order = (finish step, env), samplers/vectorized_sampler.py:80-105)."""
    dg = g["dones_grid"].astype(bool)
    Tn, B = dg.shape
    obs = np.zeros((Tn, B, S), np.float32); act = np.zeros((Tn, B, A), np.float32)
    mean = np.zeros((Tn, B, A), np.float32); rew = np.zeros((Tn, B), np.float32)
    index = -np.ones((Tn, B), np.int64)
    start = np.zeros(B, int)
    o = 0
    for t in range(Tn):
        for b in np.nonzero(dg[t])[0]:
            L = t + 1 - start[b]
            sl = slice(start[b], t + 1)
            obs[sl, b] = g["observations"][o:o + L]; act[sl, b] = g["actions"][o:o + L]
            mean[sl, b] = g["mean"][o:o + L]; rew[sl, b] = g["rewards"][o:o + L]
            index[sl, b] = np.arange(o, o + L)
            o += L
            start[b] = t + 1
    assert o == len(g["rewards"])
    return dict(obs=obs, act=act, mean=mean, rew=rew, done=dg.astype(np.uint8), index=index)


@pytest.mark.parametrize("env,K,seed", [("half-cheetah", 5, 41), ("ant", 3, 42)])
def test_cuda_trpo_half_matches_reference_iteration(env, K, seed):
    """metrpo_trpo_process / _loss_kl on the reference's own samples: advantages and returns of
    samplers/base.py process_samples, surrogate loss and mean KL of algos/npo.py init_opt."""
    from me_trpo_b200.trpo import PolicyUpdate
    f = fx("E_iter__%s__" % env)
    _, _, S, A, _ = RI.ENVS[env]
    pol = RI.policy_weights(seed + 1, S, (32, 32), A)
    pu = PolicyUpdate([S, 32, 32, A], device="cuda:0")
    theta0 = np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(pol["W"], pol["b"])]
                            + [pol["log_std"]]).astype(np.float32)
    prev_coeffs = None
    for j in range(int(f["cfg"][3])):
        g = fx("E_iter__%s__it%d__" % (env, j))
        grid = _grid_from_reference(g, S, A)
        d = {k: torch.tensor(v, device="cuda") for k, v in grid.items() if k != "index"}
        out = pu.process(d["obs"], d["rew"], d["done"], baseline_coeffs=prev_coeffs,
                         discount=float(f["discount"]), gae_lambda=1.0)
        valid = out["valid"].cpu().numpy().astype(bool)
        np.testing.assert_array_equal(valid, grid["index"] >= 0)          # only completed paths
        idx = grid["index"][valid]
        adv = out["adv"].cpu().numpy()[valid]; ret = out["ret"].cpu().numpy()[valid]
        assert np.abs(adv - g["advantages"][idx]).max() <= 2e-4
        assert np.abs(ret - g["returns"][idx]).max() <= 2e-4 * max(1.0, np.abs(g["returns"]).max())
        # surrogate loss / KL at the sampling policy and at the moved policy
        ls = torch.tensor(pol["log_std"], device="cuda")
        for step, lk, kk in ((None, "surr_loss", "mean_kl"), (g["param_step"], "surr_loss_moved", "mean_kl_moved")):
            th = torch.tensor(theta0 if step is None else theta0 + step, device="cuda")
            loss, kl = pu.loss_kl(th, d["obs"], d["act"], out["adv"], d["mean"], ls, valid=out["valid"])
            assert abs(loss - float(g[lk])) <= 2e-5 * max(1.0, abs(float(g[lk])))
            assert abs(kl - float(g[kk])) <= 2e-5
        prev_coeffs = g["baseline_coeffs"]
    pu.close()


class _OrderedPool:
    """The real simulator's reset(): hands out the reference's recorded reset states in call order."""

    def __init__(self, pool):
        self.pool, self.i = pool, 0

    def __call__(self, n):
        out = self.pool[self.i:self.i + n]
        self.i += n
        assert len(out) == n, "more simulator resets than the reference made"
        return out


@pytest.mark.parametrize("env,K,seed", VEC_CASES)
@pytest.mark.parametrize("sam_mode", RI.SAM_MODES)
def test_vec_env_socket_reproduces_reference_run_end_to_end(env, K, seed, sam_mode):
    """The B1 socket (me_trpo_b200.env_helpers.NeuralNetEnv / VecSimpleEnv over libmetrpo.so, fp32
    fidelity mode) driven exactly like the reference drove its own VecSimpleEnv when the fixture was
    recorded: same global NumPy seed, same actions, the simulator's reset states in call order.  The
    whole open-loop run -- model draws, timeouts, Ant's early terminations, which row received
    which reset state -- must come out the same: the socket consumes np.random and the simulator
    in the reference's order (env_helpers.py:583,590-593,619,626)."""
    from me_trpo_b200.env_helpers import NeuralNetEnv
    f = fx("D_vec__%s__%s__" % (env, sam_mode))
    K_, B, T, mpl, seed_ = [int(v) for v in f["cfg"]]
    _, _, S, A, _ = RI.ENVS[env]
    models = RI.dynamics_weights(seed, S, A, RI.DROP[env], HID, K)
    sim = _OrderedPool(f["pool"])
    nn_env = NeuralNetEnv(env, models, _norm(f), sam_mode=sam_mode, reset_sampler=sim, hidden=HID[0],
                          precision="fp32")
    np.random.seed(seed_)
    ve = nn_env.vec_env_executor(n_envs=B, max_path_length=mpl)
    obs0 = ve.reset()
    np.testing.assert_array_equal(obs0, f["obs0"].astype(np.float32))
    for t in range(T):
        o, r, d, info = ve.step(f["actions"][t].astype(np.float32))
        np.testing.assert_array_equal(d, f["dones"][t])
        assert np.abs(o - f["states"][t]).max() <= 1e-4 and np.abs(r - f["rewards"][t]).max() <= 2e-4
        assert sim.i == int(f["n_reset_calls"][t])          # same number of simulator resets so far
    ve.terminate()
