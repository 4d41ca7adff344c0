"""GPU parity tests of the TRPO half of the hot path (metrpo_trpo_* through the C ABI) against the
float64 oracle (oracle/trpo.py: samplers/base.py:48-182, algos/npo.py:33-111, rllab CG optimizer).

Tolerances: the kernels evaluate the policy in fp32 (like the reference's TF graph) and reduce in
fp64; the oracle is float64 throughout.  Gradients / Fisher-vector products: relative L2 error
<= 2e-4; loss / mean-KL: <= 2e-5 abs; advantages (post-centring): <= 1e-4 abs; the updated
parameter vector: step direction cosine >= 0.9999 and the same accepted back-track index.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _problem(env="half-cheetah", N=1000, seed=0, hidden=(32, 32), out_tanh=False):
    from oracle import envs as oe, models as om
    spec = oe.ENV_SPECS[env]
    S, A = spec["S"], spec["A"]
    rng = np.random.RandomState(seed)
    pol = om.init_policy(rng, S, hidden, A)
    pol["b"] = [rng.uniform(-0.1, 0.1, size=b.shape).astype(np.float32) for b in pol["b"]]
    pol["log_std"] = rng.uniform(-0.5, 0.1, size=A).astype(np.float32)
    obs = rng.normal(0, 1.0, (N, S)).astype(np.float32)
    mean = om.policy_forward(pol, obs, np.float32, out_tanh).astype(np.float32)
    eps = rng.normal(size=(N, A)).astype(np.float32)
    act = (mean + eps * np.exp(pol["log_std"])).astype(np.float32)
    adv = rng.normal(size=N).astype(np.float32)
    adv = ((adv - adv.mean()) / (adv.std() + 1e-8)).astype(np.float32)
    dims = [S] + list(hidden) + [A]
    return dict(pol=pol, obs=obs, act=act, adv=adv, mean=mean, log_std=pol["log_std"].copy(), dims=dims,
                out_tanh=out_tanh)


def _dev(pr):
    d = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    return dict(obs=d(pr["obs"]), act=d(pr["act"]), adv=d(pr["adv"]), old_mean=d(pr["mean"]),
                old_log_std=d(pr["log_std"]))


def _oracle_inputs(pr):
    N = len(pr["adv"])
    return (pr["obs"], pr["act"], pr["adv"], pr["mean"], np.tile(pr["log_std"], (N, 1)))


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("env,N,hidden,out_tanh", [
    ("half-cheetah", 1000, (32, 32), False),
    ("swimmer", 333, (32, 32), False),
    ("ant", 517, (32, 32), True),
    ("humanoid", 300, (100, 50, 25), False),
])
def test_gradient_fvp_loss_match_autograd(env, N, hidden, out_tanh):
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    pr = _problem(env, N, hidden=hidden, out_tanh=out_tanh)
    orc = ot.TRPOOracle(pr["dims"], out_tanh=out_tanh)
    theta = ot.flatten_params(pr["pol"])
    # perturb the parameters so that new != old (likelihood ratio != 1, KL > 0)
    rng = np.random.RandomState(1)
    theta_new = theta + rng.normal(0, 0.02, theta.shape)
    pu = PolicyUpdate(pr["dims"], out_tanh=out_tanh, device="cuda:0")
    assert pu.P == len(theta)
    d = _dev(pr)
    th_d = torch.tensor(theta_new.astype(np.float32), device="cuda")
    l_dev, k_dev = pu.loss_kl(th_d, **d)
    l_ref, k_ref = orc.loss_kl(theta_new.astype(np.float32), _oracle_inputs(pr))
    assert abs(l_dev - l_ref) <= 2e-5 and abs(k_dev - k_ref) <= 2e-5
    g_dev = pu.grad(th_d, **d)
    g_ref = orc.grad(theta_new.astype(np.float32), _oracle_inputs(pr))
    assert _rel(g_dev, g_ref) <= 2e-4
    # Fisher-vector product at old == new (how the optimizer calls it)
    th0_d = torch.tensor(theta.astype(np.float32), device="cuda")
    v = rng.normal(size=theta.shape).astype(np.float32)
    hv_dev = pu.grad(th0_d, **d, vec=torch.tensor(v, device="cuda"), reg_coeff=1e-5)
    hv_ref = orc.hvp(theta.astype(np.float32), _oracle_inputs(pr), v)
    assert _rel(hv_dev, hv_ref) <= 2e-4
    pu.close()


@pytest.mark.parametrize("impl,tol_l,tol_g", [("tf32x3", 2e-5, 2e-4), ("tf32", 2e-3, 5e-3)])
@pytest.mark.parametrize("env,N,out_tanh", [("half-cheetah", 1000, False), ("ant", 517, True), ("hopper", 95, False)])
def test_tensor_core_pass_variants_match_oracle(env, N, out_tanh, impl, tol_l, tol_g):
    """The warp-level tensor-core implementations of the sample pass (metrpo_trpo_set_pass_impl;
    not the default, see csrc/trpo_mma.cuh) against the float64 autograd oracle: the 3xTF32 split
    to the same tolerance as the fp32 SIMT pass, plain TF32 to 10-bit-mantissa tolerances."""
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    pr = _problem(env, N, hidden=(32, 32), out_tanh=out_tanh)
    orc = ot.TRPOOracle(pr["dims"], out_tanh=out_tanh)
    theta = ot.flatten_params(pr["pol"])
    rng = np.random.RandomState(1)
    theta_new = theta + rng.normal(0, 0.02, theta.shape)
    pu = PolicyUpdate(pr["dims"], out_tanh=out_tanh, device="cuda:0")
    pu.set_pass_impl(impl)
    d = _dev(pr)
    th_d = torch.tensor(theta_new.astype(np.float32), device="cuda")
    l_dev, k_dev = pu.loss_kl(th_d, **d)
    l_ref, k_ref = orc.loss_kl(theta_new.astype(np.float32), _oracle_inputs(pr))
    assert abs(l_dev - l_ref) <= tol_l and abs(k_dev - k_ref) <= tol_l
    assert _rel(pu.grad(th_d, **d), orc.grad(theta_new.astype(np.float32), _oracle_inputs(pr))) <= tol_g
    th0_d = torch.tensor(theta.astype(np.float32), device="cuda")
    v = rng.normal(size=theta.shape).astype(np.float32)
    hv_dev = pu.grad(th0_d, **d, vec=torch.tensor(v, device="cuda"), reg_coeff=1e-5)
    assert _rel(hv_dev, orc.hvp(theta.astype(np.float32), _oracle_inputs(pr), v)) <= tol_g
    pu.close()
    # wide policies are refused by this implementation, not silently routed elsewhere
    pw = PolicyUpdate([55, 100, 50, 25, 21], device="cuda:0")
    pw.set_pass_impl(impl)
    with pytest.raises(RuntimeError):
        pw.loss_kl(torch.zeros(pw.P, device="cuda"), torch.zeros(4, 55, device="cuda"), torch.zeros(4, 21, device="cuda"),
                   torch.zeros(4, device="cuda"), torch.zeros(4, 21, device="cuda"), torch.zeros(21, device="cuda"))
    pw.close()


def test_valid_mask_and_per_sample_log_std():
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    pr = _problem("half-cheetah", 700)
    rng = np.random.RandomState(3)
    valid = rng.rand(700) < 0.7
    orc = ot.TRPOOracle(pr["dims"])
    theta = (ot.flatten_params(pr["pol"]) + rng.normal(0, 0.02, 1868)).astype(np.float32)
    sub = tuple(a[valid] for a in _oracle_inputs(pr))
    pu = PolicyUpdate(pr["dims"], device="cuda:0")
    d = _dev(pr)
    d["old_log_std"] = torch.tensor(np.tile(pr["log_std"], (700, 1)), device="cuda")   # [N,A] form
    vd = torch.tensor(valid.astype(np.uint8), device="cuda")
    th_d = torch.tensor(theta, device="cuda")
    l_dev, k_dev = pu.loss_kl(th_d, **d, valid=vd)
    l_ref, k_ref = orc.loss_kl(theta, sub)
    assert abs(l_dev - l_ref) <= 2e-5 and abs(k_dev - k_ref) <= 2e-5
    assert _rel(pu.grad(th_d, **d, valid=vd), orc.grad(theta, sub)) <= 2e-4
    pu.close()


@pytest.mark.parametrize("env,N", [("half-cheetah", 4000), ("hopper", 1500)])
def test_update_matches_oracle_optimize(env, N):
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    pr = _problem(env, N, seed=5)
    orc = ot.TRPOOracle(pr["dims"])
    theta = ot.flatten_params(pr["pol"]).astype(np.float32)
    new_ref, info_ref = orc.optimize(theta, _oracle_inputs(pr))
    pu = PolicyUpdate(pr["dims"], device="cuda:0")
    d = _dev(pr)
    th_d = torch.tensor(theta, device="cuda")
    info = pu.update(th_d, **d).cpu().numpy()
    new_dev = th_d.cpu().numpy().astype(np.float64)
    assert info_ref["accepted"] and info[4] == 1.0
    assert int(info[3]) == info_ref["backtracks"]
    assert abs(info[0] - info_ref["loss_before"]) <= 2e-5
    assert abs(info[1] - info_ref["loss_after"]) <= 1e-4
    assert abs(info[2] - info_ref["kl"]) <= 1e-4 and info[2] <= 0.01
    step_dev, step_ref = new_dev - theta, new_ref - theta
    cos = step_dev.dot(step_ref) / (np.linalg.norm(step_dev) * np.linalg.norm(step_ref))
    assert cos >= 0.9999
    assert _rel(step_dev, step_ref) <= 5e-3
    # grad + finish, 10 x (FVP + cg step), step (d.Hd from the CG residual: no 11th Fisher-vector pass),
    # 15 x (prepare + loss + check), finalize
    assert pu.last_launches() == 2 + 10 * 2 + 1 + 15 * 3 + 1
    pu.close()


def test_update_rejects_when_no_improvement_possible():
    """advantages all zero -> gradient zero -> loss cannot decrease -> parameters restored."""
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    pr = _problem("half-cheetah", 512, seed=7)
    pr["adv"][:] = 0
    theta = ot.flatten_params(pr["pol"]).astype(np.float32)
    pu = PolicyUpdate(pr["dims"], device="cuda:0")
    th_d = torch.tensor(theta, device="cuda")
    info = pu.update(th_d, **_dev(pr)).cpu().numpy()
    assert info[4] == 0.0
    assert np.array_equal(th_d.cpu().numpy(), theta)
    pu.close()


def _flat_case(T=37, B=50, S=18, seed=0, p_done=0.04, T_max=15):
    rng = np.random.RandomState(seed)
    obs = rng.normal(0, 3.0, (T, B, S)).astype(np.float32)
    obs[0, 0, 0] = 25.0   # exercises the clip(obs, -10, 10) of the baseline features
    rew = rng.normal(-1, 1, (T, B)).astype(np.float32)
    done = rng.rand(T, B) < p_done
    ts = np.zeros(B, int)
    for t in range(T):   # add timeouts like VecSimpleEnv (env_helpers.py:604)
        ts += 1
        done[t] |= ts >= T_max
        ts[done[t]] = 0
    return dict(obs=obs, rew=rew, done=done.astype(np.uint8))


@pytest.mark.parametrize("with_baseline", [False, True])
@pytest.mark.parametrize("gamma,lam", [(1.0, 1.0), (0.99, 0.95)])
def test_process_matches_oracle(with_baseline, gamma, lam):
    from oracle import trpo as ot
    from me_trpo_b200.trpo import PolicyUpdate
    S = 18
    fl = _flat_case(S=S)
    rng = np.random.RandomState(2)
    coeffs = rng.normal(0, 0.1, 2 * S + 4) if with_baseline else None
    ref = ot.process_flat(fl, coeffs, gamma, lam)
    v = ref["valid"]
    adv_ref = np.zeros_like(ref["adv_raw"])
    adv_ref[v] = ot.center_advantages(ref["adv_raw"][v])
    pu = PolicyUpdate([S, 32, 32, 6], device="cuda:0")
    d = {k: torch.tensor(a, device="cuda") for k, a in fl.items()}
    out = pu.process(d["obs"], d["rew"], d["done"], baseline_coeffs=coeffs, discount=gamma, gae_lambda=lam)
    assert np.array_equal(out["valid"].cpu().numpy().astype(bool), v)      # bit-exact mask
    assert np.max(np.abs(out["ret"].cpu().numpy() - ref["ret"])) <= 1e-4
    assert np.max(np.abs(out["adv"].cpu().numpy() - adv_ref)) <= 1e-4
    st = out["stats"].cpu().numpy()
    assert st[0] == v.sum() and abs(st[4] - ref["adv_raw"][v].mean()) <= 1e-9 * max(1, abs(st[4])) + 1e-9
    assert v.sum() < v.size            # the case contains unfinished paths
    pu.close()


def test_fit_baseline_matches_oracle_predictions():
    from oracle import trpo as ot, rollout as orl
    from me_trpo_b200.trpo import PolicyUpdate
    S = 11
    fl = _flat_case(T=60, B=64, S=S, seed=9, T_max=20)
    ref = ot.process_flat(fl, None, 0.99, 1.0)
    flat = dict(obs=fl["obs"], rew=fl["rew"], done=fl["done"], act=np.zeros((60, 64, 3), np.float32),
                mean=np.zeros((60, 64, 3), np.float32))
    paths = orl.paths_from_flat(flat, np.zeros(3, np.float32))
    bl = ot.LinearFeatureBaselineOracle()
    ot.process_samples(paths, bl, 0.99, 1.0)       # fits bl on the returns
    pu = PolicyUpdate([S, 32, 32, 3], device="cuda:0")
    d = {k: torch.tensor(a, device="cuda") for k, a in fl.items()}
    out = pu.process(d["obs"], d["rew"], d["done"], discount=0.99)
    coeffs = pu.fit_baseline(d["obs"], out["ret"], out["valid"], d["done"]).cpu().numpy()
    # compare predictions (the coefficient vector itself is ill-conditioned)
    pred_ref = np.concatenate([bl.predict(p) for p in paths])
    bl2 = ot.LinearFeatureBaselineOracle(); bl2.coeffs = coeffs
    pred_dev = np.concatenate([bl2.predict(p) for p in paths])
    scale = np.abs(pred_ref).max()
    assert np.max(np.abs(pred_dev - pred_ref)) <= 1e-4 * scale
    # and the next iteration's advantages computed with the device coefficients
    out2 = pu.process(d["obs"], d["rew"], d["done"], baseline_coeffs=coeffs, discount=0.99)
    ref2 = ot.process_flat(fl, bl.coeffs, 0.99, 1.0)
    v = ref2["valid"]
    adv_ref = np.zeros_like(ref2["adv_raw"]); adv_ref[v] = ot.center_advantages(ref2["adv_raw"][v])
    assert np.max(np.abs(out2["adv"].cpu().numpy() - adv_ref)) <= 2e-4
    pu.close()


def test_trpo_iteration_on_rollout_buffers_full_size():
    """End to end at BASELINE size on the device: fused rollout -> process -> update, checked by
    size-independent properties (the oracle cannot run 4.1 M samples in seconds)."""
    from me_trpo_b200 import synthetic
    from me_trpo_b200.rollout import EnsembleRollout
    from me_trpo_b200.trpo import PolicyUpdate
    K, B, T = 5, 4096, 200
    spec, models, pol, norm, init, pool = synthetic.make_problem("half-cheetah", K, B, hidden=1024)
    ro = EnsembleRollout("half-cheetah", K, B, T, hidden=1024, device="cuda:0")
    ro.set_dynamics_ensemble(models); ro.set_normalization(**norm)
    ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    out = ro.run(T, init, pool, seed=3)
    ro.synchronize()
    dims = [spec["S"], 32, 32, spec["A"]]
    pu = PolicyUpdate(dims, device="cuda:0")
    pr = pu.process(out["obs"], out["rew"], out["done"], discount=1.0)
    assert int(pr["valid"].sum().item()) == T * B            # every row times out at T: all paths complete
    adv = pr["adv"]
    assert abs(float(adv.double().mean())) < 1e-5 and abs(float(adv.double().std(unbiased=False)) - 1) < 1e-4
    # returns: reverse cumulative sum of the rewards (discount 1)
    ret_ref = torch.flip(torch.cumsum(torch.flip(out["rew"].double(), [0]), 0), [0])
    assert float((pr["ret"].double() - ret_ref).abs().max()) < 1e-2 * max(1.0, float(ret_ref.abs().max())) * 1e-2
    parts = []
    for W, b in zip(pol["W"], pol["b"]):
        parts += [W.ravel(), b.ravel()]
    parts.append(pol["log_std"])
    theta0 = np.concatenate(parts).astype(np.float32)
    theta = torch.tensor(theta0, device="cuda")
    ls = torch.tensor(pol["log_std"], device="cuda")
    l0, k0 = pu.loss_kl(theta, out["obs"], out["act"], adv, out["mean"], ls, valid=pr["valid"])
    assert abs(k0) < 1e-6 and abs(l0) < 1e-4                 # old == new: ratio 1, centred advantages
    info = pu.update(theta, out["obs"], out["act"], adv, out["mean"], ls, valid=pr["valid"]).cpu().numpy()
    assert info[4] == 1.0 and info[1] < info[0] and 0 < info[2] <= 0.01
    l1, k1 = pu.loss_kl(theta, out["obs"], out["act"], adv, out["mean"], ls, valid=pr["valid"])
    assert abs(l1 - info[1]) < 1e-6 and abs(k1 - info[2]) < 1e-6
    assert not np.array_equal(theta.cpu().numpy(), theta0)
    # the baseline fit on 0.8 M samples predicts the returns better than a constant
    coeffs = pu.fit_baseline(out["obs"], pr["ret"], pr["valid"], out["done"])
    pr2 = pu.process(out["obs"], out["rew"], out["done"], baseline_coeffs=coeffs, discount=1.0)
    assert float(pr2["stats"][5]) < float(pr["stats"][5])
    pu.close(); ro.close()
