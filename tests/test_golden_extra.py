"""Golden fixtures for R12 (per-model cost), R10-R11 (TRPO half) and N3 (ensemble fit):
tests/golden/extra_golden.npz, written by tests/golden/make_golden_extra.py from the oracle.

CPU: the oracle still reproduces the committed vectors (pins the restatement against drift).
GPU: the CUDA path, through the C ABI, against the same vectors (tolerances as in the per-path
test files: per-model cost 2e-4*T vs the bf16-arithmetic vectors; TRPO loss / KL 2e-5 abs, gradient /
FVP 2e-4 rel, update step within 5e-3 of the oracle's step; fit (fp32 GEMMs) losses 1e-5 rel,
weights 2e-5 abs)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_extra as mx  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "extra_golden.npz"))


def test_oracle_reproduces_extra_golden():
    out = mx.compute()
    assert sorted(out.keys()) == sorted(GOLD.files)
    for k, v in out.items():
        np.testing.assert_allclose(np.asarray(v, np.float64), np.asarray(GOLD[k], np.float64), rtol=1e-6, atol=1e-7, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("case", mx.MC_CASES, ids=[c[0] for c in mx.MC_CASES])
def test_model_costs_match_golden(case):
    from me_trpo_b200.rollout import EnsembleRollout
    name, env, K, n, T, hidden, gamma = case
    inp, init = mx.mc_inputs(case)
    ro = EnsembleRollout(env, K, n, 1000, hidden=hidden)
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    costs, rows = ro.model_costs(T, init, gamma, return_rows=True)
    ro.synchronize()
    assert np.max(np.abs(rows.cpu().numpy() - GOLD["mc/%s/bf16/rows" % name])) <= 2e-4 * T
    assert np.max(np.abs(costs.cpu().numpy() - GOLD["mc/%s/bf16/costs" % name])) <= 1e-4 * T
    assert np.max(np.abs(costs.cpu().numpy() - GOLD["mc/%s/fp32/costs" % name])) <= 2e-3 * T
    ro.close()


@pytest.mark.gpu
def test_trpo_matches_golden():
    torch = pytest.importorskip("torch")
    from me_trpo_b200.trpo import PolicyUpdate
    pr = mx.trpo_problem()
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    pu = PolicyUpdate(mx.TRPO_DIMS, device="cuda:0")
    args = (dev(pr["obs"]), dev(pr["act"]), dev(pr["adv"]), dev(pr["mean"]), dev(pr["pol"]["log_std"]))
    l, k = pu.loss_kl(dev(pr["theta_new"]), *args)
    assert abs(l - GOLD["trpo/loss_kl"][0]) <= 2e-5 and abs(k - GOLD["trpo/loss_kl"][1]) <= 2e-5
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(pu.grad(dev(pr["theta_new"]), *args), GOLD["trpo/grad"]) <= 2e-4
    assert rel(pu.grad(dev(pr["theta"]), *args, vec=dev(pr["v"]), reg_coeff=1e-5), GOLD["trpo/hvp"]) <= 2e-4
    th = dev(pr["theta"]).clone()
    info = pu.update(th, *args).cpu().numpy()
    step_ref = GOLD["trpo/theta_after_update"] - pr["theta"]
    step_dev = th.cpu().numpy().astype(np.float64) - pr["theta"]
    assert GOLD["trpo/update_info"][3] == 1.0 and info[4] == 1.0
    assert np.linalg.norm(step_dev - step_ref) <= 5e-3 * np.linalg.norm(step_ref)
    assert abs(info[2] - GOLD["trpo/update_info"][2]) <= 2e-5          # mean KL after the step
    # flat sample processing
    fl, coeffs = mx.flat_case()
    pp = PolicyUpdate([18, 32, 32, 6], device="cuda:0")
    out = pp.process(dev(fl["obs"]), dev(fl["rew"]), dev(fl["done"]), baseline_coeffs=coeffs, discount=0.99, gae_lambda=0.97)
    assert np.array_equal(out["valid"].cpu().numpy(), GOLD["proc/valid"])
    assert np.max(np.abs(out["adv"].cpu().numpy() - GOLD["proc/adv_centered"])) <= 1e-4
    assert np.max(np.abs(out["ret"].cpu().numpy() - GOLD["proc/ret"])) <= 1e-4 * max(1.0, np.abs(GOLD["proc/ret"]).max())
    pu.close(); pp.close()


@pytest.mark.gpu
def test_fit_matches_golden():
    torch = pytest.importorskip("torch")
    from me_trpo_b200.dynamics import EnsembleFit
    c = mx.FIT
    models, norm, x, y, idx = mx.fit_problem()
    fit = EnsembleFit(c["S"], c["A"], c["drop"], c["H"], c["K"], max_rows=128, precision="fp32")
    fit.set_ensemble(models); fit.set_normalization(**norm); fit.reset_adam()
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    for j in range(c["steps"]):
        l = fit.step(xd, yd, c["batch"], 1e-3, idx=idx[j]).cpu().numpy()
        np.testing.assert_allclose(l, GOLD["fit/losses"][j], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fit.eval(xd, yd)[0].cpu().numpy(), GOLD["fit/val_after"], rtol=2e-5)
    for k in range(c["K"]):
        w = fit.get_weights(k)
        for key in w:
            assert np.max(np.abs(w[key].cpu().numpy() - GOLD["fit/model%d/%s" % (k, key)])) <= 2e-5, key
    fit.close()
