"""The fp32 fidelity mode of the rollout path (cfg.precision = METRPO_PREC_FP32,
csrc/rollout_fp32.cuh) against the fp32 oracle (= the reference's arithmetic, pinned by
tests/test_ref_fixtures.py), and what it is for: measuring ON THE DEVICE, at the benchmarked size
and the full horizon, what the bf16 tensor-core operands of the fast path cost.

Tolerances: fp32 mode vs fp32 oracle 2e-5 (only the summation order of the dot products differs);
done flags exact."""
import json
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from oracle import rollout as orl  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "rollout_golden.npz"))
TOL = 2e-5


def _run(case, precision, inp=None, **kw):
    from me_trpo_b200.rollout import EnsembleRollout
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    inp = inp or mg.make_inputs(env, K, B, T, hidden)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode, precision=precision)
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    rk = dict(seed=1234, offset=7)
    if noise_kind == "explicit":
        rk.update(eps=inp["eps"], model_idx=inp["mi"], std_noise=inp["sn"] if sam_mode == "model_mean_std" else None)
    rk.update(kw)
    out = ro.run(T, inp["init"], inp["pool"], **rk)
    ro.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    kern = ro.last_kernel()
    ro.close()
    return res, inp, kern


@pytest.mark.parametrize("case", mg.CASES, ids=[c[0] for c in mg.CASES])
def test_fp32_mode_matches_fp32_golden(case):
    """All six sam_modes, five envs, resets and Ant's early termination: fp32 golden vectors."""
    dev, _, kern = _run(case, "fp32")
    assert kern == 3
    for k in ("obs", "act", "mean", "rew", "final_states"):
        ref = GOLD["%s/fp32/%s" % (case[0], k)]
        assert np.max(np.abs(dev[k] - ref)) <= TOL, (case[0], k, float(np.max(np.abs(dev[k] - ref))))
    assert np.array_equal(dev["done"], GOLD["%s/fp32/done" % case[0]])


def test_fp32_mode_humanoid_and_odd_width():
    """The fidelity path has no tile-shape constraints: humanoid dims, hidden = 200, K = 3."""
    case = ("h", "humanoid", 3, 70, 4, 3, 200, "step_rand", "explicit")
    dev, inp, _ = _run(case, "fp32")
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"],
                           orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"]), T, T_max, sam_mode, mma="fp32")
    for k in ("obs", "act", "mean", "rew", "final_states"):
        assert np.max(np.abs(dev[k] - ref[k])) <= TOL, k
    assert np.array_equal(dev["done"], ref["done"])


def test_fp32_mode_chained_and_stepwise_equal_fused():
    from me_trpo_b200.rollout import EnsembleRollout
    case = mg.CASES[1]
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    one, inp, _ = _run(case, "fp32")
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode, precision="fp32")
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    host, _ = ro.run_to_host(T, inp["init"], inp["pool"], seed=1234, offset=7, n_chunks=3)
    ro.synchronize()
    for k in ("obs", "act", "rew", "done", "final_states"):
        np.testing.assert_array_equal(host[k].numpy(), one[k])
    ro.close()


@pytest.mark.parametrize("env,K,T", [("half-cheetah", 3, 12), ("ant", 2, 8)])
def test_fp32_mode_model_costs_match_oracle(env, K, T):
    from me_trpo_b200.rollout import EnsembleRollout
    inp = mg.make_inputs(env, K, 96, T, 256)
    ro = EnsembleRollout(env, K, 96, T, hidden=256, precision="fp32")
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    costs, rows = ro.model_costs(T, inp["init"], gamma=0.97, return_rows=True)
    ro.synchronize()
    ref, ref_rows = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], inp["init"], T, gamma=0.97,
                                    return_rows=True)
    assert np.max(np.abs(rows.cpu().numpy() - ref_rows)) <= 5e-5 * T
    assert np.max(np.abs(costs.cpu().numpy() - ref)) <= 5e-5 * T
    ro.close()


def test_bf16_vs_fp32_on_device_full_size_full_horizon():
    """The benchmarked workload (half-cheetah, 5 models, 4096 rows, H = 1024, horizon 1000, Philox):
    the fast bf16 path against the fp32 fidelity path, same noise, open loop.  The CPU oracle
    cannot run this size; the device fp32 path can (a few seconds).  Writes the drift table."""
    from me_trpo_b200 import synthetic
    from me_trpo_b200.rollout import EnsembleRollout
    env, K, B, hidden, T = "half-cheetah", 5, 4096, 1024, 1000
    spec, models, pol, norm, init, pool = synthetic.make_problem(env, K, B, hidden=hidden, seed=0)
    outs = {}
    for prec in ("bf16", "fp32"):
        ro = EnsembleRollout(env, K, B, T, hidden=hidden, precision=prec)
        ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ro.run(2, init, pool, seed=1); ro.synchronize()
        e0.record()
        out = ro.run(T, init, pool, seed=1, want=("obs", "rew", "done"))
        e1.record(); ro.synchronize()
        outs[prec] = dict(obs=out["obs"], rew=out["rew"], done=out["done"], ms=e0.elapsed_time(e1))
        ro.close()
    a, b = outs["bf16"], outs["fp32"]
    assert torch.equal(a["done"], b["done"])
    d = (a["obs"] - b["obs"]).abs()
    scale = float(b["obs"].abs().max())
    table = {}
    for t in (1, 10, 100, 999):
        table[str(t)] = dict(max=float(d[t].max()), median=float(d[t].flatten().median()))
    ret_a, ret_b = a["rew"].double().sum(0), b["rew"].double().sum(0)
    res = dict(workload="half-cheetah K=5 B=4096 H=1024 T=1000 philox (bench shape)", state_scale=scale,
               abs_state_diff_bf16_vs_fp32=table,
               return_mean_fp32=float(ret_b.mean()), return_abs_diff_max=float((ret_a - ret_b).abs().max()),
               return_abs_diff_median=float((ret_a - ret_b).abs().median()), return_std=float(ret_b.std()),
               ms_bf16=a["ms"], ms_fp32=b["ms"], slowdown_fp32=b["ms"] / a["ms"])
    print("device bf16 vs device fp32:", json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "drift_table.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old["device_bf16_vs_device_fp32_full_size"] = res
    json.dump(old, open(path, "w"), indent=1)
    assert table["1"]["max"] <= 1e-3 and table["10"]["max"] <= 5e-3 * max(1.0, scale)
    assert table["999"]["median"] <= 5e-2 * max(1.0, scale)
    assert res["return_abs_diff_median"] <= 0.05 * max(1.0, res["return_std"])
