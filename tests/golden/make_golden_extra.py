"""Generates tests/golden/extra_golden.npz from the NumPy / torch-float64 oracle (oracle/): known
answers for the per-model validation cost (R12), the TRPO half (R10-R11: flat sample processing,
loss / KL, gradient, Fisher-vector product, one whole update) and the ensemble fit (N3: losses
and weights after 4 Adam steps).  Like rollout_golden.npz these pin the ORACLE (the reference ships
no vectors for this path and cannot run here); the GPU tests compare the CUDA path with them.

    python tests/golden/make_golden_extra.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402
from oracle import fit as of, models as om, rollout as orl, trpo as ot  # noqa: E402

MC_CASES = [("hc", "half-cheetah", 3, 64, 6, 256, 0.98), ("ant", "ant", 2, 48, 5, 256, 1.0)]
TRPO_DIMS = [18, 32, 32, 6]
FIT = dict(S=11, A=3, drop=0, H=64, K=2, n=300, batch=50, steps=4)


def mc_inputs(case):
    name, env, K, n, T, hidden, gamma = case
    inp = mg.make_inputs(env, K, n, 1, hidden)
    init = inp["init"].copy()
    if env == "ant":
        init[:, 2] = 0.5
        init[::5, 2] = 3.0
    return inp, init


def trpo_problem(N=400, seed=11):
    rng = np.random.RandomState(seed)
    S, A = TRPO_DIMS[0], TRPO_DIMS[-1]
    pol = om.init_policy(rng, S, TRPO_DIMS[1:-1], A)
    pol["b"] = [rng.uniform(-0.1, 0.1, b.shape).astype(np.float32) for b in pol["b"]]
    pol["log_std"] = rng.uniform(-0.4, 0.1, A).astype(np.float32)
    obs = rng.normal(0, 1, (N, S)).astype(np.float32)
    mean = om.policy_forward(pol, obs).astype(np.float32)
    act = (mean + np.exp(pol["log_std"]) * rng.normal(size=(N, A))).astype(np.float32)
    adv = rng.normal(size=N).astype(np.float32)
    adv = ((adv - adv.mean()) / (adv.std() + 1e-8)).astype(np.float32)
    theta = ot.flatten_params(pol)
    theta_new = (theta + rng.normal(0, 0.02, theta.shape)).astype(np.float32)
    v = rng.normal(size=theta.shape).astype(np.float32)
    return dict(pol=pol, obs=obs, act=act, adv=adv, mean=mean, theta=theta.astype(np.float32), theta_new=theta_new, v=v)


def flat_case(T=24, B=12, S=18, seed=4, T_max=9):
    rng = np.random.RandomState(seed)
    obs = rng.normal(0, 2.0, (T, B, S)).astype(np.float32)
    rew = rng.normal(-1, 1, (T, B)).astype(np.float32)
    done = rng.rand(T, B) < 0.05
    ts = np.zeros(B, int)
    for t in range(T):
        ts += 1
        done[t] |= ts >= T_max
        ts[done[t]] = 0
    coeffs = rng.normal(0, 0.05, 2 * S + 4)
    return dict(obs=obs, rew=rew, done=done.astype(np.uint8)), coeffs


def fit_problem(seed=21):
    c = FIT
    rng = np.random.RandomState(seed)
    models = om.init_dynamics(rng, c["S"], c["A"], c["drop"], c["H"], c["K"], out_scale=1.0)
    norm = dict(in_mean=rng.normal(0, 0.2, c["S"] + c["A"]).astype(np.float32),
                in_std=rng.uniform(0.5, 1.5, c["S"] + c["A"]).astype(np.float32),
                diff_mean=rng.normal(0, 0.05, c["S"]).astype(np.float32),
                diff_std=rng.uniform(0.1, 0.5, c["S"]).astype(np.float32))
    x = rng.normal(0, 1, (c["n"], c["S"] + c["A"])).astype(np.float32)
    y = (x[:, :c["S"]] + rng.normal(0, 0.1, (c["n"], c["S"]))).astype(np.float32)
    idx = rng.randint(0, c["n"], (c["steps"], c["batch"] * c["K"])).astype(np.int32)
    return models, norm, x, y, idx


def compute():
    out = {}
    for case in MC_CASES:
        inp, init = mc_inputs(case)
        name, env, K, n, T, hidden, gamma = case
        for mma in ("fp32", "bf16"):
            c, r = orl.model_costs(env, inp["pol"], inp["models"], inp["norm"], init, T, gamma, mma=mma, return_rows=True)
            out["mc/%s/%s/costs" % (name, mma)] = c
            out["mc/%s/%s/rows" % (name, mma)] = r
    pr = trpo_problem()
    N = len(pr["adv"])
    orc = ot.TRPOOracle(TRPO_DIMS)
    inputs = (pr["obs"], pr["act"], pr["adv"], pr["mean"], np.tile(pr["pol"]["log_std"], (N, 1)))
    l, k = orc.loss_kl(pr["theta_new"], inputs)
    out["trpo/loss_kl"] = np.array([l, k])
    out["trpo/grad"] = orc.grad(pr["theta_new"], inputs)
    out["trpo/hvp"] = orc.hvp(pr["theta"], inputs, pr["v"])
    new, info = orc.optimize(pr["theta"], inputs)
    out["trpo/theta_after_update"] = new
    out["trpo/update_info"] = np.array([info["loss_before"], info["loss_after"], info["kl"], float(info["accepted"])])
    fl, coeffs = flat_case()
    ref = ot.process_flat(fl, coeffs, 0.99, 0.97)
    out["proc/adv_centered"] = np.where(ref["valid"], (ref["adv_raw"] - ref["adv_raw"][ref["valid"]].mean())
                                        / (ref["adv_raw"][ref["valid"]].std() + 1e-8), 0.0)
    out["proc/ret"] = ref["ret"]
    out["proc/valid"] = ref["valid"].astype(np.uint8)
    models, norm, x, y, idx = fit_problem()
    c = FIT
    adam = of.Adam(models)
    losses = [of.train_step(models, adam, norm, x, y, idx[j], c["batch"], 1e-3, c["S"], c["drop"]) for j in range(c["steps"])]
    out["fit/losses"] = np.asarray(losses)
    out["fit/val_after"] = of.validation_losses(models, norm, x, y, c["S"], c["drop"])
    for k_, m in enumerate(models):
        for key, v_ in m.items():
            out["fit/model%d/%s" % (k_, key)] = v_
    return out


def main():
    out = compute()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "extra_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
