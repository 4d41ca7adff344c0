"""Generates tests/golden/rollout_golden.npz from the NumPy oracle (oracle/).

The reference (TF 1.4 + rllab + MuJoCo) cannot be imported in this environment and ships no
golden vectors, so these fixtures pin the ORACLE: they catch accidental
changes of the restatement and give the GPU parity tests fixed targets.  Inputs are regenerated
from seeds (numpy RandomState is stable across versions); only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import envs as oe, models as om, rollout as orl  # noqa: E402

CASES = [
    # name, env, K, B, T, T_max, hidden, sam_mode, noise
    ("hc_step_rand", "half-cheetah", 3, 8, 6, 4, 256, "step_rand", "explicit"),
    ("hc_philox", "half-cheetah", 3, 8, 6, 4, 256, "step_rand", "philox"),
    ("hc_eps_rand", "half-cheetah", 3, 8, 6, 4, 256, "eps_rand", "philox"),
    ("hc_mean", "half-cheetah", 3, 8, 5, 100, 256, "model_mean", "explicit"),
    ("hc_med", "half-cheetah", 4, 8, 5, 100, 256, "model_med", "explicit"),
    ("hc_mean_std", "half-cheetah", 3, 8, 5, 100, 256, "model_mean_std", "explicit"),
    ("hc_one_model", "half-cheetah", 2, 8, 5, 100, 256, "one_model", "explicit"),
    ("swimmer", "swimmer", 5, 8, 6, 5, 512, "step_rand", "explicit"),
    ("hopper", "hopper", 2, 8, 5, 3, 256, "step_rand", "explicit"),
    ("ant", "ant", 2, 8, 6, 100, 256, "step_rand", "explicit"),
    ("snake", "snake", 2, 8, 5, 4, 256, "step_rand", "explicit"),
]


def make_inputs(env, K, B, T, hidden, seed=0):
    """Deterministic synthetic problem (also used by the GPU parity tests)."""
    spec = oe.ENV_SPECS[env]
    S, A, drop = spec["S"], spec["A"], spec["drop"]
    rng = np.random.RandomState(seed)
    models = om.init_dynamics(rng, S, A, drop, hidden, K)
    pol = om.init_policy(rng, S, spec["policy_hidden"], A)
    pol["b"] = [rng.uniform(-0.1, 0.1, size=b.shape).astype(np.float32) for b in pol["b"]]
    pol["log_std"] = rng.uniform(-0.5, 0.1, size=A).astype(np.float32)
    norm = dict(in_mean=rng.normal(0, 0.1, S + A).astype(np.float32),
                in_std=rng.uniform(0.5, 1.5, S + A).astype(np.float32),
                diff_mean=rng.normal(0, 0.01, S).astype(np.float32),
                diff_std=rng.uniform(0.1, 0.2, S).astype(np.float32))
    init = rng.normal(0, 0.1, (B, S)).astype(np.float32)
    pool = rng.normal(0, 0.1, (2 * B + 3, S)).astype(np.float32)
    if env == "ant":   # keep most rows alive (done = z outside [0.2, 1.0])
        init[:, 2] = 0.6
        pool[:, 2] = 0.6
        init[0, 2] = 0.19   # one row that terminates early
    eps = rng.normal(size=(T, B, A)).astype(np.float32)
    mi = rng.randint(K, size=(T, B)).astype(np.int32)
    sn = rng.normal(size=(T, B, S)).astype(np.float32)
    return dict(spec=spec, models=models, pol=pol, norm=norm, init=init, pool=pool, eps=eps, mi=mi, sn=sn)


def run_case(case, mma):
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    inp = make_inputs(env, K, B, T, hidden)
    noise = (orl.PhiloxNoise(1234, 7, 0, sam_mode) if noise_kind == "philox"
             else orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"]))
    return orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T,
                            T_max, sam_mode, mma=mma)


def main():
    out = {}
    for case in CASES:
        for mma in ("fp32", "bf16"):
            res = run_case(case, mma)
            for k, v in res.items():
                out["%s/%s/%s" % (case[0], mma, k)] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rollout_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
