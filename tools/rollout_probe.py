"""GPU dev probe: fused rollout vs the NumPy oracle on small configs (prints max errors)."""
import os, sys, json, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200.rollout import EnsembleRollout
from oracle import models as om, rollout as orl, envs as oe

def run_case(env, K, B, T, T_max, hidden, sam_mode="step_rand", seed=0, philox=False, teacher=True):
    spec = oe.ENV_SPECS[env]
    S, A, drop = spec["S"], spec["A"], spec["drop"]
    rng = np.random.RandomState(seed)
    models = om.init_dynamics(rng, S, A, drop, hidden, K)
    pol = om.init_policy(rng, S, spec["policy_hidden"], A)
    pol["b"] = [rng.uniform(-0.1, 0.1, size=b.shape).astype(np.float32) for b in pol["b"]]
    pol["log_std"] = rng.uniform(-0.5, 0.1, size=A).astype(np.float32)
    norm = om.default_norm(S, A)
    norm["in_mean"] = rng.normal(0, 0.1, S + A).astype(np.float32)
    norm["in_std"] = rng.uniform(0.5, 1.5, S + A).astype(np.float32)
    norm["diff_mean"] = rng.normal(0, 0.01, S).astype(np.float32)
    norm["diff_std"] = rng.uniform(0.1, 0.2, S).astype(np.float32)
    init = rng.normal(0, 0.1, (B, S)).astype(np.float32)
    if env == "ant":
        init[:, 2] = 0.6
    pool = rng.normal(0, 0.1, (2 * B + 3, S)).astype(np.float32)
    if env == "ant":
        pool[:, 2] = 0.6
    if philox:
        noise = orl.PhiloxNoise(1234, 7, 0, sam_mode)
        eps = mi = sn = None
    else:
        eps = rng.normal(size=(T, B, A)).astype(np.float32)
        mi = rng.randint(K, size=(T, B)).astype(np.int32)
        sn = rng.normal(size=(T, B, S)).astype(np.float32)
        noise = orl.ExplicitNoise(eps, mi, sn)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode)
    ro.set_dynamics_ensemble(models)
    ro.set_normalization(**norm)
    ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    t0 = time.time()
    out = ro.run(T, init, pool, eps=eps, model_idx=mi, std_noise=sn if sam_mode == "model_mean_std" else None,
                 seed=1234, offset=7)
    ro.synchronize()
    dt = time.time() - t0
    dev = {k: v.cpu().numpy() for k, v in out.items()}
    res = dict(env=env, K=K, B=B, T=T, hidden=hidden, sam_mode=sam_mode, philox=philox, wall_s=round(dt, 4))
    # open-loop vs oracle (bf16-emulating and fp32)
    for mma in ("bf16", "fp32"):
        ref = orl.rollout_flat(env, pol, models, norm, init, pool, noise, T, T_max, sam_mode, mma=mma)
        for key in ("obs", "act", "mean", "rew", "final_states"):
            res["open_%s_%s" % (mma, key)] = float(np.max(np.abs(dev[key] - ref[key])))
        res["open_%s_done_mismatch" % mma] = int(np.sum(dev["done"] != ref["done"]))
    if teacher:
        # teacher-forced: feed the DEVICE's own pre-step observations to the oracle, compare one-step results
        ref = orl.rollout_flat(env, pol, models, norm, init, pool, noise, T, T_max, sam_mode, mma="bf16",
                               teacher_states=dev["obs"])
        res["tf_bf16_rew"] = float(np.max(np.abs(dev["rew"] - ref["rew"])))
        res["tf_bf16_act"] = float(np.max(np.abs(dev["act"] - ref["act"])))
        nxt = np.concatenate([dev["obs"][1:], dev["final_states"][None]], 0)
        res["tf_bf16_next"] = float(np.max(np.abs(nxt - np.concatenate([ref["obs"][1:], ref["final_states"][None]], 0))))
        ref = orl.rollout_flat(env, pol, models, norm, init, pool, noise, T, T_max, sam_mode, mma="fp32",
                               teacher_states=dev["obs"])
        res["tf_fp32_rew"] = float(np.max(np.abs(dev["rew"] - ref["rew"])))
    res["nan"] = int(np.isnan(dev["obs"]).sum())
    ro.close()
    print(json.dumps(res), flush=True)
    return res

if __name__ == "__main__":
    cases = [
        dict(env="half-cheetah", K=1, B=128, T=3, T_max=100, hidden=256),
        dict(env="half-cheetah", K=2, B=128, T=4, T_max=100, hidden=256),
        dict(env="half-cheetah", K=5, B=200, T=6, T_max=4, hidden=512),
        dict(env="half-cheetah", K=5, B=300, T=5, T_max=100, hidden=1024),
        dict(env="half-cheetah", K=5, B=300, T=5, T_max=100, hidden=1024, philox=True),
        dict(env="hopper", K=3, B=130, T=4, T_max=3, hidden=256),
        dict(env="swimmer", K=5, B=100, T=6, T_max=5, hidden=512),
        dict(env="ant", K=4, B=256, T=5, T_max=100, hidden=256),
        dict(env="half-cheetah", K=3, B=128, T=3, T_max=100, hidden=256, sam_mode="model_mean"),
        dict(env="half-cheetah", K=4, B=128, T=3, T_max=100, hidden=256, sam_mode="model_med"),
        dict(env="half-cheetah", K=3, B=128, T=3, T_max=100, hidden=256, sam_mode="model_mean_std"),
        dict(env="half-cheetah", K=5, B=128, T=3, T_max=100, hidden=256),   # 12
        dict(env="half-cheetah", K=1, B=128, T=3, T_max=100, hidden=512),   # 13
        dict(env="half-cheetah", K=1, B=256, T=3, T_max=100, hidden=256),   # 14
        # more tiles than gang slots -> split chains
        dict(env="half-cheetah", K=5, B=4096, T=12, T_max=100, hidden=256),
    ]
    only = int(sys.argv[1]) if len(sys.argv) > 1 else None
    allres = []
    for i, c in enumerate(cases):
        if only is not None and i != only:
            continue
        try:
            allres.append(run_case(**c))
        except RuntimeError as ex:
            print("CASE %d FAILED: %s" % (i, ex), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allres, open(os.path.join(ROOT, "gpurun_out", "rollout_probe.json"), "w"), indent=1)
