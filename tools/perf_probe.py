"""GPU dev probe: time the fused rollout at a given config (CUDA events)."""
import os, sys, json, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200.rollout import EnsembleRollout
from oracle import models as om, envs as oe

def main(env="half-cheetah", K=5, B=4096, T=1000, hidden=1024, reps=3):
    spec = oe.ENV_SPECS[env]
    S, A, drop = spec["S"], spec["A"], spec["drop"]
    rng = np.random.RandomState(0)
    models = om.init_dynamics(rng, S, A, drop, hidden, K)
    pol = om.init_policy(rng, S, spec["policy_hidden"], A)
    norm = om.default_norm(S, A)
    init = rng.normal(0, 0.1, (B, S)).astype(np.float32)
    pool = rng.normal(0, 0.1, (B, S)).astype(np.float32)
    ro = EnsembleRollout(env, K, B, T, hidden=hidden)
    ro.set_dynamics_ensemble(models)
    ro.set_normalization(**norm)
    ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    init_d = torch.tensor(init).cuda(); pool_d = torch.tensor(pool).cuda()
    out = None
    times = []
    for i in range(reps + 1):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ro.run(T, init_d, pool_d, seed=1, offset=i * T, out=out)
        e1.record()
        ro.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    units = K * B * T
    din = S + A - drop
    fl = units * 2.0 * (din * hidden + hidden * hidden + hidden * S)
    print(json.dumps(dict(env=env, K=K, B=B, T=T, hidden=hidden, ms=times, best_ms=ms,
                          munits_per_s=units / ms / 1e3, tflops=fl / ms / 1e9,
                          us_per_step=ms * 1e3 / T, finite=bool(torch.isfinite(out["obs"]).all().item()),
                          rew_mean=float(out["rew"].mean().item()))), flush=True)

if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = v if k == "env" else int(v)
    main(**kw)
