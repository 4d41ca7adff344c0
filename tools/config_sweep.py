"""GPU probe: the per-GPU share of every BASELINE.json config through the fused rollout
(T = 1000, Philox noise, synthetic nets): time per launch, units/s, achieved TFLOP/s, finiteness and
launch-to-launch determinism.  Writes gpurun_out/config_sweep.json."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.envs import ENV_SPECS
from me_trpo_b200.rollout import EnsembleRollout

CONFIGS = [  # name, env, K, rows on ONE GPU, T, hidden, gpus in BASELINE
    ("swimmer (reference JSON shape)", "swimmer", 5, 100, 200, 512, 1),
    ("half-cheetah 1xB200", "half-cheetah", 5, 4096, 1000, 1024, 1),
    ("hopper 1xB200", "hopper", 10, 4096, 1000, 1024, 1),
    ("ant 4xB200 (4096 of 16384 rows)", "ant", 20, 4096, 1000, 1024, 4),
    ("humanoid 8xB200 (8192 of 65536 rows)", "humanoid", 20, 8192, 1000, 1024, 8),
]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("bf16_tflops", 1590.0))
res = []
for name, env, K, B, T, hidden, gpus in CONFIGS:
    spec, models, pol, norm, init, pool = synthetic.make_problem(env, K, B, hidden=hidden)
    ro = EnsembleRollout(env, K, B, T, hidden=hidden)
    ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    want = ("obs", "rew", "done", "act", "mean")
    out = ro.run(T, init, pool, seed=1, want=want); ro.synchronize()
    ref = {k: v.clone() for k, v in out.items()}
    ms = []
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ro.run(T, init, pool, seed=1, out=out, want=want); e1.record(); ro.synchronize()
        ms.append(e0.elapsed_time(e1))
    same = all(torch.equal(ref[k], out[k]) for k in ref)
    finite = bool(torch.isfinite(out["obs"]).all().item() and torch.isfinite(out["rew"]).all().item())
    din = spec["S"] + spec["A"] - spec["drop"]
    f_dyn = 2.0 * (din * hidden + hidden * hidden + hidden * spec["S"])
    dims = [spec["S"]] + list(spec["policy_hidden"]) + [spec["A"]]
    f_pol = 2.0 * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    flops = K * B * T * f_dyn + B * T * f_pol
    t = min(ms)
    r = dict(config=name, env=env, K=K, rows_per_gpu=B, T=T, hidden=hidden, baseline_gpus=gpus, ms_per_launch=t,
             units_per_s_per_gpu=K * B * T / t * 1e3, tflops=flops / t / 1e9, frac_of_bf16_peak=flops / t / 1e9 / peak,
             finite=finite, deterministic=same, done_count=int(out["done"].sum().item()))
    res.append(r); print(json.dumps(r), flush=True)
    ro.close()
    del out, ref
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
