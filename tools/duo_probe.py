"""GPU dev probe: single-stream vs two-stream rollout kernel over a few shapes (ms per launch,
cycles per tile-step at the measured SM clock is left to the reader: 1 ms = ~1.93 M cycles)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.rollout import EnsembleRollout

CASES = [tuple(c.split(":")) for c in (sys.argv[1] if len(sys.argv) > 1 else
         "half-cheetah:5:4096,half-cheetah:10:4096,hopper:5:4096,hopper:10:4096,half-cheetah:5:8192").split(",")]
T = int(os.environ.get("PROBE_T", "300"))
res = []
for env, K, B in CASES:
    K, B = int(K), int(B)
    spec, models, pol, norm, init, pool = synthetic.make_problem(env, K, B, hidden=1024)
    for mode in (0, 1, 2):
        os.environ["METRPO_DUO"] = str(mode)
        ro = EnsembleRollout(env, K, B, T, hidden=1024)
        ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
        out = ro.run(T, init, pool, seed=1); ro.synchronize()
        ms = []
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ro.run(T, init, pool, seed=1, out=out); e1.record(); ro.synchronize()
            ms.append(e0.elapsed_time(e1))
        r = dict(env=env, K=K, B=B, T=T, mode=mode, kernel=ro.last_kernel(), ms=min(ms),
                 us_per_tile_step=min(ms) * 1e3 / (T * ((B + 127) // 128) * K) * 148)
        res.append(r); print(json.dumps(r), flush=True)
        ro.close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "duo_probe.json"), "w"), indent=1)
