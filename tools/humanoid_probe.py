"""GPU dev probe: humanoid (wide instantiation) per-GPU share at a short horizon."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.rollout import EnsembleRollout
K, B, T = 20, 8192, int(os.environ.get("PROBE_T", "100"))
spec, models, pol, norm, init, pool = synthetic.make_problem("humanoid", K, B, hidden=1024)
ro = EnsembleRollout("humanoid", K, B, T, hidden=1024)
ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
out = ro.run(T, init, pool, seed=1); ro.synchronize()
ms = []
for i in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ro.run(T, init, pool, seed=1, out=out); e1.record(); ro.synchronize()
    ms.append(e0.elapsed_time(e1))
print(json.dumps(dict(env="humanoid", K=K, B=B, T=T, ms=min(ms), ms_per_step=min(ms) / T)))
