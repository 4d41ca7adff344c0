"""dev: one small ant run on the two-stream kernel (shared Z slot) vs the single-stream kernel"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg
from me_trpo_b200.rollout import EnsembleRollout
def run(mode, K, B, T, T_max, hidden):
    os.environ["METRPO_DUO"] = str(mode)
    inp = mg.make_inputs("ant", K, B, T, hidden)
    ro = EnsembleRollout("ant", K, B, T_max, hidden=hidden, sam_mode="step_rand")
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    out = ro.run(T, inp["init"], inp["pool"], seed=1234, offset=7)
    ro.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}; kern = ro.last_kernel(); ro.close()
    return res, kern
for (K, B, T, T_max, hidden) in [(4, 600, 7, 100, 512), (20, 1024, 4, 3, 1024)]:
    a, k0 = run(0, K, B, T, T_max, hidden)
    b, k1 = run(1, K, B, T, T_max, hidden)
    print("K", K, "kernels", k0, k1, {k: float(np.max(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)))) for k in a}, flush=True)
