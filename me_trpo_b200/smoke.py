"""One small invocation of the hot path on cuda:0, checked against the NumPy oracle
(called by __graft_entry__.smoke())."""
import os
import sys

import numpy as np


def run():
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    import make_golden as mg                     # deterministic synthetic problem
    from oracle import rollout as orl            # the checker (test infrastructure)
    from me_trpo_b200.rollout import EnsembleRollout

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device: the rollout path has no CPU fallback")
    torch.cuda.set_device(0)
    env, K, B, T, T_max, hidden = "half-cheetah", 5, 300, 6, 4, 1024
    inp = mg.make_inputs(env, K, B, T, hidden)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, device="cuda:0")
    ro.set_dynamics_ensemble(inp["models"])
    ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    out = ro.run(T, inp["init"], inp["pool"], eps=inp["eps"], model_idx=inp["mi"])
    ro.synchronize()
    dev = {k: v.cpu().numpy() for k, v in out.items()}
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"],
                           orl.ExplicitNoise(inp["eps"], inp["mi"]), T, T_max, mma="bf16")
    errs = {k: float(np.max(np.abs(dev[k] - ref[k]))) for k in ("obs", "act", "rew", "final_states")}
    ok = all(e < 1e-4 for e in errs.values()) and np.array_equal(dev["done"], ref["done"])
    print("smoke: fused rollout half-cheetah K=%d B=%d T=%d hidden=%d  max|err| vs oracle: %s  done exact: %s"
          % (K, B, T, hidden, errs, np.array_equal(dev["done"], ref["done"])))
    ro.close()
    if not ok:
        raise RuntimeError("smoke: device rollout deviates from the oracle: %s" % errs)
