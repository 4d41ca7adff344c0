"""GPU dev probe: event trace of one CTA for a couple of steps -> per-chunk timing table."""
import os, sys, json, ctypes
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200.rollout import EnsembleRollout
from me_trpo_b200 import lib as L
from oracle import models as om, envs as oe

env, K, B, T, hidden = "half-cheetah", 5, 4096, 12, 1024
if len(sys.argv) > 1:
    env, K, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec = oe.ENV_SPECS[env]; S, A, drop = spec["S"], spec["A"], spec["drop"]
rng = np.random.RandomState(0)
models = om.init_dynamics(rng, S, A, drop, hidden, K); pol = om.init_policy(rng, S, spec["policy_hidden"], A)
norm = om.default_norm(S, A)
init = rng.normal(0, 0.1, (B, S)).astype(np.float32)
ro = EnsembleRollout(env, K, B, T, hidden=hidden)
ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
ro.run(T, init, init, seed=1); ro.synchronize()
lib = L.load()
L.check(lib.metrpo_rollout_set_trace(ro._h, 7, 5, 7), "set_trace")
ro.run(T, init, init, seed=1); ro.synchronize()
buf = np.zeros(3 * 4096, np.uint64)
L.check(lib.metrpo_rollout_get_trace(ro._h, buf.ctypes.data_as(ctypes.c_void_p)), "get_trace")
buf = buf.reshape(3, 4096)
out = {}
for role, name in enumerate(["producer", "mma", "epilogue"]):
    ev = buf[role][buf[role] != 0]
    codes = (ev >> np.uint64(40)).astype(np.int64); clk = (ev & np.uint64(0xFFFFFFFFFF)).astype(np.int64)
    out[name] = [(int(c), int(t)) for c, t in zip(codes, clk)]
t00 = min(v[0][1] for v in out.values() if v)
for name, evs in out.items():
    print("==", name, len(evs))
    prev = None
    line = []
    for c, t in evs[:700]:
        d = 0 if prev is None else t - prev
        line.append("%x@%d(+%d)" % (c, t - t00, d)); prev = t
        if len(line) == 8:
            print(" ".join(line)); line = []
    if line: print(" ".join(line))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "trace.json"), "w"))
