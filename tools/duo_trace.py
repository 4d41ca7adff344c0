"""GPU dev probe: event trace of one CTA of the two-stream rollout kernel (needs the -DMETRPO_TRACE
build: METRPO_LIB=me_trpo_b200/libmetrpo_trace.so).  Prints the MMA warp's and both epilogue groups'
events (code@cycles since the first event (+delta))."""
import os, sys, json, ctypes
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.rollout import EnsembleRollout
from me_trpo_b200 import lib as L

env, K, B = "half-cheetah", 5, 4096
if len(sys.argv) > 1:
    env, K, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
T = int(os.environ.get("PROBE_T", "16"))
cta = int(os.environ.get("PROBE_CTA", "7"))
spec, models, pol, norm, init, pool = synthetic.make_problem(env, K, B, hidden=1024)
ro = EnsembleRollout(env, K, B, T, hidden=1024)
ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
ro.run(T, init, pool, seed=1); ro.synchronize()
lib = L.load()
L.check(lib.metrpo_rollout_set_trace(ro._h, cta, 6, 9), "set_trace")
ro.run(T, init, pool, seed=1); ro.synchronize()
print("kernel variant", ro.last_kernel())
buf = np.zeros(4 * 4096, np.uint64)
L.check(lib.metrpo_rollout_get_trace(ro._h, buf.ctypes.data_as(ctypes.c_void_p)), "get_trace")
buf = buf.reshape(4, 4096)
out = {}
for role, name in enumerate(["producer", "mma", "epi0", "epi1"]):
    ev = buf[role][buf[role] != 0]
    codes = (ev >> np.uint64(40)).astype(np.int64); clk = (ev & np.uint64(0xFFFFFFFFFF)).astype(np.int64)
    out[name] = [(int(c), int(t)) for c, t in zip(codes, clk)]
t00 = min(v[0][1] for v in out.values() if v)
for name, evs in out.items():
    print("==", name, len(evs))
    prev = None
    line = []
    for c, t in evs[:900]:
        d = 0 if prev is None else t - prev
        line.append("%x@%d(+%d)" % (c, t - t00, d)); prev = t
        if len(line) == 8:
            print(" ".join(line)); line = []
    if line: print(" ".join(line))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "duo_trace_%s_%d.json" % (env, K)), "w"))
