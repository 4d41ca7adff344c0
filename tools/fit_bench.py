"""GPU dev probe: ensemble-fit step timing at the BASELINE shape (half-cheetah, K = 5, H = 1024, batch 1000).
usage: fit_bench.py [iters] [precision]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200.dynamics import EnsembleFit
from oracle import models as om
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tf32", "fp32"]
S, A, drop, H, K, batch, n = 18, 6, 1, 1024, 5, 1000, 200000
rng = np.random.RandomState(0)
models = om.init_dynamics(rng, S, A, drop, H, K, out_scale=1.0)
for prec in precs:
    fit = EnsembleFit(S, A, drop, H, K, max_rows=8192, precision=prec)
    fit.set_ensemble(models); fit.set_normalization(**om.default_norm(S, A)); fit.reset_adam()
    xd = torch.randn(n, S + A, device="cuda"); yd = xd[:, :S] + 0.1 * torch.randn(n, S, device="cuda")
    for j in range(3):
        fit.step(xd, yd, batch, 1e-3, seed=1, offset=j, want_losses=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(iters):
        fit.step(xd, yd, batch, 1e-3, seed=1, offset=5 + j, want_losses=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 6.0 * batch * (23 * H + H * H + H * S) * K
    print("fit step %s: %.3f ms/iter, %.1f TFLOP/s (3x fwd flops), %.2f M samples/s, launches %d" % (
        prec, ms, flop / ms / 1e9, K * batch / ms / 1e3, fit.last_launches()))
    fit.close()
