"""Dev tool: time N fit iterations at the BASELINE shape (for ncu launch lists)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from me_trpo_b200.dynamics import EnsembleFit
from me_trpo_b200 import synthetic as syn
S, A, drop, H, K, batch, n = 18, 6, 1, 1024, 5, 1000, 200000
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
rng = np.random.RandomState(0)
models = syn.init_dynamics(rng, S, A, drop, H, K, out_scale=1.0)
fit = EnsembleFit(S, A, drop, H, K, max_rows=8192, precision=prec)
fit.set_ensemble(models); fit.set_normalization(**syn.default_norm(S, A)); fit.reset_adam()
xd = torch.randn(n, S + A, device="cuda"); yd = xd[:, :S] + 0.1 * torch.randn(n, S, device="cuda")
for j in range(3):
    fit.step(xd, yd, batch, 1e-3, seed=1, offset=j, want_losses=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for j in range(iters):
    fit.step(xd, yd, batch, 1e-3, seed=1, offset=5 + j, want_losses=False)
e1.record(); torch.cuda.synchronize()
print("fit step %s: %.3f ms/iter" % (prec, e0.elapsed_time(e1) / iters))
