"""Dev probe: tf32 vs oracle weight deviation statistics + fit step timing at the BASELINE shape."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import fit as of
import test_fit_gpu as tg

for dims in [(18, 6, 1, 256), (11, 3, 0, 128), (29, 8, 2, 64), (55, 21, 0, 96)]:
    S, A, drop, H = dims
    K, batch, steps = 3, 200, 5
    models, norm, x, y = tg._problem(1, S, A, drop, H, K)
    for prec in ("fp32", "tf32"):
        fit = tg._fit(models, norm, S, A, drop, H, prec)
        xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
        ref = [{k: v.copy() for k, v in m.items()} for m in models]
        adam = of.Adam(ref)
        rng = np.random.RandomState(7)
        lerr = 0
        for j in range(steps):
            idx = rng.randint(0, len(x), batch * K)
            l_dev = fit.step(xd, yd, batch, 1e-3, idx=idx).cpu().numpy()
            l_ref = of.train_step(ref, adam, norm, x, y, idx, batch, 1e-3, S, drop)
            lerr = max(lerr, np.max(np.abs(l_dev - l_ref) / np.abs(l_ref)))
        for key in ("W0", "W1", "W2", "b1"):
            w = fit.get_weights(0)[key].cpu().numpy()
            d = w - ref[0][key]; u = ref[0][key] - models[0][key]
            print(dims, prec, key, "maxabs %.2e rms %.2e rms_update %.2e ratio %.3f frac>1e-4 %.4f lerr %.1e" % (
                np.abs(d).max(), np.sqrt((d ** 2).mean()), np.sqrt((u ** 2).mean()),
                np.sqrt((d ** 2).mean()) / np.sqrt((u ** 2).mean()), (np.abs(d) > 1e-4).mean(), lerr))
        vl = fit.eval(xd, yd)[0].cpu().numpy(); vr = of.validation_losses(ref, norm, x, y, S, drop)
        print("   val loss rel err", np.max(np.abs(vl - vr) / vr))
        fit.close()

# timing at the BASELINE shape: half-cheetah, K=5, H=1024, batch 1000
from me_trpo_b200.dynamics import EnsembleFit
from oracle import models as om
S, A, drop, H, K, batch, n = 18, 6, 1, 1024, 5, 1000, 200000
rng = np.random.RandomState(0)
models = om.init_dynamics(rng, S, A, drop, H, K, out_scale=1.0)
for prec in ("tf32", "fp32"):
    fit = EnsembleFit(S, A, drop, H, K, max_rows=8192, precision=prec)
    fit.set_ensemble(models); fit.set_normalization(**om.default_norm(S, A)); fit.reset_adam()
    xd = torch.randn(n, S + A, device="cuda"); yd = xd[:, :S] + 0.1 * torch.randn(n, S, device="cuda")
    for j in range(5):
        fit.step(xd, yd, batch, 1e-3, seed=1, offset=j, want_losses=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(50):
        fit.step(xd, yd, batch, 1e-3, seed=1, offset=5 + j, want_losses=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    flop = 6.0 * batch * (23 * H + H * H + H * S) * K
    print("fit step %s: %.3f ms/iter, %.1f TFLOP/s (3x fwd flops), %.2f M samples/s" % (prec, ms, flop / ms / 1e9, K * batch / ms / 1e3))
    e0.record()
    l, _ = fit.eval(xd[:100000], yd[:100000])
    e1.record(); torch.cuda.synchronize()
    print("   eval 100k rows x K: %.2f ms" % e0.elapsed_time(e1))
    fit.close()
