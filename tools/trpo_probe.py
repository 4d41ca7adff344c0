"""GPU dev probe: wall time (CUDA events) of the TRPO half at BASELINE size:
process_samples, baseline fit and the full natural-gradient update on T x B samples."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.rollout import EnsembleRollout
from me_trpo_b200.trpo import PolicyUpdate

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K, B = 5, 4096
spec, models, pol, norm, init, pool = synthetic.make_problem("half-cheetah", K, B, hidden=1024)
ro = EnsembleRollout("half-cheetah", K, B, T, hidden=1024, device="cuda:0")
ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
out = ro.run(T, init, pool, seed=3); ro.synchronize()
pu = PolicyUpdate([spec["S"], 32, 32, spec["A"]], device="cuda:0")
IMPL = sys.argv[2] if len(sys.argv) > 2 else "auto"
pu.set_pass_impl(IMPL)
parts = []
for W, b in zip(pol["W"], pol["b"]):
    parts += [W.ravel(), b.ravel()]
parts.append(pol["log_std"])
theta0 = torch.tensor(np.concatenate(parts).astype(np.float32), device="cuda")
ls = torch.tensor(pol["log_std"], device="cuda")

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r

t_proc, pr = timed(lambda: pu.process(out["obs"], out["rew"], out["done"], discount=1.0))
t_fit, coeffs = timed(lambda: pu.fit_baseline(out["obs"], pr["ret"], pr["valid"], out["done"]))
t_proc2, pr2 = timed(lambda: pu.process(out["obs"], out["rew"], out["done"], baseline_coeffs=coeffs, discount=1.0))
def upd():
    th = theta0.clone()
    return pu.update(th, out["obs"], out["act"], pr2["adv"], out["mean"], ls, valid=pr2["valid"])
t_upd, info = timed(upd)
th = theta0.clone()
t_grad, _ = timed(lambda: pu.grad(th, out["obs"], out["act"], pr2["adv"], out["mean"], ls, valid=pr2["valid"]))
v = torch.randn_like(th)
t_fvp, _ = timed(lambda: pu.grad(th, out["obs"], out["act"], pr2["adv"], out["mean"], ls, valid=pr2["valid"], vec=v))
t_loss, _ = timed(lambda: pu.loss_kl(th, out["obs"], out["act"], pr2["adv"], out["mean"], ls, valid=pr2["valid"]))
N = T * B
bytes_pass = N * 4 * (spec["S"] + 2 * spec["A"] + 1) + N
res = dict(impl=IMPL, T=T, B=B, N=N, ms_process=t_proc, ms_process_with_baseline=t_proc2, ms_fit_baseline=t_fit,
           ms_update=t_upd, ms_grad=t_grad, ms_fvp=t_fvp, ms_loss=t_loss,
           pass_GBps=dict(grad=bytes_pass / t_grad / 1e6, fvp=bytes_pass / t_fvp / 1e6, loss=bytes_pass / t_loss / 1e6),
           info=info.cpu().tolist(), launches=pu.last_launches())
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "trpo_probe_%s.json" % IMPL), "w"), indent=1)
