"""ORACLE (test infrastructure, not product code): float64 restatement of the TRPO half of the hot
path -- sample processing and the natural-gradient policy update.

PARITY: process_samples and the surrogate / mean-KL composition are PINNED against the reference's
own samplers/base.py and algos/npo.py executed under shims (tests/test_ref_fixtures.py section E).
The arithmetic that lives in rllab (un-vendored, unpinned dependency of the reference,
README.md:7,10: DiagonalGaussian, LinearFeatureBaseline, the ConjugateGradientOptimizer's CG /
line search) is restated from SURVEY.md Appendix A.2-A.5 and stays UNPINNED; the reference ships no
test vectors for it.  Call sites in the reference that fix the composition:

  process_samples           samples/base.py:48-182 (non-recurrent branch :74-105,167)
  surrogate / mean_kl       algos/npo.py:33-92  (kl_sym :68, likelihood_ratio_sym :69,
                            mean_kl :74, surr_loss :75, update_opt :85-91)
  optimize                  algos/npo.py:94-111 -> rllab ConjugateGradientOptimizer.optimize
                            with all-default arguments (algos/trpo.py:17-20)
  policy distribution       rllab DiagonalGaussian; mean network = policy_model (training.py:96-117)

The Hessian-vector product is evaluated the way rllab's PerlmutterHvp does -- gradient of
(grad(mean_kl) . v) by reverse-over-reverse automatic differentiation -- using torch.autograd in
float64 on the CPU (independent of the analytic Gauss-Newton form the CUDA kernels use).
Only tests/ and __graft_entry__.smoke() may import this module.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# rllab.misc.special.discount_cumsum, rllab.algos.util.center_advantages  (Appendix A.5)
# ------------------------------------------------------------------------------------------------
def discount_cumsum(x, discount):
    """lfilter([1], [1, -discount], x[::-1])[::-1]: y_t = x_t + discount * y_{t+1}."""
    x = np.asarray(x, np.float64)
    y = np.zeros_like(x)
    run = 0.0
    for t in range(len(x) - 1, -1, -1):
        run = x[t] + discount * run
        y[t] = run
    return y


def center_advantages(adv):
    return (adv - np.mean(adv)) / (adv.std() + 1e-8)


# ------------------------------------------------------------------------------------------------
# rllab LinearFeatureBaseline(reg_coeff=1e-5)  (Appendix A.4; training.py:355-357)
# ------------------------------------------------------------------------------------------------
def baseline_features(obs, L):
    o = np.clip(np.asarray(obs, np.float64), -10, 10)
    al = np.arange(L).reshape(-1, 1) / 100.0
    return np.concatenate([o, o ** 2, al, al ** 2, al ** 3, np.ones((L, 1))], axis=1)


class LinearFeatureBaselineOracle:
    def __init__(self, reg_coeff=1e-5):
        self.coeffs = None
        self.reg_coeff = reg_coeff

    def predict(self, path):
        L = len(path["rewards"])
        if self.coeffs is None:
            return np.zeros(L)
        return baseline_features(path["observations"], L).dot(self.coeffs)

    def fit(self, paths):
        F = np.concatenate([baseline_features(p["observations"], len(p["rewards"])) for p in paths])
        ret = np.concatenate([p["returns"] for p in paths])
        reg = self.reg_coeff
        for _ in range(5):
            self.coeffs = np.linalg.lstsq(F.T.dot(F) + reg * np.identity(F.shape[1]), F.T.dot(ret),
                                          rcond=None)[0]
            if not np.any(np.isnan(self.coeffs)):
                break
            reg *= 10


# ------------------------------------------------------------------------------------------------
# BaseSampler.process_samples  (samplers/base.py:48-182, feed-forward branch)
# ------------------------------------------------------------------------------------------------
def process_samples(paths, baseline, discount, gae_lambda=1.0, center_adv=True, positive_adv=False):
    for p in paths:
        b = np.append(baseline.predict(p), 0.0)                                   # :55-56
        deltas = np.asarray(p["rewards"], np.float64) + discount * b[1:] - b[:-1]  # :57-59
        p["advantages"] = discount_cumsum(deltas, discount * gae_lambda)           # :60-61
        p["returns"] = discount_cumsum(p["rewards"], discount)                     # :62
    cat = lambda k: np.concatenate([np.asarray(p[k]) for p in paths])
    adv = cat("advantages")
    if center_adv:
        adv = center_advantages(adv)                                               # :82-83
    if positive_adv:
        adv = adv - adv.min() + 1e-8                                               # :85-86
    data = dict(observations=cat("observations"), actions=cat("actions"), rewards=cat("rewards"),
                returns=cat("returns"), advantages=adv,
                agent_infos={k: np.concatenate([p["agent_infos"][k] for p in paths])
                             for k in paths[0]["agent_infos"]}, paths=paths)
    baseline.fit(paths)                                                            # :167
    return data


def process_flat(flat, coeffs, discount, gae_lambda=1.0):
    """Same arithmetic on the time-major buffers [T,B] the fused sampler produces, keeping the
    layout: returns dict(adv_raw[T,B], ret[T,B], valid[T,B], base[T,B]); samples that belong to a
    path still open at the end of the buffer are invalid (obtain_samples returns only completed
    paths, samplers/vectorized_sampler.py:80-105)."""
    rew = np.asarray(flat["rew"], np.float64)
    done = np.asarray(flat["done"]).astype(bool)
    T, B = rew.shape
    adv = np.zeros((T, B)); ret = np.zeros((T, B)); base = np.zeros((T, B))
    valid = np.zeros((T, B), bool)
    for b in range(B):
        start = 0
        for t in range(T):
            if done[t, b]:
                L = t + 1 - start
                sl = slice(start, t + 1)
                path = dict(observations=flat["obs"][sl, b], rewards=rew[sl, b])
                bl = np.zeros(L) if coeffs is None else baseline_features(path["observations"], L).dot(coeffs)
                b1 = np.append(bl, 0.0)
                deltas = path["rewards"] + discount * b1[1:] - b1[:-1]
                adv[sl, b] = discount_cumsum(deltas, discount * gae_lambda)
                ret[sl, b] = discount_cumsum(path["rewards"], discount)
                base[sl, b] = bl
                valid[sl, b] = True
                start = t + 1
    return dict(adv_raw=adv, ret=ret, valid=valid, base=base)


# ------------------------------------------------------------------------------------------------
# policy / distribution in torch float64
# ------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def flatten_params(pol):
    """rllab get_params(trainable=True) order: mean-net (W, b per layer) then log_std (A.1)."""
    parts = []
    for W, b in zip(pol["W"], pol["b"]):
        parts += [np.asarray(W, np.float64).ravel(), np.asarray(b, np.float64).ravel()]
    parts.append(np.asarray(pol["log_std"], np.float64).ravel())
    return np.concatenate(parts)


def unflatten_params(theta, dims):
    out_W, out_b, o = [], [], 0
    for i in range(len(dims) - 1):
        n = dims[i] * dims[i + 1]
        out_W.append(theta[o:o + n].reshape(dims[i], dims[i + 1])); o += n
        out_b.append(theta[o:o + dims[i + 1]]); o += dims[i + 1]
    log_std = theta[o:o + dims[-1]]; o += dims[-1]
    assert o == len(theta)
    return out_W, out_b, log_std


class TRPOOracle:
    """surr_loss / mean_kl graph of NPO.init_opt + rllab ConjugateGradientOptimizer (A.2)."""

    def __init__(self, dims, out_tanh=False, step_size=0.01, cg_iters=10, reg_coeff=1e-5,
                 backtrack_ratio=0.8, max_backtracks=15):
        self.dims = list(dims)
        self.out_tanh = out_tanh
        self.max_kl = step_size
        self.cg_iters, self.reg_coeff = cg_iters, reg_coeff
        self.backtrack_ratio, self.max_backtracks = backtrack_ratio, max_backtracks

    # -- graph ----------------------------------------------------------------------------------
    def _dist(self, theta_t, obs_t):
        torch = _torch()
        Ws, bs, log_std = unflatten_params(theta_t, self.dims)
        h = obs_t
        n = len(Ws)
        for i in range(n):
            h = h @ Ws[i] + bs[i]
            if i < n - 1 or self.out_tanh:
                h = torch.tanh(h)
        log_std = torch.clamp(log_std, min=float(np.log(1e-6)))           # min_std (A.1)
        return h, log_std.expand_as(h)

    def _loss_kl(self, theta_t, inp):
        torch = _torch()
        obs, act, adv, old_mean, old_log_std = inp
        mean, log_std = self._dist(theta_t, obs)
        old_std, new_std = torch.exp(old_log_std), torch.exp(log_std)
        # DiagonalGaussian.kl_sym (A.3)
        kl = torch.sum(((old_mean - mean) ** 2 + old_std ** 2 - new_std ** 2) / (2 * new_std ** 2 + 1e-8)
                       + log_std - old_log_std, dim=-1)
        # likelihood_ratio_sym = exp(ll_new - ll_old)
        zn = (act - mean) / new_std
        zo = (act - old_mean) / old_std
        A = act.shape[-1]
        ll_new = -log_std.sum(-1) - 0.5 * (zn ** 2).sum(-1) - 0.5 * A * np.log(2 * np.pi)
        ll_old = -old_log_std.sum(-1) - 0.5 * (zo ** 2).sum(-1) - 0.5 * A * np.log(2 * np.pi)
        lr = torch.exp(ll_new - ll_old)
        return -(lr * adv).mean(), kl.mean()                              # npo.py:74-75

    def _inputs(self, obs, act, adv, old_mean, old_log_std):
        torch = _torch()
        t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
        return tuple(t(a) for a in (obs, act, adv, old_mean, old_log_std))

    def loss_kl(self, theta, inputs):
        torch = _torch()
        with torch.no_grad():
            l, k = self._loss_kl(torch.as_tensor(np.asarray(theta, np.float64)), self._inputs(*inputs))
        return float(l), float(k)

    def grad(self, theta, inputs):
        torch = _torch()
        th = torch.tensor(np.asarray(theta, np.float64), requires_grad=True)
        l, _ = self._loss_kl(th, self._inputs(*inputs))
        (g,) = torch.autograd.grad(l, th)
        return g.numpy()

    def hvp(self, theta, inputs, v):
        """PerlmutterHvp: d/dtheta (grad(mean_kl) . v) + reg_coeff * v."""
        torch = _torch()
        th = torch.tensor(np.asarray(theta, np.float64), requires_grad=True)
        _, k = self._loss_kl(th, self._inputs(*inputs))
        (gk,) = torch.autograd.grad(k, th, create_graph=True)
        vt = torch.as_tensor(np.asarray(v, np.float64))
        (hv,) = torch.autograd.grad((gk * vt).sum(), th)
        return hv.numpy() + self.reg_coeff * np.asarray(v, np.float64)

    # -- rllab.misc.krylov.cg (A.2) ---------------------------------------------------------------
    def cg(self, f_Ax, b, residual_tol=1e-10):
        p = b.copy(); r = b.copy(); x = np.zeros_like(b)
        rdotr = r.dot(r)
        for _ in range(self.cg_iters):
            z = f_Ax(p)
            v = rdotr / p.dot(z)
            x += v * p
            r -= v * z
            newrdotr = r.dot(r)
            mu = newrdotr / rdotr
            p = r + mu * p
            rdotr = newrdotr
            if rdotr < residual_tol:
                break
        return x

    # -- ConjugateGradientOptimizer.optimize (A.2) ---------------------------------------------------
    def optimize(self, theta, inputs):
        """Returns (new_theta, info)."""
        prev = np.asarray(theta, np.float64).copy()
        loss_before, _ = self.loss_kl(prev, inputs)
        g = self.grad(prev, inputs)
        Hx = lambda x: self.hvp(prev, inputs, x)
        d = self.cg(Hx, g)
        step0 = np.sqrt(2.0 * self.max_kl * (1.0 / (d.dot(Hx(d)) + 1e-8)))
        if np.isnan(step0):
            step0 = 1.0
        descent = step0 * d
        n_iter, loss, kl, cur = 0, loss_before, 0.0, prev
        accepted = False
        for n_iter, ratio in enumerate(self.backtrack_ratio ** np.arange(self.max_backtracks)):
            cur = prev - ratio * descent
            loss, kl = self.loss_kl(cur, inputs)
            if loss < loss_before and kl <= self.max_kl:
                accepted = True
                break
        if (np.isnan(loss) or np.isnan(kl) or loss >= loss_before or kl >= self.max_kl):
            cur = prev                                                     # accept_violation=False
            accepted = False
        return cur, dict(loss_before=loss_before, loss_after=loss, kl=kl, backtracks=n_iter,
                         accepted=accepted, grad=g, direction=d, step0=step0)
