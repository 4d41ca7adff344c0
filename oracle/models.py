"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's model
functions on the imaginary-rollout path.

  dynamics_forward   training.py:218-269 (dynamics_model) + :125-169 (prepare_input, column drop)
                     + :171-214 (build_ff_neural_net; y = x @ W + b, W[in,out])
  policy_forward     training.py:96-117 (policy_model over the rllab mean-network layers)
  RunningMeanStd     running_mean_std.py:3-42
  xavier_uniform     tf.contrib.layers.xavier_initializer() as used for W *and* b
                     (training.py:179,187-194)

PINNED (round 2): dynamics_forward / policy_forward / RunningMeanStd are compared with the
reference's own closures and class (training.py, running_mean_std.py executed under shims) in
tests/test_ref_fixtures.py sections B and C.
rllab's GaussianMLPPolicy is NOT vendored in the reference (README.md:7) -> its semantics here
follow SURVEY.md Appendix A.1 ("parity unpinned" for that part).

`mma` selects how matrix products are evaluated:
  "fp32"  operands and accumulation in `dtype` (the reference's tf.matmul semantics)
  "bf16"  operands (activations AND weights) rounded to bfloat16 (round-to-nearest-even), products
          accumulated in float64 and rounded to fp32 -- the arithmetic the tensor-core kernel
          implements (fp32 accumulate; bias / normalisation / residual in fp32).
"""
import numpy as np


def bf16_round(x):
    """Round fp32 -> bf16 (RNE) and return as fp32."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).reshape(a.shape)
    # NaN stays NaN
    return np.where(np.isnan(a), a, out)


def xavier_uniform(rng, shape):
    """TF xavier_initializer(uniform=True): U(+-sqrt(6/(fan_in+fan_out))); for a 1-D shape (n,)
    TF uses fan_in = fan_out = n."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        fan_in, fan_out = shape[0], shape[1]
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_dynamics(rng, S, A, drop, hidden, K, out_scale=0.1):
    """K independent dynamics MLPs Din -> hidden -> hidden -> S (params 'hidden_layers').
    out_scale multiplies the last layer so that long open-loop rollouts of a random net stay
    finite; 1.0 reproduces the reference initialiser exactly."""
    din = S + A - drop
    models = []
    for _ in range(K):
        m = dict(
            W0=xavier_uniform(rng, (din, hidden)), b0=xavier_uniform(rng, (hidden,)),
            W1=xavier_uniform(rng, (hidden, hidden)), b1=xavier_uniform(rng, (hidden,)),
            W2=xavier_uniform(rng, (hidden, S)) * np.float32(out_scale),
            b2=xavier_uniform(rng, (S,)) * np.float32(out_scale))
        models.append(m)
    return models


def init_policy(rng, S, hidden, A, init_std=1.0):
    """rllab GaussianMLPPolicy init: Xavier-uniform W, zero b, log_std = log(init_std)."""
    dims = [S] + list(hidden) + [A]
    W = [xavier_uniform(rng, (dims[i], dims[i + 1])) for i in range(len(dims) - 1)]
    b = [np.zeros(dims[i + 1], np.float32) for i in range(len(dims) - 1)]
    return dict(W=W, b=b, log_std=np.full(A, np.log(init_std), np.float32))


def default_norm(S, A):
    """mu_in = 0, sigma_in = 1, mu_delta = 0, sigma_delta = 0.1 (= the RMS floor, :25)."""
    return dict(in_mean=np.zeros(S + A, np.float32), in_std=np.ones(S + A, np.float32),
                diff_mean=np.zeros(S, np.float32), diff_std=np.full(S, 0.1, np.float32))


_BF16_W_CACHE = {}


def _bf16_weight(W):
    """bf16-rounded float64 copy of a weight matrix, memoised per array object (the rollout calls
    this K x T times with the same matrices; rounding 1 M elements each time dominated the run)."""
    key = id(W)
    hit = _BF16_W_CACHE.get(key)
    if hit is None or hit[0] is not W:
        if len(_BF16_W_CACHE) > 256:
            _BF16_W_CACHE.clear()
        hit = (W, bf16_round(W).astype(np.float64))
        _BF16_W_CACHE[key] = hit
    return hit[1]


def _matmul(h, W, dtype, mma):
    if mma == "bf16":
        return (bf16_round(h).astype(np.float64) @ _bf16_weight(W)).astype(np.float32)
    return h.astype(dtype) @ W.astype(dtype)


def dynamics_forward(m, norm, xu, S, drop, dtype=np.float32, mma="fp32"):
    """next_state[B,S] of one model for xu = concat([state, action], 1)  (training.py:218-269)."""
    xu = xu.astype(dtype)
    if mma == "bf16":   # the kernel multiplies by the fp32 reciprocal of in_std (1 ulp from the quotient)
        z = (xu - norm["in_mean"].astype(dtype)) * (np.float32(1.0) / norm["in_std"].astype(np.float32))
    else:
        z = (xu - norm["in_mean"].astype(dtype)) / norm["in_std"].astype(dtype)   # :228
    z = z[:, drop:]                                                           # :146-154
    b0 = m["b0"].astype(dtype)
    if mma == "bf16":
        # the kernel folds b0 into the layer-0 MMA as bf16 hi + lo parts (two constant-1 columns)
        hi = bf16_round(m["b0"])
        b0 = hi + bf16_round(m["b0"] - hi)
    h = np.maximum(_matmul(z, m["W0"], dtype, mma) + b0, 0)                     # :207-208 relu
    h = np.maximum(_matmul(h, m["W1"], dtype, mma) + m["b1"].astype(dtype), 0)
    o = _matmul(h, m["W2"], dtype, mma) + m["b2"].astype(dtype)               # identity (:166)
    # tf.add(diff_rms.mean + diff_rms.std * nn_output, x)   (:257)
    return (norm["diff_mean"].astype(dtype) + norm["diff_std"].astype(dtype) * o) + xu[:, :S]


def ensemble_forward(models, norm, xu, S, drop, dtype=np.float32, mma="fp32"):
    """[K,B,S]: all K models on the same input (env_helpers.py:612-616)."""
    return np.stack([dynamics_forward(m, norm, xu, S, drop, dtype, mma) for m in models])


def policy_forward(pol, x, dtype=np.float32, out_tanh=False):
    """mean[B,A]: tanh hidden layers, identity (or tanh) output (training.py:99-103)."""
    h = x.astype(dtype)
    n = len(pol["W"])
    for i in range(n):
        h = h @ pol["W"][i].astype(dtype) + pol["b"][i].astype(dtype)
        if i < n - 1 or out_tanh:
            h = np.tanh(h)
    return h


class RunningMeanStd:
    """running_mean_std.py:3-42: cumulative sum / sumsq / count, all initialised so that an empty
    tracker has mean 0, std = sqrt(max(1 - 0, 1e-2)) = 1; std floor sqrt(1e-2) = 0.1 (:22-27)."""

    def __init__(self, epsilon=1e-2, shape=()):
        self._sum = np.zeros(shape, np.float32)
        self._sumsq = np.full(shape, epsilon, np.float32)
        self._count = np.float32(epsilon)

    @property
    def mean(self):
        return (self._sum / self._count).astype(np.float32)

    @property
    def std(self):
        return np.sqrt(np.maximum((self._sumsq / self._count).astype(np.float32)
                                  - np.square(self.mean), np.float32(1e-2)))

    def update(self, x):
        x = np.asarray(x, np.float32)
        self._sum = self._sum + np.sum(x, axis=0)
        self._sumsq = self._sumsq + np.sum(np.square(x), axis=0)
        self._count = np.float32(self._count + len(x))
