"""Deterministic synthetic inputs shared by tests/golden/make_ref_fixtures.py (which feeds them to
the REFERENCE's code) and by the tests (which feed them to the oracle and to the CUDA path), so
that the committed fixtures only need to hold the reference's OUTPUTS.  NumPy's legacy
RandomState stream is stable across NumPy versions.  Test infrastructure, not product code."""
import numpy as np

# name -> (reference env module, class, S, A, params json, get_env spelling)
ENVS = {
    "swimmer": ("com_swimmer_env", "SwimmerEnv", 10, 2, "params-swimmer.json"),
    "half-cheetah": ("com_half_cheetah_env", "HalfCheetahEnv", 18, 6, "params-half-cheetah.json"),
    "hopper": ("com_hopper_env", "HopperEnv", 11, 3, "params-hopper.json"),
    "ant": ("com_ant_env", "AntEnv", 29, 8, "params-ant.json"),
    "humanoid": ("com_simple_humanoid_env", "SimpleHumanoidEnv", 55, 21, "params-humanoid.json"),
    "snake": ("com_snake_env", "SnakeEnv", 14, 4, "params-snake.json"),
}
DROP = {"swimmer": 2, "half-cheetah": 1, "hopper": 0, "ant": 2, "humanoid": 0, "snake": 2}
SAM_MODES = ["step_rand", "eps_rand", "model_mean_std", "model_mean", "model_med", "one_model"]


def _xavier(rs, shape):
    fi, fo = (shape[0], shape[0]) if len(shape) == 1 else shape
    lim = np.sqrt(6.0 / (fi + fo))
    return rs.uniform(-lim, lim, size=shape).astype(np.float32)


def dynamics_weights(seed, S, A, drop, hidden, K, out_scale=1.0):
    """K models [W0,b0,W1,b1,W2,b2]: Xavier-uniform W and b (training.py:179,187-194)."""
    rs = np.random.RandomState(seed)
    din = S + A - drop
    dims = [din] + list(hidden) + [S]
    models = []
    for _ in range(K):
        m = {}
        for i in range(len(dims) - 1):
            sc = np.float32(out_scale if i == len(dims) - 2 else 1.0)
            m["W%d" % i] = _xavier(rs, (dims[i], dims[i + 1])) * sc
            m["b%d" % i] = _xavier(rs, (dims[i + 1],)) * sc
        models.append(m)
    return models


def policy_weights(seed, S, hidden, A, bias_scale=0.1):
    """Mean-network weights; non-zero biases so that the bias path is exercised."""
    rs = np.random.RandomState(seed)
    dims = [S] + list(hidden) + [A]
    W = [_xavier(rs, (dims[i], dims[i + 1])) for i in range(len(dims) - 1)]
    b = [(rs.uniform(-1, 1, size=dims[i + 1]) * bias_scale).astype(np.float32)
         for i in range(len(dims) - 1)]
    log_std = (rs.uniform(-1.0, 0.2, size=A)).astype(np.float32)
    return dict(W=W, b=b, log_std=log_std)


def rms_data(seed, S, A, n=400):
    """Transitions used to drive RunningMeanStd.update: xu[n,S+A], diff[n,S]."""
    rs = np.random.RandomState(seed)
    xu = (rs.randn(n, S + A) * rs.uniform(0.02, 2.0, size=S + A) + rs.randn(S + A) * 0.3)
    diff = (rs.randn(n, S) * rs.uniform(0.01, 0.5, size=S) + rs.randn(S) * 0.05)
    return xu.astype(np.float32), diff.astype(np.float32)


def states(seed, n, S, scale=0.3):
    return (np.random.RandomState(seed).randn(n, S) * scale).astype(np.float32)


def actions(seed, shape, scale=1.2):
    """float64 like rllab's get_actions output; some entries beyond [-1, 1]."""
    return np.random.RandomState(seed).randn(*shape) * scale
