"""Generate tests/golden/ref_fixtures.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_ref_fixtures.py

The reference's TensorFlow-1.4 / rllab imports are satisfied by tests/golden/ref_shims.py (a lazy
NumPy graph evaluator + import stubs; see its docstring for what is the reference's code and what
is restated rllab).  Nothing from the reference is copied into the repository: the files are
imported / AST-extracted where they lie and only their OUTPUTS are stored.

What is executed (reference file:line) and which scope row it pins:
  A  envs/com_*_env.py  cost_np_vec / cost_np / is_done / cost_tf / is_done_tf   R5, R12
  B  running_mean_std.py RunningMeanStd (+ its own test_runningmeanstd, verbatim)  R9
  C  training.py:125-270 prepare_input / build_ff_neural_net / dynamics_model        R7
     training.py:74-118  build_policy_from_rllab -> policy_model                    R8
  D  env_helpers.py:530-635 NeuralNetEnv / VecSimpleEnv (all six sam_modes), behind
     envs/base.py TfEnv / VecTfEnv, fed by the graph of C                           R3, R4, R6
  E  algos/trpo.py TRPO -> algos/npo.py init_opt / optimize_policy,
     algos/batch_polopt.py, samplers/vectorized_sampler.py start_worker /
     obtain_samples, samplers/base.py process_samples -- the TRPO inner iteration
     of model_based_rl.py:1171-1180, two iterations                              R1, R2, R10, R11
  F  model_based_rl.py:106-151 build_policy_graph (per-model validation cost),
     :1339-1371 is_done, :1403-1419 update_stats, utils.py:285-296 stop_critereon   R12
  G  model_based_rl.py:23-103 build_dynamics_graph (fit loss), utils.py:44-131
     data_collection, utils.py:366-369 get_ith_tensor                               N3
"""
import ast
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import ref_inputs as RI          # noqa: E402
import ref_shims                 # noqa: E402

tf = ref_shims.install()
sys.path.insert(0, REF)

OUT = {}
NOTES = {}


def put(key, value, note=None):
    OUT[key] = np.asarray(value)
    if note:
        NOTES[key] = note


# --------------------------------------------------------------------------------------------------
# reference modules
# --------------------------------------------------------------------------------------------------
def load_envs_package():
    """`envs` as shipped does not import (envs/__init__.py -> com_humanoid_env.py:1 imports the
    missing `private_examples`).  Build the package object by hand from the env
    files themselves; HumanoidEnv is the SimpleHumanoidEnv of envs/com_simple_humanoid_env.py."""
    pkg = types.ModuleType("envs")
    pkg.__path__ = [os.path.join(REF, "envs")]
    sys.modules["envs"] = pkg
    classes = {}
    for name, (modname, clsname, S, A, _) in RI.ENVS.items():
        m = importlib.import_module("envs." + modname)
        cls = getattr(m, clsname)
        cls._OBS_DIM, cls._ACT_DIM = S, A
        classes[name] = cls
        setattr(pkg, clsname, cls)
    pkg.HumanoidEnv = classes["humanoid"]
    pkg.__all__ = ["AntEnv", "HalfCheetahEnv", "HopperEnv", "HumanoidEnv", "SnakeEnv", "SwimmerEnv"]
    return classes


def patch_mujoco_env():
    """The real simulator's spaces and reset(): dimensions from SURVEY section 6, reset states
    from a pool handed over by the generator (consumed in call order = the reference's row
    order, env_helpers.py:590-593)."""
    MujocoEnv = sys.modules["rllab.envs.mujoco.mujoco_env"].MujocoEnv
    Box = ref_shims.Box
    MujocoEnv.observation_space = property(lambda self: Box(-np.inf, np.inf, (self._OBS_DIM,)))
    MujocoEnv.action_space = property(lambda self: Box(-2.0, 2.0, (self._ACT_DIM,)))

    def reset(self):
        self.n_reset_calls = getattr(self, "n_reset_calls", 0) + 1
        return np.array(next(self._reset_iter), np.float64)
    MujocoEnv.reset = reset


ENV_CLASSES = load_envs_package()
patch_mujoco_env()
ref_envs_base = importlib.import_module("envs.base")
sys.modules["sandbox.rocky.tf.envs.base"] = types.ModuleType("sandbox.rocky.tf.envs.base")
sys.modules["sandbox.rocky.tf.envs.base"].TfEnv = ref_envs_base.TfEnv
ref_utils = importlib.import_module("utils")
ref_env_helpers = importlib.import_module("env_helpers")
ref_rms = importlib.import_module("running_mean_std")
ref_mbrl = importlib.import_module("model_based_rl")
ref_namedtuples = importlib.import_module("namedtuples")
ref_trpo = importlib.import_module("algos.trpo")

GET_ENV_NAME = {"half-cheetah": "half_cheetah"}


class RecordingRandom:
    """np.random stand-in for the reference modules: same global MT19937 stream, every draw
    recorded so that the oracle / the kernel can be given the identical noise as inputs."""

    def __init__(self):
        self.log = []

    def randint(self, *a, **k):
        v = np.random.randint(*a, **k)
        self.log.append(("randint", np.array(v)))
        return v

    def normal(self, *a, **k):
        v = np.random.normal(*a, **k)
        self.log.append(("normal", np.array(v)))
        return v

    def __getattr__(self, name):
        return getattr(np.random, name)


class _Cast(dict):
    """np.cast[...] of NumPy 1.12 (the reference's pin, tf14.yml:61); removed in NumPy 2."""

    def __missing__(self, key):
        return lambda x: np.asarray(x, dtype=key)


class NumpyProxy:
    def __init__(self, rnd):
        self.random = rnd
        self.cast = _Cast()

    def __getattr__(self, name):
        return getattr(np, name)


def extract_train_closures(params, env, policy_opt_params, S, A):
    """Compile the closures nested in training.py train() (74-283) in a namespace holding the
    free variables train() would have bound."""
    src = open(os.path.join(REF, "training.py")).read()
    tree = ast.parse(src)
    wanted = ["build_policy_from_rllab", "get_value", "prepare_input", "build_ff_neural_net",
              "build_dynamics_model", "get_regularizer_loss"]
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in wanted and node.name not in found:
            found[node.name] = node
    ns = dict(tf=tf, layers=tf.contrib.layers, np=np, params=params, env=env,
              policy_opt_params=policy_opt_params, n_states=S, n_actions=A, n_goals=0,
              get_scope_variable=ref_utils.get_scope_variable,
              variable_summaries=lambda *a, **k: None, data_summaries=lambda *a, **k: None)
    for name in wanted:
        mod = ast.Module(body=[found[name]], type_ignores=[])
        exec(compile(mod, os.path.join(REF, "training.py"), "exec"), ns)
    return ns


def make_params(env_name, hidden=None, K=None, T=None, sam_mode=None):
    with open(os.path.join(REF, "params", RI.ENVS[env_name][4])) as f:
        params = json.load(f)
    if hidden is not None:
        params["dynamics_model"]["hidden_layers"] = list(hidden)
    if K is not None:
        params["n_models"] = K
    if T is not None:
        params["policy_opt_params"]["T"] = T
    if sam_mode is not None:
        params["policy_opt_params"]["sam_mode"] = sam_mode
    return params


def policy_opt_namedtuple(params):
    """training.py:55-61."""
    pop = dict(params["policy_opt_params"])
    sc = pop["stop_critereon"]
    pop["stop_critereon"] = ref_utils.stop_critereon(
        threshold=sc["threshold"], offset=sc["offset"],
        percent_models_threshold=sc["percent_models_threshold"])
    return ref_namedtuples.Policy_opt_params(**pop)


class World:
    """One reference graph: env (get_env), policy (build_policy_from_rllab), RMS normalisers,
    dynamics_model closures, dynamics_in / dynamics_outs, with weights from ref_inputs."""

    def __init__(self, env_name, hidden, K, seed, out_scale=1.0, T=None, sam_mode=None):
        tf.reset_default_graph()
        tf.set_random_seed(seed)
        self.env_name = env_name
        _, _, S, A, _ = RI.ENVS[env_name]
        self.S, self.A, self.K = S, A, K
        self.params = make_params(env_name, hidden, K, T, sam_mode)
        self.policy_opt_params = policy_opt_namedtuple(self.params)
        self.sess = tf.Session()
        self.sess.__enter__()
        # training.py:295 (get_env asserts for 'half-cheetah', quirk 7 -> use its own spelling)
        self.env = ref_env_helpers.get_env(GET_ENV_NAME.get(env_name, env_name))
        self.inner_env = self.env._wrapped_env._wrapped_env          # training.py:300
        assert self.inner_env.observation_space.shape[0] == S
        ns = extract_train_closures(self.params, self.env, self.policy_opt_params, S, A)
        self.ns = ns
        self.training_policy, self.policy_model = ns["build_policy_from_rllab"]()   # :298
        with tf.variable_scope("input_rms"):                                        # :320-323
            self.input_rms = ref_rms.RunningMeanStd(epsilon=0.0, shape=(S + A))
        with tf.variable_scope("diff_rms"):
            self.diff_rms = ref_rms.RunningMeanStd(epsilon=0.0, shape=(S))
        self.dynamics_model = ns["build_dynamics_model"](
            n_states=S, n_actions=A, n_goals=0, dt=None, input_rms=self.input_rms,
            diff_rms=self.diff_rms)                                                 # :324-330
        self.get_regularizer_loss = ns["get_regularizer_loss"]
        # model_based_rl.py:262-300 (placeholders + build_dynamics_graph)
        self.dynamics_in = tf.placeholder(tf.float32, shape=[None, S + A], name="dynamics_in")
        self.dynamics_in_full = tf.placeholder(tf.float32, shape=[None, K * (S + A)])
        self.y_training_full = tf.placeholder(tf.float32, shape=[None, K * S])

        class _Log:
            def info(self, *a, **k):
                pass
            debug = info
        self.logger = _Log()
        (self.dynamics_loss, self.prediction_loss, self.regularizer_loss, self.dynamics_outs,
         self.dynamics_losses) = ref_mbrl.build_dynamics_graph(
            "training_dynamics", self.dynamics_model, self.dynamics_in, self.dynamics_in_full,
            self.y_training_full, S + A, K, self.get_regularizer_loss, S, self.logger)
        # weights
        drop = RI.DROP[env_name]
        self.models = RI.dynamics_weights(seed, S, A, drop, hidden, K, out_scale)
        for i, m in enumerate(self.models):
            for j in range(len(hidden) + 1):
                w = ref_utils.get_scope_variable("training_dynamics", "model%d/layer%d/weights" % (i, j))
                b = ref_utils.get_scope_variable("training_dynamics", "model%d/layer%d/biases" % (i, j))
                assert w.value.shape == m["W%d" % j].shape, (w.value.shape, m["W%d" % j].shape)
                w.load(m["W%d" % j])
                b.load(m["b%d" % j])
        self.pol = RI.policy_weights(seed + 1, S, self.params["policy"]["hidden_layers"], A)
        layers = self.training_policy._mean_network.layers[1:]
        for l, W, b in zip(layers, self.pol["W"], self.pol["b"]):
            l.W.load(W)
            l.b.load(b)
        self.training_policy._l_std_param.param.load(self.pol["log_std"])
        # normalisers through the reference's own update()
        xu, diff = RI.rms_data(seed + 2, S, A)
        self.input_rms.update(xu[:150])
        self.diff_rms.update(diff[:150])
        self.rms_first = self.sess.run([self.input_rms.mean, self.input_rms.std,
                                        self.diff_rms.mean, self.diff_rms.std])
        self.input_rms.update(xu[150:])
        self.diff_rms.update(diff[150:])
        self.rms = self.sess.run([self.input_rms.mean, self.input_rms.std,
                                  self.diff_rms.mean, self.diff_rms.std])
        self.policy_in = tf.placeholder(tf.float32, shape=[None, S], name="policy_in")
        self.policy_out = self.policy_model(self.policy_in)

    def close(self):
        self.sess.__exit__(None, None, None)

    def set_reset_pool(self, pool):
        self.inner_env._reset_iter = iter(pool)
        self.inner_env.n_reset_calls = 0

    def neural_net_env(self, sam_mode):
        """model_based_rl.py:373-380."""
        cost_np_vec = self.inner_env.cost_np_vec
        return ref_envs_base.TfEnv(ref_env_helpers.NeuralNetEnv(
            env=self.env, inner_env=self.inner_env, cost_np=cost_np_vec,
            dynamics_in=self.dynamics_in, dynamics_outs=self.dynamics_outs, sam_mode=sam_mode))


# --------------------------------------------------------------------------------------------------
# A. cost / done functions
# --------------------------------------------------------------------------------------------------
def craft_next_states(env_name, rs, B, S):
    xn = rs.randn(B, S)
    if env_name == "half-cheetah":
        xn[:, 9] *= 8.0                                   # reward clip at +-10 active for some rows
    if env_name == "hopper":
        xn[:, 0] = 0.45 + rs.randn(B) * 0.3
        xn[:, 1] = rs.randn(B) * 0.3
        big = rs.rand(B, S) < 0.1
        xn = np.where(big, xn * 300.0, xn)
    if env_name == "ant":
        xn[:, 2] = rs.uniform(0.0, 1.2, size=B)
        xn[0, 2], xn[1, 2] = 0.2, 1.0                     # boundaries are NOT done
        xn[2, 5] = np.nan
        xn[3, 7] = np.inf
        xn[4, 2] = np.nan
    if env_name == "humanoid":
        xn[:, -1] = 1.5 + rs.randn(B) * 0.5
    return xn.astype(np.float32)


def gen_costs():
    tf.reset_default_graph()
    with tf.Session() as sess:
        for env_name, (_, _, S, A, _) in RI.ENVS.items():
            inner = ENV_CLASSES[env_name]()
            rs = np.random.RandomState(100 + S)
            B = 48
            x = (rs.randn(B, S) * 0.5).astype(np.float32)
            u = np.clip(rs.randn(B, A) * 0.8, -1, 1)                   # f64, clipped (:599)
            xn = craft_next_states(env_name, rs, B, S)
            k = "A_costs__%s__" % env_name
            put(k + "x", x); put(k + "u", u); put(k + "x_next", xn)
            with np.errstate(invalid="ignore"):
                put(k + "cost_np_vec", inner.cost_np_vec(x, u, xn))
                put(k + "cost_np", inner.cost_np(x, u, xn))
            # NeuralNetEnv's is_done default (env_helpers.py:537)
            is_done = getattr(inner, "is_done", lambda x, y: np.asarray([False] * len(x)))
            put(k + "is_done", is_done(x, xn))
            xp = tf.placeholder(tf.float32, [None, S])
            up = tf.placeholder(tf.float32, [None, A])
            xnp = tf.placeholder(tf.float32, [None, S])
            feed = {xp: x, up: u, xnp: xn}
            with np.errstate(invalid="ignore"):
                if hasattr(inner, "is_done_tf"):
                    d_tf = inner.is_done_tf(xp, xnp)
                    put(k + "is_done_tf", sess.run(d_tf, feed))
                    dones = (rs.rand(B) < 0.3).astype(np.float32)
                    dp = tf.placeholder(tf.float32, [None])
                    feed[dp] = dones
                    put(k + "dones_in", dones)
                    finite = np.isfinite(xn).all(axis=1)
                    f2 = {xp: x[finite], up: u[finite], xnp: xn[finite], dp: dones[finite]}
                    put(k + "cost_tf", sess.run(inner.cost_tf(xp, up, xnp, dp), f2))
                    put(k + "cost_tf_rows", finite)
                else:
                    put(k + "cost_tf", sess.run(inner.cost_tf(xp, up, xnp), feed))


# --------------------------------------------------------------------------------------------------
# B. RunningMeanStd
# --------------------------------------------------------------------------------------------------
def gen_rms():
    tf.reset_default_graph()
    np.random.seed(7)
    ref_rms.test_runningmeanstd()            # the reference's own self-test (asserts inside)
    tf.reset_default_graph()
    with tf.Session() as sess:
        rms = ref_rms.RunningMeanStd(epsilon=1e-2, shape=[5])       # default epsilon
        put("B_rms__default_empty_mean", sess.run(rms.mean))
        put("B_rms__default_empty_std", sess.run(rms.std))
        rs = np.random.RandomState(11)
        x1 = (rs.randn(37, 5) * [1, 0.01, 3, 0.2, 10] + [0, 1, -2, 0.5, 100]).astype(np.float32)
        x2 = (rs.randn(64, 5) * [2, 0.01, 1, 0.2, 1] + [1, 1, 0, 0.5, 100]).astype(np.float32)
        put("B_rms__x1", x1); put("B_rms__x2", x2)
        rms.update(x1)
        put("B_rms__mean1", sess.run(rms.mean)); put("B_rms__std1", sess.run(rms.std))
        rms.update(x2)
        put("B_rms__mean2", sess.run(rms.mean)); put("B_rms__std2", sess.run(rms.std))


# --------------------------------------------------------------------------------------------------
# C. dynamics_model / policy_model
# --------------------------------------------------------------------------------------------------
MODEL_CASES = [   # (tag, env, hidden, K, seed, out_scale, B)
    ("swimmer", "swimmer", (256, 256), 3, 21, 1.0, 32),
    ("half-cheetah", "half-cheetah", (256, 256), 5, 22, 1.0, 32),
    ("hopper", "hopper", (256, 256), 3, 23, 1.0, 32),
    ("ant", "ant", (256, 256), 3, 24, 1.0, 32),
    ("humanoid", "humanoid", (256, 256), 2, 25, 1.0, 32),
    ("snake", "snake", (256, 256), 2, 26, 1.0, 32),
    ("half-cheetah-h1024", "half-cheetah", (1024, 1024), 2, 27, 1.0, 16),   # the JSON's own width
]


def gen_models():
    for tag, env_name, hidden, K, seed, out_scale, B in MODEL_CASES:
        w = World(env_name, hidden, K, seed, out_scale)
        k = "C_models__%s__" % tag
        for nm, v in zip(["in_mean", "in_std", "diff_mean", "diff_std"], w.rms):
            put(k + nm, v)
        for nm, v in zip(["in_mean", "in_std", "diff_mean", "diff_std"], w.rms_first):
            put(k + nm + "_first", v)
        s = RI.states(seed + 3, B, w.S)
        a = np.clip(RI.actions(seed + 4, (B, w.A)), -1, 1)
        xu = np.concatenate([s, a], axis=1)                      # f64 concat fed to an f32 placeholder
        put(k + "dyn_out", w.sess.run(w.dynamics_outs, {w.dynamics_in: xu}))
        put(k + "policy_mean", w.sess.run(w.policy_out, {w.policy_in: s}))
        acts, infos = None, None
        np.random.seed(seed + 5)
        acts, infos = w.training_policy.get_actions(s)
        put(k + "get_actions", acts, "rllab GaussianMLPPolicy.get_actions restated (App. A.1)")
        put(k + "get_actions_log_std", infos["log_std"])
        np.random.seed(seed + 5)
        put(k + "get_actions_rnd", np.random.normal(size=(B, w.A)))
        # G: fit loss through build_dynamics_graph (model_based_rl.py:23-103)
        rs = np.random.RandomState(seed + 6)
        xfull = (rs.randn(B, K * (w.S + w.A)) * 0.5).astype(np.float32)
        yfull = (rs.randn(B, K * w.S) * 0.5).astype(np.float32)
        put(k + "fit_x_full", xfull); put(k + "fit_y_full", yfull)
        feed = {w.dynamics_in_full: xfull, w.y_training_full: yfull}
        put(k + "fit_losses", w.sess.run(w.dynamics_losses, feed))
        put(k + "fit_prediction_loss", w.sess.run(w.prediction_loss, feed))
        w.close()


# --------------------------------------------------------------------------------------------------
# D. VecSimpleEnv, all sam_modes
# --------------------------------------------------------------------------------------------------
VEC_CASES = [   # (env, K, seed, B, T_steps, max_path_length)
    ("half-cheetah", 5, 31, 24, 8, 3),
    ("ant", 4, 32, 24, 8, 5),
    ("hopper", 3, 33, 16, 6, 4),
]


def ant_pool(rs, n, S):
    p = (rs.randn(n, S) * 0.1).astype(np.float32)
    p[:, 2] = rs.uniform(0.35, 0.85, size=n)
    return p


def gen_vec_env():
    for env_name, K, seed, B, T, mpl in VEC_CASES:
        for sam_mode in RI.SAM_MODES:
            w = World(env_name, (256, 256), K, seed, 1.0)
            rs = np.random.RandomState(seed + 10)
            n_pool = B * (T + 2)
            pool = ant_pool(rs, n_pool, w.S) if env_name == "ant" else RI.states(seed + 11, n_pool, w.S)
            w.set_reset_pool(pool)
            acts = RI.actions(seed + 12, (T, B, w.A))
            rec = RecordingRandom()
            ref_env_helpers.np = NumpyProxy(rec)
            try:
                np.random.seed(seed)
                vec = w.neural_net_env(sam_mode).vec_env_executor(n_envs=B, max_path_length=mpl)
                inner_vec = vec.vec_env
                obs0 = vec.reset()
                k = "D_vec__%s__%s__" % (env_name, sam_mode)
                put(k + "actions", acts)
                put_norm(k, w)
                put(k + "obs0", obs0)
                put(k + "cfg", np.array([K, B, T, mpl, seed]))
                states, rewards, dones, idx_used, std_noise, n_resets = [], [], [], [], [], []
                for t in range(T):
                    n_log = len(rec.log)
                    cur_before = inner_vec.cur_model_idx.copy()
                    with np.errstate(invalid="ignore"):
                        s, r, d, info = vec.step(acts[t])
                    assert info == {}
                    new = rec.log[n_log:]
                    if sam_mode == "step_rand":
                        idx_used.append(new[0][1])           # randint(K, size=B) (:619)
                    elif sam_mode == "eps_rand":
                        idx_used.append(cur_before)          # cur_model_idx (:622)
                    if sam_mode == "model_mean_std":
                        std_noise.append(new[0][1])          # normal(size=std.shape) (:626)
                    states.append(np.array(s)); rewards.append(np.array(r)); dones.append(np.array(d))
                    n_resets.append(w.inner_env.n_reset_calls)
                put(k + "states", np.stack(states)); put(k + "rewards", np.stack(rewards))
                put(k + "dones", np.stack(dones)); put(k + "n_reset_calls", np.array(n_resets))
                if idx_used:
                    put(k + "model_idx", np.stack(idx_used))
                if std_noise:
                    put(k + "std_noise", np.stack(std_noise))
                put(k + "pool", pool[:w.inner_env.n_reset_calls])     # the consumed prefix
                put(k + "states_dtype", np.array(str(np.stack(states).dtype)))
                put(k + "rewards_dtype", np.array(str(np.stack(rewards).dtype)))
                if env_name == "ant":
                    frac = np.stack(dones).mean()
                    assert 0.05 < frac < 0.95, frac
            finally:
                ref_env_helpers.np = np
            w.close()


# --------------------------------------------------------------------------------------------------
# E. the TRPO inner iteration through the reference's algo / sampler classes
# --------------------------------------------------------------------------------------------------
ITER_CASES = [   # (env, K, seed, batch_size, T, n_iters)
    ("half-cheetah", 5, 41, 60, 5, 2),      # n_envs = 60 // 5 = 12
    ("ant", 3, 42, 80, 8, 2),               # early termination: ragged paths, overshoot
]


def gen_trpo_iteration():
    for env_name, K, seed, batch_size, T, n_iters in ITER_CASES:
        w = World(env_name, (256, 256), K, seed, 1.0, T=T)
        rs = np.random.RandomState(seed + 10)
        n_pool = 4000
        pool = ant_pool(rs, n_pool, w.S) if env_name == "ant" else RI.states(seed + 11, n_pool, w.S)
        w.set_reset_pool(pool)
        LinearFeatureBaseline = ref_shims.LinearFeatureBaseline
        baseline = LinearFeatureBaseline(env_spec=w.env.spec)
        trpo = w.params["policy_opt_params"]["trpo"]
        algo = ref_trpo.TRPO(env=w.env, policy=w.training_policy, baseline=baseline,
                             batch_size=batch_size, max_path_length=T,
                             discount=0.97, step_size=trpo["step_size"])      # training.py:358-366
        algo.env = w.neural_net_env("step_rand")                            # model_based_rl.py:375
        k = "E_iter__%s__" % env_name
        put(k + "cfg", np.array([K, batch_size, T, n_iters, seed]))
        put(k + "discount", 0.97)
        put_norm(k, w)
        rec = RecordingRandom()
        ref_env_helpers.np = NumpyProxy(rec)
        sampler_mod = sys.modules["samplers.vectorized_sampler"]
        try:
            np.random.seed(seed)
            for j in range(n_iters):
                n_log0 = len(rec.log)
                resets0 = w.inner_env.n_reset_calls
                w.training_policy.noise_log = []
                algo.start_worker()                                           # :1175
                n_envs = algo.sampler.vec_env.num_envs
                dones_grid = []
                _step = algo.sampler.vec_env.step

                def _logged_step(a, _step=_step, dones_grid=dones_grid):
                    out = _step(a)
                    dones_grid.append(np.array(out[2]))
                    return out
                algo.sampler.vec_env.step = _logged_step
                with np.errstate(invalid="ignore"):
                    paths = algo.obtain_samples(j)                            # :1177
                    samples_data = algo.process_samples(j, paths)             # :1178
                    algo.optimize_policy(j, samples_data)                     # :1179
                kk = k + "it%d__" % j
                put(kk + "n_envs", n_envs)
                put(kk + "n_paths", len(paths))
                put(kk + "path_len", np.array([len(p["rewards"]) for p in paths]))
                put(kk + "reset_calls", np.array([resets0, w.inner_env.n_reset_calls]))
                # step_rand indices: the randint(K,size=B) draws (first randint of the iteration is
                # VecSimpleEnv.__init__'s cur_model_idx, then one per reset row, then per step)
                draws = [v for (kind, v) in rec.log[n_log0:] if kind == "randint" and v.shape == (n_envs,)]
                put(kk + "model_idx", np.stack(draws[1:]))
                put(kk + "eps", np.stack(w.training_policy.noise_log))      # get_actions draws
                put(kk + "dones_grid", np.stack(dones_grid))                 # [steps, n_envs]
                for key in ("observations", "actions", "rewards", "returns", "advantages"):
                    put(kk + key, samples_data[key])
                put(kk + "mean", samples_data["agent_infos"]["mean"])
                put(kk + "log_std", samples_data["agent_infos"]["log_std"])
                put(kk + "baseline_coeffs", baseline._coeffs,
                    "rllab LinearFeatureBaseline restated (App. A.4)")
                call = algo.optimizer.calls[-1]
                put(kk + "surr_loss", call["loss"],
                    "composition algos/npo.py:68-75 is the reference's; DiagonalGaussian restated (A.3)")
                put(kk + "mean_kl", call["constraint"])
                # the same loss / KL after moving the policy (so that KL != 0 is pinned too)
                prev = w.training_policy.get_param_values()
                step = np.random.RandomState(seed + 20 + j).randn(len(prev)).astype(np.float32) * 0.02
                w.training_policy.set_param_values(prev + step)
                put(kk + "param_step", step)
                put(kk + "surr_loss_moved", algo.optimizer.loss(call["inputs"]))
                put(kk + "mean_kl_moved", algo.optimizer.constraint_val(call["inputs"]))
                w.training_policy.set_param_values(prev)
                put(kk + "dtypes", np.array([str(samples_data[key].dtype) for key in
                                             ("observations", "actions", "rewards", "advantages")]))
        finally:
            ref_env_helpers.np = np
        del sampler_mod
        put(k + "pool", pool[:w.inner_env.n_reset_calls])             # the consumed prefix
        put(k + "max_constraint_val", algo.optimizer.max_constraint_val)
        w.close()


# --------------------------------------------------------------------------------------------------
# F. per-model validation cost graph + stop logic
# --------------------------------------------------------------------------------------------------
COST_CASES = [("half-cheetah", 5, 51, 20, 12, 0.99), ("ant", 4, 52, 20, 10, 1.0),
              ("hopper", 3, 53, 12, 8, 1.0), ("humanoid", 2, 54, 8, 6, 1.0)]


def gen_model_costs():
    for env_name, K, seed, B, T, gamma in COST_CASES:
        w = World(env_name, (256, 256), K, seed, 1.0, T=T)
        pop = w.policy_opt_params._replace(gamma=gamma)
        policy_training_init = tf.placeholder(tf.float32, shape=[None, w.S])
        is_done_tf = getattr(w.inner_env, "is_done_tf", None)           # model_based_rl.py:302-303
        costs, n_sat = ref_mbrl.build_policy_graph(
            "training_policy", "training_dynamics", policy_training_init, K, pop,
            w.policy_model, w.dynamics_model, w.env, w.inner_env.cost_tf, w.logger,
            is_done_tf, 0.0)
        rs = np.random.RandomState(seed + 10)
        init = ant_pool(rs, B, w.S) if env_name == "ant" else RI.states(seed + 11, B, w.S)
        k = "F_costs__%s__" % env_name
        put(k + "cfg", np.array([K, B, T, seed])); put(k + "gamma", gamma)
        put_norm(k, w)
        put(k + "init", init)
        with np.errstate(invalid="ignore"):
            put(k + "policy_costs", w.sess.run(costs, {policy_training_init: init}))
        w.close()
    # stop_critereon (utils.py:285-296), is_done (model_based_rl.py:1339-1371), update_stats (:1403-1419)
    rs = np.random.RandomState(60)
    f = ref_utils.stop_critereon(threshold=0.1, offset=1e-5, percent_models_threshold=0.3)
    old = rs.randn(40, 5)
    new = old + rs.randn(40, 5) * 0.5
    put("F_stop__old", old); put("F_stop__new", new)
    put("F_stop__vector", np.array([f(o, n, mode="vector") for o, n in zip(old, new)]))
    put("F_stop__scalar", np.array([f(float(o[0]), float(n[0])) for o, n in zip(old, new)]))

    class _L:
        def info(self, *a, **k):
            pass
    pop = types.SimpleNamespace(mode="estimated", stop_critereon=f)
    dec = []
    for o, n in zip(old, new):
        dec.append(ref_mbrl.is_done(pop, {"real": 0.0, "estimated": o.copy()},
                                    {"real": 1.0, "estimated": n.copy()}, _L()))
    put("F_stop__is_done_estimated", np.array(dec))
    pop_real = types.SimpleNamespace(mode="real", stop_critereon=f)
    put("F_stop__is_done_real", np.array([ref_mbrl.is_done(
        pop_real, {"real": float(o[0])}, {"real": float(n[0])}, _L()) for o, n in zip(old, new)]))
    upd_whole, upd_part = [], []
    for o, n in zip(old, new):
        m = {"real": float(o[0]), "estimated": o.copy()}
        ref_mbrl.update_stats(m, {"real": float(n[0]), "estimated": n.copy()}, whole=True)
        upd_whole.append(np.append(m["estimated"], m["real"]))
        m = {"real": float(o[0]), "estimated": o.copy()}
        ref_mbrl.update_stats(m, {"real": float(n[0]), "estimated": n.copy()}, whole=False)
        upd_part.append(np.append(m["estimated"], m["real"]))
    put("F_stop__update_whole", np.stack(upd_whole)); put("F_stop__update_part", np.stack(upd_part))


# --------------------------------------------------------------------------------------------------
# G. data_collection
# --------------------------------------------------------------------------------------------------
def gen_data_collection():
    dc = ref_utils.data_collection(max_size=50)
    rs = np.random.RandomState(70)
    log = []
    np.random.seed(70)
    for step, n in enumerate([20, 25, 30, 7]):
        x = rs.randn(n, 3).astype(np.float32) + step
        y = rs.randn(n, 2).astype(np.float32) + step
        dc.add_data(x, y)
        put("G_dc__add%d_x" % step, x); put("G_dc__add%d_y" % step, y)
        put("G_dc__after%d_x" % step, dc.x); put("G_dc__after%d_y" % step, dc.y)
        log.append([dc.get_num_data(), dc.cur_idx])
        bx, by = dc.get_next_batch(16)
        put("G_dc__next%d_x" % step, bx)
        log.append([dc.get_num_data(), dc.cur_idx])
        u_state = np.random.get_state()
        sx, sy = dc.sample(9)
        np.random.set_state(u_state)
        put("G_dc__sample%d_u" % step, np.random.uniform(0.0, 1.0, size=9))
        put("G_dc__sample%d_x" % step, sx); put("G_dc__sample%d_y" % step, sy)
    put("G_dc__log", np.array(log))
    t = rs.randn(6, 12).astype(np.float32)
    put("G_ith__t", t)
    put("G_ith__out", np.stack([ref_utils.get_ith_tensor(t, i, 4) for i in range(3)]))


def put_norm(k, w):
    for nm, v in zip(["in_mean", "in_std", "diff_mean", "diff_std"], w.rms):
        put(k + nm, v)


def main():
    gen_costs()
    gen_rms()
    gen_models()
    gen_vec_env()
    gen_trpo_iteration()
    gen_model_costs()
    gen_data_collection()
    out = os.path.join(HERE, "ref_fixtures.npz")
    np.savez_compressed(out, **OUT)
    with open(os.path.join(HERE, "ref_fixtures_notes.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_ref_fixtures.py",
                       reference="/root/reference (thanard/me-trpo @ 7dad9cd)",
                       n_arrays=len(OUT), restated_rllab_dependencies=NOTES), f, indent=1, sort_keys=True)
    print("wrote %s: %d arrays, %.1f KB" % (out, len(OUT), os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main()
