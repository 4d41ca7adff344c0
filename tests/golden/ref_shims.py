"""Shim modules that let the reference's OWN source files (/root/reference, read-only) be imported
and executed in this container, where TensorFlow 1.4 / rllab / MuJoCo are absent.

Used ONLY by tests/golden/make_ref_fixtures.py to generate committed golden fixtures (the
reference tree does not travel to the GPU box; the fixtures do).  Nothing here is product code and
nothing under me_trpo_b200/ imports it.

Two families of stand-ins are installed into sys.modules by `install()`:

1. `tensorflow` -- a tiny LAZY graph evaluated with NumPy in float32: placeholders, variables with
   variable_scope / get_variable reuse semantics, the ~40 ops the reference's hot-path graph
   builders call, and a Session whose run(fetches, feed_dict) walks the graph.  It has no kernels
   of its own: every op is the NumPy function of the same name, so what is being executed is the
   reference's graph-construction code (training.py dynamics_model / policy_model,
   running_mean_std.py, model_based_rl.py build_policy_graph / build_dynamics_graph, envs/*
   cost_tf / is_done_tf), not a restatement of it.  Feeds are cast to the placeholder dtype like
   TF does (the f64 -> f32 cast at the feed boundary is part of the reference's behaviour).

2. `rllab.*` / `sandbox.rocky.tf.*` -- import-time stubs (Serializable, MujocoEnv, logger, ...)
   plus the handful of rllab helpers that carry arithmetic on the path.  Those helpers are NOT in
   /root/reference (un-vendored dependency); they are restated here from rllab's published source
    and every fixture that depends on one says so in its `pins` note:
       rllab.misc.special.discount_cumsum      scipy.signal.lfilter([1],[1,-d],x[::-1])[::-1]
       rllab.algos.util.center_advantages      (a - mean) / (std + 1e-8)
       rllab.algos.util.shift_advantages_to_positive   (a - min) + 1e-8
       rllab.misc.tensor_utils.*               stack / concat / split of tensor (dict) lists
       rllab.envs.normalized_env.normalize     action space Box(-1, 1)
"""
import contextlib
import sys
import types

import numpy as np
import scipy.signal

# =================================================================================================
# mini TensorFlow (lazy graph, NumPy evaluation)
# =================================================================================================
_DUMMY_BATCH = 3


class DType:
    def __init__(self, name, np_dtype):
        self.name, self.np = name, np_dtype

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", np.float32)
float64 = DType("float64", np.float64)
int32 = DType("int32", np.int32)
int64 = DType("int64", np.int64)
bool_ = DType("bool", np.bool_)


def _np_dtype(dt):
    return dt.np if isinstance(dt, DType) else np.dtype(dt).type


def _const(a):
    """tf.convert_to_tensor semantics for constants mixed into float32 graphs: NumPy float64
    scalars / arrays become float32 (Python floats are already weakly typed under NumPy 2)."""
    if isinstance(a, np.ndarray) and a.dtype == np.float64:
        return a.astype(np.float32)
    if isinstance(a, np.float64):
        return np.float32(a)
    return a


class Tensor:
    """Node of the lazy graph.  `fn(*evaluated_inputs)` produces the NumPy value."""
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), name=None):
        self._fn, self._inputs, self.name = fn, tuple(_const(a) for a in inputs), name
        self._static_shape = None

    # ---- evaluation -----------------------------------------------------------------------------
    def _eval(self, cache, feed):
        key = id(self)
        if key in cache:
            return cache[key]
        if key in feed:
            v = feed[key]
        else:
            args = [a._eval(cache, feed) if isinstance(a, Tensor) else a for a in self._inputs]
            v = self._fn(*args)
        cache[key] = v
        return v

    @property
    def shape(self):
        """Static shape by evaluating the sub-graph on zero-filled placeholders (unknown dims take a
        dummy size); enough for the reference's `x.shape[1] == n` asserts."""
        if self._static_shape is None:
            v = self._eval({}, _DummyFeed())
            self._static_shape = tuple(np.shape(v))
        return self._static_shape

    def get_shape(self):
        return self.shape

    # ---- operators ------------------------------------------------------------------------------
    def __getitem__(self, idx):
        return Tensor(lambda v: v[idx], [self])

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        return Tensor(lambda *a: ufunc(*a, **kwargs), inputs)

    def __neg__(self):
        return Tensor(np.negative, [self])

    def __abs__(self):
        return Tensor(np.abs, [self])


def _binop(npf, swap=False):
    def f(self, other):
        args = [other, self] if swap else [self, other]
        return Tensor(lambda a, b: npf(a, b), args)
    return f


for _name, _f in [("add", np.add), ("sub", np.subtract), ("mul", np.multiply),
                  ("truediv", np.true_divide), ("pow", np.power)]:
    setattr(Tensor, "__%s__" % _name, _binop(_f))
    setattr(Tensor, "__r%s__" % _name, _binop(_f, swap=True))
for _name, _f in [("ge", np.greater_equal), ("le", np.less_equal), ("gt", np.greater),
                  ("lt", np.less)]:
    setattr(Tensor, "__%s__" % _name, _binop(_f))


class _DummyFeed(dict):
    """feed used for static-shape inference: placeholders evaluate to zeros."""

    def __contains__(self, key):
        return False


class Placeholder(Tensor):
    def __init__(self, dtype, shape, name):
        self.dtype = dtype
        self._ph_shape = None if shape is None else tuple(shape) if np.ndim(shape) else (
            () if shape in ((), []) else (shape,))
        super().__init__(self._dummy, [], name)

    def _dummy(self):
        if self._ph_shape is None:
            raise ValueError("placeholder %s without a shape was not fed" % self.name)
        shp = tuple(_DUMMY_BATCH if d is None else int(d) for d in self._ph_shape)
        return np.zeros(shp, _np_dtype(self.dtype))

    def _eval(self, cache, feed):
        key = id(self)
        if key in feed:
            return feed[key]
        if isinstance(feed, _DummyFeed):
            return self._dummy()
        raise ValueError("You must feed a value for placeholder tensor '%s'" % self.name)


class Variable(Tensor):
    def __init__(self, value, name, trainable=True):
        self.value = value
        self.trainable = trainable
        super().__init__(lambda: self.value, [], name)

    def _eval(self, cache, feed):       # never cached: assign ops change it within a run
        return self.value

    def assign(self, other):
        return assign(self, other)

    def load(self, value, session=None):
        self.value = np.asarray(value, self.value.dtype).reshape(self.value.shape)


def _t(x):
    """Python lists of tensors behave like tf.stack'ed tensors when passed to an op."""
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return stack(list(x))
    return x


def _op(npf, *args):
    return Tensor(npf, [_t(a) for a in args])


def _f32(v):
    return v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v


# ---- graph construction API -----------------------------------------------------------------------
def placeholder(dtype=float32, shape=None, name=None):
    return Placeholder(dtype, shape, name)


def constant(value, dtype=float32, name=None):
    v = np.asarray(value, _np_dtype(dtype))
    return Tensor(lambda: v, [], name)


def identity(x, name=None):
    return Tensor(lambda v: v, [_t(x)], name)


def matmul(a, b, name=None):
    return _op(np.matmul, a, b)


def add(a, b, name=None):
    return _op(np.add, a, b)


def concat(values, axis, name=None):
    return Tensor(lambda *vs: np.concatenate(vs, axis=axis), list(values), name)


def stack(values, axis=0, name=None):
    return Tensor(lambda *vs: np.stack([np.asarray(v) for v in vs], axis=axis), list(values), name)


def _reduce(npf):
    def f(x, axis=None, name=None, keep_dims=False):
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return Tensor(lambda v: npf(np.asarray(v), axis=ax, keepdims=keep_dims), [_t(x)], name)
    return f


reduce_mean = _reduce(np.mean)
reduce_sum = _reduce(np.sum)
reduce_all = _reduce(np.all)
reduce_max = _reduce(np.max)
reduce_min = _reduce(np.min)


def square(x, name=None):
    return _op(np.square, x)


def sqrt(x, name=None):
    return _op(np.sqrt, x)


def exp(x, name=None):
    return _op(np.exp, x)


def log(x, name=None):
    return _op(np.log, x)


def tanh(x, name=None):
    return _op(np.tanh, x)


def abs(x, name=None):   # noqa: A001  (mirrors tf.abs)
    return _op(np.abs, x)


def maximum(a, b, name=None):
    return _op(np.maximum, a, b)


def minimum(a, b, name=None):
    return _op(np.minimum, a, b)


def equal(a, b, name=None):
    return _op(np.equal, a, b)


def logical_and(a, b, name=None):
    return _op(np.logical_and, a, b)


def logical_not(x, name=None):
    return _op(np.logical_not, x)


def logical_or(a, b, name=None):
    return _op(np.logical_or, a, b)


def is_finite(x, name=None):
    return _op(np.isfinite, x)


def clip_by_value(x, lo, hi, name=None):
    return Tensor(lambda v, a, b: np.clip(v, a, b).astype(np.asarray(v).dtype), [_t(x), lo, hi], name)


def cast(x, dtype, name=None):
    npd = _np_dtype(dtype)
    return Tensor(lambda v: np.asarray(v).astype(npd), [_t(x)], name)


def to_float(x, name=None):
    return cast(x, float32)


def shape(x, name=None):   # noqa: F811
    return Tensor(lambda v: np.asarray(np.shape(v), np.int32), [_t(x)], name)


_rng = np.random.RandomState(1234)


def set_random_seed(seed):
    global _rng
    _rng = np.random.RandomState(seed)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, name=None):   # noqa: F811
    parts = list(shape) if isinstance(shape, (list, tuple)) else [shape]

    def f(*dims):
        shp = tuple(int(np.asarray(d)) for d in dims) if isinstance(shape, (list, tuple)) \
            else tuple(int(d) for d in np.asarray(dims[0]))
        return (_rng.standard_normal(shp) * stddev + mean).astype(_np_dtype(dtype))
    return Tensor(f, parts, name)


class _AssignOp(Tensor):
    def __init__(self, var, value, mode):
        self._var, self._mode = var, mode
        super().__init__(self._apply, [_t(value)], None)

    def _apply(self, v):
        new = np.asarray(v, self._var.value.dtype)
        if self._mode == "add":
            new = self._var.value + new
        self._var.value = np.broadcast_to(new, self._var.value.shape).astype(
            self._var.value.dtype).copy()
        return self._var.value

    def _eval(self, cache, feed):
        if isinstance(feed, _DummyFeed):
            return self._var.value
        return super()._eval(cache, feed)


def assign(var, value, name=None):
    return _AssignOp(var, value, "set")


def assign_add(var, value, name=None):
    return _AssignOp(var, value, "add")


# ---- variables / scopes ---------------------------------------------------------------------------
class _Graph:
    def __init__(self):
        self.variables = {}
        self.scope_stack = []
        self.collections = {}


_graph = _Graph()


def reset_default_graph():
    global _graph
    _graph = _Graph()


class _VarScope:
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse

    def reuse_variables(self):
        self.reuse = True


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    parent_reuse = _graph.scope_stack[-1].reuse if _graph.scope_stack else False
    sc = _VarScope(name, bool(reuse) or parent_reuse)
    _graph.scope_stack.append(sc)
    try:
        yield sc
    finally:
        _graph.scope_stack.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


def _full_name(name):
    return "/".join([s.name for s in _graph.scope_stack] + [name])


def get_variable(name, shape=None, dtype=float32, initializer=None, trainable=True):   # noqa: F811
    full = _full_name(name)
    reuse = _graph.scope_stack[-1].reuse if _graph.scope_stack else False
    if full in _graph.variables:
        if not reuse:
            raise ValueError("Variable %s already exists, disallowed." % full)
        return _graph.variables[full]
    if reuse:
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()." % full)
    if shape is None or initializer is None:
        raise ValueError("Shape/initializer of a new variable (%s) must be fully defined." % full)
    shp = tuple(shape) if np.ndim(shape) else (int(shape),)
    value = np.asarray(initializer(shp), _np_dtype(dtype)).reshape(shp)
    v = Variable(value, full + ":0", trainable)
    _graph.variables[full] = v
    return v


def constant_initializer(value=0.0):
    return lambda shp: np.full(shp, value, np.float32)


def zeros_initializer():
    return constant_initializer(0.0)


def xavier_initializer(uniform=True, seed=None, dtype=float32):
    """tf.contrib.layers.xavier_initializer(): U(+-sqrt(6/(fan_in+fan_out))); TF's fan computation
    for a rank-1 shape (n,) gives fan_in = fan_out = n."""
    def init(shp):
        fan_in, fan_out = (shp[0], shp[0]) if len(shp) == 1 else (shp[0], shp[1])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _rng.uniform(-lim, lim, size=shp).astype(np.float32)
    return init


def global_variables_initializer():
    return Tensor(lambda: None, [])


def variables_initializer(var_list):
    return Tensor(lambda: None, [])


def global_variables():
    return list(_graph.variables.values())


def get_collection(key, scope=None):
    if key == "variables":
        return [v for n, v in _graph.variables.items() if scope is None or n.startswith(scope)]
    return list(_graph.collections.get(key, []))


def add_to_collection(key, value):
    _graph.collections.setdefault(key, []).append(value)


class GraphKeys:
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "variables"


# ---- session --------------------------------------------------------------------------------------
_session_stack = []


class Session:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        _session_stack.append(self)
        return self

    def __exit__(self, *exc):
        _session_stack.pop()
        return False

    def as_default(self):
        return self

    def run(self, fetches, feed_dict=None):
        feed = {}
        for ph, val in (feed_dict or {}).items():
            npd = _np_dtype(ph.dtype) if hasattr(ph, "dtype") else np.float32
            feed[id(ph)] = np.asarray(val, npd)
        cache = {}

        def ev(f):
            if isinstance(f, (list, tuple)):
                return [ev(e) for e in f]
            if isinstance(f, dict):
                return {k: ev(e) for k, e in f.items()}
            v = f._eval(cache, feed)
            return np.copy(v) if isinstance(v, np.ndarray) else v
        return ev(fetches)


InteractiveSession = Session


def get_default_session():
    return _session_stack[-1] if _session_stack else None


class _Summary:
    @staticmethod
    def histogram(*a, **k):
        return None

    @staticmethod
    def scalar(*a, **k):
        return None


def _build_tf_module():
    tf = types.ModuleType("tensorflow")
    g = globals()
    for n in ("placeholder constant identity matmul add concat stack reduce_mean reduce_sum "
              "reduce_all reduce_max reduce_min square sqrt exp log tanh abs maximum minimum equal "
              "logical_and logical_not logical_or is_finite clip_by_value cast to_float shape set_random_seed "
              "random_normal assign assign_add reset_default_graph variable_scope name_scope "
              "get_variable constant_initializer zeros_initializer global_variables_initializer "
              "variables_initializer global_variables get_collection add_to_collection GraphKeys "
              "Session InteractiveSession get_default_session float32 float64 int32 int64 Tensor "
              "Variable").split():
        setattr(tf, n, g[n])
    tf.bool = bool_
    tf.summary = _Summary
    nn = types.ModuleType("tensorflow.nn")
    nn.relu = lambda x, name=None: _op(lambda v: np.maximum(v, 0), x)
    nn.tanh = tanh
    nn.sigmoid = lambda x, name=None: _op(lambda v: 1.0 / (1.0 + np.exp(-v)), x)
    nn.l2_loss = lambda x, name=None: _op(lambda v: np.sum(np.square(v)) / 2, x)
    tf.nn = nn
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = xavier_initializer
    contrib.layers = layers
    tf.contrib = contrib
    return {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.contrib": contrib,
            "tensorflow.contrib.layers": layers}


# =================================================================================================
# rllab / sandbox stubs
# =================================================================================================
class Box:
    """rllab.spaces.Box: only what the path touches."""

    def __init__(self, low, high, shape=None):
        if shape is None:
            self.low, self.high = np.asarray(low, np.float64), np.asarray(high, np.float64)
        else:
            self.low, self.high = np.full(shape, low, np.float64), np.full(shape, high, np.float64)

    @property
    def shape(self):
        return self.low.shape

    @property
    def bounds(self):
        return self.low, self.high

    @property
    def flat_dim(self):
        return int(np.prod(self.low.shape))

    def flatten(self, x):
        return np.asarray(x).flatten()

    def flatten_n(self, xs):                     # rllab Box.flatten_n
        xs = np.asarray(xs)
        return xs.reshape((xs.shape[0], -1))


class EnvSpec:
    def __init__(self, observation_space, action_space):
        self.observation_space, self.action_space = observation_space, action_space


def _tensor_utils():
    m = types.ModuleType("rllab.misc.tensor_utils")

    def stack_tensor_list(tensor_list):
        return np.array(tensor_list)

    def stack_tensor_dict_list(tensor_dict_list):
        keys = list(tensor_dict_list[0].keys())
        ret = dict()
        for k in keys:
            example = tensor_dict_list[0][k]
            if isinstance(example, dict):
                v = stack_tensor_dict_list([x[k] for x in tensor_dict_list])
            else:
                v = stack_tensor_list([x[k] for x in tensor_dict_list])
            ret[k] = v
        return ret

    def concat_tensor_list(tensor_list):
        return np.concatenate(tensor_list, axis=0)

    def concat_tensor_dict_list(tensor_dict_list):
        keys = list(tensor_dict_list[0].keys())
        ret = dict()
        for k in keys:
            example = tensor_dict_list[0][k]
            if isinstance(example, dict):
                v = concat_tensor_dict_list([x[k] for x in tensor_dict_list])
            else:
                v = concat_tensor_list([x[k] for x in tensor_dict_list])
            ret[k] = v
        return ret

    def split_tensor_dict_list(tensor_dict):
        keys = list(tensor_dict.keys())
        ret = None
        for k in keys:
            vals = tensor_dict[k]
            if isinstance(vals, dict):
                vals = split_tensor_dict_list(vals)
            if ret is None:
                ret = [{k: v} for v in vals]
            else:
                for v, cur_dict in zip(vals, ret):
                    cur_dict[k] = v
        return ret

    for f in (stack_tensor_list, stack_tensor_dict_list, concat_tensor_list,
              concat_tensor_dict_list, split_tensor_dict_list):
        setattr(m, f.__name__, f)
    return m


def _special():
    m = types.ModuleType("rllab.misc.special")

    def discount_cumsum(x, discount):
        return scipy.signal.lfilter([1], [1, float(-discount)], x[::-1], axis=0)[::-1]

    def explained_variance_1d(ypred, y):
        vary = np.var(y)
        if np.isclose(vary, 0):
            return 1 if np.var(ypred) > 0 else 0
        return 1 - np.var(y - ypred) / (vary + 1e-8)

    m.discount_cumsum, m.explained_variance_1d = discount_cumsum, explained_variance_1d
    return m


def _algos_util():
    m = types.ModuleType("rllab.algos.util")
    m.center_advantages = lambda advantages: (advantages - np.mean(advantages)) / (advantages.std() + 1e-8)
    m.shift_advantages_to_positive = lambda advantages: (advantages - np.min(advantages)) + 1e-8
    return m


class _Logger(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def _build_rllab_modules():
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        mods[name] = m
        return m

    class Serializable:
        def __init__(self, *a, **k):
            pass

        @staticmethod
        def quick_init(self, locals_):
            pass

    class Env:
        @property
        def spec(self):
            return EnvSpec(self.observation_space, self.action_space)

    class MujocoEnv(Env):
        def __init__(self, *a, **k):
            pass

    class ProxyEnv(Env):
        def __init__(self, wrapped_env):
            self._wrapped_env = wrapped_env

        @property
        def wrapped_env(self):
            return self._wrapped_env

        def reset(self, **kwargs):
            return self._wrapped_env.reset(**kwargs)

        @property
        def action_space(self):
            return self._wrapped_env.action_space

        @property
        def observation_space(self):
            return self._wrapped_env.observation_space

        def step(self, action):
            return self._wrapped_env.step(action)

        def terminate(self):
            pass

    class NormalizedEnv(ProxyEnv):
        """rllab normalize(): action space becomes Box(-1, 1)."""

        @property
        def action_space(self):
            ub = np.ones(self._wrapped_env.action_space.shape)
            return Box(-1 * ub, ub)

    class ProgBarCounter:
        def __init__(self, total_count):
            pass

        def inc(self, n):
            pass

        def stop(self):
            pass

    def Step(observation, reward, done, **kwargs):
        return observation, reward, done, kwargs

    class _AutoArgs(types.ModuleType):
        @staticmethod
        def arg(*a, **k):
            return lambda f: f

    mod("rllab")
    mod("rllab.core")
    mod("rllab.core.serializable", Serializable=Serializable)
    mod("rllab.envs")
    mod("rllab.envs.base", Env=Env, Step=Step, EnvSpec=EnvSpec)
    mod("rllab.envs.mujoco")
    mod("rllab.envs.mujoco.mujoco_env", MujocoEnv=MujocoEnv, q_mult=None, q_inv=None)
    mod("rllab.envs.normalized_env", normalize=NormalizedEnv, NormalizedEnv=NormalizedEnv)
    mod("rllab.envs.proxy_env", ProxyEnv=ProxyEnv)
    mods["rllab.misc"] = types.ModuleType("rllab.misc")
    mods["rllab.misc.logger"] = _Logger("rllab.misc.logger")
    mods["rllab.misc.autoargs"] = _AutoArgs("rllab.misc.autoargs")
    mod("rllab.misc.overrides", overrides=lambda f: f)
    mods["rllab.misc.special"] = _special()
    mods["rllab.misc.tensor_utils"] = _tensor_utils()
    mod("rllab.misc.ext")
    mod("rllab.algos")
    mods["rllab.algos.util"] = _algos_util()
    mod("rllab.algos.base", RLAlgorithm=object)
    mod("rllab.sampler")
    mod("rllab.sampler.stateful_pool", ProgBarCounter=ProgBarCounter, singleton_pool=None)
    mod("rllab.sampler.parallel_sampler")
    mod("rllab.spaces")
    mod("rllab.spaces.box", Box=Box)
    mod("rllab.spaces.discrete", Discrete=type("Discrete", (), {}))
    mod("rllab.spaces.product", Product=type("Product", (), {}))
    mod("rllab.plotter")
    mod("rllab.config", PROJECT_PATH="/tmp")
    mod("sandbox")
    mod("sandbox.rocky")
    mod("sandbox.rocky.tf")
    mod("sandbox.rocky.tf.envs")
    mod("sandbox.rocky.tf.misc")
    mods["sandbox.rocky.tf.misc.tensor_utils"] = types.ModuleType("sandbox.rocky.tf.misc.tensor_utils")
    mod("sandbox.rocky.tf.spaces")
    mod("sandbox.rocky.tf.spaces.box", Box=Box)
    mod("sandbox.rocky.tf.spaces.discrete", Discrete=type("Discrete", (), {}))
    mod("sandbox.rocky.tf.spaces.product", Product=type("Product", (), {}))
    mod("cached_property", cached_property=property)
    mods["rllab.misc.ext"].extract = lambda x, *keys: tuple(x[k] for k in keys)
    mods["sandbox.rocky.tf.misc.tensor_utils"].new_tensor = \
        lambda name, ndim, dtype: placeholder(dtype, [None] * ndim, name=name)
    mod("sandbox.rocky.tf.policies")
    mod("sandbox.rocky.tf.policies.base", Policy=object)
    mod("sandbox.rocky.tf.policies.gaussian_mlp_policy", GaussianMLPPolicy=GaussianMLPPolicy)
    mod("sandbox.rocky.tf.distributions")
    mod("sandbox.rocky.tf.distributions.diagonal_gaussian", DiagonalGaussian=DiagonalGaussian)
    mod("sandbox.rocky.tf.optimizers")
    mod("sandbox.rocky.tf.optimizers.conjugate_gradient_optimizer",
        ConjugateGradientOptimizer=CapturingOptimizer)
    mod("sandbox.rocky.tf.optimizers.penalty_lbfgs_optimizer", PenaltyLbfgsOptimizer=CapturingOptimizer)
    mod("sandbox.rocky.tf.optimizers.first_order_optimizer", FirstOrderOptimizer=CapturingOptimizer)
    mod("rllab.baselines")
    mod("rllab.baselines.linear_feature_baseline", LinearFeatureBaseline=LinearFeatureBaseline)
    mod("joblib")
    # attribute links for `from rllab.misc import x` / `import rllab.misc.logger as logger`
    for full, m in list(mods.items()):
        if "." in full:
            parent, child = full.rsplit(".", 1)
            if parent in mods:
                setattr(mods[parent], child, m)
    return mods


# =================================================================================================
# rllab policy / distribution / baseline / optimizer stand-ins (RESTATED from rllab's published
# source, SURVEY.md Appendix A.1-A.4 -- rllab is an un-vendored, unpinned dependency of the
# reference).  They are written against the mini-tf above so that the reference's own graph
# builders (training.py policy_model, algos/npo.py init_opt) compose with them.
# =================================================================================================
class DiagonalGaussian:
    """sandbox.rocky.tf.distributions.diagonal_gaussian.DiagonalGaussian (Appendix A.3)."""

    def __init__(self, dim):
        self._dim = dim

    @property
    def dim(self):
        return self._dim

    @property
    def dist_info_keys(self):
        return ["mean", "log_std"]

    @property
    def dist_info_specs(self):
        return [("mean", (self.dim,)), ("log_std", (self.dim,))]

    def kl_sym(self, old_dist_info_vars, new_dist_info_vars):
        old_means, old_log_stds = old_dist_info_vars["mean"], old_dist_info_vars["log_std"]
        new_means, new_log_stds = new_dist_info_vars["mean"], new_dist_info_vars["log_std"]
        old_std, new_std = exp(old_log_stds), exp(new_log_stds)
        numerator = square(old_means - new_means) + square(old_std) - square(new_std)
        denominator = 2 * square(new_std) + 1e-8
        return reduce_sum(numerator / denominator + new_log_stds - old_log_stds, axis=-1)

    def log_likelihood_sym(self, x_var, dist_info_vars):
        means, log_stds = dist_info_vars["mean"], dist_info_vars["log_std"]
        zs = (x_var - means) / exp(log_stds)
        return - reduce_sum(log_stds, axis=-1) - 0.5 * reduce_sum(square(zs), axis=-1) - \
            0.5 * self.dim * np.log(2 * np.pi)

    def likelihood_ratio_sym(self, x_var, old_dist_info_vars, new_dist_info_vars):
        logli_new = self.log_likelihood_sym(x_var, new_dist_info_vars)
        logli_old = self.log_likelihood_sym(x_var, old_dist_info_vars)
        return exp(logli_new - logli_old)

    def entropy(self, dist_info):
        return np.sum(dist_info["log_std"] + np.log(np.sqrt(2 * np.pi * np.e)), axis=-1)


class _Layer:
    pass


class GaussianMLPPolicy:
    """sandbox.rocky.tf.policies.gaussian_mlp_policy.GaussianMLPPolicy (Appendix A.1), defaults
    learn_std=True, adaptive_std=False, min_std=1e-6, hidden tanh, std_parametrization='exp'.
    W Xavier-uniform, b zeros, log_std = log(init_std)."""
    vectorized = True
    recurrent = False
    state_info_specs = []
    state_info_keys = []

    def __init__(self, name, env_spec, hidden_sizes=(32, 32), init_std=1.0, min_std=1e-6,
                 hidden_nonlinearity=None, output_nonlinearity=None):
        hidden_nonlinearity = hidden_nonlinearity or tanh
        self.name = name
        obs_dim = env_spec.observation_space.flat_dim
        action_dim = env_spec.action_space.flat_dim
        self._mean_network = _Layer()
        inp = _Layer()
        inp.shape = (None, obs_dim)
        layers = [inp]
        dims = [obs_dim] + list(hidden_sizes) + [action_dim]
        xav = xavier_initializer()
        with variable_scope(name):
            with variable_scope("mean_network"):
                for i in range(len(dims) - 1):
                    lname = "hidden_%d" % i if i < len(dims) - 2 else "output"
                    with variable_scope(lname):
                        l = _Layer()
                        l.W = get_variable("W", (dims[i], dims[i + 1]), initializer=xav)
                        l.b = get_variable("b", (dims[i + 1],), initializer=zeros_initializer())
                    nl = hidden_nonlinearity if i < len(dims) - 2 else output_nonlinearity
                    l.nonlinearity = nl if nl is not None else identity
                    layers.append(l)
            with variable_scope("std_network"):
                self._l_std_param = _Layer()
                self._l_std_param.param = get_variable(
                    "output_std_param/param", (action_dim,),
                    initializer=constant_initializer(np.log(init_std)))
        self._mean_network.layers = layers
        self.min_std_param = np.log(min_std)
        self._dist = DiagonalGaussian(action_dim)
        self._obs_ph = placeholder(float32, [None, obs_dim], name="policy_obs")
        self._f_dist_out = self.dist_info_sym(self._obs_ph)

    @property
    def distribution(self):
        return self._dist

    def dist_info_sym(self, obs_var, state_info_vars=None):
        h = obs_var
        for l in self._mean_network.layers[1:]:
            h = l.nonlinearity(matmul(h, l.W) + l.b)
        mean_var = h
        # ParamLayer: the [A] vector tiled over the batch
        std_param_var = Tensor(lambda m, p: np.broadcast_to(p, np.shape(m)).astype(np.float32),
                               [mean_var, self._l_std_param.param])
        std_param_var = maximum(std_param_var, np.float32(self.min_std_param))
        return dict(mean=mean_var, log_std=std_param_var)

    def get_actions(self, observations):
        flat_obs = np.asarray(observations).reshape(len(observations), -1)
        out = get_default_session().run(self._f_dist_out, {self._obs_ph: flat_obs})
        means, log_stds = out["mean"], out["log_std"]
        rnd = np.random.normal(size=means.shape)
        if getattr(self, "noise_log", None) is not None:
            self.noise_log.append(rnd)            # lets the fixture generator store the draws
        actions = rnd * np.exp(log_stds) + means
        return actions, dict(mean=means, log_std=log_stds)

    def reset(self, dones=None):
        pass

    def get_params(self, trainable=True):
        ps = []
        for l in self._mean_network.layers[1:]:
            ps += [l.W, l.b]
        return ps + [self._l_std_param.param]

    def get_param_values(self, trainable=True):
        return np.concatenate([p.value.flatten() for p in self.get_params()])

    def set_param_values(self, flat, trainable=True):
        o = 0
        for p in self.get_params():
            n = p.value.size
            p.value = np.asarray(flat[o:o + n], np.float32).reshape(p.value.shape)
            o += n


class LinearFeatureBaseline:
    """rllab.baselines.linear_feature_baseline.LinearFeatureBaseline (Appendix A.4)."""

    def __init__(self, env_spec=None, reg_coeff=1e-5):
        self._coeffs = None
        self._reg_coeff = reg_coeff

    def _features(self, path):
        o = np.clip(path["observations"], -10, 10)
        l = len(path["rewards"])
        al = np.arange(l).reshape(-1, 1) / 100.0
        return np.concatenate([o, o ** 2, al, al ** 2, al ** 3, np.ones((l, 1))], axis=1)

    def fit(self, paths):
        featmat = np.concatenate([self._features(path) for path in paths])
        returns = np.concatenate([path["returns"] for path in paths])
        reg_coeff = self._reg_coeff
        for _ in range(5):
            self._coeffs = np.linalg.lstsq(
                featmat.T.dot(featmat) + reg_coeff * np.identity(featmat.shape[1]),
                featmat.T.dot(returns), rcond=None)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg_coeff *= 10

    def predict(self, path):
        if self._coeffs is None:
            return np.zeros(len(path["rewards"]))
        return self._features(path).dot(self._coeffs)


class CapturingOptimizer:
    """Stands where rllab's ConjugateGradientOptimizer does (algos/trpo.py:20): keeps what
    NPO.init_opt hands to update_opt (algos/npo.py:85-91) and, on optimize(inputs)
    (algos/npo.py:111), records the inputs and evaluates the reference-built loss / constraint
    tensors at the current policy parameters.  It does not step the policy."""

    def __init__(self, **kwargs):
        self.calls = []

    def update_opt(self, loss, target, leq_constraint, inputs, constraint_name="constraint", **kw):
        self.loss_t, self.target, self.inputs_t = loss, target, inputs
        self.constraint_t, self.max_constraint_val = leq_constraint
        self.constraint_name = constraint_name

    def _feed(self, inputs):
        return {ph: v for ph, v in zip(self.inputs_t, inputs)}

    def loss(self, inputs):
        return get_default_session().run(self.loss_t, self._feed(inputs))

    def constraint_val(self, inputs):
        return get_default_session().run(self.constraint_t, self._feed(inputs))

    def optimize(self, inputs):
        self.calls.append(dict(inputs=[np.array(v) for v in inputs], loss=self.loss(inputs),
                               constraint=self.constraint_val(inputs)))


def _new_tensor_variable(self, name, extra_dims):
    return placeholder(float32, [None] * extra_dims + [self.flat_dim], name=name)


Box.new_tensor_variable = _new_tensor_variable


def install():
    """Install the shim modules (idempotent) and return the fake `tensorflow` module."""
    mods = {}
    mods.update(_build_tf_module())
    mods.update(_build_rllab_modules())
    for name, m in mods.items():
        sys.modules[name] = m
    return sys.modules["tensorflow"]
