"""ctypes binding of libmetrpo.so (include/metrpo.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is no fallback: if the
shared object is missing or a symbol is absent, importing callers get a RuntimeError.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("METRPO_LIB", os.path.join(_HERE, "libmetrpo.so"))   # METRPO_LIB: dev override (trace build)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "metrpo.h")

MAX_POLICY_LAYERS = 4

ENV_IDS = {
    "swimmer": 0, "half-cheetah": 1, "half_cheetah": 1, "hopper": 2, "ant": 3,
    "humanoid": 4, "snake": 5,
}
SAM_MODES = {
    "step_rand": 0, "eps_rand": 1, "model_mean_std": 2, "model_mean": 3, "model_med": 4,
    "one_model": 5,
}


class RolloutCfg(ctypes.Structure):
    _fields_ = [
        ("state_dim", ctypes.c_int32), ("action_dim", ctypes.c_int32),
        ("drop_cols", ctypes.c_int32), ("hidden", ctypes.c_int32),
        ("n_models", ctypes.c_int32), ("n_envs", ctypes.c_int32),
        ("max_path_length", ctypes.c_int32), ("env_id", ctypes.c_int32),
        ("sam_mode", ctypes.c_int32), ("n_policy_layers", ctypes.c_int32),
        ("policy_dims", ctypes.c_int32 * (MAX_POLICY_LAYERS + 1)),
        ("policy_out_tanh", ctypes.c_int32), ("precision", ctypes.c_int32),
        ("device", ctypes.c_int32), ("row_offset", ctypes.c_int32),
    ]


class TrpoCfg(ctypes.Structure):
    _fields_ = [
        ("state_dim", ctypes.c_int32), ("action_dim", ctypes.c_int32),
        ("n_policy_layers", ctypes.c_int32),
        ("policy_dims", ctypes.c_int32 * (MAX_POLICY_LAYERS + 1)),
        ("policy_out_tanh", ctypes.c_int32), ("device", ctypes.c_int32),
    ]


class FitCfg(ctypes.Structure):
    _fields_ = [
        ("state_dim", ctypes.c_int32), ("action_dim", ctypes.c_int32), ("drop_cols", ctypes.c_int32),
        ("hidden", ctypes.c_int32), ("n_models", ctypes.c_int32), ("max_rows", ctypes.c_int32),
        ("precision", ctypes.c_int32), ("device", ctypes.c_int32),
    ]


_vp, _i, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
_ll, _d = ctypes.c_longlong, ctypes.c_double
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)
_PROTOS = {
    "metrpo_version": (ctypes.c_char_p, []),
    "metrpo_last_error": (ctypes.c_char_p, []),
    "metrpo_rollout_create": (_i, [ctypes.POINTER(RolloutCfg), ctypes.POINTER(_vp)]),
    "metrpo_rollout_destroy": (_i, [_vp]),
    "metrpo_rollout_set_dynamics": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_rollout_set_normalization": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_rollout_set_policy": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _vp]),
    "metrpo_rollout_reset": (_i, [_vp, _vp, _vp]),
    "metrpo_rollout_set_rows": (_i, [_vp, _vp, _i, _vp, _vp]),
    "metrpo_rollout_step": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "metrpo_rollout_run": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _u64, _u64, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_rollout_continue": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _u64, _u64, _i,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_rollout_model_costs": (_i, [_vp, _i, _i, _vp, _d, _vp, _vp, _vp]),
    "metrpo_rollout_last_launches": (_i, [_vp]),
    "metrpo_rollout_last_kernel": (_i, [_vp]),
    "metrpo_rollout_status": (_i, [_vp, _vp]),
    "metrpo_rollout_set_trace": (_i, [_vp, _i, _i, _i]),
    "metrpo_rollout_get_trace": (_i, [_vp, _vp]),
    "metrpo_debug_schedule": (_i, [_i, _i, _i, _vp, _i]),
    "metrpo_trpo_create": (_i, [ctypes.POINTER(TrpoCfg), ctypes.POINTER(_vp)]),
    "metrpo_trpo_destroy": (_i, [_vp]),
    "metrpo_trpo_set_allreduce": (_i, [_vp, ALLREDUCE_FN, _vp]),
    "metrpo_trpo_p2p_handle": (_i, [_vp, _vp]),
    "metrpo_trpo_enable_p2p": (_i, [_vp, _i, _i, _vp]),
    "metrpo_trpo_num_params": (_i, [_vp]),
    "metrpo_trpo_set_pass_impl": (_i, [_vp, _i]),
    "metrpo_trpo_last_launches": (_i, [_vp]),
    "metrpo_trpo_process": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_trpo_fit_baseline": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _d, _vp, _vp]),
    "metrpo_trpo_update": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _d, _i, _d, _d, _i, _vp, _vp]),
    "metrpo_trpo_loss_kl": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "metrpo_trpo_grad": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _d, _vp, _vp]),
    "metrpo_fit_create": (_i, [ctypes.POINTER(FitCfg), ctypes.POINTER(_vp)]),
    "metrpo_fit_destroy": (_i, [_vp]),
    "metrpo_fit_num_params": (_i, [_vp]),
    "metrpo_fit_last_launches": (_i, [_vp]),
    "metrpo_fit_set_weights": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_fit_get_weights": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_fit_set_normalization": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_fit_reset_adam": (_i, [_vp, _vp]),
    "metrpo_fit_step": (_i, [_vp, _vp, _vp, _i, _vp, _i, _u64, _u64, _d, _vp, _vp]),
    "metrpo_fit_eval": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "metrpo_fit_restore_best": (_i, [_vp, _vp]),
}

# development library (include/metrpo_dev.h): descriptor self-test + issue-rate micro-benchmarks
DEV_LIB_PATH = os.path.join(_HERE, "libmetrpo_dev.so")
DEV_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "metrpo_dev.h")
_DEV_PROTOS = {
    "metrpo_last_error": (ctypes.c_char_p, []),
    "metrpo_bench_mma": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "metrpo_bench_mma_sync": (_i, [_i, _i, _i, _vp, _vp]),
    "metrpo_selftest_umma": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "metrpo_dev_gemm_tf32": (_i, [_i, _i, _i, _i, _vp, _ll, _ll, _i, _vp, _ll, _ll, _i, _vp, _ll, _ll,
                                  _i, _vp, _ll, _vp, _ll, _ll, _vp, _vp]),
}

_lib = None
_dev_lib = None


def declared_symbols(header=None):
    """Every function name include/metrpo.h (or the given header) declares."""
    with open(header or HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(metrpo_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the in-tree library (once) and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmetrpo.so is not built (%s). Run `python __graft_entry__.py`; there is no "
            "CPU fallback for the rollout path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError -> symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_dev():
    """dlopen the development library (self-test / micro-benchmarks); not used by the product."""
    global _dev_lib
    if _dev_lib is None:
        if not os.path.exists(DEV_LIB_PATH):
            raise RuntimeError("libmetrpo_dev.so is not built (%s). Run `python __graft_entry__.py`." % DEV_LIB_PATH)
        lib = ctypes.CDLL(DEV_LIB_PATH)
        for name, (res, args) in _DEV_PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _dev_lib = lib
    return _dev_lib


def check_dev(status, what=""):
    if status != 0:
        raise RuntimeError("%s failed (status %d): %s" % (what or "metrpo dev call", status,
                                                        load_dev().metrpo_last_error().decode("utf-8", "replace")))


def check_exports():
    """The library exports every symbol the header declares (CPU-only check, no compute)."""
    lib = load()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError("libmetrpo.so lacks symbols declared in metrpo.h: %s" % missing)
    undeclared = [s for s in declared_symbols() if s not in _PROTOS]
    if undeclared:
        raise RuntimeError("lib.py has no prototype for: %s" % undeclared)
    leaked = [s for s in _DEV_PROTOS if s != "metrpo_last_error" and hasattr(lib, s)]
    if leaked:
        raise RuntimeError("development symbols exported by the product library: %s" % leaked)
    if os.path.exists(DEV_LIB_PATH):
        dev = load_dev()
        missing = [s for s in declared_symbols(DEV_HEADER_PATH) if not hasattr(dev, s)]
        if missing:
            raise RuntimeError("libmetrpo_dev.so lacks symbols declared in metrpo_dev.h: %s" % missing)
    return declared_symbols()


def last_error():
    return load().metrpo_last_error().decode("utf-8", "replace")


def check(status, what=""):
    """Map a negative metrpo_status_t to RuntimeError (the reference raises Python exceptions)."""
    if status != 0:
        raise RuntimeError("%s failed (status %d): %s" % (what or "metrpo call", status, last_error()))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None, device=None):
    """Raw cudaStream_t of `stream`, or of the current stream of `device` (a handle's own device,
    not whatever device happens to be current)."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return ctypes.c_void_p(s.cuda_stream)
