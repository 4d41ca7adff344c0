"""EnsembleRollout: thin Python owner of a metrpo_rollout_t handle (include/metrpo.h).

Torch owns every tensor; this class only passes raw device pointers + the current stream through
ctypes.  It is what NeuralNetEnv / VecSimpleEnv / VectorizedSampler (the mirrors of the reference
sockets) are built on.  No CPU fallback exists: construction fails on a non-sm_100 device.
"""
import ctypes

import torch

from . import lib as _lib
from .envs import ENV_SPECS, canonical_env_name


def _f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32).to(device).contiguous()


class EnsembleRollout:
    """One imaginary vec-env: K dynamics MLPs + Gaussian MLP policy + analytic reward/done for
    `n_envs` parallel rows (NeuralNetEnv + VecSimpleEnv, env_helpers.py:532-635)."""

    def __init__(self, env, n_models, n_envs, max_path_length, hidden=None, policy_hidden=None,
                 sam_mode="step_rand", drop_cols=None, policy_out_tanh=False, device=None,
                 state_dim=None, action_dim=None, row_offset=0, precision="bf16"):
        name = canonical_env_name(env)
        spec = ENV_SPECS[name]
        self.env_name = name
        self.S = int(state_dim or spec["S"])
        self.A = int(action_dim or spec["A"])
        self.drop = int(spec["drop"] if drop_cols is None else drop_cols)
        self.hidden = int(hidden or spec["hidden"])
        self.policy_hidden = tuple(policy_hidden if policy_hidden is not None else spec["policy_hidden"])
        self.K, self.B, self.T_max = int(n_models), int(n_envs), int(max_path_length)
        self.sam_mode = sam_mode
        if sam_mode not in _lib.SAM_MODES:
            raise AssertionError("sam mode %s is not defined." % sam_mode)   # env_helpers.py:634
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("EnsembleRollout needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        cfg = _lib.RolloutCfg()
        cfg.state_dim, cfg.action_dim, cfg.drop_cols = self.S, self.A, self.drop
        cfg.hidden, cfg.n_models, cfg.n_envs = self.hidden, self.K, self.B
        cfg.max_path_length = self.T_max
        cfg.env_id = _lib.ENV_IDS[name]
        cfg.sam_mode = _lib.SAM_MODES[sam_mode]
        dims = [self.S] + list(self.policy_hidden) + [self.A]
        if len(dims) - 1 > _lib.MAX_POLICY_LAYERS:
            raise RuntimeError("policy has too many layers for this build")
        cfg.n_policy_layers = len(dims) - 1
        for i, d in enumerate(dims):
            cfg.policy_dims[i] = d
        cfg.policy_out_tanh = 1 if policy_out_tanh else 0
        # "bf16": tcgen05 tensor cores (bf16 operands, fp32 accumulate); "fp32": the reference's fp32
        # arithmetic on CUDA cores, ~50x slower (fidelity mode, include/metrpo.h METRPO_PREC_FP32)
        self.precision = precision
        cfg.precision = {"bf16": 0, "fp32": 1}[precision]
        cfg.row_offset = int(row_offset)
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.metrpo_rollout_create(ctypes.byref(cfg), ctypes.byref(self._h)),
                   "metrpo_rollout_create")
        self._keep = []          # tensors that must outlive async launches
        self.log_std = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.metrpo_rollout_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters -------------------------------------------------------------------------
    def set_dynamics(self, k, W0, b0, W1, b1, W2, b2):
        """Weights of model k in the reference's TF layout W[in,out] (training.py:187-208)."""
        ts = [_f32(t, self.device) for t in (W0, b0, W1, b1, W2, b2)]
        din = self.S + self.A - self.drop
        shapes = [(din, self.hidden), (self.hidden,), (self.hidden, self.hidden), (self.hidden,),
                  (self.hidden, self.S), (self.S,)]
        for t, s in zip(ts, shapes):
            if tuple(t.shape) != s:
                raise ValueError("dynamics weight shape %s, expected %s" % (tuple(t.shape), s))
        _lib.check(self._lib.metrpo_rollout_set_dynamics(self._h, int(k), *[_lib.ptr(t) for t in ts],
                                                         _lib.stream_ptr(device=self.device)), "set_dynamics")
        self._keep.append(ts)

    def set_dynamics_ensemble(self, models):
        assert len(models) == self.K
        for k, m in enumerate(models):
            self.set_dynamics(k, m["W0"], m["b0"], m["W1"], m["b1"], m["W2"], m["b2"])

    def set_normalization(self, in_mean, in_std, diff_mean, diff_std):
        ts = [_f32(t, self.device) for t in (in_mean, in_std, diff_mean, diff_std)]
        assert ts[0].numel() == self.S + self.A and ts[2].numel() == self.S
        _lib.check(self._lib.metrpo_rollout_set_normalization(self._h, *[_lib.ptr(t) for t in ts],
                                                              _lib.stream_ptr(device=self.device)), "set_normalization")
        self._keep.append(ts)

    def set_policy(self, Ws, bs, log_std):
        n = len(self.policy_hidden) + 1
        assert len(Ws) == n and len(bs) == n
        Ws = [_f32(w, self.device) for w in Ws]
        bs = [_f32(b, self.device) for b in bs]
        ls = _f32(log_std, self.device)
        Wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in Ws])
        bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bs])
        _lib.check(self._lib.metrpo_rollout_set_policy(self._h, Wp, bp, _lib.ptr(ls), _lib.stream_ptr(device=self.device)),
                   "set_policy")
        self._keep.append((Ws, bs, ls))
        self.log_std = ls

    # -- B1: step-granular vec env ------------------------------------------------------------
    def reset(self, states):
        st = _f32(states, self.device)
        assert tuple(st.shape) == (self.B, self.S)
        _lib.check(self._lib.metrpo_rollout_reset(self._h, _lib.ptr(st), _lib.stream_ptr(device=self.device)), "reset")
        self._keep.append(st)

    def set_rows(self, rows, states):
        """Overwrite the states of the given rows (step counters untouched): the done rows' fresh
        simulator resets of the step-granular socket."""
        dev = self.device
        idx = torch.as_tensor(rows, dtype=torch.int32).to(dev).contiguous()
        st = _f32(states, dev)
        assert st.dim() == 2 and st.shape[0] == idx.numel() and st.shape[1] == self.S
        _lib.check(self._lib.metrpo_rollout_set_rows(self._h, _lib.ptr(idx), int(idx.numel()), _lib.ptr(st),
                                                     _lib.stream_ptr(device=self.device)), "set_rows")
        self._keep = self._keep[-8:] + [(idx, st)]

    def step(self, actions, reset_states, model_idx=None, std_noise=None, seed=0, offset=0):
        dev = self.device
        act = _f32(actions, dev)
        rs = _f32(reset_states, dev)
        mi = None if model_idx is None else torch.as_tensor(model_idx, dtype=torch.int32).to(dev).contiguous()
        sn = None if std_noise is None else _f32(std_noise, dev)
        obs = torch.empty(self.B, self.S, device=dev)
        rew = torch.empty(self.B, device=dev)
        done = torch.empty(self.B, dtype=torch.uint8, device=dev)
        _lib.check(self._lib.metrpo_rollout_step(self._h, _lib.ptr(act), _lib.ptr(mi), _lib.ptr(sn),
                                                 _lib.ptr(rs), int(seed), int(offset), _lib.ptr(obs),
                                                 _lib.ptr(rew), _lib.ptr(done), _lib.stream_ptr(device=self.device)), "step")
        self._keep = self._keep[-8:] + [(act, rs, mi, sn)]
        return obs, rew, done

    # -- B2: whole-horizon fused rollout ------------------------------------------------------
    def run(self, n_steps, init_states, reset_pool, eps=None, model_idx=None, std_noise=None, seed=0,
            offset=0, determ=False, out=None, want=("obs", "act", "mean", "rew", "done")):
        """Returns dict of time-major device tensors obs[T,B,S], act[T,B,A] (unclipped), mean[T,B,A],
        rew[T,B], done[T,B] (uint8) and final_states[B,S]."""
        dev, T, B, S, A = self.device, int(n_steps), self.B, self.S, self.A
        init = _f32(init_states, dev)
        pool = _f32(reset_pool, dev)
        assert tuple(init.shape) == (B, S) and pool.dim() == 2 and pool.shape[1] == S
        ep = None if eps is None else _f32(eps, dev)
        mi = None if model_idx is None else torch.as_tensor(model_idx, dtype=torch.int32).to(dev).contiguous()
        sn = None if std_noise is None else _f32(std_noise, dev)
        if ep is not None:
            assert tuple(ep.shape) == (T, B, A)
        if mi is not None:
            assert tuple(mi.shape) == (T, B)
        if out is None:
            out = {}
        shapes = dict(obs=(T, B, S), act=(T, B, A), mean=(T, B, A), rew=(T, B), done=(T, B))
        for name in want:
            if name not in out:
                dt = torch.uint8 if name == "done" else torch.float32
                out[name] = torch.empty(shapes[name], dtype=dt, device=dev)
        if "final_states" not in out:
            out["final_states"] = torch.empty(B, S, device=dev)
        g = lambda n: _lib.ptr(out.get(n)) if n in want else None
        _lib.check(self._lib.metrpo_rollout_run(
            self._h, T, _lib.ptr(init), _lib.ptr(pool), int(pool.shape[0]), _lib.ptr(ep), _lib.ptr(mi),
            _lib.ptr(sn), int(seed), int(offset), 1 if determ else 0, g("obs"), g("act"), g("mean"),
            g("rew"), g("done"), _lib.ptr(out["final_states"]), _lib.stream_ptr(device=self.device)), "run")
        self._keep = self._keep[-8:] + [(init, pool, ep, mi, sn)]
        return out

    def run_continue(self, n_steps, reset_pool, eps=None, model_idx=None, std_noise=None, seed=0, offset=0,
                     determ=False, out=None, want=("obs", "act", "mean", "rew", "done")):
        """metrpo_rollout_continue: n_steps more steps from the row state the previous call left
        (states, step counters, reset counts); `offset` = noise-stream index of its first step."""
        dev, T, B, S, A = self.device, int(n_steps), self.B, self.S, self.A
        pool = _f32(reset_pool, dev)
        ep = None if eps is None else _f32(eps, dev)
        mi = None if model_idx is None else torch.as_tensor(model_idx, dtype=torch.int32).to(dev).contiguous()
        sn = None if std_noise is None else _f32(std_noise, dev)
        out = {} if out is None else out
        shapes = dict(obs=(T, B, S), act=(T, B, A), mean=(T, B, A), rew=(T, B), done=(T, B))
        for name in want:
            if name not in out:
                out[name] = torch.empty(shapes[name], dtype=torch.uint8 if name == "done" else torch.float32, device=dev)
            assert out[name].is_contiguous() and tuple(out[name].shape) == shapes[name]
        if "final_states" not in out:
            out["final_states"] = torch.empty(B, S, device=dev)
        g = lambda n: _lib.ptr(out.get(n)) if n in want else None
        _lib.check(self._lib.metrpo_rollout_continue(
            self._h, T, _lib.ptr(pool), int(pool.shape[0]), _lib.ptr(ep), _lib.ptr(mi), _lib.ptr(sn), int(seed),
            int(offset), 1 if determ else 0, g("obs"), g("act"), g("mean"), g("rew"), g("done"),
            _lib.ptr(out["final_states"]), _lib.stream_ptr(device=self.device)), "continue")
        self._keep = self._keep[-8:] + [(pool, ep, mi, sn)]
        return out

    def run_to_host(self, n_steps, init_states, reset_pool, host_out=None, dev_out=None, n_chunks=4,
                    eps=None, model_idx=None, std_noise=None, seed=0, offset=0, determ=False,
                    want=("obs", "act", "mean", "rew", "done")):
        """run() whose trajectory lands in pinned HOST buffers: the horizon is cut into n_chunks
        launches (metrpo_rollout_continue) and the device->host copy of chunk c runs on a second
        stream while chunk c+1 is computed, so only the last (short) chunk's copy is exposed.  Results are
        identical to run().  Returns (host_out, dev_out); host buffers are valid after
        `synchronize()`."""
        dev, T, B, S, A = self.device, int(n_steps), self.B, self.S, self.A
        init = _f32(init_states, dev)
        pool = _f32(reset_pool, dev)
        ep = None if eps is None else _f32(eps, dev)
        mi = None if model_idx is None else torch.as_tensor(model_idx, dtype=torch.int32).to(dev).contiguous()
        sn = None if std_noise is None else _f32(std_noise, dev)
        shapes = dict(obs=(T, B, S), act=(T, B, A), mean=(T, B, A), rew=(T, B), done=(T, B))
        dev_out = {} if dev_out is None else dev_out
        host_out = {} if host_out is None else host_out
        for name in list(want) + ["final_states"]:
            shp = (B, S) if name == "final_states" else shapes[name]
            dt = torch.uint8 if name == "done" else torch.float32
            if name not in dev_out:
                dev_out[name] = torch.empty(shp, dtype=dt, device=dev)
            if name not in host_out:
                host_out[name] = torch.empty(shp, dtype=dt).pin_memory()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main, side = torch.cuda.current_stream(dev), self._copy_stream
        side.wait_stream(main)          # earlier copies out of dev_out / into host_out are ordered
        n_chunks = max(1, min(int(n_chunks), T))
        # Only the LAST chunk's device->host copy is exposed (a chunk's copy is ~5x faster than its
        # compute), so the tail is tapered: the last chunk is T/40 steps, the one before 3T/40, the rest equal.
        tail = max(1, T // 40)
        if n_chunks >= 3 and T - 4 * tail >= n_chunks - 2:
            head = T - 4 * tail
            bounds = [round(i * head / (n_chunks - 2)) for i in range(n_chunks - 1)] + [T - tail, T]
        else:
            bounds = [round(i * T / n_chunks) for i in range(n_chunks + 1)]
        sl = lambda t, a, b: None if t is None else t[a:b]
        for c in range(n_chunks):
            t0, t1 = bounds[c], bounds[c + 1]
            g = lambda n: _lib.ptr(dev_out[n][t0:t1]) if n in want else None
            common = (_lib.ptr(pool), int(pool.shape[0]), _lib.ptr(sl(ep, t0, t1)), _lib.ptr(sl(mi, t0, t1)),
                      _lib.ptr(sl(sn, t0, t1)), int(seed), int(offset) + t0, 1 if determ else 0, g("obs"), g("act"),
                      g("mean"), g("rew"), g("done"), _lib.ptr(dev_out["final_states"]), _lib.stream_ptr(main))
            if c == 0:
                _lib.check(self._lib.metrpo_rollout_run(self._h, t1 - t0, _lib.ptr(init), *common), "run")
            else:
                _lib.check(self._lib.metrpo_rollout_continue(self._h, t1 - t0, *common), "continue")
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for n in want:
                    host_out[n][t0:t1].copy_(dev_out[n][t0:t1], non_blocking=True)
                if c == n_chunks - 1:
                    host_out["final_states"].copy_(dev_out["final_states"], non_blocking=True)
        main.wait_stream(side)          # later work on the main stream sees the copies complete
        self._keep = self._keep[-8:] + [(init, pool, ep, mi, sn)]
        return host_out, dev_out

    # -- R12: per-model validation cost (build_policy_graph, model_based_rl.py:122-142) --------
    def model_costs(self, n_steps, init_states, gamma=1.0, return_rows=False):
        """[K] discounted cost of the deterministic policy under each model rolled forward on its
        own predictions from init_states[n_rows,S] (n_rows <= n_envs)."""
        dev = self.device
        init = _f32(init_states, dev)
        n = int(init.shape[0])
        assert init.dim() == 2 and init.shape[1] == self.S and 1 <= n <= self.B
        rows = torch.empty(self.K, n, device=dev)
        costs = torch.empty(self.K, device=dev)
        _lib.check(self._lib.metrpo_rollout_model_costs(
            self._h, int(n_steps), n, _lib.ptr(init), float(gamma), _lib.ptr(rows), _lib.ptr(costs),
            _lib.stream_ptr(device=self.device)), "model_costs")
        self._keep = self._keep[-8:] + [(init,)]
        return (costs, rows) if return_rows else costs

    def synchronize(self):
        """Wait for the stream and raise if the last kernel aborted on an internal wait timeout."""
        _lib.check(self._lib.metrpo_rollout_status(self._h, _lib.stream_ptr(device=self.device)), "rollout kernel")

    def last_launches(self):
        return int(self._lib.metrpo_rollout_last_launches(self._h))

    def last_kernel(self):
        """0: single-stream kernel, 1 / 2: two-stream kernel without / with column split."""
        return int(self._lib.metrpo_rollout_last_kernel(self._h))
