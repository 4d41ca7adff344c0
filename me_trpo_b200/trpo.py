"""PolicyUpdate: thin Python owner of a metrpo_trpo_t handle (include/metrpo.h, TRPO section).

Torch owns every tensor; this class passes raw device pointers + the current stream through
ctypes.  It backs the reference-shaped classes in me_trpo_b200/algos/ (BatchPolopt / NPO / TRPO /
ConjugateGradientOptimizer) and BaseSampler.process_samples_flat.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import lib as _lib


class PolicyUpdate:
    def __init__(self, policy_dims, out_tanh=False, device=None):
        self.dims = [int(d) for d in policy_dims]
        self.S, self.A = self.dims[0], self.dims[-1]
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("PolicyUpdate needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        cfg = _lib.TrpoCfg()
        cfg.state_dim, cfg.action_dim = self.S, self.A
        cfg.n_policy_layers = len(self.dims) - 1
        if cfg.n_policy_layers > _lib.MAX_POLICY_LAYERS:
            raise RuntimeError("policy has too many layers for this build")
        for i, d in enumerate(self.dims):
            cfg.policy_dims[i] = d
        cfg.policy_out_tanh = 1 if out_tanh else 0
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.metrpo_trpo_create(ctypes.byref(cfg), ctypes.byref(self._h)), "metrpo_trpo_create")
        self.P = int(self._lib.metrpo_trpo_num_params(self._h))
        self._cb = None
        self.allreduce_mode = None      # None (single GPU) | "p2p" | "nccl"
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.metrpo_trpo_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_pass_impl(self, impl):
        """'auto' (fastest measured: SIMT today), 'simt' (fp32 CUDA cores), 'tf32' / 'tf32x3'
        (warp-level tensor-core MMAs, single / split products)."""
        _lib.check(self._lib.metrpo_trpo_set_pass_impl(self._h, {"auto": 0, "simt": 1, "tf32": 2, "tf32x3": 3}[impl]),
                   "trpo_set_pass_impl")

    # -- multi-GPU: sum the (tiny) accumulators across ranks with torch.distributed -------------
    def enable_allreduce(self, group=None, mode="auto"):
        """Every reduction of the update (gradient, each Fisher-vector product, (loss, kl) pairs,
        advantage moments, baseline normal equations) is summed over the process group; rows are
        sharded across ranks.

        mode "p2p": the library's own one-shot all-reduce over NVLink peer memory (one small kernel
        per reduction, exchange buffers mapped through CUDA IPC; single node, summed in rank order so
        that every rank gets the bit-identical result); "nccl": a host callback into
        torch.distributed per reduction; "auto": p2p when the peer buffers can be mapped, else nccl."""
        import torch.distributed as dist
        dev = self.device
        if mode in ("auto", "p2p"):
            ok = self._enable_p2p(group)
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks or none
            if int(flag.item()) == 1:
                self.allreduce_mode = "p2p"
                return
            if mode == "p2p":
                raise RuntimeError("peer-memory all-reduce unavailable: %s" % _lib.last_error())
            self._lib.metrpo_trpo_enable_p2p(self._h, 0, 1, ctypes.create_string_buffer(64))   # back to world 1
        self.allreduce_mode = "nccl"

        def _ar(user, buf, n, stream):
            try:
                t = _wrap_device_doubles(buf, n, dev)
                dist.all_reduce(t, group=group)
                return 0
            except Exception:   # never raise through the C ABI
                return 1

        self._cb = _lib.ALLREDUCE_FN(_ar)
        _lib.check(self._lib.metrpo_trpo_set_allreduce(self._h, self._cb, None), "set_allreduce")

    def _enable_p2p(self, group):
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = ctypes.create_string_buffer(64)
        if self._lib.metrpo_trpo_p2p_handle(self._h, mine) != 0:
            handles = [None] * world
            dist.all_gather_object(handles, None, group=group)
            return False
        handles = [None] * world
        dist.all_gather_object(handles, mine.raw, group=group)
        if any(hd is None for hd in handles):
            return False
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * world)
        return self._lib.metrpo_trpo_enable_p2p(self._h, rank, world, blob) == 0

    # -- R10 ------------------------------------------------------------------------------------
    def process(self, obs, rew, done, baseline_coeffs=None, discount=1.0, gae_lambda=1.0,
                center_adv=True, positive_adv=False):
        """obs[T,B,S], rew[T,B], done[T,B] (uint8) device tensors -> dict(adv, ret, valid, stats)."""
        T, B = rew.shape
        dev = self.device
        if positive_adv and self.allreduce_mode is not None:
            # the shift needs the GLOBAL minimum; the accumulator all-reduce is a sum
            raise NotImplementedError("positive_adv with row-sharded samples (sum all-reduce only); "
                                      "the reference's TRPO path never sets it (algos/batch_polopt.py:33)")
        assert obs.is_contiguous() and rew.is_contiguous() and done.is_contiguous()
        assert obs.dtype == torch.float32 and rew.dtype == torch.float32 and done.dtype == torch.uint8
        adv = torch.empty(T, B, device=dev)
        ret = torch.empty(T, B, device=dev)
        valid = torch.empty(T, B, dtype=torch.uint8, device=dev)
        stats = torch.empty(8, dtype=torch.float64, device=dev)
        co = None if baseline_coeffs is None else torch.as_tensor(baseline_coeffs, dtype=torch.float64).to(dev).contiguous()
        _lib.check(self._lib.metrpo_trpo_process(
            self._h, int(T), int(B), _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(done), _lib.ptr(co),
            float(discount), float(gae_lambda), 1 if center_adv else 0, 1 if positive_adv else 0,
            _lib.ptr(adv), _lib.ptr(ret), _lib.ptr(valid), _lib.ptr(stats), _lib.stream_ptr(device=self.device)), "trpo_process")
        self._keep = self._keep[-4:] + [co]
        return dict(adv=adv, ret=ret, valid=valid, stats=stats)

    def fit_baseline(self, obs, ret, valid, done, reg_coeff=1e-5):
        T, B = ret.shape
        coeffs = torch.empty(2 * self.S + 4, dtype=torch.float64, device=self.device)
        _lib.check(self._lib.metrpo_trpo_fit_baseline(
            self._h, int(T), int(B), _lib.ptr(obs), _lib.ptr(ret), _lib.ptr(valid), _lib.ptr(done),
            float(reg_coeff), _lib.ptr(coeffs), _lib.stream_ptr(device=self.device)), "trpo_fit_baseline")
        return coeffs

    # -- R11 ------------------------------------------------------------------------------------
    def _inputs(self, obs, act, adv, old_mean, old_log_std, valid):
        N = adv.numel()
        for t in (obs, act, adv, old_mean, old_log_std):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        assert obs.numel() == N * self.S and act.numel() == N * self.A and old_mean.numel() == N * self.A
        per_sample = 0 if old_log_std.numel() == self.A else 1
        if per_sample:
            assert old_log_std.numel() == N * self.A
        if valid is not None:
            assert valid.dtype == torch.uint8 and valid.numel() == N and valid.is_contiguous()
        return N, per_sample

    def update(self, theta, obs, act, adv, old_mean, old_log_std, valid=None, step_size=0.01,
               cg_iters=10, reg_coeff=1e-5, backtrack_ratio=0.8, max_backtracks=15):
        """In-place natural-gradient step on the flat parameter vector theta[P] (device fp32).
        Returns the 8-double info tensor (device; reading it synchronises)."""
        N, per_sample = self._inputs(obs, act, adv, old_mean, old_log_std, valid)
        assert theta.is_cuda and theta.dtype == torch.float32 and theta.numel() == self.P and theta.is_contiguous()
        info = torch.zeros(8, dtype=torch.float64, device=self.device)
        _lib.check(self._lib.metrpo_trpo_update(
            self._h, int(N), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(adv), _lib.ptr(old_mean),
            _lib.ptr(old_log_std), per_sample, _lib.ptr(valid), _lib.ptr(theta), float(step_size),
            int(cg_iters), float(reg_coeff), float(backtrack_ratio), int(max_backtracks), _lib.ptr(info),
            _lib.stream_ptr(device=self.device)), "trpo_update")
        return info

    def loss_kl(self, theta, obs, act, adv, old_mean, old_log_std, valid=None):
        N, per_sample = self._inputs(obs, act, adv, old_mean, old_log_std, valid)
        out = (ctypes.c_double * 2)()
        _lib.check(self._lib.metrpo_trpo_loss_kl(
            self._h, int(N), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(adv), _lib.ptr(old_mean),
            _lib.ptr(old_log_std), per_sample, _lib.ptr(valid), _lib.ptr(theta), out, _lib.stream_ptr(device=self.device)),
            "trpo_loss_kl")
        return float(out[0]), float(out[1])

    def grad(self, theta, obs, act, adv, old_mean, old_log_std, valid=None, vec=None, reg_coeff=0.0):
        """vec is None: flat gradient of surr_loss; else Hx(vec).  Host float64 array [P]."""
        N, per_sample = self._inputs(obs, act, adv, old_mean, old_log_std, valid)
        out = np.zeros(self.P, np.float64)
        _lib.check(self._lib.metrpo_trpo_grad(
            self._h, int(N), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(adv), _lib.ptr(old_mean),
            _lib.ptr(old_log_std), per_sample, _lib.ptr(valid), _lib.ptr(theta), _lib.ptr(vec),
            float(reg_coeff), out.ctypes.data_as(ctypes.c_void_p), _lib.stream_ptr(device=self.device)), "trpo_grad")
        return out

    def last_launches(self):
        return int(self._lib.metrpo_trpo_last_launches(self._h))


def _wrap_device_doubles(addr, n, device):
    """torch view of n doubles at a raw device address (library-owned accumulator)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(addr), False), version=3)
    return torch.as_tensor(h, device=device)
