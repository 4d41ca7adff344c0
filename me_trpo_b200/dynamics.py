"""Ensemble dynamics fit: host mirrors of the reference's model-learning half around the
device fit (libmetrpo.so metrpo_fit_*, csrc/fit_kernels.cu).

  EnsembleFit       thin owner of a metrpo_fit_t handle (weights, Adam moments, best snapshots on
                    the device; one stream of kernels per training iteration for all K models)
  data_collection   utils.py:44-131 -- FIFO-capped (x, y) store, here resident on the GPU so that
                    minibatch gathering never crosses PCIe
  RunningMeanStd    running_mean_std.py:3-42 (cumulative sum / sumsq / count, std floor 0.1)
  optimize_models   model_based_rl.py:881-1051 -- validate every log_every passes, snapshot every
                    model whose own validation loss improved, scratch->refine lr drop, stop when no
                    model improved for num_passes_threshold passes, restore the best snapshots

No CPU fallback: construction fails without an sm_100 device.
"""
import ctypes

import numpy as np
import torch

from . import lib as _lib
from .synthetic import xavier_uniform


def _f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32).to(device).contiguous()


class EnsembleFit:
    WEIGHT_KEYS = ("W0", "b0", "W1", "b1", "W2", "b2")

    def __init__(self, state_dim, action_dim, drop_cols, hidden, n_models, max_rows=8192,
                 precision="tf32", device=None):
        self.S, self.A, self.drop, self.H, self.K = int(state_dim), int(action_dim), int(drop_cols), int(hidden), int(n_models)
        self.Din = self.S + self.A - self.drop
        self.max_rows = int(max_rows)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("EnsembleFit needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _lib.load()
        cfg = _lib.FitCfg()
        cfg.state_dim, cfg.action_dim, cfg.drop_cols = self.S, self.A, self.drop
        cfg.hidden, cfg.n_models, cfg.max_rows = self.H, self.K, self.max_rows
        cfg.precision = {"tf32": 0, "fp32": 1}[precision]
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.metrpo_fit_create(ctypes.byref(cfg), ctypes.byref(self._h)), "metrpo_fit_create")
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.metrpo_fit_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shapes(self):
        return [(self.Din, self.H), (self.H,), (self.H, self.H), (self.H,), (self.H, self.S), (self.S,)]

    def set_weights(self, k, m):
        ts = [_f32(m[key], self.device) for key in self.WEIGHT_KEYS]
        for t, s in zip(ts, self._shapes()):
            if tuple(t.shape) != s:
                raise ValueError("dynamics weight shape %s, expected %s" % (tuple(t.shape), s))
        _lib.check(self._lib.metrpo_fit_set_weights(self._h, int(k), *[_lib.ptr(t) for t in ts], _lib.stream_ptr(device=self.device)),
                   "fit_set_weights")
        self._keep = self._keep[-16:] + [ts]

    def set_ensemble(self, models):
        assert len(models) == self.K
        for k, m in enumerate(models):
            self.set_weights(k, m)

    def get_weights(self, k):
        ts = [torch.empty(s, device=self.device) for s in self._shapes()]
        _lib.check(self._lib.metrpo_fit_get_weights(self._h, int(k), *[_lib.ptr(t) for t in ts], _lib.stream_ptr(device=self.device)),
                   "fit_get_weights")
        return dict(zip(self.WEIGHT_KEYS, ts))

    def get_ensemble(self):
        return [self.get_weights(k) for k in range(self.K)]

    def set_normalization(self, in_mean, in_std, diff_mean, diff_std):
        ts = [_f32(t, self.device) for t in (in_mean, in_std, diff_mean, diff_std)]
        assert ts[0].numel() == self.S + self.A and ts[2].numel() == self.S
        _lib.check(self._lib.metrpo_fit_set_normalization(self._h, *[_lib.ptr(t) for t in ts], _lib.stream_ptr(device=self.device)),
                   "fit_set_normalization")
        self._keep = self._keep[-16:] + [ts]

    def reset_adam(self):
        _lib.check(self._lib.metrpo_fit_reset_adam(self._h, _lib.stream_ptr(device=self.device)), "fit_reset_adam")

    def step(self, x, y, batch, lr, idx=None, seed=0, offset=0, want_losses=True):
        """One Adam step of all K models.  x[n,S+A], y[n,S] device tensors; idx[batch*K] int32 sample
        indices or None (Philox on the device).  Returns device losses [K] (pre-update) or None."""
        assert x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32
        assert x.is_contiguous() and y.is_contiguous() and x.shape[1] == self.S + self.A and y.shape[1] == self.S
        n = int(x.shape[0])
        ix = None
        if idx is not None:
            ix = torch.as_tensor(idx, dtype=torch.int32).to(self.device).contiguous()
            assert ix.numel() == batch * self.K
        losses = torch.empty(self.K, device=self.device) if want_losses else None
        _lib.check(self._lib.metrpo_fit_step(self._h, _lib.ptr(x), _lib.ptr(y), n, _lib.ptr(ix), int(batch),
                                             int(seed), int(offset), float(lr), _lib.ptr(losses), _lib.stream_ptr(device=self.device)),
                   "fit_step")
        self._keep = self._keep[-16:] + [ix]
        return losses

    def eval(self, x, y, snapshot=0):
        """(losses [K], improved [K] uint8) device tensors; snapshot: 0 none, 1 save improved models,
        2 save all + initialise the minima."""
        assert x.is_cuda and y.is_cuda and x.is_contiguous() and y.is_contiguous()
        losses = torch.empty(self.K, device=self.device)
        improved = torch.empty(self.K, dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.metrpo_fit_eval(self._h, _lib.ptr(x), _lib.ptr(y), int(x.shape[0]), int(snapshot),
                                             _lib.ptr(losses), _lib.ptr(improved), _lib.stream_ptr(device=self.device)), "fit_eval")
        return losses, improved

    def restore_best(self):
        _lib.check(self._lib.metrpo_fit_restore_best(self._h, _lib.stream_ptr(device=self.device)), "fit_restore_best")

    def last_launches(self):
        return int(self._lib.metrpo_fit_last_launches(self._h))

    def num_params(self):
        return int(self._lib.metrpo_fit_num_params(self._h))


class data_collection:
    """utils.py:44-131 with x, y as device tensors (same FIFO cap, same cur_idx bookkeeping)."""

    def __init__(self, max_size=int(5e4), device="cuda"):
        self.cur_idx, self.x, self.y, self.n_data, self.max_size = 0, None, None, None, int(max_size)
        self.device = torch.device(device)

    def cap_data_size(self):
        new_start_idx = self.x.shape[0] - self.max_size
        if new_start_idx > 0:
            self.x = self.x[new_start_idx:].contiguous()
            self.y = self.y[new_start_idx:].contiguous()
            self.n_data = self.max_size
            self.cur_idx -= new_start_idx

    def set_data(self, x, y):
        x, y = _f32(x, self.device), _f32(y, self.device)
        assert x.shape[0] == y.shape[0]
        self.n_data, self.x, self.y = x.shape[0], x, y
        self.cur_idx %= self.n_data
        self.cap_data_size()

    def add_data(self, x_new, y_new):
        x_new, y_new = _f32(x_new, self.device), _f32(y_new, self.device)
        assert x_new.shape[0] == y_new.shape[0]
        if self.x is not None:
            self.cur_idx = self.x.shape[0]
            self.x = torch.cat([self.x, x_new], 0)
            self.y = torch.cat([self.y, y_new], 0)
        else:
            self.cur_idx, self.x, self.y = 0, x_new, y_new
        self.n_data = self.x.shape[0]
        self.cap_data_size()

    def get_num_data(self):
        return 0 if self.n_data is None else self.n_data

    def get_next_batch(self, batch_size):
        """utils.py:107-124: sequential window with wrap-around ('next_batch' sample_mode)."""
        assert batch_size <= self.n_data, \
            "Batch size %d is larger than n_data %d" % (batch_size, self.n_data)
        start_idx, end_idx = self.cur_idx, self.cur_idx + batch_size
        if end_idx > self.n_data:
            head = batch_size - (self.n_data - start_idx)
            idx = torch.cat([torch.arange(start_idx, self.n_data), torch.arange(0, head)])
            self.cur_idx = head
        else:
            idx = torch.arange(start_idx, end_idx)
            self.cur_idx = end_idx
        idx = idx.to(self.device)
        return self.x[idx], self.y[idx]

    def sample_indices(self, batch_size, rng=np.random):
        """Indices of `sample` (utils.py:129-131): uniform with replacement."""
        return np.floor(self.n_data * rng.uniform(0.0, 1.0, size=batch_size)).astype(np.int32)

    def sample(self, batch_size, rng=np.random):
        idx = torch.as_tensor(self.sample_indices(batch_size, rng), device=self.device, dtype=torch.long)
        return self.x[idx], self.y[idx]


class RunningMeanStd:
    """running_mean_std.py:3-42: fp32 cumulative sums; mean = sum/count,
    std = sqrt(max(sumsq/count - mean^2, 1e-2))."""

    def __init__(self, epsilon=1e-2, shape=(), device="cuda"):
        self.device = torch.device(device)
        self._sum = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self._sumsq = torch.full(shape, epsilon, dtype=torch.float32, device=self.device)
        self._count = torch.tensor(epsilon, dtype=torch.float32, device=self.device)

    @property
    def mean(self):
        return self._sum / self._count

    @property
    def std(self):
        return torch.sqrt(torch.clamp(self._sumsq / self._count - torch.square(self.mean), min=1e-2))

    def update(self, x):
        x = _f32(x, self.device)
        self._sum += x.sum(0)
        self._sumsq += torch.square(x).sum(0)
        self._count += float(len(x))


def add_rollout_data(x_all, y_all, dynamics_data, dynamics_validation, input_rms, output_rms, split_ratio):
    """The `use_same_dataset` branch of collect_data (model_based_rl.py:822-835): the first
    split_ratio fraction of the new transitions goes to validation, the rest to training, and the
    normalisers see the training part only."""
    x_all, y_all = np.asarray(x_all, np.float32), np.asarray(y_all, np.float32)
    n = int(round(split_ratio * len(x_all)))
    dynamics_validation.add_data(x_all[:n], y_all[:n])
    dynamics_data.add_data(x_all[n:], y_all[n:])
    input_rms.update(x_all[n:])
    output_rms.update(y_all[n:] - x_all[n:, :y_all.shape[1]])


def reinitialize_models(fit, rng):
    """sess.run(dynamics_initializer): Xavier-uniform W and b (training.py:179,187-194)."""
    for k in range(fit.K):
        shapes = dict(zip(fit.WEIGHT_KEYS, fit._shapes()))
        fit.set_weights(k, {key: xavier_uniform(rng, s) for key, s in shapes.items()})


def optimize_models(fit, dynamics_data, dynamics_validation, batch_size, learning_rate, log_every,
                    num_passes_threshold, max_passes, reinitialize, rng=None, index_source=None, seed=0,
                    logger=None):
    """optimize_models (model_based_rl.py:881-1051) for one scope.  learning_rate is the JSON dict
    {"scratch", "refine"} or a float.  Minibatch indices come from `index_source(j, n)` (parity
    tests), from `rng` (NumPy, like the reference) or -- default -- from Philox on the device, in
    which case a training iteration involves no host->device traffic at all.  The host reads back
    K floats + K flags only at validation points (every log_every passes)."""
    lr = learning_rate if isinstance(learning_rate, dict) else dict(scratch=learning_rate, refine=learning_rate)
    K = fit.K
    cur_lr = lr["scratch"] if reinitialize else lr["refine"]
    if reinitialize:                                                   # :906-912
        reinitialize_models(fit, rng if rng is not None else np.random.RandomState(seed))
    fit.reset_adam()                                                   # :913-918
    xv, yv = dynamics_validation.x, dynamics_validation.y
    xt, yt = dynamics_data.x, dynamics_data.y
    losses0, _ = fit.eval(xv, yv, snapshot=2)                          # :925-946
    min_validation_losses = losses0.cpu().numpy().astype(np.float32)
    min_sum_validation_loss = float(min_validation_losses.sum())
    recover_indices = np.zeros(K)
    refine_idx = -1
    n_data = dynamics_data.n_data
    iter_const = n_data / batch_size                                   # :954
    max_iters = int(max_passes * iter_const)
    log_it = max(1, int(log_every * iter_const))
    thresh = int(num_passes_threshold * iter_const)
    training_losses, validation_losses = [], []
    best_j, j = 0, 0
    for j in range(1, max_iters + 1):
        idx = None
        if index_source is not None:
            idx = index_source(j, n_data)
        elif rng is not None:
            idx = dynamics_data.sample_indices(batch_size * K, rng)
        want = (j % log_it == 0)
        tl = fit.step(xt, yt, batch_size, cur_lr, idx=idx, seed=seed, offset=j, want_losses=want)
        if want:
            vl_d, imp_d = fit.eval(xv, yv, snapshot=1)                 # :973-1007 (snapshot on device)
            vl = vl_d.cpu().numpy().astype(np.float32)
            imp = imp_d.cpu().numpy().astype(bool)
            training_losses.append(float(tl.sum().item()))
            validation_losses.append(float(vl.sum()))
            if min_sum_validation_loss > vl.sum():
                min_sum_validation_loss, best_j = float(vl.sum()), j
            min_validation_losses[imp] = vl[imp]
            recover_indices[imp] = j
            if logger:
                logger.info("iter %d val %.5f saved %d models" % (j, vl.sum(), int(imp.sum())))
            if j - max(np.amax(recover_indices), refine_idx) >= thresh:    # :1022-1031
                if reinitialize and refine_idx < 0 and lr["scratch"] > lr["refine"]:
                    fit.restore_best()
                    cur_lr = lr["refine"]
                    refine_idx = j
                    continue
                break
    fit.restore_best()                                                 # :1034
    return {"training_losses": training_losses, "validation_losses": validation_losses,
            "best_index": best_j, "n_updates": j, "min_validation_losses": min_validation_losses,
            "min_sum_validation_loss": min_sum_validation_loss, "recover_indices": recover_indices}
