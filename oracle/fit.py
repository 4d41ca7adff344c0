"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's ensemble
dynamics fit -- build_dynamics_graph (model_based_rl.py:23-103), get_dynamics_optimizer
(:154-183), optimize_models (:881-1051) and data_collection (utils.py:44-131).

PARITY: the prediction loss (build_dynamics_graph), get_ith_tensor and data_collection are PINNED
against the reference's own code executed under shims (tests/test_ref_fixtures.py sections C, G).
UNPINNED: the optimiser -- TensorFlow 1.4 cannot be imported here; tf.train.AdamOptimizer's update rule is restated from its documentation
(lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); theta -= lr_t * m / (sqrt(v) + epsilon));
tests/test_fit_oracle.py pins the hand-written backward pass against torch.autograd in float64.
"""
import numpy as np

from . import models as _models


def get_ith_tensor(tensor, i, sliced_length):          # utils.py:366-369
    assert tensor.shape[1] % sliced_length == 0
    return tensor[:, i * sliced_length:(i + 1) * sliced_length]


def minibatches(x_batch, y_batch, batch_size, K):
    """model_based_rl.py:966-969: a (batch*K)-row sample reshaped to (batch, K*(S+A)); model i reads
    column block i, i.e. rows i, i+K, i+2K, .. of the sample."""
    xf = np.reshape(x_batch, (batch_size, -1))
    yf = np.reshape(y_batch, (batch_size, -1))
    SA, S = x_batch.shape[1], y_batch.shape[1]
    return [(get_ith_tensor(xf, i, SA), get_ith_tensor(yf, i, S)) for i in range(K)]


def forward(m, norm, xu, S, drop, dtype=np.float32):
    """dynamics_model (training.py:218-269) keeping the intermediates the backward pass needs."""
    xu = xu.astype(dtype)
    z = ((xu - norm["in_mean"].astype(dtype)) / norm["in_std"].astype(dtype))[:, drop:]
    h0 = np.maximum(z @ m["W0"].astype(dtype) + m["b0"].astype(dtype), 0)
    h1 = np.maximum(h0 @ m["W1"].astype(dtype) + m["b1"].astype(dtype), 0)
    o = h1 @ m["W2"].astype(dtype) + m["b2"].astype(dtype)
    pred = (norm["diff_mean"].astype(dtype) + norm["diff_std"].astype(dtype) * o) + xu[:, :S]
    return pred, (z, h0, h1)


def prediction_loss(m, norm, xu, y, S, drop, dtype=np.float32):
    """tf.reduce_mean(tf.reduce_sum(tf.square(y_predicted - y), axis=[1]))  (:57-71)."""
    pred, _ = forward(m, norm, xu, S, drop, dtype)
    return dtype(np.mean(np.sum(np.square(pred - y.astype(dtype)), axis=1)))


def loss_and_grads(m, norm, xu, y, S, drop, dtype=np.float32):
    pred, (z, h0, h1) = forward(m, norm, xu, S, drop, dtype)
    B = len(xu)
    diff = pred - y.astype(dtype)
    loss = dtype(np.mean(np.sum(np.square(diff), axis=1)))
    dO = (dtype(2.0 / B) * norm["diff_std"].astype(dtype)) * diff
    g = dict(W2=h1.T @ dO, b2=dO.sum(0))
    dh1 = (dO @ m["W2"].astype(dtype).T) * (h1 > 0)
    g.update(W1=h0.T @ dh1, b1=dh1.sum(0))
    dh0 = (dh1 @ m["W1"].astype(dtype).T) * (h0 > 0)
    g.update(W0=z.T @ dh0, b0=dh0.sum(0))
    return loss, g


class Adam:
    """tf.train.AdamOptimizer(learning_rate) with TF defaults beta1=0.9, beta2=0.999, eps=1e-8."""

    def __init__(self, models, dtype=np.float32):
        self.dtype = dtype
        self.t = 0
        self.m = [{k: np.zeros_like(v, dtype) for k, v in mod.items()} for mod in models]
        self.v = [{k: np.zeros_like(v, dtype) for k, v in mod.items()} for mod in models]

    def apply(self, models, grads, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.t += 1
        dt = self.dtype
        lr_t = dt(lr * np.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t))
        for i, (mod, g) in enumerate(zip(models, grads)):
            for k in mod:
                gk = g[k].astype(dt)
                self.m[i][k] = dt(b1) * self.m[i][k] + dt(1 - b1) * gk
                self.v[i][k] = dt(b2) * self.v[i][k] + dt(1 - b2) * gk * gk
                mod[k] = (mod[k].astype(dt) - lr_t * self.m[i][k] / (np.sqrt(self.v[i][k]) + dt(eps))).astype(dt)


def train_step(models, adam, norm, x_data, y_data, idx, batch_size, lr, S, drop, dtype=np.float32):
    """One iteration of the training loop (:957-970) with the sample indices `idx` [batch*K] given
    (the reference draws them with np.random.uniform, utils.py:129-131).  Returns the K per-model
    training losses evaluated BEFORE the update (sess.run([opt_op, loss]) fetches both from the
    same forward pass)."""
    K = len(models)
    parts = minibatches(x_data[idx], y_data[idx], batch_size, K)
    losses, grads = [], []
    for mod, (xb, yb) in zip(models, parts):
        l, g = loss_and_grads(mod, norm, xb, yb, S, drop, dtype)
        losses.append(l)
        grads.append(g)
    adam.apply(models, grads, lr)
    return np.asarray(losses, dtype)


def validation_losses(models, norm, x_val, y_val, S, drop, dtype=np.float32):
    """dynamics_losses on np.tile(x_val, K) (:934-946): every model sees the whole set."""
    return np.asarray([prediction_loss(m, norm, x_val, y_val, S, drop, dtype) for m in models], dtype)


def optimize_models(models, norm, x_train, y_train, x_val, y_val, S, drop, batch_size, lr_scratch,
                    lr_refine, log_every, num_passes_threshold, max_passes, reinitialize, index_source,
                    dtype=np.float32):
    """optimize_models (:881-1051) for one scope.  `index_source(j, n)` returns the batch*K sample
    indices of iteration j.  Returns (models restored to their best snapshots, info dict)."""
    K = len(models)
    adam = Adam(models, dtype)
    lr = lr_scratch if reinitialize else lr_refine
    best = [{k: v.copy() for k, v in m.items()} for m in models]             # :925-930
    min_losses = validation_losses(models, norm, x_val, y_val, S, drop, dtype)
    min_sum = float(np.sum(min_losses))
    recover = np.zeros(K)
    refine_idx = -1
    n_data = len(x_train)
    iter_const = n_data / batch_size                                          # :954
    max_iters = int(max_passes * iter_const)
    log_it = int(log_every * iter_const)
    thresh = int(num_passes_threshold * iter_const)
    best_j, j = 0, 0
    val_hist = []
    for j in range(1, max_iters + 1):
        idx = index_source(j, n_data)
        train_step(models, adam, norm, x_train, y_train, idx, batch_size, lr, S, drop, dtype)
        if j % log_it == 0:
            vl = validation_losses(models, norm, x_val, y_val, S, drop, dtype)
            val_hist.append(float(np.sum(vl)))
            if min_sum > np.sum(vl):
                min_sum, best_j = float(np.sum(vl)), j
            upd = min_losses > vl
            min_losses[upd] = vl[upd]
            for i in np.nonzero(upd)[0]:
                best[i] = {k: v.copy() for k, v in models[i].items()}
                recover[i] = j
            if j - max(np.amax(recover), refine_idx) >= thresh:              # :1022-1031
                if reinitialize and refine_idx < 0 and lr_scratch > lr_refine:
                    for i in range(K):
                        models[i] = {k: v.copy() for k, v in best[i].items()}
                    lr = lr_refine
                    refine_idx = j
                    continue
                break
    for i in range(K):                                                        # :1034
        models[i] = {k: v.copy() for k, v in best[i].items()}
    return models, dict(n_updates=j, best_index=best_j, min_validation_losses=min_losses,
                        min_sum_validation_loss=min_sum, recover_indices=recover, validation_sums=val_hist)


class DataCollection:
    """utils.py:44-131 data_collection (FIFO-capped store + sampling with replacement)."""

    def __init__(self, max_size=int(5e4)):
        self.cur_idx, self.x, self.y, self.n_data, self.max_size = 0, None, None, None, max_size

    def cap_data_size(self):
        new_start_idx = self.x.shape[0] - self.max_size
        if new_start_idx > 0:
            self.x, self.y = self.x[new_start_idx:], self.y[new_start_idx:]
            self.n_data = self.max_size
            self.cur_idx -= new_start_idx

    def add_data(self, x_new, y_new):
        assert x_new.shape[0] == y_new.shape[0]
        if self.x is not None:
            self.cur_idx = self.x.shape[0]
            self.x = np.concatenate([self.x, x_new], axis=0)
            self.y = np.concatenate([self.y, y_new], axis=0)
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
            indices = list(range(start_idx, self.n_data)) + list(range(0, batch_size - (self.n_data - start_idx)))
            self.cur_idx = batch_size - (self.n_data - start_idx)
        else:
            indices = list(range(start_idx, end_idx))
            self.cur_idx = end_idx
        return self.x[indices, :], self.y[indices, :]

    def sample_indices(self, batch_size, rng):
        return np.floor(self.n_data * rng.uniform(0.0, 1.0, size=batch_size)).astype(np.intp)

    def sample(self, batch_size, rng):
        idx = self.sample_indices(batch_size, rng)
        return self.x[idx, :], self.y[idx, :]
