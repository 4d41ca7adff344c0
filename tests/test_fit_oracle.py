"""CPU checks that pin the ensemble-fit oracle (oracle/fit.py): hand-written backward pass against
torch.autograd in float64, Adam against a direct transcription of TF's documented update, the
minibatch column-block layout of the reference's reshape, and data_collection's FIFO cap."""
import numpy as np
import pytest

from oracle import fit as of
from oracle import models as om

torch = pytest.importorskip("torch")


def _problem(seed=0, S=5, A=2, drop=1, H=16, K=3, n=64):
    rng = np.random.RandomState(seed)
    models = om.init_dynamics(rng, S, A, drop, H, K, out_scale=1.0)
    norm = dict(in_mean=rng.normal(0, 0.2, S + A).astype(np.float32),
                in_std=rng.uniform(0.5, 1.5, S + A).astype(np.float32),
                diff_mean=rng.normal(0, 0.1, S).astype(np.float32),
                diff_std=rng.uniform(0.1, 0.5, S).astype(np.float32))
    x = rng.normal(0, 1, (n, S + A)).astype(np.float32)
    y = (x[:, :S] + rng.normal(0, 0.1, (n, S))).astype(np.float32)
    return models, norm, x, y, S, drop


def test_backward_matches_autograd_float64():
    models, norm, x, y, S, drop = _problem()
    m = models[0]
    loss, g = of.loss_and_grads(m, norm, x, y, S, drop, np.float64)
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in m.items()}
    nt = {k: torch.tensor(v, dtype=torch.float64) for k, v in norm.items()}
    xu = torch.tensor(x, dtype=torch.float64)
    z = ((xu - nt["in_mean"]) / nt["in_std"])[:, drop:]
    h = torch.relu(z @ t["W0"] + t["b0"])
    h = torch.relu(h @ t["W1"] + t["b1"])
    pred = nt["diff_mean"] + nt["diff_std"] * (h @ t["W2"] + t["b2"]) + xu[:, :S]
    L = ((pred - torch.tensor(y, dtype=torch.float64)) ** 2).sum(1).mean()
    L.backward()
    assert abs(L.item() - loss) <= 1e-12 * max(1, abs(loss))
    for k in m:
        np.testing.assert_allclose(g[k], t[k].grad.numpy(), rtol=1e-10, atol=1e-12)


def test_adam_is_tf_formulation():
    rng = np.random.RandomState(1)
    w = rng.normal(size=(4, 3))
    models = [dict(W=w.copy())]
    adam = of.Adam(models, np.float64)
    m = np.zeros_like(w); v = np.zeros_like(w); ref = w.copy()
    for t in range(1, 6):
        g = rng.normal(size=w.shape)
        adam.apply(models, [dict(W=g)], 1e-3)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        ref = ref - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        np.testing.assert_allclose(models[0]["W"], ref, rtol=1e-13)
    # first step moves every weight by ~lr against the gradient sign
    models = [dict(W=w.copy())]
    of.Adam(models, np.float64).apply(models, [dict(W=np.ones_like(w))], 1e-3)
    np.testing.assert_allclose(w - models[0]["W"], 1e-3, rtol=1e-6)


def test_minibatch_layout_is_the_reference_reshape():
    K, batch, SA, S = 3, 4, 5, 2
    x = np.arange(K * batch * SA, dtype=np.float32).reshape(K * batch, SA)
    y = np.arange(K * batch * S, dtype=np.float32).reshape(K * batch, S)
    parts = of.minibatches(x, y, batch, K)
    for i in range(K):
        np.testing.assert_array_equal(parts[i][0], x[i::K])
        np.testing.assert_array_equal(parts[i][1], y[i::K])


def test_train_step_reduces_loss_and_optimize_models_restores_best():
    models, norm, x, y, S, drop = _problem(n=256, K=2)
    rng = np.random.RandomState(5)
    adam = of.Adam(models)
    l0 = of.validation_losses(models, norm, x, y, S, drop)
    for j in range(60):
        of.train_step(models, adam, norm, x, y, rng.randint(0, len(x), 32 * 2), 32, 1e-3, S, drop)
    l1 = of.validation_losses(models, norm, x, y, S, drop)
    assert np.all(l1 < l0)
    models2, norm, x, y, S, drop = _problem(n=256, K=2)
    rng = np.random.RandomState(6)
    out, info = of.optimize_models(models2, norm, x[:192], y[:192], x[192:], y[192:], S, drop, batch_size=32,
                                   lr_scratch=1e-3, lr_refine=1e-3, log_every=1, num_passes_threshold=2,
                                   max_passes=8, reinitialize=True,
                                   index_source=lambda j, n: rng.randint(0, n, 64))
    vl = of.validation_losses(out, norm, x[192:], y[192:], S, drop)
    np.testing.assert_allclose(vl, info["min_validation_losses"], rtol=1e-6)


def test_data_collection_fifo_cap_and_sampling():
    dc = of.DataCollection(max_size=10)
    dc.add_data(np.arange(8, dtype=np.float32)[:, None], np.arange(8, dtype=np.float32)[:, None])
    dc.add_data(np.arange(8, 14, dtype=np.float32)[:, None], np.arange(8, 14, dtype=np.float32)[:, None])
    assert dc.get_num_data() == 10 and dc.x[0, 0] == 4 and dc.x[-1, 0] == 13   # oldest rows dropped
    idx = dc.sample_indices(1000, np.random.RandomState(0))
    assert idx.min() >= 0 and idx.max() <= 9 and len(np.unique(idx)) == 10
