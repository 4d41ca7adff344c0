"""GPU parity of the ensemble dynamics fit (metrpo_fit_*, through the C ABI) against oracle/fit.py.

Stated tolerances:
  precision "fp32" (true fp32 GEMMs, the reference's arithmetic): losses 1e-5 relative; weights
      after 5 Adam steps 2e-5 absolute (lr 1e-3 => steps of ~1e-3 per weight, so this is 2 % of ONE
      step; Adam's m/sqrt(v) normalisation amplifies rounding of tiny gradients, hence not tighter);
  precision "tf32" (tensor cores, 10-bit operand mantissa, fp32 accumulate): losses 2e-3 relative;
      weights after 5 steps: RMS deviation <= 10 % of the RMS update (measured 1-4 %).  A max-abs
      bound is meaningless here: Adam's m/sqrt(v) turns the sign of a rounding-level gradient into
      a full lr-sized step for the few weights whose gradient is ~0; validation loss of the trained
      models within 2e-3 relative of the oracle's.
"""
import numpy as np
import pytest

from oracle import fit as of
from oracle import models as om

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _problem(seed=0, S=18, A=6, drop=1, H=256, K=3, n=700):
    rng = np.random.RandomState(seed)
    models = om.init_dynamics(rng, S, A, drop, H, K, out_scale=1.0)
    norm = dict(in_mean=rng.normal(0, 0.2, S + A).astype(np.float32),
                in_std=rng.uniform(0.5, 1.5, S + A).astype(np.float32),
                diff_mean=rng.normal(0, 0.05, S).astype(np.float32),
                diff_std=rng.uniform(0.1, 0.5, S).astype(np.float32))
    x = rng.normal(0, 1, (n, S + A)).astype(np.float32)
    y = (x[:, :S] + rng.normal(0, 0.1, (n, S))).astype(np.float32)
    return models, norm, x, y


def _fit(models, norm, S, A, drop, H, precision, max_rows=512):
    from me_trpo_b200.dynamics import EnsembleFit
    fit = EnsembleFit(S, A, drop, H, len(models), max_rows=max_rows, precision=precision)
    fit.set_ensemble(models)
    fit.set_normalization(**norm)
    fit.reset_adam()
    return fit


@pytest.mark.parametrize("precision,ltol,wtol", [("fp32", 1e-5, 2e-5), ("tf32", 2e-3, None)])
@pytest.mark.parametrize("dims", [(18, 6, 1, 256), (11, 3, 0, 128), (29, 8, 2, 64), (55, 21, 0, 96)],
                         ids=["half-cheetah", "hopper", "ant", "humanoid"])
def test_train_steps_match_oracle(dims, precision, ltol, wtol):
    S, A, drop, H = dims
    K, batch, steps = 3, 200, 5
    models, norm, x, y = _problem(1, S, A, drop, H, K)
    fit = _fit(models, norm, S, A, drop, H, precision)
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    ref = [{k: v.copy() for k, v in m.items()} for m in models]
    adam = of.Adam(ref)
    rng = np.random.RandomState(7)
    for j in range(steps):
        idx = rng.randint(0, len(x), batch * K)
        l_dev = fit.step(xd, yd, batch, 1e-3, idx=idx).cpu().numpy()
        l_ref = of.train_step(ref, adam, norm, x, y, idx, batch, 1e-3, S, drop)
        np.testing.assert_allclose(l_dev, l_ref, rtol=ltol, atol=ltol)
    assert fit.last_launches() == (13 if precision == "tf32" else (15 if H % 128 == 0 else 17))
    for k in range(K):
        w = fit.get_weights(k)
        for key in w:
            d = w[key].cpu().numpy() - ref[k][key]
            if wtol is not None:
                assert np.max(np.abs(d)) <= wtol, (key, np.max(np.abs(d)))
            else:
                upd = ref[k][key] - models[k][key]
                assert np.sqrt(np.mean(d ** 2)) <= 0.1 * np.sqrt(np.mean(upd ** 2)), key
    vl = fit.eval(xd, yd)[0].cpu().numpy()
    np.testing.assert_allclose(vl, of.validation_losses(ref, norm, x, y, S, drop), rtol=ltol if wtol else 2e-3)
    fit.close()


def test_eval_chunks_snapshot_and_restore():
    S, A, drop, H, K = 18, 6, 1, 128, 4
    models, norm, x, y = _problem(2, S, A, drop, H, K, n=1100)
    fit = _fit(models, norm, S, A, drop, H, "fp32", max_rows=256)     # 1100 rows -> 5 chunks, ragged tail
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    l0, imp0 = fit.eval(xd, yd, snapshot=2)
    np.testing.assert_allclose(l0.cpu().numpy(), of.validation_losses(models, norm, x, y, S, drop), rtol=2e-5)
    assert imp0.cpu().numpy().all()
    # make models 0 and 2 worse, 1 and 3 better (smaller output layer => pred ~ x, loss ~ 0.01*S)
    worse = [{k: v.copy() for k, v in m.items()} for m in models]
    for k in (0, 2):
        worse[k]["W2"] = worse[k]["W2"] * 3
    for k in (1, 3):
        worse[k]["W2"] = worse[k]["W2"] * 0.1
        worse[k]["b2"] = worse[k]["b2"] * 0.1
    fit.set_ensemble(worse)
    l1, imp1 = fit.eval(xd, yd, snapshot=1)
    assert list(imp1.cpu().numpy()) == [0, 1, 0, 1]
    fit.restore_best()        # models 0, 2 return to the initial snapshot; 1, 3 keep the improved weights
    for k in range(K):
        w = fit.get_weights(k)
        exp = models[k] if k in (0, 2) else worse[k]
        for key in w:
            np.testing.assert_array_equal(w[key].cpu().numpy(), exp[key])
    fit.close()


def test_philox_minibatches_train_and_are_reproducible():
    S, A, drop, H, K = 18, 6, 1, 128, 2
    models, norm, x, y = _problem(3, S, A, drop, H, K, n=600)
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    outs = []
    for rep in range(2):
        fit = _fit(models, norm, S, A, drop, H, "tf32")
        first = fit.eval(xd, yd)[0].cpu().numpy()
        for j in range(40):
            fit.step(xd, yd, 128, 1e-3, seed=5, offset=j, want_losses=False)
        last = fit.eval(xd, yd)[0].cpu().numpy()
        assert np.all(last < first), (first, last)
        outs.append(fit.get_weights(0)["W1"].cpu().numpy())
        fit.close()
    np.testing.assert_array_equal(outs[0], outs[1])


def test_optimize_models_matches_oracle_schedule():
    """Whole optimize_models loop (validate / snapshot / lr drop / stop / restore) against the
    oracle with the same index stream; fp32 GEMMs so that the snapshot decisions coincide."""
    from me_trpo_b200.dynamics import data_collection, optimize_models
    S, A, drop, H, K = 11, 3, 0, 64, 2
    models, norm, x, y = _problem(4, S, A, drop, H, K, n=512)
    xt, yt, xv, yv = x[:384], y[:384], x[384:], y[384:]
    fit = _fit(models, norm, S, A, drop, H, "fp32")
    dt, dv = data_collection(), data_collection()
    dt.add_data(xt, yt); dv.add_data(xv, yv)
    mk = lambda: (lambda r: (lambda j, n: r.randint(0, n, 64 * K)))(np.random.RandomState(9))
    kw = dict(batch_size=64, log_every=1, num_passes_threshold=2, max_passes=6)
    res = optimize_models(fit, dt, dv, learning_rate=dict(scratch=1e-3, refine=1e-3), reinitialize=False,
                          index_source=mk(), **kw)
    ref = [{k: v.copy() for k, v in m.items()} for m in models]
    ref, info = of.optimize_models(ref, norm, xt, yt, xv, yv, S, drop, lr_scratch=1e-3, lr_refine=1e-3,
                                   reinitialize=False, index_source=mk(), **kw)
    assert res["n_updates"] == info["n_updates"] and res["best_index"] == info["best_index"]
    np.testing.assert_array_equal(res["recover_indices"], info["recover_indices"])
    np.testing.assert_allclose(res["min_validation_losses"], info["min_validation_losses"], rtol=1e-4)
    for k in range(K):
        w = fit.get_weights(k)
        for key in w:
            assert np.max(np.abs(w[key].cpu().numpy() - ref[k][key])) <= 2e-4
    fit.close()


def test_running_mean_std_and_data_split():
    from me_trpo_b200.dynamics import RunningMeanStd, data_collection, add_rollout_data
    rng = np.random.RandomState(0)
    # the reference's own test_runningmeanstd property (running_mean_std.py:44-60): statistics of
    # the concatenation of all updates
    rms = RunningMeanStd(epsilon=0.0, shape=(2,))
    xs = [rng.normal(m, s, (n, 2)).astype(np.float32) for m, s, n in ((2.0, 1.0, 50), (1.0, 2.0, 70), (0.0, 3.0, 30))]
    for xx in xs:
        rms.update(xx)
    cat = np.concatenate(xs)
    np.testing.assert_allclose(rms.mean.cpu().numpy(), cat.mean(0), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rms.std.cpu().numpy(), cat.std(0), rtol=1e-3)
    assert np.allclose(RunningMeanStd(shape=(3,)).std.cpu().numpy(), 1.0)           # empty tracker: std 1
    small = RunningMeanStd(shape=(1,)); small.update(np.full((100, 1), 5.0, np.float32))
    assert abs(small.std.item() - 0.1) < 1e-6                                        # floor sqrt(1e-2)
    dt, dv = data_collection(max_size=50), data_collection(max_size=20)
    irms, orms = RunningMeanStd(shape=(4,)), RunningMeanStd(shape=(3,))
    x = rng.normal(size=(90, 4)).astype(np.float32); y = rng.normal(size=(90, 3)).astype(np.float32)
    add_rollout_data(x, y, dt, dv, irms, orms, 1.0 / 3)
    assert dv.get_num_data() == 20 and dt.get_num_data() == 50          # 30 -> cap 20, 60 -> cap 50
    np.testing.assert_array_equal(dt.x.cpu().numpy(), x[40:])           # FIFO: oldest rows dropped
    np.testing.assert_allclose(orms.mean.cpu().numpy(), (y[30:] - x[30:, :3]).sum(0) / (60 + 1e-2), rtol=1e-4, atol=1e-5)


def test_graph_replay_equals_eager_iterations():
    """TF32 mode, Philox minibatches: the CUDA-graph replay of the iteration (second call onwards) leaves
    bit-identical weights to the eager launch sequence, with and without a per-call loss buffer."""
    import os
    S, A, drop, H, K = 18, 6, 1, 256, 3
    models, norm, x, y = _problem(5, S, A, drop, H, K, n=2000)
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    res = {}
    for flag in ("0", "1"):
        old = os.environ.get("METRPO_FIT_GRAPH")
        os.environ["METRPO_FIT_GRAPH"] = flag
        try:
            fit = _fit(models, norm, S, A, drop, H, "tf32")
        finally:
            if old is None:
                os.environ.pop("METRPO_FIT_GRAPH", None)
            else:
                os.environ["METRPO_FIT_GRAPH"] = old
        losses = []
        for j in range(6):
            l = fit.step(xd, yd, 200, 1e-3, seed=3, offset=j, want_losses=(j % 2 == 0))
            if l is not None:
                losses.append(l.cpu().numpy())
        res[flag] = ([{k: v.cpu().numpy() for k, v in fit.get_weights(m).items()} for m in range(K)], losses)
        fit.close()
    for m in range(K):
        for key in res["0"][0][m]:
            np.testing.assert_array_equal(res["0"][0][m][key], res["1"][0][m][key])
    np.testing.assert_allclose(np.array(res["0"][1]), np.array(res["1"][1]), rtol=1e-6)


def test_train_steps_match_oracle_at_the_bench_shape():
    """The shape bench.py's fit_iteration times (half-cheetah: K = 5 models, H = 1024, batch 1000, TF32
    tcgen05 GEMMs with 256 x 256 tiles, split-K narrow products, side-stream wgrads) against the float32
    NumPy oracle: same tolerances as the small shapes."""
    S, A, drop, H, K, batch, steps = 18, 6, 1, 1024, 5, 1000, 3
    models, norm, x, y = _problem(11, S, A, drop, H, K, n=4000)
    fit = _fit(models, norm, S, A, drop, H, "tf32", max_rows=1024)
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    ref = [{k: v.copy() for k, v in m.items()} for m in models]
    adam = of.Adam(ref)
    rng = np.random.RandomState(3)
    for j in range(steps):
        idx = rng.randint(0, len(x), batch * K)
        l_dev = fit.step(xd, yd, batch, 1e-3, idx=idx).cpu().numpy()
        l_ref = of.train_step(ref, adam, norm, x, y, idx, batch, 1e-3, S, drop)
        np.testing.assert_allclose(l_dev, l_ref, rtol=2e-3, atol=2e-3)
    for k in range(K):
        w = fit.get_weights(k)
        for key in w:
            d = w[key].cpu().numpy() - ref[k][key]
            upd = ref[k][key] - models[k][key]
            assert np.sqrt(np.mean(d ** 2)) <= 0.1 * np.sqrt(np.mean(upd ** 2)), key
    vl = fit.eval(xd, yd)[0].cpu().numpy()
    np.testing.assert_allclose(vl, of.validation_losses(ref, norm, x, y, S, drop), rtol=2e-3)
    fit.close()
