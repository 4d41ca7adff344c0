"""GPU parity tests of the fused ensemble rollout (through the C ABI) against the NumPy oracle.

Stated tolerances (fp32 state units; synthetic nets with unit-scale normalised inputs):
  * vs the oracle evaluated with the kernel's arithmetic (bf16 operands, fp32 accumulation,
    mma="bf16"): 1e-4 on obs / act / mean / rew over the short open-loop horizons used here;
    done flags exact;
  * vs the oracle in the reference's arithmetic (all fp32, mma="fp32"): 1e-3;
  * one step, teacher-forced with the device's own pre-step states, vs mma="bf16": 5e-5.
"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402
from oracle import rollout as orl  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "rollout_golden.npz"))
TOL_BF16, TOL_FP32, TOL_TF = 1e-4, 1e-3, 5e-5


def _device_run(case, inp=None, **run_kw):
    from me_trpo_b200.rollout import EnsembleRollout
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    inp = inp or mg.make_inputs(env, K, B, T, hidden)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode)
    ro.set_dynamics_ensemble(inp["models"])
    ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    kw = dict(seed=1234, offset=7)
    if noise_kind == "explicit":
        kw.update(eps=inp["eps"], model_idx=inp["mi"], std_noise=inp["sn"] if sam_mode == "model_mean_std" else None)
    kw.update(run_kw)
    out = ro.run(T, inp["init"], inp["pool"], **kw)
    ro.synchronize()
    assert ro.last_launches() == 1
    res = {k: v.cpu().numpy() for k, v in out.items()}
    ro.close()
    return res, inp


@pytest.mark.parametrize("case", mg.CASES, ids=[c[0] for c in mg.CASES])
def test_rollout_matches_golden(case):
    dev, _ = _device_run(case)
    for mma, tol in (("bf16", TOL_BF16), ("fp32", TOL_FP32)):
        for k in ("obs", "act", "mean", "rew", "final_states"):
            ref = GOLD["%s/%s/%s" % (case[0], mma, k)]
            err = np.max(np.abs(dev[k] - ref))
            assert err <= tol, (case[0], mma, k, err)
        assert np.array_equal(dev["done"], GOLD["%s/%s/done" % (case[0], mma)])


BIG = [
    ("hc_b300_h1024", "half-cheetah", 5, 300, 5, 100, 1024, "step_rand", "explicit"),
    ("hc_ragged_reset", "half-cheetah", 5, 200, 6, 4, 512, "step_rand", "explicit"),
    ("hopper_b130", "hopper", 3, 130, 4, 3, 256, "step_rand", "explicit"),
    ("ant_b256", "ant", 4, 256, 5, 100, 256, "step_rand", "explicit"),
    ("hc_split_chains", "half-cheetah", 5, 4096, 12, 100, 256, "step_rand", "philox"),
    ("hc_k1", "half-cheetah", 1, 128, 3, 100, 256, "one_model", "explicit"),
    # humanoid dims: S=55, A=21, policy 100-50-25 -> the <64,24,128> instantiation (128-column passes)
    ("humanoid_b200", "humanoid", 3, 200, 5, 3, 256, "step_rand", "explicit"),
    ("humanoid_k20_philox", "humanoid", 20, 1100, 4, 100, 128, "step_rand", "philox"),
    ("humanoid_mean_std", "humanoid", 3, 128, 3, 100, 128, "model_mean_std", "explicit"),
]


@pytest.mark.parametrize("case", BIG, ids=[c[0] for c in BIG])
def test_rollout_matches_oracle_larger_cases(case):
    """Ragged last tile (B % 128 != 0), resets inside the horizon, several row tiles per gang slot,
    more tiles than slots (chains split across CTAs), Ant's early termination, K = 1."""
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    dev, inp = _device_run(case)
    noise = (orl.PhiloxNoise(1234, 7, 0, sam_mode) if noise_kind == "philox"
             else orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"]))
    for mma, tol in (("bf16", TOL_BF16), ("fp32", TOL_FP32)):
        ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T,
                               T_max, sam_mode, mma=mma)
        for k in ("obs", "act", "mean", "rew", "final_states"):
            err = np.max(np.abs(dev[k] - ref[k]))
            assert err <= tol, (name, mma, k, err)
        assert np.array_equal(dev["done"], ref["done"])
    # teacher-forced single-step parity
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T, T_max,
                           sam_mode, mma="bf16", teacher_states=dev["obs"])
    nxt_dev = np.concatenate([dev["obs"][1:], dev["final_states"][None]], 0)
    nxt_ref = np.concatenate([ref["obs"][1:], ref["final_states"][None]], 0)
    assert np.max(np.abs(nxt_dev - nxt_ref)) <= TOL_TF
    assert np.max(np.abs(dev["rew"] - ref["rew"])) <= TOL_TF
    assert np.isfinite(dev["obs"]).all()


def test_philox_offsets_beyond_32_bits():
    """The step counter of the noise streams is 64 bits wide: offsets 2**32 apart give different
    noise, the carry into the high word is handled, and the kernel still matches the oracle."""
    case = [c for c in mg.CASES if c[8] == "philox" and c[7] == "step_rand"][0]
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    lo, inp = _device_run(case, offset=5)
    hi, _ = _device_run(case, inp=inp, offset=5 + (1 << 32))
    assert not np.array_equal(lo["act"], hi["act"])
    off = (1 << 32) - 2                                   # steps 2.. run with the high word = 1
    dev, _ = _device_run(case, inp=inp, offset=off)
    noise = orl.PhiloxNoise(1234, off, 0, sam_mode)
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T,
                           T_max, sam_mode=sam_mode, mma="bf16")
    for k in ("obs", "act", "rew"):
        assert np.max(np.abs(dev[k] - ref[k])) <= TOL_BF16, k
    assert np.array_equal(dev["done"], ref["done"])


def test_determ_mode_actions_equal_mean():
    case = mg.CASES[0]
    dev, inp = _device_run(case, determ=True)
    np.testing.assert_array_equal(dev["act"], dev["mean"])      # obtain_samples(determ=True) (:64-65)


def test_full_size_properties_half_cheetah():
    """BASELINE shape (5 models, 4096 rollouts) with size-independent checks: determinism,
    timeout placement, reward recomputed from the trajectory, reset rule, finiteness."""
    from me_trpo_b200.rollout import EnsembleRollout
    from oracle import models as om, envs as oe
    env, K, B, T, T_max, hidden = "half-cheetah", 5, 4096, 120, 50, 1024
    spec = oe.ENV_SPECS[env]
    rng = np.random.RandomState(0)
    models = om.init_dynamics(rng, spec["S"], spec["A"], spec["drop"], hidden, K)
    pol = om.init_policy(rng, spec["S"], spec["policy_hidden"], spec["A"])
    norm = om.default_norm(spec["S"], spec["A"])
    init = rng.normal(0, 0.1, (B, spec["S"])).astype(np.float32)
    pool = rng.normal(0, 0.1, (2 * B, spec["S"])).astype(np.float32)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden)
    ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    a = ro.run(T, init, pool, seed=3); ro.synchronize()
    a = {k: v.clone() for k, v in a.items()}
    b = ro.run(T, init, pool, seed=3); ro.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k                        # bitwise deterministic
    c = ro.run(T, init, pool, seed=4); ro.synchronize()
    assert not torch.equal(a["act"], c["act"])                   # the seed matters
    done = a["done"].cpu().numpy()
    want = np.zeros((T, B), np.uint8); want[T_max - 1::T_max] = 1
    assert np.array_equal(done, want)                            # only the timeout ends paths (:604)
    obs, act, rew = a["obs"], a["act"], a["rew"]
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    # reward = clip(x'[9] - 0.05 * sum(clip(a)^2), -10, 10) with x' = next pre-step obs (non-done steps)
    u = act.clamp(-1, 1)
    r = (obs[1:, :, 9] - 0.05 * (u[:-1] ** 2).sum(-1)).clamp(-10, 10)
    keep = torch.as_tensor(done[:-1] == 0, device=obs.device)
    assert (r - rew[:-1]).abs()[keep].max().item() < 1e-5
    # after a timeout the next observation is the reset state pool[(n*B + i) % R]   (:605-607)
    pool_d = torch.as_tensor(pool, device=obs.device)
    assert torch.equal(obs[T_max], pool_d[:B]) and torch.equal(obs[2 * T_max], pool_d[B:2 * B])
    assert torch.equal(obs[0], torch.as_tensor(init, device=obs.device))
    ro.close()


def test_fused_horizon_equals_single_steps_on_device():
    """test_policy_cost-style equivalence on the GPU (env_helpers.py:271-305): the persistent
    kernel's T-step rollout equals T calls of the step-granular socket (B1) fed with the recorded
    unclipped actions and model indices -- bitwise, same arithmetic."""
    from me_trpo_b200.rollout import EnsembleRollout
    case = ("x", "half-cheetah", 5, 600, 9, 4, 512, "step_rand", "explicit")
    dev, inp = _device_run(case)
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode)
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.reset(inp["init"])
    R = len(inp["pool"])
    nres = np.zeros(B, np.int64)
    for t in range(T):
        reset_states = inp["pool"][(nres * B + np.arange(B)) % R]
        obs, rew, done = ro.step(dev["act"][t], reset_states, model_idx=inp["mi"][t])
        ro.synchronize()
        nxt = dev["obs"][t + 1] if t + 1 < T else dev["final_states"]
        np.testing.assert_array_equal(obs.cpu().numpy(), nxt)
        np.testing.assert_array_equal(rew.cpu().numpy(), dev["rew"][t])
        np.testing.assert_array_equal(done.cpu().numpy(), dev["done"][t])
        nres += dev["done"][t]
    ro.close()


def test_row_sharding_reproduces_unsharded_run():
    """Multi-GPU decomposition exercised on one device: two handles owning row blocks
    [0,B/2) and [B/2,B) with row_offset reproduce the unsharded rollout bitwise (Philox streams are
    keyed by global row; rows never interact)."""
    from me_trpo_b200.rollout import EnsembleRollout
    from me_trpo_b200.parallel import shard_rows
    case = ("x", "half-cheetah", 3, 512, 7, 3, 256, "step_rand", "philox")
    full, inp = _device_run(case)
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case
    R = len(inp["pool"]); n_res = -(-T // T_max)
    parts = []
    for r in range(2):
        lo, hi = shard_rows(B, r, 2)
        local_pool = np.stack([inp["pool"][(n * B + i) % R] for n in range(n_res) for i in range(lo, hi)])
        ro = EnsembleRollout(env, K, hi - lo, T_max, hidden=hidden, sam_mode=sam_mode, row_offset=lo)
        ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
        ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
        out = ro.run(T, inp["init"][lo:hi], local_pool, seed=1234, offset=7); ro.synchronize()
        parts.append({k: v.cpu().numpy() for k, v in out.items()})
        ro.close()
    for k in ("obs", "act", "rew", "done"):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts], axis=1), full[k])


def test_call_order_errors():
    from me_trpo_b200.rollout import EnsembleRollout
    ro = EnsembleRollout("half-cheetah", 2, 128, 10, hidden=256)
    init = np.zeros((128, 18), np.float32)
    with pytest.raises(RuntimeError, match="policy was never set|dynamics model"):
        ro.run(2, init, init)
    with pytest.raises(RuntimeError, match="reset"):
        ro.step(np.zeros((128, 6), np.float32), init)
    ro.close()
    with pytest.raises(AssertionError):
        EnsembleRollout("half-cheetah", 2, 128, 10, sam_mode="bogus")


@pytest.mark.parametrize("n_chunks", [1, 3, 8])
def test_run_to_host_chunked_equals_single_launch(n_chunks):
    """metrpo_rollout_continue: a horizon cut into chained launches (with the D2H copy overlapped)
    reproduces the single launch bit for bit -- resets inside and across chunk boundaries, chains
    split across gang slots, Philox noise keyed by global step."""
    from me_trpo_b200.rollout import EnsembleRollout
    env, K, B, T, T_max, hidden = "half-cheetah", 5, 4096, 24, 7, 256
    inp = mg.make_inputs(env, K, B, 1, hidden)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden)
    ro.set_dynamics_ensemble(inp["models"])
    ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    ref = {k: v.cpu().numpy() for k, v in ro.run(T, inp["init"], inp["pool"], seed=3, offset=11).items()}
    ro.synchronize()
    host, _ = ro.run_to_host(T, inp["init"], inp["pool"], seed=3, offset=11, n_chunks=n_chunks)
    ro.synchronize()
    for k in ref:
        np.testing.assert_array_equal(host[k].numpy(), ref[k], err_msg=k)
    assert ref["done"].sum() == B * (T // T_max)
    ro.close()


# ---------------------------------------------------------------------------------------------
# two-stream kernel (csrc/rollout_duo.cuh): forced through the METRPO_DUO dev switch, which the
# library reads when a handle is created
# ---------------------------------------------------------------------------------------------
DUO = [
    # ragged: 5 tiles -> 3 tile pairs (the last pair has one dummy tile), resets inside the horizon
    ("duo_hc_philox", "half-cheetah", 5, 600, 7, 4, 512, "step_rand", "philox"),
    ("duo_hc_explicit", "half-cheetah", 3, 256, 5, 100, 512, "step_rand", "explicit"),
    ("duo_hc_eps_rand", "half-cheetah", 4, 384, 6, 3, 1024, "eps_rand", "philox"),
    ("duo_hopper", "hopper", 3, 300, 5, 3, 512, "step_rand", "explicit"),
    ("duo_swimmer", "swimmer", 5, 700, 6, 5, 512, "step_rand", "philox"),
    # more pairs than gang slots: chains cut across slots, tile state handed over through row_state
    ("duo_hc_many_tiles", "half-cheetah", 5, 4096, 9, 100, 512, "step_rand", "philox"),
    # ant: K0 = 48 -> ONE Z slot in TMEM taken in turns by the streams (z_shared), 8-row policy passes
    # (shared memory), early-terminating paths (is_done) with resets from the pool
    ("duo_ant", "ant", 4, 600, 7, 100, 512, "step_rand", "philox"),
    ("duo_ant_k20", "ant", 20, 1024, 4, 3, 1024, "step_rand", "explicit"),
]


def _with_duo(mode, fn):
    old = os.environ.get("METRPO_DUO")
    os.environ["METRPO_DUO"] = str(mode)
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop("METRPO_DUO", None)
        else:
            os.environ["METRPO_DUO"] = old


def _device_run_kernel(case, inp=None, **kw):
    """(results, inputs, kernel id) of one fused launch."""
    from me_trpo_b200.rollout import EnsembleRollout
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    inp = inp or mg.make_inputs(env, K, B, T, hidden)
    ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode)
    ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
    ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
    rk = dict(seed=1234, offset=7)
    if noise_kind == "explicit":
        rk.update(eps=inp["eps"], model_idx=inp["mi"])
    rk.update(kw)
    out = ro.run(T, inp["init"], inp["pool"], **rk)
    ro.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    kern = ro.last_kernel()
    ro.close()
    return res, inp, kern


@pytest.mark.parametrize("cs", [1, 2])
@pytest.mark.parametrize("case", DUO, ids=[c[0] for c in DUO])
def test_duo_kernel_matches_oracle_and_single_stream(case, cs):
    name, env, K, B, T, T_max, hidden, sam_mode, noise_kind = case
    if cs == 2 and (hidden // 256) % 2:
        pytest.skip("column split needs an even number of layer-1 passes")
    duo, inp, kern = _with_duo(cs, lambda: _device_run_kernel(case))
    assert kern == cs, "the two-stream kernel was not selected"
    single, _, kern0 = _with_duo(0, lambda: _device_run_kernel(case, inp=inp))
    assert kern0 == 0
    noise = (orl.PhiloxNoise(1234, 7, 0, sam_mode) if noise_kind == "philox"
             else orl.ExplicitNoise(inp["eps"], inp["mi"], inp["sn"]))
    ref = orl.rollout_flat(env, inp["pol"], inp["models"], inp["norm"], inp["init"], inp["pool"], noise, T,
                           T_max, sam_mode, mma="bf16")
    for k in ("obs", "act", "mean", "rew", "final_states"):
        assert np.max(np.abs(duo[k] - ref[k])) <= TOL_BF16, (name, cs, k)
        if cs == 1:     # same arithmetic in the same order: bit-identical to the single-stream kernel
            np.testing.assert_array_equal(duo[k], single[k])
        else:           # layer-2 partial sums of the two column halves are added in a different order
            assert np.max(np.abs(duo[k] - single[k])) <= 2e-5, (name, k)
    assert np.array_equal(duo["done"], ref["done"]) and np.array_equal(duo["done"], single["done"])


@pytest.mark.parametrize("cs", [1, 2])
def test_duo_kernel_chained_launches_equal_single_launch(cs):
    """metrpo_rollout_continue through the two-stream kernel: the state handed over in row_state
    makes an 3-chunk run bit-identical to one launch."""
    from me_trpo_b200.rollout import EnsembleRollout
    case = DUO[0]
    name, env, K, B, T, T_max, hidden, sam_mode, _ = case

    def go():
        inp = mg.make_inputs(env, K, B, T, hidden)
        ro = EnsembleRollout(env, K, B, T_max, hidden=hidden, sam_mode=sam_mode)
        ro.set_dynamics_ensemble(inp["models"]); ro.set_normalization(**inp["norm"])
        ro.set_policy(inp["pol"]["W"], inp["pol"]["b"], inp["pol"]["log_std"])
        one = ro.run(T, inp["init"], inp["pool"], seed=9, offset=3); ro.synchronize()
        one = {k: v.cpu().numpy() for k, v in one.items()}
        host, _ = ro.run_to_host(T, inp["init"], inp["pool"], seed=9, offset=3, n_chunks=3); ro.synchronize()
        assert ro.last_kernel() == cs
        ro.close()
        return one, {k: v.numpy() for k, v in host.items()}
    one, chunked = _with_duo(cs, go)
    for k in ("obs", "act", "mean", "rew", "done", "final_states"):
        np.testing.assert_array_equal(one[k], chunked[k])
