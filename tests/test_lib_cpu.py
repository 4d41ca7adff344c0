"""CPU-only checks of the C-ABI library: it builds, loads, exports every symbol the header declares,
rejects bad arguments with the documented status codes, fails loudly without a GPU (no fallback),
and its gang schedule has the properties the kernel relies on."""
import ctypes

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(metrpo_lib):
    syms = metrpo_lib.check_exports()
    for s in ["metrpo_rollout_create", "metrpo_rollout_run", "metrpo_rollout_step", "metrpo_rollout_reset",
              "metrpo_rollout_set_dynamics", "metrpo_rollout_set_policy", "metrpo_rollout_set_normalization",
              "metrpo_rollout_destroy", "metrpo_last_error", "metrpo_version"]:
        assert s in syms
    assert b"sm_100a" in metrpo_lib.load().metrpo_version()


def _cfg(lib, **kw):
    cfg = lib.RolloutCfg()
    d = dict(state_dim=18, action_dim=6, drop_cols=1, hidden=1024, n_models=5, n_envs=256,
             max_path_length=100, env_id=1, sam_mode=0, n_policy_layers=3, policy_out_tanh=0, precision=0,
             device=0, row_offset=0)
    d.update({k: v for k, v in kw.items() if k != "policy_dims"})
    for k, v in d.items():
        setattr(cfg, k, v)
    for i, v in enumerate(kw.get("policy_dims", [d["state_dim"], 32, 32, d["action_dim"]])):
        cfg.policy_dims[i] = v
    return cfg


@pytest.mark.parametrize("kw,status,needle", [
    (dict(state_dim=1), -1, "S>=2"),
    (dict(drop_cols=3), -1, "drop_cols"),
    (dict(env_id=9), -1, "env_id"),
    (dict(sam_mode=7), -1, "sam_mode"),
    (dict(hidden=300), -3, "multiple of 256"),
    (dict(state_dim=70, action_dim=21, policy_dims=[70, 32, 32, 21]), -3, "S <= 64"),
    (dict(state_dim=55, action_dim=21, hidden=192, policy_dims=[55, 32, 32, 21]), -3, "multiple of 128"),
    (dict(policy_dims=[18, 200, 32, 6]), -3, "policy hidden width"),
    (dict(precision=3), -3, "PREC_BF16"),
    (dict(policy_dims=[17, 32, 32, 6]), -1, "policy_dims"),
    (dict(env_id=3, state_dim=10, action_dim=2, policy_dims=[10, 32, 32, 2]), -1, "needs state_dim"),
])
def test_create_rejects_bad_config(metrpo_lib, kw, status, needle):
    lib = metrpo_lib.load()
    h = ctypes.c_void_p()
    cfg = _cfg(metrpo_lib, **kw)
    st = lib.metrpo_rollout_create(ctypes.byref(cfg), ctypes.byref(h))
    assert st == status, metrpo_lib.last_error()
    assert needle in metrpo_lib.last_error()
    assert not h.value


def test_no_cpu_fallback(metrpo_lib):
    """Without a CUDA device a valid config must fail with METRPO_ERR_CUDA, never run on the host."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = metrpo_lib.load()
    h = ctypes.c_void_p()
    cfg = _cfg(metrpo_lib)
    st = lib.metrpo_rollout_create(ctypes.byref(cfg), ctypes.byref(h))
    assert st == -2 and not h.value
    with pytest.raises(RuntimeError):
        metrpo_lib.check(st, "create")


def test_null_handles_are_errors_not_crashes(metrpo_lib):
    lib = metrpo_lib.load()
    assert lib.metrpo_rollout_run(None, 1, None, None, 0, None, None, None, 0, 0, 0, None, None, None, None,
                                  None, None, None) == -1
    assert lib.metrpo_rollout_step(None, None, None, None, None, 0, 0, None, None, None, None) == -1
    assert lib.metrpo_rollout_reset(None, None, None) == -1
    assert lib.metrpo_rollout_destroy(None) == 0
    assert lib.metrpo_rollout_last_launches(None) == 0


def _schedule(metrpo_lib, n_tiles, n_slots, T):
    lib = metrpo_lib.load()
    buf = np.full(n_slots * 256 * 4, -7, np.int32)
    ms = lib.metrpo_debug_schedule(n_tiles, n_slots, T, buf.ctypes.data_as(ctypes.c_void_p), n_slots * 256)
    assert ms > 0, metrpo_lib.last_error()
    return buf[:n_slots * ms * 4].reshape(n_slots, ms, 4)


@pytest.mark.parametrize("n_tiles,n_slots,T", [(32, 29, 1000), (32, 29, 12), (3, 3, 5), (64, 7, 1000),
                                               (32, 14, 1), (1, 1, 100), (33, 29, 7)])
def test_gang_schedule_properties(metrpo_lib, n_tiles, n_slots, T):
    segs = _schedule(metrpo_lib, n_tiles, n_slots, T)
    cover = np.zeros((n_tiles, T), np.int32)
    loads = []
    for j in range(n_slots):
        used = [tuple(q) for q in segs[j] if q[0] >= 0]
        loads.append(sum(t1 - t0 for _, t0, t1, _ in used))
        seen_tail = False
        for tile, t0, t1, wait in used:
            assert 0 <= t0 < t1 <= T
            cover[tile, t0:t1] += 1
            assert wait == (1 if t0 > 0 else 0)          # only tails wait for another slot
            if t0 > 0:
                seen_tail = True
            else:
                assert not seen_tail                     # pieces that start a chain run before tails
    assert (cover == 1).all()                            # every (tile, step) exactly once
    assert sum(loads) == n_tiles * T
    assert max(loads) == -(-n_tiles * T // n_slots)      # balanced to the ceiling
    # a chain is cut at most once, and the slot holding its head is a different slot
    for tile in range(n_tiles):
        owners = [(j, tuple(q)) for j in range(n_slots) for q in segs[j] if q[0] == tile]
        assert 1 <= len(owners) <= 2
        if len(owners) == 2:
            assert owners[0][0] != owners[1][0]


def test_gang_schedule_head_finishes_before_tail_starts(metrpo_lib):
    """Timing argument behind the deadlock-freedom claim: with equal per-step cost, the head piece of
    a split chain ends (it runs first in its slot) before the tail piece starts (it runs last)."""
    for n_tiles, n_slots, T in [(32, 29, 1000), (64, 7, 1000), (33, 29, 7)]:
        segs = _schedule(metrpo_lib, n_tiles, n_slots, T)
        end_of_head, start_of_tail = {}, {}
        for j in range(n_slots):
            clock = 0
            for tile, t0, t1, wait in segs[j]:
                if tile < 0:
                    continue
                if t0 == 0 and t1 < T:
                    end_of_head[tile] = clock + (t1 - t0)
                if t0 > 0:
                    start_of_tail[tile] = clock
                clock += t1 - t0
        for tile, st in start_of_tail.items():
            assert end_of_head[tile] <= st
