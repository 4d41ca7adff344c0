"""N4 host surface: params/*.json + -replace, the run_model_based_rl.py command line, the
real-env adapter, and (GPU) two full sweeps of the ME-TRPO loop through the entry point."""
import csv
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENVS = ["half-cheetah", "swimmer", "hopper", "ant", "humanoid", "snake"]


def test_params_files_match_table_and_replace():
    from me_trpo_b200 import params as P
    for env in ENVS:
        with open(os.path.join(ROOT, "params", "params-%s.json" % env)) as f:
            assert json.load(f) == P.default_params(env)
        p = P.load_params(env)
        assert p["env"] == env and p["algo"] == "trpo" and p["policy_opt_params"]["sam_mode"] == "step_rand"
    hc, sw, hu = P.default_params("half-cheetah"), P.default_params("swimmer"), P.default_params("humanoid")
    assert hc["n_models"] == 5 and hc["dynamics_model"]["hidden_layers"] == [1024, 1024]
    assert hc["policy_opt_params"]["trpo"] == dict(init_std=1.0, step_size=0.01, discount=1.0, batch_size=50000, reset=True)
    assert sw["dynamics_model"]["hidden_layers"] == [512, 512] and sw["dynamics_model"]["ignore_xy_input"] is True
    assert sw["policy_opt_params"]["T"] == 200 and hu["policy"]["hidden_layers"] == [100, 50, 25]
    P.replace_dict(hc, {"n_models": 3, "policy_opt_params": {"trpo": {"batch_size": 10}}})
    assert hc["n_models"] == 3 and hc["policy_opt_params"]["trpo"]["batch_size"] == 10
    assert hc["policy_opt_params"]["trpo"]["step_size"] == 0.01            # siblings untouched
    with pytest.raises(KeyError):
        P.replace_dict(hc, {"no_such_key": 1})
    with pytest.raises(ValueError):
        P.default_params("point2D")
    from me_trpo_b200.envs import drop_cols_from_params
    assert [drop_cols_from_params(P.default_params(e)["dynamics_model"]) for e in ENVS] == [1, 2, 0, 2, 0, 2]


def test_cli_argument_errors():
    sys.path.insert(0, ROOT)
    import run_model_based_rl as cli
    with pytest.raises(ValueError):
        cli.main(["trpo", "-env", "point2D"])
    with pytest.raises(NotImplementedError):
        cli.main(["trpo", "-env", "swimmer", "-ec2"])
    with pytest.raises(SystemExit):
        cli.main([])                                                          # algo is positional


def test_env_costs_and_synthetic_env():
    from me_trpo_b200 import env_costs, real_env
    from oracle import envs as oe
    rng = np.random.RandomState(0)
    for env in ENVS:
        S, A = oe.ENV_SPECS[env]["S"], oe.ENV_SPECS[env]["A"]
        x, xn = rng.normal(size=(7, S)), rng.normal(size=(7, S))
        u = np.clip(rng.normal(size=(7, A)), -1, 1)
        np.testing.assert_array_equal(env_costs.cost_np_vec(env, x, u, xn), oe.cost_np_vec(env, x, u, xn))
        np.testing.assert_array_equal(env_costs.is_done(env, x, xn), oe.is_done(env, x, xn))
        e = real_env.make_real_env(env, seed=1)
        o = e.reset()
        o2, r, d, info = e.step(rng.normal(size=A) * 3)                      # clipped inside
        assert o.shape == (S,) and o2.shape == (S,) and np.isfinite(r) and isinstance(d, bool)
        assert abs(r + env_costs.cost_np_vec(env, o[None], np.clip(np.zeros((1, A)), -1, 1), o2[None])[0]) < 10
    real_env.register("swimmer", lambda: "custom")
    assert real_env.make_real_env("swimmer") == "custom"
    real_env.REGISTRY.clear()


@pytest.mark.gpu
def test_two_sweeps_through_the_entry_point(tmp_path):
    sys.path.insert(0, ROOT)
    import run_model_based_rl as cli
    replace = {
        "sample_size": 400, "n_models": 3,
        "dynamics_model": {"hidden_layers": [256, 256]},
        "dynamics_opt_params": {"max_passes": 4, "log_every": 1, "num_passes_threshold": 2, "batch_size": 128},
        "policy_opt_params": {"T": 25, "max_iters": 4, "log_every": 2, "num_iters_threshold": 4, "batch_size": 64,
                              "trpo": {"batch_size": 2500}},
        "rollout_params": {"max_timestep": 25, "training_data_size": 2000, "validation_data_size": 1000},
    }
    out = cli.main(["trpo", "-env", "half-cheetah", "-seed", "3", "-sweeps", "2", "-replace", repr(replace),
                    "-snapshot_dir", str(tmp_path)])
    rows = out["progress"]
    assert len(rows) == 2
    with open(os.path.join(str(tmp_path), "progress.csv")) as f:
        got = list(csv.DictReader(f))
    assert len(got) == 2
    for key in ("collect_data_time", "model_opt_time", "policy_opt_time", "Time", "ItrTime", "EpisodesCollected",
                "TimeStepsCollected", "# model updates", "training_dynamics_min_sum_validation_loss",
                "estimated_policy_mean_min_validation_cost", "real_current_validation_cost", "# policy updates",
                "MaxPolicyWeightDiff"):
        assert key in got[0], key
    assert all(np.isfinite(float(r["training_dynamics_min_sum_validation_loss"])) for r in got)
    assert int(got[0]["TimeStepsCollected"]) >= 400
    # the policy moves unless every candidate of a sweep was rejected by the validation check and the
    # best (initial) policy restored (model_based_rl.py:1286-1299)
    for r in got:
        assert (float(r["MaxPolicyWeightDiff"]) > 0) or int(r["# policy updates"]) == 0 or r is got[1]
    assert os.path.exists(os.path.join(str(tmp_path), "params.json"))
    # the fitted ensemble predicts the stand-in simulator better than at initialisation
    assert float(got[1]["training_dynamics_min_sum_validation_loss"]) < 3 * 18
