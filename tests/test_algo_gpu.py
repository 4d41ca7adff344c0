"""The reference's TRPO inner iteration (model_based_rl.py:1171-1180:
start_worker -> obtain_samples -> process_samples -> optimize_policy) driven through the mirrored
classes, in the reference's list-of-paths form and in the device-resident flat form; both must
produce the same policy update, and the update must equal the float64 oracle's."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402


class _PoolSampler:
    def __init__(self, pool):
        self.pool, self.i = pool, 0

    def __call__(self, n):
        idx = (self.i + np.arange(n)) % len(self.pool)
        self.i += n
        return self.pool[idx]


def _make_algo(B=128, T_max=20, hidden=256, K=3, seed=0):
    from me_trpo_b200.algos import TRPO
    from me_trpo_b200.baselines import LinearFeatureBaseline
    from me_trpo_b200.env_helpers import NeuralNetEnv
    from me_trpo_b200.policies import GaussianMLPPolicy
    env = "half-cheetah"
    inp = mg.make_inputs(env, K, B, 1, hidden)
    nn_env = NeuralNetEnv(env, inp["models"], inp["norm"], reset_sampler=_PoolSampler(inp["pool"]))
    policy = GaussianMLPPolicy(18, 6, (32, 32), seed=seed)
    algo = TRPO(env=nn_env, policy=policy, baseline=LinearFeatureBaseline(env_spec=nn_env.spec),
                batch_size=B * T_max, max_path_length=T_max, discount=0.99, step_size=0.01,
                sampler_args=dict(n_envs=B, seed=11))
    return algo


def test_trpo_iteration_paths_and_flat_agree_and_match_oracle():
    from oracle import trpo as ot
    algo_p, algo_f = _make_algo(), _make_algo()
    theta0 = algo_p.policy.get_param_values().astype(np.float64)
    # reference-style iteration (list of paths on the host)
    algo_p.start_worker()
    paths = algo_p.obtain_samples(1)
    samples_data = algo_p.process_samples(1, paths)
    algo_p.optimize_policy(1, samples_data)
    # device-resident iteration
    algo_f.start_worker()
    flat = algo_f.obtain_samples_flat(1)
    data = algo_f.process_samples_flat(1, flat)
    algo_f.optimize_policy(1, data)
    th_p = algo_p.policy.get_param_values().astype(np.float64)
    th_f = algo_f.policy.get_param_values().astype(np.float64)
    assert not np.allclose(th_p, theta0)
    assert np.max(np.abs(th_p - th_f)) <= 2e-5 * max(1.0, np.abs(th_p - theta0).max() / 1e-2)
    info = algo_f.optimizer.last_info.cpu().numpy()
    assert info[4] == 1.0 and info[2] <= 0.01 and info[1] < info[0]
    # oracle update from the same processed samples (float64, autograd HVP)
    orc = ot.TRPOOracle([18, 32, 32, 6])
    inputs = (samples_data["observations"], samples_data["actions"], samples_data["advantages"],
              samples_data["agent_infos"]["mean"], samples_data["agent_infos"]["log_std"])
    new_ref, info_ref = orc.optimize(theta0.astype(np.float32), inputs)
    step_dev, step_ref = th_p - theta0, new_ref - theta0
    cos = step_dev.dot(step_ref) / (np.linalg.norm(step_dev) * np.linalg.norm(step_ref))
    assert info_ref["accepted"] and cos >= 0.9999
    assert np.linalg.norm(step_dev - step_ref) <= 5e-3 * np.linalg.norm(step_ref)
    # the baseline was refitted after the advantages (samplers/base.py:167), on host and device alike
    c_host = algo_p.baseline._coeffs
    pred_h = algo_p.baseline.predict(paths[0])
    pred_d = algo_f.baseline.predict(paths[0])
    assert c_host is not None and np.max(np.abs(pred_h - pred_d)) <= 1e-3 * max(1.0, np.abs(pred_h).max())
    algo_p.shutdown_worker(); algo_f.shutdown_worker()


def test_optimizer_socket_loss_and_constraint_val():
    algo = _make_algo(B=64, T_max=10)
    algo.start_worker()
    data = algo.process_samples_flat(0, algo.obtain_samples_flat(0))
    inputs = (data["observations"], data["actions"], data["advantages"], data["agent_infos"]["mean"],
              data["agent_infos"]["log_std"], data["valids"])
    kl0 = algo.optimizer.constraint_val(inputs)
    loss0 = algo.optimizer.loss(inputs)
    assert abs(kl0) < 1e-6 and abs(loss0) < 1e-5     # old == new, centred advantages
    algo.optimize_policy(0, data)
    assert algo.optimizer.loss(inputs) < loss0 and 0 < algo.optimizer.constraint_val(inputs) <= 0.01
    algo.shutdown_worker()


def test_several_iterations_run_and_keep_the_trust_region():
    algo = _make_algo(B=256, T_max=25)
    for j in range(1, 4):
        algo.start_worker()                                   # the reference rebuilds the vec env per iteration
        data = algo.process_samples_flat(j, algo.obtain_samples_flat(j))
        algo.optimize_policy(j, data)
        info = algo.optimizer.last_info.cpu().numpy()
        assert info[2] <= 0.01 + 1e-9 and np.isfinite(info).all()
        algo.shutdown_worker()
