"""Dimensions and config of the reference's environments as seen by the imaginary-rollout path.

S / A follow the obs builders of envs/com_*_env.py + the MuJoCo XML;
drop_cols mirrors `ignore_x_input` / `ignore_xy_input` of params/params-<env>.json
(training.py:146-154); hidden sizes are the shipped JSON values.  The analytic cost / done
functions themselves are fused into the CUDA kernel (csrc/rollout_kernel.cuh env_cost/env_is_done).
"""

ENV_SPECS = {
    "swimmer": dict(S=10, A=2, drop=2, hidden=512, policy_hidden=(32, 32)),
    "half-cheetah": dict(S=18, A=6, drop=1, hidden=1024, policy_hidden=(32, 32)),
    "hopper": dict(S=11, A=3, drop=0, hidden=1024, policy_hidden=(32, 32)),
    "ant": dict(S=29, A=8, drop=2, hidden=1024, policy_hidden=(32, 32)),
    "humanoid": dict(S=55, A=21, drop=0, hidden=1024, policy_hidden=(100, 50, 25)),
    "snake": dict(S=14, A=4, drop=2, hidden=1024, policy_hidden=(32, 32)),
}


def canonical_env_name(env):
    """Accept both the CLI/JSON spelling ('half-cheetah', run_model_based_rl.py:70) and get_env's
    ('half_cheetah', env_helpers.py:18)."""
    name = str(env).replace("_", "-") if str(env).startswith("half") else str(env)
    if name not in ENV_SPECS:
        raise AssertionError("unknown env %r" % (env,))   # env_helpers.py:31-32 asserts False
    return name


def drop_cols_from_params(dynamics_model_params):
    """training.py:146-154."""
    if dynamics_model_params.get("ignore_xy_input"):
        return 2
    if dynamics_model_params.get("ignore_x_input"):
        return 1
    return 0
