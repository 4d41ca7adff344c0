"""params/params-<env>.json config surface (run_model_based_rl.py:70-86, training.py:26-70).

The shipped JSON files are generated from the table below (same keys and values as the
reference's params/*.json; `python -m me_trpo_b200.params` rewrites them), and user-supplied
files in the reference's format load unchanged.  `replace_dict` mirrors the `-replace` option
(run_model_based_rl.py:35-51)."""
import copy
import json
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS_DIR = os.path.join(_ROOT, "params")

_BASE = {
    "algo": "trpo",
    "env": "half-cheetah",
    "policy": {"hidden_layers": [32, 32], "output_nonlinearity": "tf.identity"},
    "n_models": 5,
    "sample_size": 3000,
    "sweep_iters": 100,
    "dynamics_model": {
        "hidden_layers": [1024, 1024],
        "regularization": {"method": "tf.nn.l2_loss", "constant": 0.0},
        "nonlinearity": ["tf.nn.relu", "tf.nn.relu"],
        "ignore_x_input": True,
        "prediction_type": "state_change",
        "use_logit_weights": False,
    },
    "dynamics_opt_params": {
        "learning_rate": {"scratch": 1e-3, "refine": 1e-3},
        "log_every": 5,
        "max_passes": 2000,
        "stop_critereon": {"offset": 1e-05, "threshold": 0.10},
        "batch_size": 1000,
        "sample_mode": "random",
        "reinitialize": 5,
        "num_passes_threshold": 25,
    },
    "policy_opt_params": {
        "mode": "estimated", "whole": True, "T": 100, "gamma": 1.0, "grad_norm_clipping": 10,
        "learning_rate": 1e-3, "log_every": 5, "num_iters_threshold": 25, "max_iters": 400,
        "oracle_maxtimestep": 100,
        "stop_critereon": {"offset": 1e-05, "threshold": 0.10, "percent_models_threshold": 0.30},
        "validation_init_path": "data_upload/policy_validation_inits_half_cheetah.save",
        "validation_reset_init_path": "data_upload/policy_validation_reset_inits_half_cheetah.save",
        "trpo": {"init_std": 1.0, "step_size": 0.01, "discount": 1.0, "batch_size": 50000, "reset": True},
        "vpg": {"init_std": 1.0, "discount": 1.0, "batch_size": 50000, "reset": True},
        "batch_size": 500,
        "sam_mode": "step_rand",
    },
    "rollout_params": {
        "training_data_size": 200000, "validation_data_size": 100000, "split_ratio": 0.33333333,
        "splitting_mode": "trajectory", "use_same_dataset": True,
        "exploration": {"initial_param_std": 0.0, "param_noise": 3.0, "action_noise": 3.0,
                        "vary_trajectory_noise": True},
        "datapath": "", "is_monitored": False, "max_timestep": 100, "render_every": None,
        "load_rollout_data": False,
    },
}


def _xy(d):      # ignore_xy_input replaces ignore_x_input (training.py:146-154)
    d["dynamics_model"].pop("ignore_x_input", None)
    d["dynamics_model"]["ignore_xy_input"] = True


def _horizon(d, T):
    d["policy_opt_params"]["T"] = T
    d["policy_opt_params"]["oracle_maxtimestep"] = T
    d["rollout_params"]["max_timestep"] = T


def _swimmer(d):
    d["dynamics_model"]["hidden_layers"] = [512, 512]
    d["dynamics_model"].pop("use_logit_weights", None)
    d["rollout_params"].update(training_data_size=100000, validation_data_size=50000)
    # (sic) the shipped swimmer file points both entries at the same pickle
    d["policy_opt_params"]["validation_reset_init_path"] = "data_upload/policy_validation_inits_swimmer.save"
    _xy(d); _horizon(d, 200)


def _hopper(d):
    d["dynamics_model"]["ignore_x_input"] = False
    d["dynamics_model"].pop("use_logit_weights", None)


def _ant(d):
    d["sweep_iters"] = 200
    _xy(d)


def _humanoid(d):
    d["sweep_iters"] = 400
    d["sample_size"] = 6000
    d["policy"]["hidden_layers"] = [100, 50, 25]
    d["policy_opt_params"]["batch_size"] = 32
    d["rollout_params"].update(training_data_size=400000, split_ratio=0.2)
    d["dynamics_model"].pop("ignore_x_input", None)


def _snake(d):
    d["dynamics_model"].pop("use_logit_weights", None)
    _xy(d); _horizon(d, 200)


_ENVS = {"half-cheetah": lambda d: None, "swimmer": _swimmer, "hopper": _hopper, "ant": _ant,
         "humanoid": _humanoid, "snake": _snake}


def default_params(env):
    if env not in _ENVS:
        raise ValueError("Value Error: not implemented.")          # run_model_based_rl.py:79
    d = copy.deepcopy(_BASE)
    d["env"] = env
    tag = env.replace("-", "_")
    d["policy_opt_params"]["validation_init_path"] = "data_upload/policy_validation_inits_%s.save" % tag
    d["policy_opt_params"]["validation_reset_init_path"] = "data_upload/policy_validation_reset_inits_%s.save" % tag
    _ENVS[env](d)
    return d


def load_params(env, param_path=None):
    """params/params-<env>.json if present (or an explicit path), else the built-in table."""
    path = param_path or os.path.join(PARAMS_DIR, "params-%s.json" % env)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return default_params(env)


def replace_dict(main_dict, input_dict):
    """run_model_based_rl.py:35-51: recursive in-place override; unknown keys are errors."""
    for key, value in input_dict.items():
        if key not in main_dict:
            raise KeyError("replace: key %r is not a parameter" % (key,))
        if isinstance(value, dict) and isinstance(main_dict[key], dict):
            replace_dict(main_dict[key], value)
        else:
            main_dict[key] = value


def write_all(directory=PARAMS_DIR):
    os.makedirs(directory, exist_ok=True)
    for env in _ENVS:
        with open(os.path.join(directory, "params-%s.json" % env), "w") as f:
            json.dump(default_params(env), f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    write_all()
    print("wrote", sorted(os.listdir(PARAMS_DIR)))
