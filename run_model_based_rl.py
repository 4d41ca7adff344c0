#!/usr/bin/env python
"""Entry point with the reference's command line (run_model_based_rl.py:54-184):

    python run_model_based_rl.py trpo -env half-cheetah [-seed 0] [-prefix NAME] [-replace "{...}"] [-f]

Loads params/params-<env>.json, applies -replace, and runs the ME-TRPO loop on the B200-native
components (me_trpo_b200/training.py).  Differences, all explicit: no rllab run_experiment_lite
(the experiment runs in-process, snapshots under data/local/<prefix>/<prefix>_seed<seed>), -ec2 is
refused (the EC2 launcher is out of scope), and the real simulator is whatever
me_trpo_b200.real_env.make_real_env returns (MuJoCo adapters are registered there by the user;
offline it is a synthetic stand-in).  Extra options: -param_path, -sweeps, -n_envs, -snapshot_dir."""
import argparse
import ast
import logging
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS = ["half-cheetah", "snake", "hopper", "ant", "swimmer", "humanoid"]


def main(argv=None):
    parser = argparse.ArgumentParser(description="run experiment options")
    parser.add_argument("algo")
    parser.add_argument("-ec2", action="store_true", default=False)
    parser.add_argument("-env")
    parser.add_argument("-prefix")
    parser.add_argument("-n", type=int, default=10)
    parser.add_argument("-seed", type=int, default=0)
    parser.add_argument("-replace", type=str, default="{}")
    parser.add_argument("-f", action="store_true", default=False, help="force")
    parser.add_argument("-param_path", default=None)
    parser.add_argument("-sweeps", type=int, default=None, help="override sweep_iters")
    parser.add_argument("-n_envs", type=int, default=None, help="parallel imaginary rollouts (default: reference rule)")
    parser.add_argument("-snapshot_dir", default=None)
    options = parser.parse_args(argv)
    from me_trpo_b200 import params as P
    if options.env not in ENVS:
        raise ValueError("Value Error: not implemented.")                      # run_model_based_rl.py:79
    if options.ec2:
        raise NotImplementedError("-ec2: the EC2 launcher is outside this repository's scope")
    params = P.load_params(options.env, options.param_path)
    P.replace_dict(params, ast.literal_eval(options.replace))                  # :92-93 (eval in the reference)
    if params["algo"] != options.algo:                                         # :95-107
        if not options.f:
            response = input("The algo option in params is %s. Are you sure you want to run %s [y/N]?"
                             % (params["algo"], options.algo))
            if response not in ("Y", "y"):
                sys.exit()
        params["algo"] = options.algo
    if params["env"] != options.env:                                           # :116-124
        params["env"] = options.env
    prefix = options.prefix or params["env"]
    snapshot_dir = options.snapshot_dir or os.path.join(ROOT, "data", "local", prefix, "%s_seed%d" % (prefix, options.seed))
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    from me_trpo_b200.training import train
    out = train(dict(mode="local", params=params, seed=options.seed), snapshot_dir=snapshot_dir,
                sampler_n_envs=options.n_envs, sweep_iters=options.sweeps)
    print("done: %d sweeps, progress.csv in %s" % (len(out["progress"]), snapshot_dir))
    return out


if __name__ == "__main__":
    main()
