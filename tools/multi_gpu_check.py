"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL):
  1. the row-sharded fused rollout of all ranks, gathered, equals rank 0's unsharded rollout bit for bit
     (global-row Philox keys; no data-path collective);
  2. the TRPO update with its accumulators all-reduced over NCCL (gradient, every Fisher-vector product,
     (loss, kl) pairs, advantage moments, baseline normal equations) yields the same new parameters on
     every rank, equal (to fp64-accumulation round-off) to the single-GPU update on the whole batch."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from me_trpo_b200 import synthetic
from me_trpo_b200.parallel import shard_rows
from me_trpo_b200.rollout import EnsembleRollout
from me_trpo_b200.trpo import PolicyUpdate

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
env, K, B, T, T_max, hidden = "half-cheetah", 5, 1024, 40, 20, 256
spec, models, pol, norm, init, pool = synthetic.make_problem(env, K, B, hidden=hidden, pool_rows=2 * B)
lo, hi = shard_rows(B, rank, world)
n_res = T // T_max


def rollout(rows_lo, rows_hi):
    nb = rows_hi - rows_lo
    local_pool = np.stack([pool[(n * B + i) % len(pool)] for n in range(n_res + 1) for i in range(rows_lo, rows_hi)])
    ro = EnsembleRollout(env, K, nb, T_max, hidden=hidden, device=dev, row_offset=rows_lo)
    ro.set_dynamics_ensemble(models); ro.set_normalization(**norm); ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    out = ro.run(T, init[rows_lo:rows_hi], local_pool, seed=5)
    ro.synchronize()
    return ro, out


def update(out, allreduce, mode="auto", reps=1):
    pu = PolicyUpdate([spec["S"], 32, 32, spec["A"]], device=dev)
    if allreduce:
        pu.enable_allreduce(mode=mode)
        modes_used.append(pu.allreduce_mode)
    pr = pu.process(out["obs"], out["rew"], out["done"], discount=0.99)
    coeffs = pu.fit_baseline(out["obs"], pr["ret"], pr["valid"], out["done"])
    pr = pu.process(out["obs"], out["rew"], out["done"], baseline_coeffs=coeffs, discount=0.99)
    parts = []
    for W, b in zip(pol["W"], pol["b"]):
        parts += [W.ravel(), b.ravel()]
    parts.append(pol["log_std"])
    theta = torch.tensor(np.concatenate(parts).astype(np.float32), device=dev)
    N = out["rew"].numel()
    ls = torch.tensor(pol["log_std"], device=dev)
    theta0 = theta.clone()
    info = pu.update(theta, out["obs"].reshape(N, -1), out["act"].reshape(N, -1), pr["adv"].reshape(N),
                     out["mean"].reshape(N, -1), ls, valid=pr["valid"].reshape(N))
    torch.cuda.synchronize()
    if reps > 1:      # timing of the whole update (device events, max over ranks)
        ms = []
        for _ in range(reps):
            th = theta0.clone()
            if allreduce:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pu.update(th, out["obs"].reshape(N, -1), out["act"].reshape(N, -1), pr["adv"].reshape(N),
                      out["mean"].reshape(N, -1), ls, valid=pr["valid"].reshape(N))
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ms)], device=dev, dtype=torch.float64)
        if allreduce:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timings.append(float(t.item()))
    return theta.cpu().numpy(), info.cpu().numpy(), coeffs.cpu().numpy()

modes_used, timings = [], []

ro_s, out_s = rollout(lo, hi)
res = {}
# 1. gather the shards on rank 0 and compare with the unsharded rollout
sizes = [shard_rows(B, r, world)[1] - shard_rows(B, r, world)[0] for r in range(world)]
gathered = [torch.empty(T, n, spec["S"], device=dev) for n in sizes]
dist.all_gather(gathered, out_s["obs"].contiguous())
if rank == 0:
    ro_f, out_f = rollout(0, B)
    res["rollout_bitexact"] = bool(torch.equal(torch.cat(gathered, 1), out_f["obs"]))
# 2. sharded TRPO update with NCCL all-reduce vs single-GPU update on everything
theta_n, info_n, coeffs_n = update(out_s, True, mode="nccl", reps=5)       # host callback into NCCL per reduction
theta_s, info_s, coeffs_s = update(out_s, True, mode="p2p", reps=5)        # in-library one-shot NVLink all-reduce
all_theta = [torch.empty(len(theta_s), device=dev) for _ in range(world)]
dist.all_gather(all_theta, torch.tensor(theta_s, device=dev))
if rank == 0:
    theta_f, info_f, coeffs_f = update(out_f, False)
    res["theta_identical_across_ranks"] = bool(all(torch.equal(all_theta[0], t) for t in all_theta))
    res["theta_max_abs_diff_vs_single_gpu"] = float(np.max(np.abs(theta_s - theta_f)))
    res["theta_step_norm"] = float(np.linalg.norm(theta_f - np.concatenate([p.ravel() for wb in zip(pol["W"], pol["b"]) for p in wb] + [pol["log_std"]])))
    res["baseline_coeffs_max_abs_diff"] = float(np.max(np.abs(coeffs_s - coeffs_f)))
    res["accepted"] = [float(info_s[4]), float(info_f[4])]
    res["mean_kl"] = [float(info_s[2]), float(info_f[2])]
    res["world_size"] = world
    res["allreduce_modes"] = modes_used
    res["update_ms_nccl_callback"], res["update_ms_p2p"] = timings[0], timings[1]
    res["theta_max_abs_diff_p2p_vs_nccl"] = float(np.max(np.abs(theta_s - theta_n)))
    res["samples_per_rank"] = int(out_s["rew"].numel())
    ok = (res["rollout_bitexact"] and res["theta_identical_across_ranks"]
          and res["theta_max_abs_diff_vs_single_gpu"] <= 1e-5 * max(1.0, res["theta_step_norm"]))
    res["ok"] = bool(ok)
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "multi_gpu_check.json"), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
