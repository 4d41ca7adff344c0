#!/usr/bin/env python
"""Benchmark of the ensemble imaginary-rollout hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference path's CPU restatement

One "step" = one pass of the hot path over one batch = the whole-horizon fused rollout of
BASELINE.json configs[1]: half-cheetah, 5-model ensemble, 4096 parallel rollouts (per GPU),
horizon 1000, sam_mode step_rand, Philox noise on device.  Unit = one (model, row, timestep)
dynamics evaluation; value = units of all ranks / time (weak scaling: 4096 rows per GPU).

For N > 1 launch with torch.distributed.run (one rank per GPU); rows are sharded by rank with
global-row noise keys and NO data-path collective; only the timing is reduced (max).
"""
import argparse
import json
import os
import sys
import threading
import time


def _cpu_threads_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; NumPy's BLAS reads the
    variable when it is first imported.  The CPU legs of this file (--impl reference, and the
    cpu_baseline of the CUDA arm at N = 1) must see all host cores, so the variables are set
    explicitly BEFORE numpy is imported; the thread count actually used is read back through
    threadpoolctl and reported in the JSON line."""
    want_all = "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1
    if want_all:
        n = str(os.cpu_count() or 1)
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = n


_cpu_threads_env()
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs: name -> (env, K, total rows, horizon, hidden, GPUs the config names)
WORKLOADS = {
    "swimmer": ("swimmer", 5, 100, 200, 512, 1),           # configs[0]: the reference's own JSON shape
    "half-cheetah": ("half-cheetah", 5, 4096, 1000, 1024, 1),     # configs[1]  (the headline metric)
    "hopper": ("hopper", 10, 4096, 1000, 1024, 1),                # configs[2]
    "ant": ("ant", 20, 16384, 1000, 1024, 4),                     # configs[3]
    "humanoid": ("humanoid", 20, 65536, 1000, 1024, 8),           # configs[4]
}
ENV, K_MODELS, B_ROWS, HORIZON, HIDDEN = "half-cheetah", 5, 4096, 1000, 1024
E2E_CHUNKS = int(os.environ.get("METRPO_E2E_CHUNKS", "4"))


def select_workload(args, world):
    """Sets the module-level shape from --config / --scaling.  weak (default): the per-GPU share of
    the config is fixed (total rows / the GPU count the config names) and the job grows with N;
    strong: the config's TOTAL rows are split over the N ranks."""
    global ENV, K_MODELS, B_ROWS, HORIZON, HIDDEN
    env, K, rows, T, hidden, gpus = WORKLOADS[args.config]
    per_gpu = rows // gpus if args.scaling == "weak" else -(-rows // world)
    ENV, K_MODELS, B_ROWS, HORIZON, HIDDEN = env, K, per_gpu, T, hidden
    return workload_config(args, world)


def workload_config(args, world):
    """The `config` object of the JSON line -- identical for both arms (--impl cuda / reference)."""
    env, K, rows, T, hidden, gpus = WORKLOADS[args.config]
    total = B_ROWS * world
    return {"workload": "%s K=%d B=%d/GPU horizon=%d step_rand (BASELINE configs[%d])"
                        % (env, K, B_ROWS, T, list(WORKLOADS).index(args.config)),
            "hidden": hidden, "rows_per_gpu": B_ROWS, "rows_total": total, "noise": "philox",
            "l2": "512 MiB flush between timed iterations; weights (bf16) are re-streamed from L2 by design "
                  "inside each launch", "parallelism": "rows sharded, no collective"}
METRIC = "simulated env steps/sec (ensemble x batch x horizon)"
UNIT = "units/s"


def flops_per_step(spec, K, B, T, hidden):
    din = spec["S"] + spec["A"] - spec["drop"]
    f_dyn = 2.0 * (din * hidden + hidden * hidden + hidden * spec["S"])
    dims = [spec["S"]] + list(spec["policy_hidden"]) + [spec["A"]]
    f_pol = 2.0 * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    return K * B * T * f_dyn + B * T * f_pol


def make_problem(seed=0, B=None):
    """Synthetic weights / states of the BASELINE shape: Xavier-uniform nets with the
    dynamics output layer scaled by 0.1 so that a 1000-step rollout of a random net stays finite.
    Product-side generator (me_trpo_b200/synthetic.py): the CUDA arm never imports oracle/."""
    from me_trpo_b200 import synthetic
    return synthetic.make_problem(ENV, K_MODELS, B_ROWS if B is None else B, hidden=HIDDEN, seed=seed)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def blas_info():
    try:
        from threadpoolctl import threadpool_info
        return [{"lib": i.get("internal_api"), "threads": i.get("num_threads")} for i in threadpool_info()]
    except Exception:
        return []


def cpu_reference_rate(steps_sample, threads=None, seed=0, rows=None, horizon=None):
    """Times the reference path's CPU restatement (oracle/, reference-faithful structure:
    per-step Python loop, all-K fp32 NumPy forward, per-env bookkeeping) on `steps_sample` steps of
    the workload (optionally with `rows` parallel envs instead of the workload's).  `threads`
    limits the BLAS pool (None = all host cores).  Returns (units/s, seconds, BLAS threads used)."""
    from oracle import rollout as orl
    from threadpoolctl import threadpool_limits
    B = int(rows or B_ROWS)
    T = int(horizon or HORIZON)
    spec, models, pol, norm, init, pool = make_problem(seed, B)
    noise = orl.PhiloxNoise(1, 0, 0, "step_rand")
    ve = orl.VecSimpleEnvOracle(ENV, models, norm, B, T, "step_rand", noise, pool,
                                spec["S"], spec["A"], spec["drop"], np.float32, "fp32")
    n_threads = int(threads or os.cpu_count() or 1)
    with threadpool_limits(limits=n_threads):
        t0 = time.perf_counter()
        orl.obtain_samples(ve, pol, init, batch_size=B * T, max_steps=steps_sample)
        dt = time.perf_counter() - t0
    used = max([i["threads"] or 1 for i in blas_info()] + [1])
    return K_MODELS * B * steps_sample / dt, dt, min(n_threads, used)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (restated in oracle/ and pinned to the
    reference's code by tests/test_ref_fixtures.py; TF1.4 + rllab + MuJoCo cannot run here) on the
    host cores of the box.  Headline value: all BLAS threads, BASELINE shape.  Also reported: one
    BLAS thread (the reference pins TF to 1 thread, utils.py:229-232) and the reference's own
    sampler shape (100 parallel envs, samplers/vectorized_sampler.py:26-27)."""
    if rank != 0:
        return
    config = select_workload(args, world)
    sample = 10 if B_ROWS >= 1024 else 100   # env-steps per bench step (bounded sample; steps are homogeneous)
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt, cores = cpu_reference_rate(sample, seed=i)
        if i >= args.warmup:
            rates.append((r, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates]) * 1e3)
    r1, dt1, _ = cpu_reference_rate(max(1, sample // 5), threads=1)
    r100, dt100, _ = cpu_reference_rate(100, rows=100)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "host_cores": os.cpu_count(), "blas": blas_info(),
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "sample": "%d env-steps x %d models x %d rows per bench step (ms_per_step is for this "
                                   "sample, not a %d-step horizon), NumPy fp32, reference-faithful Python loop "
                                   "(oracle/rollout.py)" % (sample, K_MODELS, B_ROWS, HORIZON),
                         "one_thread": {"value": r1, "unit": UNIT, "cores": 1, "seconds": dt1,
                                        "why": "the reference pins TF to 1 intra/inter-op thread (utils.py:229-232)"},
                         "reference_sampler_shape": {"value": r100, "unit": UNIT, "rows": 100, "seconds": dt100,
                                                     "why": "n_envs is capped at 100 (samplers/vectorized_sampler.py:26-27)"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cuda(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from me_trpo_b200.rollout import EnsembleRollout

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the rollout path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    config = select_workload(args, world)
    spec, models, pol, norm, init, pool = make_problem(0)
    T = HORIZON
    row_offset = rank * B_ROWS                         # rows sharded by rank, global-row noise keys
    ro = EnsembleRollout(ENV, K_MODELS, B_ROWS, T, hidden=HIDDEN, device=dev, row_offset=row_offset)
    ro.set_dynamics_ensemble(models)
    ro.set_normalization(**norm)
    ro.set_policy(pol["W"], pol["b"], pol["log_std"])
    init_d, pool_d = torch.tensor(init, device=dev), torch.tensor(pool, device=dev)
    init_h, pool_h = torch.tensor(init).pin_memory(), torch.tensor(pool).pin_memory()
    out = ro.run(T, init_d, pool_d, seed=1, offset=0)   # allocates the trajectory buffers once
    ro.synchronize()
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i, e2e):
        flush.fill_(i & 0xFF)                           # L2 flush between iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:                                         # host buffers in, host buffers out
            a = init_h.to(dev, non_blocking=True)
            b = pool_h.to(dev, non_blocking=True)
            # public API of the host-side sampler: chunked launches with the D2H copy of finished
            # steps overlapped with the rest of the horizon; returns after queuing, the end event
            # below covers the last copy
            ro.run_to_host(T, a, b, host_out=host_out, dev_out=out, seed=1, offset=i * T, n_chunks=E2E_CHUNKS)
        else:
            ro.run(T, init_d, pool_d, seed=1, offset=i * T, out=out)
        e1.record()
        return e0, e1

    def timed(e2e):
        for i in range(args.warmup):
            one_step(i, e2e)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        evs = [one_step(args.warmup + i, e2e) for i in range(args.steps)]
        barrier()
        sampler.stop_flag = True
        sampler.join()
        ro.synchronize()
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), ms, sampler.summary()

    total_ms, per_step, clocks = timed(False)
    kernel_variant = ro.last_kernel()
    total_ms_e2e, _, _ = timed(True)
    finite = bool(torch.isfinite(out["obs"]).all().item()) and bool(torch.isfinite(host_out["obs"]).all().item())

    # The rest of the TRPO inner iteration on the trajectory of the last step (reported, not part of
    # `value`): process_samples + baseline fit + natural-gradient update, all on the device.
    trpo_info = None
    try:
        from me_trpo_b200.trpo import PolicyUpdate
        pu = PolicyUpdate([spec["S"]] + list(spec["policy_hidden"]) + [spec["A"]], device=dev)
        if world > 1:
            pu.enable_allreduce()
        parts = []
        for W_, b_ in zip(pol["W"], pol["b"]):
            parts += [W_.ravel(), b_.ravel()]
        parts.append(pol["log_std"])
        theta = torch.tensor(np.concatenate(parts).astype(np.float32), device=dev)
        ls = torch.tensor(pol["log_std"], device=dev)
        N = B_ROWS * T

        def iteration():
            pr = pu.process(out["obs"], out["rew"], out["done"], discount=1.0)
            pu.fit_baseline(out["obs"], pr["ret"], pr["valid"], out["done"])
            th = theta.clone()
            return pu.update(th, out["obs"].reshape(N, -1), out["act"].reshape(N, -1), pr["adv"].reshape(N),
                             out["mean"].reshape(N, -1), ls, valid=pr["valid"].reshape(N))
        iteration()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = iteration()
        e1.record()
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        iv = info.cpu().numpy()
        trpo_info = {"ms": float(tms.item()), "samples_per_gpu": N, "accepted": bool(iv[4] == 1.0),
                     "mean_kl": float(iv[2]), "allreduce": pu.allreduce_mode,
                     "what": "process_samples + baseline fit + TRPO update (1 gradient, 10 "
                     "Fisher-vector products, line search) on the last step's trajectory, device resident"}
        pu.close()
    except Exception as exc:   # reported, never fatal for the headline metric
        trpo_info = {"error": repr(exc)}

    # ---- the ensemble-fit iteration of the same config: K models x batch 1000, all
    #      contractions on the tcgen05 TF32 GEMM; reported next to the headline, never part of it ----
    fit_info = None
    if world == 1:      # single-GPU runs only: no rank may run ahead of a collective tear-down
        try:
            from me_trpo_b200.dynamics import EnsembleFit
            S_, A_, drop_ = spec["S"], spec["A"], spec["drop"]
            fb, fn = 1000, 200000
            fit = EnsembleFit(S_, A_, drop_, HIDDEN, K_MODELS, max_rows=1024, precision="tf32")
            fit.set_ensemble(models)                 # the synthetic ensemble the rollout above used
            fit.set_normalization(**norm); fit.reset_adam()
            xd = torch.randn(fn, S_ + A_, device=dev); yd = xd[:, :S_] + 0.1 * torch.randn(fn, S_, device=dev)
            for j in range(5):
                fit.step(xd, yd, fb, 1e-3, seed=1, offset=j, want_losses=False)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for j in range(50):
                fit.step(xd, yd, fb, 1e-3, seed=1, offset=5 + j, want_losses=False)
            f1.record(); torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / 50
            fit_info = {"ms": fms, "samples_per_s": K_MODELS * fb / (fms * 1e-3), "launches": fit.last_launches(),
                        "what": "one Adam step of all %d models on independent minibatches of %d rows (gather, "
                                "forward, MSE, backward, Adam; TF32 tcgen05 GEMMs, fp32 accumulate)" % (K_MODELS, fb)}
            fit.close()
        except Exception as exc:
            fit_info = {"error": repr(exc)}

    units_per_step = K_MODELS * B_ROWS * T * world
    value = units_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = units_per_step * args.steps / (total_ms_e2e * 1e-3)
    h2d = init_h.numel() * 4 + pool_h.numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in out.values())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        peak_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops (burst, cuBLAS bf16)" if peaks else "fallback 1.59 PFLOP/s"
        kernel_ms = float(np.mean(per_step))            # the step IS one launch of the persistent kernel
        achieved_tf = flops_per_step(spec, K_MODELS, B_ROWS, T, HIDDEN) / (kernel_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        if args.config == "half-cheetah":
            for name in ("r2_rollout_ncu_summary.json", "r1_rollout_ncu_summary.json"):
                try:
                    traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["dram_bytes_per_launch_T1000_est"]
                    traffic_src = "profiles/%s (ncu --set full capture of this kernel and shape; not re-measured in this run)" % name
                    break
                except Exception:
                    pass
        cpu = None
        if world == 1:            # the contract: rank 0 at N = 1 only
            cpu_rate, cpu_dt, cores = cpu_reference_rate(40 if B_ROWS >= 1024 else 200)
            cpu = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port", "host_cores": os.cpu_count(),
                   "blas": blas_info(),
                   "sample": "%d of %d env-steps of the same workload (homogeneous steps), NumPy fp32, %d BLAS "
                             "threads, reference-faithful loop; %.1f s" % (40 if B_ROWS >= 1024 else 200, T, cores, cpu_dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "metrpo::rollout_kernel", "kernel_ms": kernel_ms,
                         # every launch runs ~50-90 ms back to back under sw_power_cap (see `clocks`): the
                         # sustained cuBLAS figure is the power-limited reference for such a kernel
                         "peak_sustained": peak_sus, "frac_of_sustained": achieved_tf / peak_sus},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "gpu_launches_per_step": E2E_CHUNKS,
                    "what": "pinned host init/reset states -> device, fused rollout in %d chained launches, whole "
                            "trajectory (obs, act, mean, rew, done) -> pinned host, copy of finished steps "
                            "overlapped with the remaining horizon (EnsembleRollout.run_to_host)" % E2E_CHUNKS},
            "gpu_launches": args.steps * ro.last_launches(),
            "rollout_kernel_variant": {0: "single-stream", 1: "two-stream", 2: "two-stream, column split"}[kernel_variant],
            "clocks": clocks, "finite": finite, "trpo_half_of_iteration": trpo_info,
            "fit_iteration": fit_info,
        }
        print(json.dumps(line), flush=True)
    ro.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="half-cheetah", choices=list(WORKLOADS),
                    help="BASELINE.json config (default: configs[1], the one the metric is quoted on)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's per-GPU share on every rank; strong: its total rows split over N")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            sys.stderr.write("bench.py: --gpus %d needs torch.distributed.run (one rank per GPU); running 1 GPU\n" % args.gpus)
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
