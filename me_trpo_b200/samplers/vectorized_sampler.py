"""VectorizedSampler (samplers/vectorized_sampler.py): socket B2.

`obtain_samples` keeps the reference's contract -- a list of COMPLETED path dicts with
observations (pre-step), UNCLIPPED actions, rewards, agent_infos{mean, log_std} -- but produces it
with ONE persistent kernel launch for the whole batch instead of a Python loop with two session
round trips per env-step.  `obtain_samples_flat` returns the time-major device buffers directly
(what an on-device policy update consumes)."""
import numpy as np
import torch

from .base import BaseSampler
from ..parallel import local_reset_pool
from ..rollout import EnsembleRollout


class VectorizedSampler(BaseSampler):
    """`dist_ctx` (parallel.DistContext): with G > 1 ranks the `n_envs` parallel rollouts are
    block-sharded over the ranks -- rank r owns rows [lo, hi) -- with global-row noise keys and the
    matching slice of the reset pool, so the union of the ranks' buffers is bit-identical to a
    single-GPU run; rank 0 draws the global start states from the (real) simulator and broadcasts
    them, every rank keeps its slice."""

    def __init__(self, algo, n_envs=None, seed=0, reset_pool_size=None, dist_ctx=None):
        super().__init__(algo, dist_ctx=dist_ctx)
        self.n_envs = n_envs
        self.seed = int(seed)
        self.reset_pool_size = reset_pool_size
        self.rollout = None
        self._calls = 0

    def start_worker(self):
        algo = self.algo
        n_envs = self.n_envs
        if n_envs is None:   # reference rule (:26-27); pass n_envs explicitly to use the whole GPU
            n_envs = max(1, min(int(algo.batch_size / algo.max_path_length), 100))
        env = algo.env
        if not getattr(env, "vectorized", False):
            raise RuntimeError("VectorizedSampler needs a vectorized (NeuralNetEnv) environment")
        pol = algo.policy
        if self.rollout is not None:      # the reference rebuilds the vec env every iteration (:23-40)
            self.rollout.close()
        ctx = self.dist_ctx
        self._lo, self._hi = ctx.shard_rows(n_envs) if ctx is not None else (0, n_envs)
        if self._hi <= self._lo:
            raise RuntimeError("more ranks than parallel rollouts (n_envs=%d)" % n_envs)
        self.rollout = EnsembleRollout(env.env_name, env.n_models, self._hi - self._lo, algo.max_path_length,
                                       hidden=env.hidden, policy_hidden=pol.hidden_sizes,
                                       sam_mode=env.sam_mode, policy_out_tanh=pol.output_tanh,
                                       device=env.device, row_offset=self._lo,
                                       precision=getattr(env, "precision", "bf16"))
        self.rollout.set_dynamics_ensemble(env.models)
        self.rollout.set_normalization(**env.norm)
        self.env_spec = env.spec
        self._n_envs = n_envs             # GLOBAL number of parallel rollouts

    def shutdown_worker(self):
        if self.rollout is not None:
            self.rollout.close()
            self.rollout = None

    def _steps_for_batch(self):
        """Number of kernel steps so that completed paths cover batch_size samples: with a common
        timeout every row completes a path each max_path_length steps (:60 loops until then)."""
        algo = self.algo
        per_round = self._n_envs * algo.max_path_length
        rounds = -(-int(algo.batch_size) // per_round)
        return rounds * algo.max_path_length

    def _start_states(self, T):
        """(init[B_local,S], pool) of this rank: the global draw, sliced to the rank's rows."""
        algo, env = self.algo, self.algo.env
        B = self._n_envs
        n_resets = -(-T // algo.max_path_length)
        if self._early_termination():
            n_resets *= 4         # paths may end long before the timeout: more simulator resets per row
        R = int(self.reset_pool_size or B * n_resets)
        ctx = self.dist_ctx
        if ctx is None or not ctx.distributed or ctx.rank == 0:      # only rank 0 touches the simulator
            init = np.asarray(env.reset_sampler(B), np.float32)      # the initial reset() (:49)
            pool = np.asarray(env.reset_sampler(R), np.float32)
        else:
            init = pool = None
        if ctx is not None and ctx.distributed:
            init, pool = ctx.broadcast_arrays([init, pool], env.device or "cuda")
        if self._hi - self._lo != B:
            pool = local_reset_pool(pool, B, self._lo, self._hi, n_resets)
            init = init[self._lo:self._hi]
        return init, pool

    @staticmethod
    def completed_samples_per_step(done):
        """cum[t] = number of samples in paths COMPLETED by the end of step t (the `n_samples` of
        samplers/vectorized_sampler.py:96-104), from the time-major done flags [T,B] (device)."""
        T, B = done.shape
        d = done.bool()
        t_idx = torch.arange(T, device=done.device, dtype=torch.int64)[:, None].expand(T, B)
        last = torch.cummax(torch.where(d, t_idx, torch.full_like(t_idx, -1)), dim=0).values   # last done <= t
        prev = torch.cat([torch.full((1, B), -1, dtype=torch.int64, device=done.device), last[:-1]], 0)
        length = torch.where(d, t_idx - prev, torch.zeros_like(t_idx))                         # path length at its end
        return torch.cumsum(length.sum(dim=1), dim=0)

    def _early_termination(self):
        """Envs whose is_done ends paths before the timeout (only Ant defines is_done;
        env_helpers.py:537): the number of steps the reference loop takes then depends on the data."""
        return self.algo.env.env_name == "ant"

    def obtain_samples_flat(self, itr, determ=False, n_steps=None, check=True):
        """Time-major device buffers of one batch (this rank's rows).  `check` waits for the kernel
        and raises if it aborted on an internal wait timeout (the buffers would be partially
        written); pass False only if the caller checks `self.rollout.synchronize()` itself before
        consuming them."""
        pol, algo = self.algo.policy, self.algo
        T = int(n_steps or self._steps_for_batch())
        self.rollout.set_policy(pol.W, pol.b, pol.log_std)
        if n_steps is None and self._early_termination():
            return self._obtain_until_batch_size(T, determ)
        init, pool = self._start_states(T)
        out = self.rollout.run(T, init, pool, seed=self.seed, offset=self._calls * (1 << 20),
                               determ=determ)
        self._calls += 1
        if check:
            self.rollout.synchronize()
        return out

    def _obtain_until_batch_size(self, T, determ):
        """The reference's stop rule for early-terminating paths (`while n_samples < batch_size`,
        samplers/vectorized_sampler.py:60,96-105): the loop ends after the first step at which the
        COMPLETED paths hold batch_size samples.  With paths ending before the timeout that step
        is not a multiple of the horizon and is not known in advance: the kernel runs in chunks
        (metrpo_rollout_continue) until the count is reached and the buffers are cut there; paths
        still open at that step are dropped by process_samples_flat's validity mask, like the
        reference drops its running_paths."""
        algo = self.algo
        T_max = int(algo.max_path_length)
        cap = T + 4 * T_max
        B = self._hi - self._lo
        dev = self.rollout.device
        S, A = self.rollout.S, self.rollout.A
        init, pool = self._start_states(cap)
        full = dict(obs=torch.empty(cap, B, S, device=dev), act=torch.empty(cap, B, A, device=dev),
                    mean=torch.empty(cap, B, A, device=dev), rew=torch.empty(cap, B, device=dev),
                    done=torch.empty(cap, B, dtype=torch.uint8, device=dev))
        base = self._calls * (1 << 20)
        self._calls += 1
        t0, chunk, final = 0, T, None
        while True:
            view = {k: v[t0:t0 + chunk] for k, v in full.items()}
            if t0 == 0:
                res = self.rollout.run(chunk, init, pool, seed=self.seed, offset=base, determ=determ, out=view)
            else:
                res = self.rollout.run_continue(chunk, pool, seed=self.seed, offset=base + t0, determ=determ, out=view)
            self.rollout.synchronize()
            final = res["final_states"]
            t0 += chunk
            cum = self.completed_samples_per_step(full["done"][:t0])
            ctx = self.dist_ctx
            if ctx is not None and ctx.distributed:          # n_samples counts the paths of ALL ranks
                import torch.distributed as dist
                dist.all_reduce(cum, group=ctx.group)
            hit = torch.nonzero(cum >= int(algo.batch_size))
            if hit.numel():
                t_stop = int(hit[0].item()) + 1
                break
            if t0 + T_max > cap:
                raise RuntimeError("obtain_samples: batch_size not reached within %d steps" % cap)
            chunk = T_max
        out = {k: v[:t_stop] for k, v in full.items()}
        out["final_states"] = final
        return out

    def obtain_samples(self, itr, determ=False):
        """The reference's contract (list of completed path dicts on the HOST).  The device->host
        copy of the trajectory is overlapped with the rollout itself (EnsembleRollout.run_to_host)."""
        pol = self.algo.policy
        T = int(self._steps_for_batch())
        self.rollout.set_policy(pol.W, pol.b, pol.log_std)
        if self._early_termination():     # data-dependent number of steps (reference stop rule)
            flat = self._obtain_until_batch_size(T, determ)
            log_std = self.algo.policy.log_std.clamp(min=float(np.log(1e-6))).cpu().numpy()
            return paths_from_flat({k: v.cpu().numpy() for k, v in flat.items()}, log_std)
        init, pool = self._start_states(T)
        host, self._dev_out = self.rollout.run_to_host(T, init, pool, seed=self.seed,
                                                       offset=self._calls * (1 << 20), determ=determ,
                                                       n_chunks=max(1, min(8, T // 32)))
        self._calls += 1
        self.rollout.synchronize()
        log_std = self.algo.policy.log_std.clamp(min=float(np.log(1e-6))).cpu().numpy()
        return paths_from_flat({k: v.numpy() for k, v in host.items()}, log_std)


def paths_from_flat(flat, log_std):
    """Cut time-major buffers [T,B,...] into the reference's list of completed path dicts, in the
    order the reference appends them (by finishing step, then row)."""
    T, B = flat["rew"].shape
    done = flat["done"].astype(bool)
    start = np.zeros(B, np.int64)
    paths = []
    for t in np.nonzero(done.any(axis=1))[0]:
        for b in np.nonzero(done[t])[0]:
            sl = slice(start[b], t + 1)
            L = t + 1 - start[b]
            paths.append(dict(
                observations=flat["obs"][sl, b], actions=flat["act"][sl, b], rewards=flat["rew"][sl, b],
                env_infos={}, agent_infos=dict(mean=flat["mean"][sl, b],
                                               log_std=np.tile(log_std, (L, 1)))))
            start[b] = t + 1
    return paths
