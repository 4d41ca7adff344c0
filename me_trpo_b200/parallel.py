"""Multi-GPU decomposition of the ME-TRPO inner loop, one process per GPU over
torch.distributed (NCCL on GPUs; the same host logic runs over gloo in the CPU tests).

  rollout          rows are independent for the whole horizon: rank r of G owns a contiguous block
                   of rows and the rollout needs NO collective.  Noise streams are keyed by GLOBAL
                   row (cfg.row_offset) and each rank builds the slice of the reset pool its rows
                   will consume, so a sharded run equals the unsharded one bit for bit.
  TRPO update      every rank reduces its own samples; the (tiny) accumulators -- advantage moments,
                   baseline normal equations, gradient, each Fisher-vector product, (loss, kl)
                   pairs -- are all-reduced (PolicyUpdate.enable_allreduce), so every rank takes
                   the identical step.
  ensemble fit     the K models are independent with independent minibatches
                   (model_based_rl.py:964-970): rank r fits models k = r (mod G); the owners then
                   broadcast their weights so that every rank holds the whole ensemble again.
  real-env data    collected by rank 0 and broadcast.
"""
import numpy as np


def shard_rows(n_rows, rank, world_size):
    """Contiguous [lo, hi) block of rank; the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_reset_pool(pool, n_rows, lo, hi, n_resets):
    """Rows [lo, hi) of an unsharded run take their n-th reset from pool[(n * n_rows + i) % R]
    (per-row rule of the kernel); the sharded handle indexes its pool with LOCAL sizes, so hand it
    exactly those entries, ordered (n, local row)."""
    pool = np.asarray(pool)
    R = len(pool)
    idx = [(n * n_rows + i) % R for n in range(int(n_resets)) for i in range(lo, hi)]
    return pool[idx]


class DistContext:
    """rank / world / process group + the handful of collectives the loop needs.  world_size == 1
    makes every method a no-op, so single-GPU callers never touch torch.distributed."""

    def __init__(self, rank=0, world_size=1, group=None):
        self.rank, self.world_size, self.group = int(rank), int(world_size), group

    @classmethod
    def from_env(cls, device=None):
        """Under torchrun (WORLD_SIZE > 1): join (or reuse) the default process group."""
        import os
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return cls()
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            on_gpu = device is not None and torch.device(device).type == "cuda"
            kw = dict(device_id=torch.device(device)) if on_gpu else {}
            dist.init_process_group("nccl" if on_gpu else "gloo", **kw)
        return cls(dist.get_rank(), dist.get_world_size())

    @property
    def distributed(self):
        return self.world_size > 1

    # -- partitioning ---------------------------------------------------------------------------
    def shard_rows(self, n_rows):
        return shard_rows(n_rows, self.rank, self.world_size)

    def owner_of_model(self, k):
        return int(k) % self.world_size

    def my_models(self, n_models):
        return [k for k in range(int(n_models)) if self.owner_of_model(k) == self.rank]

    # -- collectives ----------------------------------------------------------------------------
    def broadcast_(self, tensor, src=0):
        if self.distributed:
            import torch.distributed as dist
            dist.broadcast(tensor, src=src, group=self.group)
        return tensor

    def barrier(self):
        if self.distributed:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def broadcast_policy(self, policy, src=0):
        """Rank `src`'s policy parameters on every rank (before each rollout)."""
        if self.distributed:
            flat = policy.flat_params()
            self.broadcast_(flat, src)
            policy.set_flat_params(flat)

    def broadcast_arrays(self, arrays, device, src=0):
        """NumPy arrays known on rank `src` (None elsewhere) -> the same arrays on every rank."""
        if not self.distributed:
            return arrays
        import torch
        import torch.distributed as dist
        meta = [[(a.shape, str(a.dtype)) for a in arrays]] if self.rank == src else [None]
        dist.broadcast_object_list(meta, src=src, group=self.group)
        out = []
        for i, (shape, dtype) in enumerate(meta[0]):
            t = torch.as_tensor(np.ascontiguousarray(arrays[i])) if self.rank == src \
                else torch.empty(shape, dtype=getattr(torch, dtype))
            t = t.to(device)
            self.broadcast_(t, src)
            out.append(t.cpu().numpy())
        return out

    def gather_models(self, local_models, n_models, like=None):
        """Every rank ends up with all `n_models` weight dicts: model k is broadcast by its owner
        (k mod G), where it is local_models[k // G]."""
        if not self.distributed:
            return list(local_models)
        import torch
        mine = self.my_models(n_models)
        assert len(local_models) == len(mine)
        template = local_models[0] if local_models else like
        out = []
        for k in range(int(n_models)):
            src = self.owner_of_model(k)
            m = {}
            for key, ref in template.items():
                t = local_models[mine.index(k)][key].clone() if src == self.rank else torch.empty_like(ref)
                m[key] = self.broadcast_(t.contiguous(), src)
            out.append(m)
        return out
