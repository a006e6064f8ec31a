"""Batch-sharded sampling across the GPUs of one box (one process per GPU, torch.distributed over NCCL/NVLink).

The sampling path has no data dependence between samples (the only batch coupling in the reference, the Langevin
corrector's batch-mean norms, is kept per rank: each rank is one independent reference run with its own batch --
SURVEY.md section 8e), so there is NO collective inside the loop: every rank runs its shard with its own Philox
stream (stream_id = rank) and ONE all-gather of the final samples closes the job.
"""
import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process == 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str = None):
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)      # binds the communicator to this rank's GPU (no device guessing)
        dist.init_process_group(backend=backend, rank=rank, world_size=ws, **kw)
    return rank, ws, local


def shard_sizes(total: int, world_size: int) -> List[int]:
    """Contiguous shards, sizes differ by at most one (ragged totals allowed, empty shards when total < world)."""
    base, rem = divmod(total, world_size)
    return [base + (1 if r < rem else 0) for r in range(world_size)]


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    sizes = shard_sizes(total, world_size)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def all_gather_samples(local: torch.Tensor, total: int = None) -> torch.Tensor:
    """The single collective of the job: concatenate every rank's final samples along the batch axis.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    ws = dist.get_world_size()
    sizes = shard_sizes(total, ws) if total is not None else [local.shape[0]] * ws
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((ws * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    chunks = [out[r * mx:r * mx + sizes[r]] for r in range(ws)]
    return torch.cat(chunks)


def sample_sharded(sampler, model, shape, device, **kw):
    """Run `sampler.sample` on this rank's shard of the batch and all-gather the final [0,1] images (device tensor)."""
    rank, ws, _ = world()
    lo, hi = shard_range(shape[0], rank, ws)
    local_shape = [hi - lo] + list(shape[1:])
    if hi > lo:
        imgs = sampler.sample(model, local_shape, device=device, **kw)
        local = imgs[-1].to(device)
    else:
        local = torch.empty([0] + list(shape[1:]), device=device)
    return all_gather_samples(local, total=shape[0])
