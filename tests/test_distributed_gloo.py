"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding, per-rank RNG stream ids, the single all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_model_nemo_b200 import distributed as D


def test_shard_arithmetic():
    assert D.shard_sizes(256, 8) == [32] * 8
    assert D.shard_sizes(10, 4) == [3, 3, 2, 2]
    assert D.shard_sizes(2, 4) == [1, 1, 0, 0]
    assert D.shard_sizes(0, 2) == [0, 0]
    cover = []
    for r in range(4):
        lo, hi = D.shard_range(10, r, 4)
        cover += list(range(lo, hi))
    assert cover == list(range(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, ws)
    from diffusion_model_nemo_b200.modules import _runtime as R
    assert R.rank_stream_id() == rank                      # independent Philox stream per rank
    os.environ.pop("RANK")                                 # launchers that do not export RANK (mp.spawn, ddp_spawn, tcp:// init):
    assert R.rank_stream_id() == rank                      # the process-group rank still separates the streams
    os.environ["RANK"] = str(rank)
    lo, hi = D.shard_range(total, rank, ws)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()
    full = D.all_gather_samples(local, total=total)
    q.put((rank, full[:, 0, 0, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_all_gather_of_final_samples_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r] == [float(i) for i in range(total)]      # every rank holds the whole batch, in order
