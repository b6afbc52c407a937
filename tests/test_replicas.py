"""N>1 host logic (replicas only) on CPU with the gloo backend, world_size 2."""
import os
import socket

import pytest


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from fasttrack_b200 import replicas
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, lr, w = replicas.env_rank()
    assert (r, lr, w) == (rank, rank, world)
    seed = replicas.sequence_seed(rank)
    replicas.barrier(dist, world)
    elapsed = 0.010 * (rank + 1)            # rank 1 is the slower replica
    mx, = replicas.reduce_max(dist, world, [elapsed])
    value = replicas.aggregate_throughput(world, 100, mx)
    q.put((rank, seed, mx, value))
    dist.destroy_process_group()


def test_two_replicas_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == [5, 6]                      # independent sequences: seed = 5 + rank
    assert all(abs(o[2] - 0.020) < 1e-12 for o in out)        # max over ranks
    assert all(abs(o[3] - 2 * 100 / 0.020) < 1e-6 for o in out)  # whole-job throughput, identical on every rank


def test_single_process_is_identity():
    from fasttrack_b200 import replicas
    assert replicas.reduce_max(None, 1, [1.5, 2.5]) == [1.5, 2.5]
    assert replicas.aggregate_throughput(1, 10, 0.5) == 20.0
