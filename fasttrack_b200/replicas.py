"""Multi-GPU story of the front-end: replicas only.

Tracking is frame-sequential, so N GPUs run N independent sequences, one process + one context
per GPU, with no collective on the data path. torch.distributed is used for exactly two things: the barrier that
brackets the timed region and the MAX reduction of the per-rank elapsed time. Works with any backend (nccl on
GPUs, gloo in the CPU tests).
"""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def sequence_seed(rank, base=5):
    """one independent synthetic sequence per GPU, seed = 5 + gpu_id"""
    return base + rank


def barrier(dist, world):
    if world > 1:
        dist.barrier()


def reduce_max(dist, world, values, device="cpu"):
    """element-wise max over ranks of a list of floats"""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(world, steps, max_elapsed_s):
    """whole-job frames/s: every rank processed `steps` frames of its own sequence in (at most) max_elapsed_s"""
    return world * steps / max_elapsed_s
