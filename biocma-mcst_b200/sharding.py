"""Particle sharding across GPUs (one process per GPU).

The reference shards particles across MPI ranks with a load balancer
(apps/core/src/load_balancing/iload_balancer.cpp:7-49: every rank gets
floor(n * ratio), rank 0 additionally takes the remainder) and replicates the
compartment state; the only per-step exchange is the sum of the source terms
(apps/core/src/sync.cpp:57-78).  Same here: each rank owns an independent
ParticleLoop (own container, division buffer, compaction, Philox `rank` word)
and one all-reduce of n_species x n_compartments doubles per step.
"""
import numpy as np


def uniform_ratio(world_size):
    return 1.0 / world_size


def shard_count(n_total, rank, world_size):
    """UniformLoadBalancer + ILoadBalancer::balance."""
    base = int(float(n_total) * uniform_ratio(world_size))
    if rank != 0:
        return base
    return base + (n_total - base * world_size)


def shard_offsets(n_total, world_size):
    counts = [shard_count(n_total, r, world_size) for r in range(world_size)]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return counts, offs


def allreduce_sources_torch(sources, group=None):
    """Sum the per-rank source vectors through torch.distributed (gloo on CPU for
    the host-logic tests; the GPU path uses bmc_allreduce_sources / NCCL on the
    context stream)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(sources, np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
