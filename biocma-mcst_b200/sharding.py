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


def setup_peer_allreduce(loop, world, rank, device=None, group=None, log=None):
    """Collective set-up of the peer-memory all-reduce (bmc_p2p_export / bmc_p2p_attach) over torch.distributed: every
    rank exports its 64-byte IPC handle, the handles are all-gathered, every rank attaches.  The decision is collective:
    if ANY rank fails to export or attach, every rank calls p2p_disable() and stays on the NCCL communicator.
    Returns True when the peer path is active on all ranks.  `device`: where the exchanged tensors live (cuda for NCCL
    groups, None/cpu for gloo)."""
    import torch
    import torch.distributed as dist
    ok = 1
    try:
        mine = torch.from_numpy(np.ascontiguousarray(loop.p2p_export(), np.uint8).copy())
    except RuntimeError as e:
        if log:
            log(f"rank {rank}: peer export failed ({e})")
        ok, mine = 0, torch.zeros(64, dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    if ok:
        try:
            loop.p2p_attach(world, rank, torch.stack(gathered).cpu().numpy())
        except RuntimeError as e:
            if log:
                log(f"rank {rank}: peer attach failed ({e})")
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32)
    if device is not None:
        flag = flag.to(device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    active = int(flag.item()) == 1
    if not active:
        loop.p2p_disable()
    dist.barrier(group=group)
    return active


def global_init_weight(local_total_mass, x0, total_volume, device=None, group=None):
    """post_init_weight on a sharded population: every rank initialises its shard, the total masses are summed
    (MPI all-reduce in global_initaliser.cpp:311) and every rank uses w = X0 * V_tot / m_tot (mc/src/unit.cpp:232-257)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_total_mass)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return x0 * float(total_volume) / float(t.item())


def global_repartition(local_repartition, device=None, group=None):
    """records/number_particle of the whole job: the per-rank getRepartition() vectors summed"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_repartition, np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().astype(np.uint64)
