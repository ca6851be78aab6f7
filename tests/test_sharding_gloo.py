"""N>1 host logic on CPU: two gloo ranks, particles sharded with the reference's
balancing rule, replicated compartment state, one all-reduce of the source vector.
The per-rank compute is the oracle (tests may use it); the GPU path replaces it by
ParticleLoop + bmc_allreduce_sources (NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_counts_follow_load_balancer(bmc):
    from biocma_mcst_b200 import sharding
    for n, w in ((10_000_000, 10), (1_000_003, 8), (7, 2), (5_000_000, 3)):
        counts, offs = sharding.shard_offsets(n, w)
        assert sum(counts) == n and offs[-1] == n                       # test_load_balancing.cpp:19-20
        base = int(float(n) * (1.0 / w))
        assert counts[1:] == [base] * (w - 1) and counts[0] == base + (n - base * w)  # remainder on rank 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _bmc_loader import load_pkg, load_synth
    load_pkg(); synth = load_synth()
    from biocma_mcst_b200 import sharding
    import oracle, util
    case = util.make_case(synth, "monod", 60_001, 100, dt=5.0, near_division=0.5)
    counts, offs = sharding.shard_offsets(case["n"], world)
    lo, hi = offs[rank], offs[rank + 1]
    o = oracle.OracleLoop("monod", 1, 100, rank=rank)
    o.set_particles(case["props"][:, lo:hi], case["pos"][lo:hi])
    o.set_weight(case["weight"])
    fm = case["fm"]
    o.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    o.set_leaving_flows(case["flows"]); o.set_concentrations(case["conc"])
    o.cycle(case["dt"])
    total = sharding.allreduce_sources_torch(o.get_sources())
    n_tot = np.array([o.counters()["n_used"]], np.float64)
    n_tot = sharding.allreduce_sources_torch(n_tot)
    # job-wide weight and repartition (global_initaliser.cpp:311)
    lin_density = np.float32(1000.0) * np.float32(np.pi) * np.float32(0.6e-6) * np.float32(0.6e-6) / np.float32(4.0)
    local_mass = float(np.sum(case["props"][0, lo:hi].astype(np.float64) * float(lin_density)))
    w = sharding.global_init_weight(local_mass, 0.5, float(np.sum(fm["volumes"])))
    rep = sharding.global_repartition(o.repartition())
    if rank == 0:
        q.put((total, float(n_tot[0]), counts, w, rep))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_source_allreduce_matches_single_rank(orc, synth):
    import util
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, n_tot, counts, w, rep = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference: contributions are taken at the pre-move position of this step, so the
    # sharded sum must equal the unsharded one up to fp64 summation order
    case = util.make_case(synth, "monod", 60_001, 100, dt=5.0, near_division=0.5)
    o = orc.OracleLoop("monod", 1, 100)
    util.load_case(o, case)
    o.cycle(case["dt"])
    ref = o.get_sources()
    assert counts == [30_001, 30_000]
    assert np.allclose(total, ref, rtol=1e-12, atol=0)
    assert n_tot == o.counters()["n_used"]  # same number of divisions in total
    # the job-wide weight is the single-rank one (post_init_weight on the summed mass), and so is the repartition: the
    # set of particles per compartment does not depend on the sharding (moves draw from (seed, rank, slot, step), so the
    # two runs are different realisations — only the totals are comparable after a step)
    assert abs(w - case["weight"]) <= 1e-12 * case["weight"]
    assert int(rep.sum()) == o.counters()["n_used"] - o.counters()["n_inactive"] and rep.shape == (100,)


# ---- collective set-up of the peer-memory all-reduce (sharding.setup_peer_allreduce): handle exchange and the
# all-or-nothing fallback, with a stand-in for the context (the real one needs a GPU: tests/test_p2p_allreduce_gpu.py)
class _FakeLoop:
    def __init__(self, rank, fail_attach):
        self.rank, self.fail_attach, self.attached, self.disabled = rank, fail_attach, None, False

    def p2p_export(self):
        return np.full(64, 17 + self.rank, np.uint8)

    def p2p_attach(self, world, rank, handles):
        if self.fail_attach:
            raise RuntimeError("cudaIpcOpenMemHandle: simulated failure")
        self.attached = np.asarray(handles).reshape(world, 64).copy()

    def p2p_disable(self):
        self.disabled = True


def _peer_worker(rank, world, port, fail_rank, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _bmc_loader import load_pkg
    load_pkg()
    from biocma_mcst_b200 import sharding
    loop = _FakeLoop(rank, fail_attach=(rank == fail_rank))
    active = sharding.setup_peer_allreduce(loop, world, rank)
    rows = None if loop.attached is None else loop.attached[:, 0].tolist()
    q.put((rank, active, loop.disabled, rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_peer_allreduce_setup_is_all_or_nothing(fail_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() + 7 * (fail_rank + 2)) % 2000
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, active, disabled, rows in got:
        if fail_rank < 0:   # every rank sees every handle, in rank order, and keeps the peer path
            assert active and not disabled and rows == [17, 18]
        else:               # one rank could not attach: nobody uses the peer path
            assert not active and disabled
