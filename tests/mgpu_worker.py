"""Worker of tests/test_multigpu_gpu.py (one process per GPU, launched with torch.distributed.run): a sharded, fully
device-resident coupled loop — liquid step (which finishes the all-reduce of the previous cycle's sources) -> cycle
(whose first block would finish it otherwise, and whose commit block publishes the new sources) -> bmc_allreduce_sources
(lazy) — over the REAL peer-memory path (cudaIpc mappings between processes), compared with the same loop over NCCL."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from _bmc_loader import load_pkg, load_synth
import util

pkg, synth = load_pkg(), load_synth()
from biocma_mcst_b200 import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_comp, n_total, dt, steps = 64, 120_000, 5.0, 12
case = util.make_case(synth, "monod", n_total, n_comp, dt=dt, near_division=0.6, p_move=0.2, p_exit=0.0, outlet=False)
counts, offs = sharding.shard_offsets(n_total, world)
sl = slice(int(offs[rank]), int(offs[rank + 1]))
fm = case["fm"]
q = 0.02 * fm["volumes"][n_comp - 1] / dt
feeds = [dict(species=0, input_position=0, flow=q, concentration=8.0, output_position=n_comp - 1)]


def run(use_p2p):
    g = pkg.ParticleLoop("monod", 1, n_comp, device=local, seed=11, rank=rank)
    g.set_particles(case["props"][:, sl], case["pos"][sl]); g.set_weight(case["weight"] * 2e3)
    g.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    g.liquid_set_transition(fm["coo"]); g.set_concentrations(np.full(n_comp, 2.0)); g.liquid_set_feeds(feeds)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(g.nccl_unique_id().copy())
    uid = uid.to(dev); dist.broadcast(uid, 0)
    g.comm_init(world, rank, uid.cpu().numpy())
    if use_p2p:
        assert sharding.setup_peer_allreduce(g, world, rank, device=dev), "peer attach failed"
    traj = []
    for step in range(steps):
        g.liquid_step(dt)          # consumes (and, on the peer path, finishes the all-reduce of) the last cycle's sources
        g.cycle(dt)
        g.allreduce_sources()
        if step % 4 == 3:
            traj.append(g.get_concentrations())
    last = g.get_sources()         # the all-reduced sources of the last step (finished by the small kernel)
    c = g.counters()
    g.close()
    return np.array(traj), last, c


tp, sp, cp = run(True)
tn, sn, cn = run(False)
ok = True
msg = []
# every rank holds bitwise the same concentrations and sources on the peer path (rank-ordered sums)
for arr, name in ((tp, "traj"), (sp, "sources")):
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        ok = False; msg.append(f"{name} differs between ranks")
# ... and they agree with the NCCL run (other summation order) to rounding
e_traj = float(np.max(np.abs(tp - tn) / np.abs(tn)))
e_src = float(np.max(np.abs(sp - sn)) / np.max(np.abs(sn)))
if e_traj > 1e-12 or e_src > 1e-12:
    ok = False; msg.append(f"peer vs NCCL: traj {e_traj:.2e} sources {e_src:.2e}")
if cp["n_used"] != cn["n_used"] or cp["total_new"] != cn["total_new"]:
    ok = False; msg.append("particle counters differ between the two runs")
moved = float(np.max(np.abs(tp[-1] - 2.0)))
if not (moved > 1e-3 and cp["total_new"] > 0):
    ok = False; msg.append("nothing happened")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ok": bool(flag.item()), "world": world, "max_rel_traj": e_traj, "max_rel_sources": e_src, "msg": msg, "conc_change": moved}))
dist.barrier(); dist.destroy_process_group()
