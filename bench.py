#!/usr/bin/env python
"""bench.py — particle-steps/s of the Monte-Carlo particle loop on N B200s.

A "step" is one cycleProcess (model update + division, contribution scatter,
compartment move, outlet exit, compaction, spawn) over this rank's particles.

Headline workload at every N ("ns", the north-star case of BASELINE.json): the stirred-tank
CMA of configs[1] — 500 compartments, Monod uptake model, one outlet, dt = 0.1 s, fixed
synthetic flow map — at 1.25e8 particles PER GPU (weak scaling), so that --gpus 8 IS the
1e9-particle target case and the 1 -> 8 curve compares like with like.  At N = 1 the line also
carries, under "configs", one sub-record per BASELINE configuration measured the same way in
the same run: c2 (configs[1] at its own 1e7 particles), c2_eager (the same with ages updated
eagerly, i.e. the 57 B/particle layout of SURVEY.md §8d), c3 (1e8 particles, division + outlet
+ compaction inside the timed region), c4 (10 000 compartments, 1.25e8 particles = the 8-GPU
shard of configs[3]), c5 (32-property UDF model, 2.5e8 particles = the shard of configs[4]), and fl_ns / sa_ns
(fixed_length and simple_acetate, the two models of the reference's default build, at the headline's
1.25e8 particles).

  value     : whole-job particle-steps/s, state resident in HBM, no host sync
              inside the timed region, one all-reduce of the source vector
              per step when N > 1.
  e2e       : same metric through the host-buffer C ABI a reference caller would
              use each step: concentrations H2D -> bmc_cycle -> sources D2H.
  roofline  : dominant kernel (fused cycle) algorithmic bytes / CUDA-event time.
  cpu_baseline / --impl reference : the REFERENCE'S OWN kernels on the host cores, on a bounded
              sample: oracle/_ref/libbmc_ref_release.so = the reference's hot-path sources
              compiled with its release flags over oracle/kokkos_shim, leagues and ranges spread
              over all host threads like the Kokkos OpenMP backend (kind "reference").  Where
              that library is absent, or for models the reference does not have, the oracle's
              OpenMP restatement (kind "port").  See DESIGN.md §2/§6.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic bytes per particle-step of THIS design: 4*(R + W) property bytes + position
# read+write (8) + status (1).  Ages cost nothing per step (step stamps), which is 16 B less
# than the figure SURVEY.md §8d derives for an eagerly updated layout (kept below as B_SURVEY; it IS what the
# eager-age kernel moves: 2 x 4 B read + 2 x 4 B written per particle on top).
B_ALG = {"monod": 41, "fixed_length": 21, "simple_acetate": 49, "wide_udf": 8 * 32 + 9}      # multi-compartment
B_ALG_0D = {"monod": 37, "fixed_length": 17, "simple_acetate": 45, "wide_udf": 8 * 32 + 5}  # 0D: position only read
B_SURVEY = {"monod": 57, "fixed_length": 37, "simple_acetate": 65, "wide_udf": 8 * 32 + 25}
N_SPECIES = {"monod": 1, "fixed_length": 1, "simple_acetate": 2, "wide_udf": 4}

WORKLOADS = {
    # name: (model, n_comp, particles per GPU, dt, near_division, p_exit)
    # north-star target: 1e9 particles, 500 compartments, monod on 8 GPUs -> 1.25e8 per GPU
    "ns": ("monod", 500, 125_000_000, 0.1, 0.0, 1e-3),
    "c2": ("monod", 500, 10_000_000, 0.1, 0.0, 1e-3),
    "c3": ("monod", 500, 100_000_000, 1.0, 0.5, 1e-3),
    "c1": ("monod", 1, 100_000, 0.1, 0.0, 1e-3),
    # configs[3]: 10k compartments, 1e9 particles over 8 GPUs -> 1.25e8 per GPU
    "c4": ("monod", 10_000, 125_000_000, 0.1, 0.0, 1e-3),
    # configs[4]: wide UDF (32 properties), 2e9 particles over 8 GPUs -> 2.5e8 per GPU
    "c5": ("wide_udf", 500, 250_000_000, 0.1, 0.0, 1e-3),
    # other built-in models on the c2 shape
    "fl": ("fixed_length", 500, 10_000_000, 0.1, 0.0, 1e-3),
    "sa": ("simple_acetate", 500, 10_000_000, 0.1, 0.0, 1e-3),
    # the two models of the reference's default build (meson_options.txt: model_list_name) at the north-star population
    "fl_ns": ("fixed_length", 500, 125_000_000, 0.1, 0.0, 1e-3),
    "sa_ns": ("simple_acetate", 500, 125_000_000, 0.1, 0.0, 1e-3),
}
# measured after the headline at N = 1 (each on a fresh context)
SUB_RECORDS = ("c2", "c2_eager", "c3", "c4", "c5", "fl_ns", "sa_ns")


def workload_label(wl, n_per_gpu=None):
    model, n_comp, n_full, dt, _, _ = WORKLOADS[wl]
    n = n_per_gpu or n_full
    return f"{wl}: {n_comp}-compartment stirred-tank CMA, {model}, {n} particles/GPU, dt={dt}"


def workload_config(wl, n_per_gpu, n_var):
    """the `config` object, identical in both arms (it describes the workload, not the implementation)"""
    return {"workload": workload_label(wl, n_per_gpu),
            "l2_policy": f"inputs larger than L2 ({n_per_gpu * (n_var * 4 + 13) / 1e6:.0f} MB of particle state per GPU)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

        def collect(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons
        # samples are printed ~one period after they were taken: accept [begin, end + 0.15 s]
        inside = [x for x in self.lines if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or 1e30) + 0.15]
        sm, mx, reasons = collect(inside)
        scope = "timed region"
        if not sm:
            sm, mx, reasons = collect(self.lines)
            scope = "whole run (timed region shorter than the sampling period)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def host_threads():
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def build_case(synth, model, n_comp, dt, p_exit):
    fm = synth.make_flowmap(n_comp, dt, p_move=0.01, seed=2024)
    flows = []
    if p_exit > 0:
        o = n_comp - 1
        flows = [(o, p_exit * fm["volumes"][o] / dt, fm["volumes"][o])]
    ns = N_SPECIES[model]
    conc = np.full(ns * n_comp, 3.0) * (0.8 + 0.4 * np.random.default_rng(5).random(ns * n_comp))
    return fm, flows, conc


def setup_loop(loop, fm, flows, conc, n_comp):
    if n_comp > 1:
        loop.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    else:
        loop.domain_update(fm["volumes"], None, fm["out_flows"], None)
    loop.set_leaving_flows(flows)
    loop.set_concentrations(conc)


def make_cpu_loop(model, n_species, n_comp, threads):
    """(loop, kind, label): the reference's own kernels when oracle/_ref holds the timing build, else the port"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if os.environ.get("BMC_CPU_BASELINE", "reference") == "reference":
        try:
            import ref
            if model in ref.RELEASE_MODELS and (os.path.exists(ref.RELEASE_LIB_PATH) or ref.can_build()):
                return (ref.RefLoop(model, n_species, n_comp, release=True, n_threads=threads), "reference",
                        "reference kernels (oracle/_ref, release flags, Kokkos shim with OpenMP leagues)")
        except Exception as e:  # noqa: BLE001 - fall back to the port, say why
            print(f"[bench] reference build unavailable ({e}); timing the oracle port", file=sys.stderr)
    import oracle
    return oracle.OracleLoop(model, n_species, n_comp, n_threads=threads), "port", "oracle OpenMP restatement"


N_VAR = {"monod": 6, "fixed_length": 2, "simple_acetate": 9, "wide_udf": 32}
# c3 exercises removal: exits at 5 % per step in the outlet compartment and a compaction threshold low enough that
# several compactions fall inside the timed region
LOOP_KW = {"c3": dict(dead_ratio=5e-4)}
P_EXIT = {"c3": 0.05}
# c3 grows by ~1 % per step: like a production run that knows its horizon, the arrays are reserved up front (bmc_reserve)
# so that no reallocation (a one-off of tens of ms) lands inside a 30-step timed region
RESERVE = {"c3": 2.4}
PROFILE_EVERY = 4


def run_reference(args, wl):
    """Reference arm: the reference's own kernels on the host cores (oracle/_ref timing build, else the oracle port) on
    the same config; each step is a bounded sample of the workload (--cpu-sample particles)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # torchrun exports OMP_NUM_THREADS=1; libgomp reads it when it is loaded (nothing has loaded it yet in this
    # process: torch is only imported by the GPU arm) and runs oversized teams badly when it says 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    # (no OMP_PROC_BIND/OMP_PLACES: binding as tools/exec.py:47-49 does measured 20 % SLOWER on the GPU boxes' cgroups)
    from _bmc_loader import load_synth
    synth = load_synth()
    model, n_comp, n_full, dt, near, p_exit = WORKLOADS[wl]
    p_exit = P_EXIT.get(wl, p_exit)
    if args.particles:
        n_full = args.particles
    n = min(n_full, args.cpu_sample)
    fm, flows, conc = build_case(synth, model, n_comp, dt, p_exit)

    def make(n):
        props, pos = synth.make_population(model, n, n_comp, seed=11, near_division=near)
        o, kind, label = make_cpu_loop(model, N_SPECIES[model], n_comp, threads)
        o.set_particles(props, pos)
        o.set_weight(synth.initial_weight(props, 0.5, float(fm["volumes"].sum())))
        setup_loop(o, fm, flows, conc, n_comp)
        return o, kind, label
    o, kind, label = make(n)
    # bound the whole --steps/--warmup run to a few minutes: calibrate on two steps, shrink the sample if needed
    o.cycle(dt)
    t0 = time.perf_counter(); o.cycle(dt); t_step = time.perf_counter() - t0
    budget = float(os.environ.get("BMC_REF_BUDGET_S", "120"))
    if t_step * (args.steps + args.warmup) > budget:
        n = max(65_536, int(n * budget / (t_step * (args.steps + args.warmup))))
        o.close()
        o, kind, label = make(n)
    for _ in range(args.warmup):
        o.cycle(dt)
    live = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c = o.counters()
        live += c["n_used"]
        o.cycle(dt)
    el = time.perf_counter() - t0
    v = live / el
    sample = f"{n} of {n_full} particles/GPU x {args.steps} steps, {threads} OpenMP threads, {label}"
    line = {"metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(wl, n_full, N_VAR[model]),
            "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def traffic_entry(wl, n_per_gpu, kcfg, eager):
    """measured DRAM bytes per launch of the step kernel for this workload, from the committed ncu capture
    (profiles/traffic.json: one entry per workload with the kernel instantiation and the commit it was taken on);
    None when the entry does not describe the kernel that just ran"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(wl + ("_eager" if eager else ""))
        if t and t["particles"] == n_per_gpu and t.get("block") == (kcfg["block_eager"] if eager else kcfg["block"]) and t.get("vec") == kcfg["vec"]:
            return t["bytes_per_launch"], f"profiles/traffic.json ({t.get('kernel')}, commit {t.get('commit')}, {t.get('capture')})"
    except Exception:
        pass
    return None, None


class Dist:
    """the torch.distributed plumbing of one bench process"""
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the particle loop has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.device = torch.device("cuda", self.local)

    def barrier(self, loop=None):
        if loop is not None:
            loop.sync()
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()


def measure(D, pkg, synth, wl, n_per_gpu, steps, warmup, e2e_steps, eager=False, sampler=None):
    """one workload on this job's GPUs: device-resident leg, end-to-end leg, roofline of the step kernel.
    Returns the record on rank 0 (None elsewhere); the context is destroyed before returning."""
    torch, dist, world, rank, local = D.torch, D.dist, D.world, D.rank, D.local
    model, n_comp, _, dt, near, p_exit = WORKLOADS[wl]
    p_exit = P_EXIT.get(wl, p_exit)
    fm, flows, conc = build_case(synth, model, n_comp, dt, p_exit)
    if eager:
        os.environ["BMC_EAGER_AGES"] = "1"
    try:
        loop = pkg.ParticleLoop(model, N_SPECIES[model], n_comp, device=local, seed=2024, rank=rank, **LOOP_KW.get(wl, {}))
        # population: device-side mc_init_first (monod draws its TruncatedNormal lengths on the GPU)
        linit = None
        if model in ("fixed_length", "wide_udf"):  # configurable models take their lengths from a Config view
            linit = (1e-6 + 1e-6 * np.random.default_rng(3 + rank).random(n_per_gpu)).astype(np.float32)
        total_mass = loop.init_particles(n_per_gpu, uniform_position=True, linit=linit)
        del linit
        if near > 0:  # c3: bring cells close to division through the host path once
            props, pos = synth.make_population(model, n_per_gpu, n_comp, seed=11 + rank, near_division=near)
            loop.set_particles(props, pos)
            total_mass = float(np.sum(props[0].astype(np.float64))) * 2.8274e-10
            del props, pos
    finally:
        os.environ.pop("BMC_EAGER_AGES", None)
    if wl in RESERVE:
        loop.reserve(int(RESERVE[wl] * n_per_gpu))
    v_tot = float(fm["volumes"].sum())
    if dist is not None:
        # post_init_weight on a sharded population: the masses are all-reduced, every rank uses the same weight
        # (global_initaliser.cpp:311, mc/src/unit.cpp:232-257)
        from biocma_mcst_b200 import sharding
        loop.set_weight(sharding.global_init_weight(total_mass, 0.5, v_tot, device=D.device))
    else:
        loop.set_weight(0.5 * v_tot / total_mass)
    setup_loop(loop, fm, flows, conc, n_comp)
    collective, check = "", None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(loop.nccl_unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        loop.comm_init(world, rank, uid.cpu().numpy())
        # one-shot all-reduce over NVLink peer mappings (bmc_p2p_*): the 64-byte IPC handles are gathered here; if any
        # rank cannot attach, every rank stays on the NCCL communicator.  BMC_P2P=0 forces NCCL, BMC_P2P=1 forces the
        # peer-memory path; by default it is used from BMC_P2P_MIN_RANKS ranks on (see DESIGN.md §7 for the measurements).
        collective = "1 NCCL all-reduce/step"
        want_p2p = os.environ.get("BMC_P2P", "auto")
        use_p2p = want_p2p == "1" or (want_p2p == "auto" and world >= int(os.environ.get("BMC_P2P_MIN_RANKS", "2")))
        if use_p2p and world <= 16:
            from biocma_mcst_b200 import sharding
            if sharding.setup_peer_allreduce(loop, world, rank, device=D.device,
                                             log=lambda m: print(f"[bench] {m}; NCCL", file=sys.stderr)):
                collective = "1 peer-memory all-reduce/step (one-shot over NVLink, NCCL only for set-up)"
    stream = torch.cuda.ExternalStream(loop.stream_handle(), device=D.device)

    def step_resident():
        loop.cycle(dt)
        if world > 1:
            loop.allreduce_sources()

    for _ in range(warmup):
        step_resident()
    D.barrier(loop)
    if world > 1:
        # one-time check on the real exchange path: the all-reduced vector equals the NCCL sum of the per-rank vectors
        loop.cycle(dt)
        mine = torch.from_numpy(loop.get_sources().copy()).to(D.device)
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        loop.allreduce_sources()
        got = loop.get_sources()
        ref_sum = mine.cpu().numpy()
        err = float(np.max(np.abs(got - ref_sum)) / (np.max(np.abs(ref_sum)) + 1e-300))
        check = {"allreduce_matches_nccl_sum": bool(err <= 1e-12), "max_rel_err": err}
        if err > 1e-12:
            raise SystemExit(f"[bench] rank {rank}: all-reduced sources differ from the NCCL sum of the per-rank vectors ({err:.3e})")
        D.barrier(loop)
    # ---------------- value: device-resident steps -----------------------------
    c0 = loop.counters()
    launches0 = loop.launch_count()
    # CUDA events around every PROFILE_EVERY-th step kernel of the timed region (around every one they cost ~7 % of a
    # 90 us step: two event records per launch on the launching stream)
    loop.profile_enable(PROFILE_EVERY if steps >= 4 * PROFILE_EVERY else 1)  # at least 4 samples
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier(loop)
    if sampler is not None:
        sampler.mark_begin()
    e0.record(stream)
    for _ in range(steps):
        step_resident()
    e1.record(stream)
    D.barrier(loop)
    if sampler is not None:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_n = loop.profile_read()
    loop.profile_enable(0)
    launches = loop.launch_count() - launches0
    c1 = loop.counters()
    live_avg = 0.5 * (c0["n_used"] + c1["n_used"])

    # ---------------- e2e: host-buffer ABI every step ---------------------------
    conc_host = np.ascontiguousarray(conc, np.float64)
    for _ in range(3):
        loop.set_concentrations(conc_host); step_resident(); loop.get_sources()
    D.barrier(loop)
    n_e0 = loop.counters()["n_used"]
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-step inputs prepared outside the timed region (the liquid solver of the caller owns them);
    # what is timed is the ABI: host buffer in -> step -> host buffer out, every step
    conc_steps = [np.ascontiguousarray(conc * (1.0 + 0.01 * np.sin(0.1 * s)), np.float64) for s in range(16)]
    src = np.empty(conc_host.size, np.float64)
    f0.record(stream)
    for s in range(e2e_steps):
        loop.set_concentrations(conc_steps[s & 15])   # H2D of this step's inputs
        step_resident()
        loop.get_sources(src)                        # D2H of this step's result (synchronises)
    f1.record(stream)
    D.barrier(loop)
    ms_e2e = f0.elapsed_time(f1)
    c2 = loop.counters()
    live_e2e = 0.5 * (n_e0 + c2["n_used"])
    kcfg = loop.kernel_config()
    n_var = loop.n_var

    # ---------------- reduce over ranks (max time, sum particles) ---------------
    stats = torch.tensor([ms, ms_e2e, live_avg, live_e2e, kernel_ms / max(1, kernel_n), float(launches)], dtype=torch.float64,
                         device="cuda")
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e, k_ms = mx[0].item(), mx[1].item(), mx[4].item()
        live_tot, live_e2e_tot, launches_tot = sm[2].item(), sm[3].item(), sm[5].item()
    else:
        k_ms = stats[4].item(); live_tot, live_e2e_tot, launches_tot = live_avg, live_e2e, float(launches)
    loop.close()
    del loop
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    b_tab = B_SURVEY if eager else (B_ALG if n_comp > 1 else B_ALG_0D)
    b_alg = b_tab[model]
    achieved = (live_avg * b_alg) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    nb = N_SPECIES[model] * n_comp * 8
    traffic, traffic_src = traffic_entry(wl, n_per_gpu, kcfg, eager)
    block = kcfg["block_eager"] if eager else kcfg["block"]
    rec = {
        "value": live_tot * steps / (ms * 1e-3), "unit": "particle-steps/s", "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "config": workload_config(wl, n_per_gpu, n_var),
        # particles actually stepped (mean over the timed region, all ranks): what `value` and `roofline` are computed from.
        # Differs from the configured population where the model's own initial state divides at once (simple_acetate
        # draws length and l_max from the same law, simple_acetate.hpp:132-152: half the cells divide in the first step)
        "live_particles": live_tot,
        "parallelism": f"particle-sharded x{world}, replicated liquid state, {collective}" if world > 1 else "single GPU",
        "e2e": {"value": live_e2e_tot * e2e_steps / (ms_e2e * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": nb,
                "d2h_bytes_per_step": nb, "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": int(launches_tot),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": f"cycle_kernel<{model}, VEC={kcfg['vec']}, {block} threads, {'eager' if eager else 'stamped'} ages> "
                               "(whole step: particle pass + post-cycle phase)",
                     "kernel_ms": k_ms, "kernel_ms_samples": int(kernel_n), "bytes_per_particle": b_alg,
                     "bytes_per_particle_note": ("ages loaded and stored every step (the layout SURVEY.md 8d counts)" if eager else
                                                 "4*(R+W) property bytes + position 8 + status 1; ages are step stamps, not rewritten"),
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "kernel_share_of_step": k_ms * steps / ms},
        "events_in_timed_region": {"divisions": c1["total_new"] - c0["total_new"], "exits": c1["total_out"] - c0["total_out"],
                                   "compactions": c1["n_compactions"] - c0["n_compactions"], "reallocations": c1["n_reallocations"] - c0["n_reallocations"]},
    }
    if check is not None:
        rec["collective_check"] = check
    return rec


def cpu_baseline(synth, wl, n_per_gpu, cpu_sample):
    model, n_comp, _, dt, near, p_exit = WORKLOADS[wl]
    fm, flows, conc = build_case(synth, model, n_comp, dt, P_EXIT.get(wl, p_exit))
    threads = host_threads()
    n = min(n_per_gpu, cpu_sample)
    props, pos = synth.make_population(model, n, n_comp, seed=11, near_division=near)
    o, kind, label = make_cpu_loop(model, N_SPECIES[model], n_comp, threads)
    o.set_particles(props, pos)
    o.set_weight(1.0)
    setup_loop(o, fm, flows, conc, n_comp)
    o.cycle(dt)
    steps_cpu, t0, done = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < 10.0 and steps_cpu < 400:
        done += o.counters()["n_used"]; o.cycle(dt); steps_cpu += 1
    el = time.perf_counter() - t0
    o.close()
    return {"value": done / el, "unit": "particle-steps/s", "cores": threads, "kind": kind,
            "sample": f"{n} of {n_per_gpu} particles x {steps_cpu} steps, {label}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--eager", action="store_true", help="eager float ages (BMC_EAGER_AGES=1): the 57 B/particle kernel")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="particles in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration sub-records (N = 1)")
    ap.add_argument("--configs-steps", type=int, default=30, help="timed steps of each sub-record")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (default: --steps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = args.workload
    if args.impl == "reference":
        return run_reference(args, wl)

    from _bmc_loader import load_pkg, load_synth
    pkg, synth = load_pkg(), load_synth()
    D = Dist()
    n_per_gpu = args.particles or WORKLOADS[wl][2]
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
    rec = measure(D, pkg, synth, wl, n_per_gpu, args.steps, args.warmup, args.e2e_steps or args.steps, eager=args.eager, sampler=sampler)
    clocks = sampler.stop() if D.rank == 0 else None
    if D.rank == 0:
        line = {"metric": "particle-steps/sec", "value": rec["value"], "unit": rec["unit"], "n_gpus": D.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": rec["config"], "parallelism": rec["parallelism"],
                "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "live_particles": rec["live_particles"], "roofline": rec["roofline"],
                "events_in_timed_region": rec["events_in_timed_region"], "clocks": clocks}
        if "collective_check" in rec:
            line["collective_check"] = rec["collective_check"]
    if D.world == 1 and wl == "ns" and not args.particles and not args.eager and not args.no_configs:
        # one sub-record per BASELINE configuration, same measurement, fresh context each
        subs = {}
        for name in SUB_RECORDS:
            swl, eager = (name[:-6], True) if name.endswith("_eager") else (name, False)
            try:
                r = measure(D, pkg, synth, swl, WORKLOADS[swl][2], args.configs_steps, max(3, min(args.warmup, 5)), args.configs_steps, eager=eager)
                subs[name] = {k: r[k] for k in ("value", "unit", "steps", "ms_per_step", "live_particles", "config", "e2e", "roofline",
                                                "events_in_timed_region")}
            except Exception as e:  # noqa: BLE001 - a sub-record must not take the headline down
                subs[name] = {"error": f"{type(e).__name__}: {e}"}
        line["configs"] = subs
    if D.rank == 0:
        if D.world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(synth, wl, n_per_gpu, args.cpu_sample)
        print(json.dumps(line), flush=True)
    if D.dist is not None:
        D.dist.barrier()
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
