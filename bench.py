#!/usr/bin/env python
"""bench.py — particle-steps/s of the Monte-Carlo particle loop on N B200s.

A "step" is one cycleProcess (model update + division, contribution scatter,
compartment move, outlet exit, compaction, spawn) over this rank's particles.
Workload at any N: BASELINE.json configs[1] per GPU — stirred-tank CMA with 500
compartments, Monod uptake model, 1e7 particles per GPU (weak scaling), fixed
synthetic flow map, one outlet, dt = 0.1 s; synthetic population initialised on
the device.  `--workload c3` is configs[2] (1e8 particles per GPU, faster
growth so division/removal/compaction are exercised every step).

  value     : whole-job particle-steps/s, state resident in HBM, no host sync
              inside the timed region, one NCCL all-reduce of the source vector
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
# than the figure SURVEY.md §8d derives for an eagerly updated layout (kept below as B_SURVEY).
B_ALG = {"monod": 41, "fixed_length": 21, "simple_acetate": 49, "wide_udf": 8 * 32 + 9}      # multi-compartment
B_ALG_0D = {"monod": 37, "fixed_length": 17, "simple_acetate": 45, "wide_udf": 8 * 32 + 5}  # 0D: position only read
B_SURVEY = {"monod": 57, "fixed_length": 37, "simple_acetate": 65, "wide_udf": 8 * 32 + 25}
N_SPECIES = {"monod": 1, "fixed_length": 1, "simple_acetate": 2, "wide_udf": 4}

WORKLOADS = {
    # name: (model, n_comp, particles per GPU, dt, near_division, p_exit)
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
}


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


def run_reference(args, wl):
    """Reference arm: the reference path's CPU implementation (oracle restatement, OpenMP,
    all host threads) on the same config; each step is a bounded sample of the workload."""
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
            "config": {"workload": f"{wl}: {n_comp}-compartment stirred-tank CMA, {model}, {n_full} particles/GPU, dt={dt}"},
            "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="particles in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (default: --steps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = args.workload
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    from _bmc_loader import load_pkg, load_synth
    pkg, synth = load_pkg(), load_synth()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the particle loop has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    model, n_comp, n_per_gpu, dt, near, p_exit = WORKLOADS[wl]
    if args.particles:
        n_per_gpu = args.particles
    fm, flows, conc = build_case(synth, model, n_comp, dt, p_exit)
    loop = pkg.ParticleLoop(model, N_SPECIES[model], n_comp, device=local, seed=2024, rank=rank)
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
    loop.set_weight(0.5 * float(fm["volumes"].sum()) / (total_mass * world))
    setup_loop(loop, fm, flows, conc, n_comp)
    collective = ""
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
        use_p2p = want_p2p == "1" or (want_p2p == "auto" and world >= int(os.environ.get("BMC_P2P_MIN_RANKS", "4")))
        if use_p2p and world <= 16:
            from biocma_mcst_b200 import sharding
            if sharding.setup_peer_allreduce(loop, world, rank, device=torch.device("cuda", local),
                                             log=lambda m: print(f"[bench] {m}; NCCL", file=sys.stderr)):
                collective = "1 peer-memory all-reduce/step (one-shot over NVLink, NCCL only for set-up)"
    stream = torch.cuda.ExternalStream(loop.stream_handle(), device=torch.device("cuda", local))

    def barrier():
        loop.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        loop.cycle(dt)
        if world > 1:
            loop.allreduce_sources()

    # ---------------- value: device-resident steps -----------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    n_live0 = loop.counters()["n_used"]
    launches0 = loop.launch_count()
    loop.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_n = loop.profile_read()
    loop.profile_enable(False)
    launches = loop.launch_count() - launches0
    n_live1 = loop.counters()["n_used"]
    live_avg = 0.5 * (n_live0 + n_live1)

    # ---------------- e2e: host-buffer ABI every step ---------------------------
    e2e_steps = args.e2e_steps or args.steps
    conc_host = np.ascontiguousarray(conc, np.float64)
    for _ in range(3):
        loop.set_concentrations(conc_host); step_resident(); loop.get_sources()
    barrier()
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
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    n_e1 = loop.counters()["n_used"]
    live_e2e = 0.5 * (n_e0 + n_e1)

    clocks = sampler.stop() if rank == 0 else None

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

    if rank == 0:
        value = live_tot * args.steps / (ms * 1e-3)
        e2e_v = live_e2e_tot * e2e_steps / (ms_e2e * 1e-3)
        peak, peak_src = peaks()
        b_alg = B_ALG[model] if n_comp > 1 else B_ALG_0D[model]
        achieved = (live_avg * b_alg) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        nb = N_SPECIES[model] * n_comp * 8
        traffic = None  # DRAM bytes per launch of the cycle kernel from the committed ncu capture of this workload
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f).get(wl)
            if t and t["particles"] == n_per_gpu:
                traffic = t["bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl}: {n_comp}-compartment stirred-tank CMA, {model}, {n_per_gpu} particles/GPU, dt={dt}",
                       "parallelism": f"particle-sharded x{world}, replicated liquid state, {collective}" if world > 1 else "single GPU",
                       "l2_policy": f"inputs larger than L2 ({n_per_gpu * (loop.n_var * 4 + 13) / 1e6:.0f} MB of particle state per GPU)"},
            "e2e": {"value": e2e_v, "unit": "particle-steps/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches_tot),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": f"cycle_kernel<{model}> (whole step: particle pass + post-cycle phase)", "kernel_ms": k_ms,
                         "bytes_per_particle": b_alg, "bytes_per_particle_survey_8d": B_SURVEY[model],
                         "frac_survey_8d": (live_avg * B_SURVEY[model]) / (k_ms * 1e-3) / 1e9 / peak if k_ms > 0 else 0.0,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "kernel_share_of_step": k_ms * args.steps / ms},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            n = min(n_per_gpu, args.cpu_sample)
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
            line["cpu_baseline"] = {"value": done / el, "unit": "particle-steps/s", "cores": threads, "kind": kind,
                                    "sample": f"{n} of {n_per_gpu} particles x {steps_cpu} steps, {label}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
