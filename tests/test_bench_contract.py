"""bench.py contract of the reference arm, which runs without a GPU: one JSON line on stdout with the keys the driver
reads, the reference's own kernels as the thing timed (kind "reference") when oracle/_ref holds the timing build, and the
behaviour under torchrun's environment (OMP_NUM_THREADS=1; non-zero ranks stay silent)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "impl", "cpu_baseline", "e2e")


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                           "--cpu-sample", "150000", *args], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")   # what a torchrun rank 0 sees
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["n_gpus"] == 2 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref
    if os.path.exists(ref.RELEASE_LIB_PATH):
        assert cb["kind"] == "reference", cb   # the reference's own kernels, not the port, are what is timed


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_port_baseline_can_be_forced():
    r = _run({"BMC_CPU_BASELINE": "port"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"
