"""Compile-time resource budget of the step kernel (no GPU needed: cuobjdump reads the built library).

The 1024-thread variant only fits one block per SM with 64 registers per thread; an innocent-looking edit that pushes
live state into local memory costs real time (a multi-step outer loop tried in session 3 took the kernel from an
8-byte to a 168-byte stack frame and the step from 108 to 118 us).  This test makes such a change visible
before it reaches a GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "biocma-mcst_b200", "libbmc_b200.so")


def _usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library not available")
    out = subprocess.run([exe, "-res-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    return res


def test_every_step_kernel_fits_one_block_per_sm():
    # cycle_kernel<Model, VEC, WB, LAZY>: one block of 256*WB threads per SM -> at most 65536 / (256*WB) registers per thread;
    # spills are tolerated for the wide models only, and bounded
    res = _usage()
    hits = {k: v for k, v in res.items() if "cycle_kernel" in k}
    assert len(hits) >= 12
    for k, v in hits.items():
        m = re.search(r"ELi(\d)ELi(\d)ELb([01])E", k)
        assert m, k
        wb = int(m.group(2))
        assert v["reg"] <= 65536 // (256 * wb), (k, v)
        assert v["stack"] <= 320 and v["local"] == 0, (k, v)
        assert v["shared"] <= 14 * 1024, (k, v)


@pytest.mark.parametrize("model", ["Monod", "FixedLength"])
def test_step_kernel_register_and_stack_budget(model):
    res = _usage()
    # cycle_kernel<Model, VEC=4, WB, LAZY=true>: WB 4 -> 1024 threads x <= 64 registers, WB 3 -> 768 x <= 80
    for wb, max_reg in ((4, 64), (3, 80)):
        hits = {k: v for k, v in res.items() if f"cycle_kernelINS_{len(model)}{model}ELi4ELi{wb}ELb1E" in k}
        if not hits:
            continue   # variant not instantiated for this model
        for k, v in hits.items():
            assert v["reg"] <= max_reg, (k, v)
            # 768 threads: spill-free.  1024 threads x 64 registers: the spills sit in the post-cycle phase (compaction,
            # commit), executed once per warp and step — the particle-pass loop itself has none (checked below)
            assert v["stack"] <= (16 if wb == 3 else 128) and v["local"] == 0, (k, v)
            assert v["shared"] <= 14 * 1024, (k, v)                      # static shared memory reserve of configure_launch
    assert any(f"{model}ELi4ELi4ELb1E" in k for k in res), "default variant missing"


def test_particle_pass_loop_of_the_headline_kernel_has_no_spills():
    # SASS of cycle_kernel<Monod, 4, 4, true>: every local-memory access lies behind the last 128-bit particle-column
    # load of the pass, i.e. in the post-cycle phase
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library not available")
    out = subprocess.run([exe, "-sass", "-fun", "_ZN3bmc12cycle_kernelINS_5MonodELi4ELi4ELb1EEEvNS_11CycleParamsE", LIB],
                         capture_output=True, text=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    assert len(lines) > 2000
    stores = [i for i, l in enumerate(lines) if "STG.E.128" in l]
    local = [i for i, l in enumerate(lines) if re.search(r"\b(STL|LDL)\b", l)]
    # the pass stores its columns with STG.E.128; the first of them marks the body, the last one its end
    assert stores and (not local or min(local) > max(stores)), (min(local), max(stores))
