"""Builds libbmc_b200.so (C-ABI + sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no fallback arch
  -fmad=false                               IEEE arithmetic without contraction:
                                            deterministic models are bit-exact
                                            against the oracle (-ffp-contract=off)
  -lineinfo                                 ncu source page maps to the .cuh files
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbmc_b200.so")
SOURCES = [os.path.join(CSRC, "bmc_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("bmc_kernels.cuh", "bmc_models.cuh", "bmc_rng.cuh")] + [
    os.path.join(HERE, "..", "include", "bmc.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    cmd = [nvcc_path(), "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-shared", "-o", out] + [f"-D{d}" for d in defines] + SOURCES + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc (no libgomp spec); nvcc is happy with the distro g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libbmc_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
