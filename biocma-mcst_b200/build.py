"""Builds libbmc_b200.so (C-ABI + sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no fallback arch
  -fmad=false                               IEEE arithmetic without contraction:
                                            deterministic models are bit-exact
                                            against the oracle (-ffp-contract=off)
  -lineinfo                                 ncu source page maps to the .cuh files
Each model's kernels live in their own translation unit; the units compile in parallel.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbmc_b200.so")
OBJ_DIR = os.path.join(HERE, "build")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "bmc.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def needs_build(out=OUT):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps() if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    common = [nvcc_path()]
    if os.path.exists("/usr/bin/g++"):  # the image exports CC/CXX=/opt/gcc; nvcc is happy with the distro g++
        common += ["-ccbin", "/usr/bin/g++"]
    common += ["-std=c++17", "-O3", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"] + [f"-D{d}" for d in defines]
    if verbose:
        common += ["-Xptxas=-v"]

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, f"{tag}_{os.path.splitext(os.path.basename(src))[0]}.o")
        r = subprocess.run(common + ["-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    r = subprocess.run(common[:3] + ["-shared", "-o", out] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libbmc_b200.so failed")
    return out


def build_host(force=False):
    """C++ host layer (biocma-mcst_b200/host): CLI driver and the container test, g++ -std=c++17,
    linked against libbmc_b200.so with an rpath to this directory."""
    host = os.path.join(HERE, "host")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    outs = []
    for src, exe in (("biocma_b200_cli.cpp", "biocma_b200"), ("test_container.cpp", "test_container")):
        out = os.path.join(host, exe)
        srcp = os.path.join(host, src)
        hdr = os.path.join(host, "bmc_host.hpp")
        if force or not os.path.exists(out) or max(os.path.getmtime(srcp), os.path.getmtime(hdr), os.path.getmtime(OUT)) > os.path.getmtime(out):
            r = subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-o", out, srcp, "-L" + HERE, "-lbmc_b200",
                                "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/..", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                               capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"g++ failed on {src}")
        outs.append(out)
    return outs


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
    if not outs:
        print(build_host(force="--force" in sys.argv))
