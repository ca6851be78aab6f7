#!/usr/bin/env python
"""Turn the ncu captures of tools/r2_profile.sh (gpurun_out/<tag>_cycle_<workload>.ncu-rep, <tag>_launches.csv) into the
committed summaries under profiles/ and into profiles/traffic.json (DRAM bytes per launch of the step kernel, per
workload, with the kernel instantiation and the commit the capture was taken on)."""
import csv
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table)

commit = os.environ.get("PROFILE_COMMIT") or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
dirty = subprocess.run(["git", "status", "--porcelain", "biocma-mcst_b200/csrc"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
commit += "+" if dirty else ""
# entries of workloads without a new capture are kept (they carry the commit and the capture they came from)
try:
    with open(os.path.join(P, "traffic.json")) as _f:
        _old_traffic = json.load(_f)
except Exception:  # noqa: BLE001
    _old_traffic = {}
traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel (ncu --set full --clock-control none, "
                   "7th launch of `bench.py --workload <w> --steps 8 --warmup 3`); bench.py reports an entry only when the kernel "
                   "instantiation it ran has the same block size and vector width"}


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def to_bytes(v, u):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v.replace(",", "")) * mult


for name in ("ns", "c2", "c2_eager", "c3", "c4", "fl_ns", "sa_ns"):
    rep = os.path.join(G, f"{tag}_cycle_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    m = raw_metrics(rep)
    kname = m["Kernel Name"][0] if "Kernel Name" in m else "cycle_kernel"
    block = int(re.sub(r"[^0-9]", "", m["Block Size"][0].split(",")[0])) if "Block Size" in m else None
    wl = name[:-6] if name.endswith("_eager") else name
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    vec = int(re.search(r"<[^,]+,\s*(\d+)", kname).group(1)) if re.search(r"<[^,]+,\s*(\d+)", kname) else 4
    traffic[name] = {"bytes_per_launch": dram, "particles": bench.WORKLOADS[wl][2], "kernel": kname.split("(")[0], "block": block, "vec": vec,
                     "commit": commit, "capture": f"profiles/{out_tag}_step_kernel_ncu_{name}.txt", "when": time.strftime("%Y-%m-%d")}
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "60"], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{out_tag}_step_kernel_ncu_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:cycle_kernel -s 6 -c 1 python bench.py --workload {wl}"
                f"{' --eager' if name.endswith('_eager') else ''} --steps 8 --warmup 3 --no-configs --no-cpu-baseline   (commit {commit})\n"
                f"# kernel: {kname}\n")
        f.write(summ)
    print(name, kname.split("(")[0], block, f"{dram/1e6:.1f} MB", f"{dram / bench.WORKLOADS[wl][2]:.2f} B/particle")

c5 = os.path.join(G, f"{tag}_ncu_c5.csv")
if os.path.exists(c5):
    rows = [r for r in csv.reader(open(c5)) if len(r) > 6 and r[0].isdigit()]
    vals = {r[-3]: float(r[-1].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[-2], 1) for r in rows}
    kname = rows[0][4]
    dram = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    traffic["c5"] = {"bytes_per_launch": dram, "particles": bench.WORKLOADS["c5"][2], "kernel": kname.split("(")[0], "block": 768, "vec": 1,
                     "commit": commit, "capture": f"profiles/{out_tag}_step_kernel_ncu_c5.txt", "when": time.strftime("%Y-%m-%d")}
    with open(os.path.join(P, f"{out_tag}_step_kernel_ncu_c5.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none "
                f"-k regex:cycle_kernel -s 6 -c 1 python bench.py --workload c5 --steps 8 --warmup 3 --no-configs --no-cpu-baseline (commit {commit})\n# kernel: {kname}\n")
        for k, v in vals.items():
            f.write(f"{k} {v}\n")
    print("c5", f"{dram/1e9:.2f} GB", f"{dram / bench.WORKLOADS['c5'][2]:.1f} B/particle")

for _k, _v in _old_traffic.items():
    if _k != "note" and _k not in traffic:
        traffic[_k] = _v
with open(os.path.join(P, "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)

lc = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 6 and r[0].isdigit()]
    per = {}
    for r in rows:
        k = r[4].split("(")[0]
        per.setdefault(k, []).append(float(r[-1].replace(",", "")) / (1e3 if r[-2] == "ns" else 1.0))
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(P, f"{out_tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 20 --warmup 3 --no-configs "
                f"--no-cpu-baseline   (commit {commit}; default workload ns = 1.25e8 particles, 500 compartments, monod)\n"
                "# per-launch times are cold-cache and serialised: the SHARE of the step is what should agree with bench.py\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:70]:70s} {len(v):8d} {sum(v):12.1f} {sum(v)/len(v):10.1f} {sum(v)/tot:7.3f}\n")
    print(open(os.path.join(P, f"{out_tag}_launches.txt")).read())
