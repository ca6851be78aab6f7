import sys, os, ctypes, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
os.environ["BMC_LIB"] = os.path.join(os.getcwd(), "biocma-mcst_b200", "lib_tl.so")
from _bmc_loader import load_pkg, load_synth
import util, torch
pkg, synth = load_pkg(), load_synth()
N = int(os.environ.get("TL_N", "10000000"))
case = util.make_case(synth, "monod", N, 500, dt=0.1)
g = pkg.ParticleLoop("monod", 1, 500)
util.load_case(g, case)
for _ in range(20): g.cycle(0.1)
g.sync()
# read DevState dbg: find offset by scanning: use cudaMemcpy of the whole struct via torch
lib = g.lib
# DevState pointer is not exported; use bmc_get_counters? add helper: read via debug export
buf = (ctypes.c_ulonglong * 16)()
rc = lib.bmc_debug_timeline(g.h, buf)
t = np.array(list(buf), dtype=np.int64)
print("rc", rc); print("block 0 stamps [us]:", np.round((t[:9] - t[0]) / 1000.0, 1))
G = 148
raw = np.zeros(4 * G, np.uint32)
lib.bmc_debug_blocks(g.h, raw.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(raw.size))
raw = raw.reshape(G, 4)
tend = (raw[:, 0].astype(np.int64) | (raw[:, 1].astype(np.int64) << 32)) - t[0]
tiles, smid = raw[:, 2], raw[:, 3]
us = tend / 1000.0
print("end of particle pass per block [us]: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (us.min(), np.percentile(us, 10), np.median(us), np.percentile(us, 90), us.max()))
for nt in np.unique(tiles): print("  tiles", nt, "blocks", (tiles == nt).sum(), "mean end %.1f" % us[tiles == nt].mean())
per_sm = {}
for b in range(G): per_sm.setdefault(int(smid[b]), []).append(us[b])
sm_mean = np.array([np.mean(v) for v in per_sm.values()]); sm_spread = np.array([np.max(v) - np.min(v) for v in per_sm.values()])
print("per-SM mean end: min %.1f max %.1f ; within-SM spread mean %.1f max %.1f ; SMs %d" % (sm_mean.min(), sm_mean.max(), sm_spread.mean(), sm_spread.max(), len(per_sm)))
order = np.argsort(us); print("slowest blocks:", [(int(b), int(smid[b]), int(tiles[b]), round(float(us[b]), 1)) for b in order[-8:]])
print("fastest blocks:", [(int(b), int(smid[b]), int(tiles[b]), round(float(us[b]), 1)) for b in order[:8]])

# per-block start / end of the last two kernels (step parity halves)
raw2 = np.zeros(12 * G, np.uint32)
lib.bmc_debug_blocks(g.h, raw2.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(raw2.size))
w = raw2[4 * G:].reshape(G, 8).astype(np.int64)
def u64(lo, hi): return lo | (hi << 32)
st = [u64(w[:, 0], w[:, 1]), u64(w[:, 4], w[:, 5])]; en = [u64(w[:, 2], w[:, 3]), u64(w[:, 6], w[:, 7])]
last = 0 if st[0].min() > st[1].min() else 1   # the later kernel
prev = 1 - last
print("previous kernel: first start 0, last start %.1f, first end %.1f, last end %.1f us" % tuple((x - st[prev].min()) / 1000.0 for x in (st[prev].max(), en[prev].min(), en[prev].max())))
print("gap: last end of previous kernel -> first start of next %.1f us, -> last start of next %.1f us" % ((st[last].min() - en[prev].max()) / 1000.0, (st[last].max() - en[prev].max()) / 1000.0))
print("next kernel: duration first start -> last end %.1f us ; period (start to start) %.1f us" % ((en[last].max() - st[last].min()) / 1000.0, (st[last].min() - st[prev].min()) / 1000.0))
