"""Developer experiment: time the count+sample pass of variant kernels on a realistic tree.
Builds the first 30 batches with the real kernel, then inserts 6 more with the variant."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from simlod_b200 import SimLOD
K, PRE = 36, 30
batches, mn, mx = bench.generate_batches(K, list(range(K)))
sim = SimLOD(1920, 1080, persistent_bytes=8 << 30)
sim.set_box(mn, mx)
dptr = sim.device_alloc(K * bench.BATCH * 16)
sim.memcpy_htod(dptr, np.concatenate(batches).view(np.uint8))
names = ["count+sample", "split", "rewalk", "deferred", "alloc", "insert", "stats", "prologue"]
def phases():
    return sim.memcpy_dtoh(sim.buffers().momentary + 160, 64).view(np.uint64).astype(np.float64) / 1e3
variants = [None] + sorted(glob.glob(os.path.join(ROOT, "tools", "exp", "*.cubin")))
for v in variants:
    for rep in range(2):
        sim.use_module(0, None)
        sim.reset()
        sim.insert_device(dptr, PRE * bench.BATCH)
        p0 = phases()
        sim.use_module(0, v)
        kms, tms = sim.insert_device(dptr + PRE * bench.BATCH * 16, (K - PRE) * bench.BATCH)
        p1 = phases() - p0
    print(os.path.basename(v) if v else "baseline", "kernel ms %.3f" % kms, {n: round(float(x) / (K - PRE), 1) for n, x in zip(names, p1)}, flush=True)
sim.close()
