"""Per-phase device time of kernel_construct over a full bench-like run (developer tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, data
K = int(sys.argv[1]) if len(sys.argv) > 1 else 36
batches, mn, mx = data.terrain_batches(K, list(range(K)))
sim = SimLOD(1920, 1080, persistent_bytes=8 << 30)
sim.set_box(mn, mx)
n = K * 1_000_000
dptr = sim.device_alloc(n * 16)
sim.memcpy_htod(dptr, np.concatenate(batches).view(np.uint8))
for rep in range(2):
    sim.reset(); sim.flush_l2()
    kms, tms = sim.insert_device(dptr, n)
    ph = sim.memcpy_dtoh(sim.buffers().momentary + 96, 64).view(np.uint64) / 1e3
    names = ["fused(alloc|count+sample|insert)", "split", "rewalk", "deferred", "final_alloc", "final_insert+stats", "split_rounds(count)", "prologue"]
    print("kernel ms %.3f total ms %.3f Mpts/s %.1f" % (kms, tms, n / kms / 1e3))
    vb = sim.memcpy_dtoh(sim.buffers().momentary + 64, 32).view(np.uint64)
    print("voxels first-visit / re-walk:", int(vb[0]), int(vb[1]), "spilled", int(vb[2]), "total", int(vb[3]))
    print("phase us:", {k: round(float(v), 1) for k, v in zip(names, ph)}, "sum", round(float(ph.sum()), 1), flush=True)
sim.close()
