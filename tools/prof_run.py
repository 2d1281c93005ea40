"""Workload for ncu captures: deterministic launches (B batches each, far below the 10 ms in-kernel
budget so that every replay pass does the same work), then a few frames.
usage: prof_run.py [num_launches] [batches_per_launch] [total_batches_of_the_stream]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, camera, data  # noqa: E402

BATCH = 1_000_000
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
TOTAL = int(sys.argv[3]) if len(sys.argv) > 3 else 36
sim = SimLOD(1920, 1080, persistent_bytes=6 << 30)
sim.set_box((0, 0, 0), data.TERRAIN_EXTENT)
n = L * B * BATCH
dptr = sim.device_alloc(n * 16)
sim.generate(sim.GEN_TERRAIN, dptr, TOTAL * BATCH, 0, n, 7)
sim.reset()
for l in range(L):
    for b in range(B):
        sim.upload_batch_device(dptr + (l * B + b) * BATCH * 16, BATCH)
    ms = sim.update_octree()
    print("launch", l, "ms", ms, "batches done", sim.stats().batchletIndex, flush=True)
ph = sim.memcpy_dtoh(sim.buffers().momentary + 96, 64).view(np.uint64) / 1e3
names = ["fused(alloc|count+sample|insert)", "split", "rewalk", "deferred", "final_alloc", "final_insert+stats", "split_rounds(count)", "prologue"]
print("phase us:", {n_: round(float(v), 1) for n_, v in zip(names, ph)}, "total", round(float(ph.sum()), 1), flush=True)
view, proj = camera.orbit_camera(width=1920, height=1080, **camera.MORRO_BIRD)
sim.set_camera(view, proj)
for i in range(4):
    print("render ms", sim.render(), flush=True)
sim.set_settings(useHighQualityShading=1)
for i in range(3):
    print("render hqs ms", sim.render(), flush=True)
sim.close()
