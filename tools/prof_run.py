"""Workload for ncu captures: deterministic launches (B batches each, far below the 10 ms in-kernel
budget so that every replay pass does the same work), then a few frames.
usage: prof_run.py [num_launches] [batches_per_launch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from simlod_b200 import SimLOD, camera  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
batches, mn, mx = bench.generate_batches(36, list(range(L * B)))
sim = SimLOD(1920, 1080, persistent_bytes=3 << 30)
sim.set_box(mn, mx)
dptr = sim.device_alloc(L * B * bench.BATCH * 16)
sim.memcpy_htod(dptr, np.concatenate(batches).view(np.uint8))
sim.reset()
for l in range(L):
    for b in range(B):
        sim.upload_batch_device(dptr + (l * B + b) * bench.BATCH * 16, bench.BATCH)
    ms = sim.update_octree()
    print("launch", l, "ms", ms, "batches done", sim.stats().batchletIndex, flush=True)
ph = sim.memcpy_dtoh(sim.buffers().momentary + 96, 64).view(np.uint64) / 1e3
names = ["fused(alloc|count+sample|insert)", "split", "rewalk", "deferred", "final_alloc", "final_insert+stats", "split_rounds(count)", "prologue"]
print("phase us:", {n: round(float(v), 1) for n, v in zip(names, ph)}, "total", round(float(ph.sum()), 1), flush=True)
view, proj = camera.orbit_camera(width=1920, height=1080, **camera.MORRO_BIRD)
sim.set_camera(view, proj)
for i in range(4):
    print("render ms", sim.render(), flush=True)
sim.set_settings(useHighQualityShading=1)
for i in range(3):
    print("render hqs ms", sim.render(), flush=True)
sim.close()
