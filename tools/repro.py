"""Scratch reproduction script for compute-sanitizer runs (developer aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, data  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_300_123
pts, mn, mx = data.terrain(n)
sim = SimLOD(320, 176, persistent_bytes=3 << 30)
sim.set_box(mn, mx)
for rep in range(3):
    sim.reset()
    if rep == 0:
        for b in data.batches(pts):
            sim.upload_batch(b)
        while sim.stats().batchletIndex < (n + 999_999) // 1_000_000:
            sim.update_octree()
    else:
        sim.insert_batches(data.batches(pts))
    st = sim.stats()
    print("build", rep, "points", st.numPoints, "nodes", st.numNodes, "dbg", st.dbg, flush=True)
sim.close()
