"""Developer tool: .simlod streamer timeline for several loader-thread counts and file sizes (SIMLOD_STREAM_TRACE=1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, data
os.environ["SIMLOD_STREAM_TRACE"] = "1"
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 48
batches, mn, mx = data.terrain_batches(NB, list(range(NB)))
for nb in (16, NB):
    path = "/dev/shm/probe_%d.simlod" % os.getpid()
    data.write_simlod(path, np.concatenate(batches[:nb]), mn, mx)
    sim = SimLOD(320, 176, persistent_bytes=max(4 << 30, nb * (220 << 20)))
    for threads in (8, 16, 24, 32):
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter(); got, kms, tms = sim.insert_simlod_file(path, loader_threads=threads); dt = time.perf_counter() - t0
            best = min(best, dt)
        print("batches", nb, "threads", threads, "best %.2f ms %.0f Mpts/s device_ms %.2f" % (best * 1e3, nb * 1_000_000 / best / 1e6, tms), flush=True)
    sim.close(); os.remove(path)
