"""Where the difference between `value` (whole insert) and kernel-only time goes: device-clock start / end of every
kernel_construct launch of one pass (Ctl::launchClock), so gaps between launches are seen as the GPU sees them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, data  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 120
n = NB * 1_000_000
sim = SimLOD(1920, 1080, persistent_bytes=max(4 << 30, NB * (60 << 20)))
sim.set_box((0, 0, 0), data.TERRAIN_EXTENT)
dptr = sim.device_alloc(n * 16)
sim.generate(sim.GEN_TERRAIN, dptr, n, 0, n, 7)
for rep in range(2):
    sim.reset(); sim.flush_l2()
    kms, tms = sim.insert_device(dptr, n)
raw = sim.memcpy_dtoh(sim.buffers().momentary + 272, 32 * 16 + 4)
clk = raw[:512].view(np.uint64).reshape(32, 2).astype(np.int64)
count = int(raw[512:516].view(np.uint32)[0])
L = min(count, 32)
order = [(count - L + i) % 32 for i in range(L)]
starts, ends = clk[order, 0], clk[order, 1]
dur = (ends - starts) / 1e3
gaps = (starts[1:] - ends[:-1]) / 1e3
print("pass: kernel-only %.3f ms, total %.3f ms, %d launches" % (kms, tms, count))
print("launch durations us:", [round(float(x)) for x in dur])
print("gaps between launches us:", [round(float(x)) for x in gaps])
print("sum of durations %.3f ms, sum of gaps %.3f ms, first start -> last end %.3f ms" % (dur.sum() / 1e3, gaps.sum() / 1e3, (ends[-1] - starts[0]) / 1e6))
sim.close()
