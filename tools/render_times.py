"""Developer tool: render timings on the 36 M terrain octree (6 cameras, both shading paths)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, camera, data
K = 36
batches, mn, mx = data.terrain_batches(K, list(range(K)))
sim = SimLOD(1920, 1080, persistent_bytes=8 << 30)
sim.set_box(mn, mx)
n = K * 1_000_000
dptr = sim.device_alloc(n * 16)
sim.memcpy_htod(dptr, np.concatenate(batches).view(np.uint8))
sim.reset(); sim.insert_device(dptr, n)
cams = [("af%d" % k, camera.autofocus(mx, 1920, 1080, yaw_offset=k * np.pi / 2)) for k in range(4)]
cams += [("bird", camera.orbit_camera(width=1920, height=1080, **camera.MORRO_BIRD)), ("close", camera.orbit_camera(width=1920, height=1080, **camera.MORRO_CLOSE))]
for hqs in (0, 1):
    sim.set_settings(useHighQualityShading=hqs)
    out = []
    for name, (v, p) in cams:
        sim.set_camera(v, p); sim.render()
        ms = min(sim.render() for _ in range(5)); s = sim.stats()
        out.append((name, round(ms, 4), s.numVisibleNodes, s.numVisiblePoints + s.numVisibleVoxels))
    tot = sum(o[3] for o in out) / sum(o[1] for o in out) / 1e3
    print("hqs", hqs, out, "Msamples/s %.0f" % tot, flush=True)
sim.close()
