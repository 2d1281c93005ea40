"""Ad-hoc GPU bring-up check (developer tool; the real checks live in tests/)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlod_b200 import SimLOD, data, camera, api
import oracle

W, H = 1920, 1080
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pts, mn, mx = data.uniform_cube(N)

def build(sim, ref):
    if ref:
        for prog in (0, 2):
            sim.use_module(prog, oracle.REF_CUBINS[prog])
    else:
        for prog in (0, 2):
            sim.use_module(prog, None)
    sim.set_box(mn, mx)
    sim.reset()
    t = time.time()
    ms = sim.insert_batches(data.batches(pts))
    s = sim.stats()
    print("ref" if ref else "ours", "insert kernel ms", ms, "wall", time.time() - t, "Mpts/s", N / ms / 1e3, flush=True)
    print({f: getattr(s, f) for f, _ in s._fields_ if not f.startswith("_")}, flush=True)
    return s

sim = SimLOD(W, H, momentary_bytes=oracle.REF_MOMENTARY_BYTES, persistent_bytes=8 << 30)
print(sim.launch_info(), flush=True)
s_ours = build(sim, False)
img = sim.download_octree()
c_ours = oracle.canon_from_image(*img)

o = oracle.Oracle(mn, mx)
for b in data.batches(pts):
    o.add_batch(b)
c_or = o.canon()
print("ours vs oracle canon:", oracle.compare_canon(c_ours, c_or) or "EQUAL")
print("ours vs oracle stats:", oracle.compare_stats(s_ours, o.stats()) or "EQUAL")
print("voxel colour violations:", o.check_voxel_colors(c_ours))

# render ours on our octree
view, proj = camera.autofocus(mx, W, H)
sim.set_camera(view, proj)
for hqs in (0, 1):
    sim.set_settings(useHighQualityShading=hqs)
    sim.use_module(1, None)
    ms = sim.render(); ms = sim.render()
    fb_o = sim.framebuffer(); su_o = sim.surface(); st_o = sim.stats()
    print("hqs", hqs, "ours render ms", ms, "visible nodes/points/voxels", st_o.numVisibleNodes, st_o.numVisiblePoints, st_o.numVisibleVoxels, flush=True)
    info = sim.launch_info()
    sim.use_module(1, oracle.REF_CUBINS[1])
    info_r = sim.launch_info()
    print("render blocks ours", info["render_blocks"], "ref", info_r["render_blocks"])
    ms = sim.render(); ms = sim.render()
    fb_r = sim.framebuffer(); su_r = sim.surface(); st_r = sim.stats()
    print("hqs", hqs, "ref render ms", ms, "visible nodes/points/voxels", st_r.numVisibleNodes, st_r.numVisiblePoints, st_r.numVisibleVoxels, flush=True)
    neq = fb_o != fb_r
    print("  fb mismatches:", int(neq.sum()), "depth mismatches:", int(((fb_o >> 32) != (fb_r >> 32)).sum()), "surface mismatches:", int((su_o != su_r).sum()))
    if neq.any():
        ys, xs = np.nonzero(neq)
        for k in range(min(5, len(ys))):
            print("   ", ys[k], xs[k], hex(int(fb_o[ys[k], xs[k]])), hex(int(fb_r[ys[k], xs[k]])))
sim.use_module(1, None)
sim.set_settings(useHighQualityShading=0)

# reference builder on the same input
s_ref = build(sim, True)
c_ref = oracle.canon_from_image(*sim.download_octree())
print("ref vs oracle canon:", oracle.compare_canon(c_ref, c_or) or "EQUAL")
print("ref vs oracle stats:", oracle.compare_stats(s_ref, o.stats()) or "EQUAL")
print("ref voxel colour violations:", o.check_voxel_colors(c_ref))
print("ours vs ref canon:", oracle.compare_canon(c_ours, c_ref) or "EQUAL")
