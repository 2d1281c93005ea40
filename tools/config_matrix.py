"""Developer tool (not part of bench.py): SURVEY.md §8(d) configs 4 and 5 on one GPU.

config 5: the octree of config 3 (350 x 1M-point terrain batches) built once by OUR builder, then
          `kernel_render` only at 1920x1080 from six cameras, both shading paths, timed with our
          render kernel and with the reference's render kernel (oracle/_ref/ref_render.cubin) swapped
          in on the SAME octree; the raw u64 framebuffers of the two are compared bit for bit.
config 4: one GPU's share of the sphere-shell weak-scaling case (250 x 1M-point batches in
          latitude/longitude tile order, cube 4096^3), inserted from HBM by our kernels and by the
          reference's kernels; Stats of the two runs are compared.

usage: python tools/config_matrix.py [terrain_batches=350] [shell_batches=250] [out.json]
"""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from simlod_b200 import SimLOD, camera, data  # noqa: E402

KT = int(sys.argv[1]) if len(sys.argv) > 1 else 350
KS = int(sys.argv[2]) if len(sys.argv) > 2 else 250
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "config_matrix.json")
W, H = 1920, 1080
res = {}

DET = ("numNodes", "numInner", "numLeaves", "numNonemptyLeaves", "numPoints", "numVoxels", "numChunksPoints",
       "numChunksVoxels", "numAllocatedChunks", "chunkPoolSize", "allocatedBytes_persistent")


def det_stats(s):
    return {k: int(getattr(s, k)) for k in DET}


# ------------------------------------------------------------------ config 5
if KT > 0:
    batches, mn, mx = data.terrain_batches(KT, list(range(KT)))
    n = KT * 1_000_000
    cams = [("morro_bird", camera.orbit_camera(width=W, height=H, **camera.MORRO_BIRD)),
            ("morro_close", camera.orbit_camera(width=W, height=H, **camera.MORRO_CLOSE))]
    cams += [("autofocus+%d*pi/2" % k, camera.autofocus(mx, W, H, yaw_offset=k * np.pi / 2)) for k in range(4)]

    def build(render_blocks_per_sm):
        sim = SimLOD(W, H, persistent_bytes=max(8 << 30, KT * (220 << 20)), render_blocks_per_sm=render_blocks_per_sm)
        sim.set_box(mn, mx)
        dptr = sim.device_alloc(n * 16)
        for b, pts in enumerate(batches):
            sim.memcpy_htod(dptr + b * 1_000_000 * 16, pts.view(np.uint8))
        sim.reset(); sim.insert_device(dptr, 3 * 1_000_000); sim.reset(); sim.flush_l2()
        kms, tms = sim.insert_device(dptr, n)
        st = sim.stats()
        assert st.numPoints == n and st.dbg == 0, (st.numPoints, st.dbg)
        sim.device_free(dptr)
        return sim, kms, tms, st

    # pass 1: both render kernels at the SAME grid (3 blocks/SM: the EDL tile walk depends on gridDim),
    # raw u64 framebuffer compared bit for bit on the same octree
    sim, kms, tms, st = build(3)
    res["config3_build"] = {"points": n, "kernel_ms": kms, "total_ms": tms, "mpoints_per_s_kernel": n / kms / 1e3,
                            "mpoints_per_s_total": n / tms / 1e3, **det_stats(st)}
    print("config3 build", json.dumps(res["config3_build"]), flush=True)
    fbs, equal = {}, {}
    for impl in ("ours", "reference"):
        if impl == "reference":
            sim.use_module(1, oracle.REF_CUBINS[1])
        for hqs in (0, 1):
            sim.set_settings(useHighQualityShading=hqs)
            for name, (v, p) in cams:
                sim.set_camera(v, p)
                sim.render()
                fb = sim.framebuffer().copy()
                if impl == "ours":
                    fbs[(hqs, name)] = fb
                else:
                    equal["hqs%d/%s" % (hqs, name)] = bool(np.array_equal(fb, fbs[(hqs, name)]))
    print("framebuffer equality (same octree, same grid):", equal, flush=True)
    res["config5_framebuffer_bit_exact_vs_reference_kernel"] = equal
    del fbs
    sim.close()

    # pass 2: timings, each kernel at its own occupancy-derived grid (main.cpp:493-497)
    sim, kms2, tms2, _ = build(0)
    res["config3_build"]["kernel_ms_second_run"] = kms2
    frames, blocks = {}, {}
    for impl in ("ours", "reference"):
        if impl == "reference":
            sim.use_module(1, oracle.REF_CUBINS[1])
        blocks[impl] = sim.launch_info()["render_blocks"]
        rows = []
        for hqs in (0, 1):
            sim.set_settings(useHighQualityShading=hqs)
            for name, (v, p) in cams:
                sim.set_camera(v, p)
                sim.render()
                ms = float(np.median([sim.render() for _ in range(7)]))
                s = sim.stats()
                samples = int(s.numVisiblePoints + s.numVisibleVoxels)
                rows.append({"hqs": hqs, "camera": name, "ms": ms, "fps": 1000.0 / ms, "visible_nodes": int(s.numVisibleNodes),
                             "visible_points": int(s.numVisiblePoints), "visible_voxels": int(s.numVisibleVoxels),
                             "msamples_per_s": samples / ms / 1e3})
        frames[impl] = rows
        for hqs in (0, 1):
            f = [r for r in rows if r["hqs"] == hqs]
            print(impl, "hqs", hqs, "Msamples/s %.0f" % (sum(r["visible_points"] + r["visible_voxels"] for r in f) / sum(r["ms"] for r in f) / 1e3),
                  [(r["camera"], round(r["ms"], 3)) for r in f], flush=True)
    res["config5_render"] = {"render_blocks": blocks, **frames}
    sim.close()
    del batches
    json.dump(res, open(out, "w"), indent=1)

# ------------------------------------------------------------------ config 4 (one GPU's share)
if KS > 0:
    n = KS * 1_000_000
    with ThreadPoolExecutor(min(16, os.cpu_count() or 4)) as ex:
        parts = list(ex.map(lambda b: data.shell(n, b * 1_000_000, 1_000_000)[0], range(KS)))
    mn, mx = (0.0, 0.0, 0.0), (data.SHELL_CUBE,) * 3
    r4 = {"workload": "sphere shell R=1800+-0.25 in 4096^3, %d x 1M-point batches in lat/lon tile order" % KS}
    for impl in ("reference", "ours"):
        sim = SimLOD(W, H, momentary_bytes=oracle.REF_MOMENTARY_BYTES, persistent_bytes=max(8 << 30, KS * (260 << 20)),
                     construct_blocks_per_sm=1 if impl == "reference" else 0)
        if impl == "reference":
            for p in (0, 1, 2):
                sim.use_module(p, oracle.REF_CUBINS[p])
        sim.set_box(mn, mx)
        dptr = sim.device_alloc(n * 16)
        for b, pts in enumerate(parts):
            sim.memcpy_htod(dptr + b * 1_000_000 * 16, pts.view(np.uint8))
        sim.reset(); sim.insert_device(dptr, 3 * 1_000_000); sim.reset(); sim.flush_l2()
        kms, tms = sim.insert_device(dptr, n)
        st = sim.stats()
        r4[impl] = {"kernel_ms": kms, "total_ms": tms, "mpoints_per_s_kernel": n / kms / 1e3, "mpoints_per_s_total": n / tms / 1e3,
                    "dbg": int(st.dbg), "memCapacityReached": int(st.memCapacityReached), **det_stats(st)}
        print("config4", impl, json.dumps(r4[impl]), flush=True)
        sim.device_free(dptr)
        sim.close()
    r4["stats_equal"] = all(r4["ours"][k] == r4["reference"][k] for k in DET)
    r4["speedup_kernel"] = r4["ours"]["mpoints_per_s_kernel"] / r4["reference"]["mpoints_per_s_kernel"]
    res["config4_shell_1gpu_share"] = r4
    print("config4 stats_equal", r4["stats_equal"], "speedup", r4["speedup_kernel"], flush=True)

json.dump(res, open(out, "w"), indent=1)
