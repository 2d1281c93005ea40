"""Developer experiment: build variant cubins of the builder (compile-time knobs of construct.cu, or any *.cubin dropped
into tools/exp/), check each against the shipped kernel (deterministic Stats + canonical octree of a full 36-batch
build) and time it: whole build, and per-phase µs per batch on a realistic tree (30 batches with the shipped kernel,
6 more with the variant).

  python tools/exp_variants.py            # on the GPU box; nvcc is in the image
"""
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402  (checker: canonical forms of the two device octrees)
from simlod_b200 import SimLOD  # noqa: E402

# name -> extra nvcc flags for simlod_b200/csrc/construct.cu
KNOBS = {
    "tile256": ["-DSIMLOD_TILE_POINTS=256"],
    "tile1024": ["-DSIMLOD_TILE_POINTS=1024"],
    "tab128": ["-DSIMLOD_VOXTAB_SIZE=128"],
    "tab32": ["-DSIMLOD_VOXTAB_SIZE=32"],
    "no_tma": ["-DSIMLOD_NO_TMA"],
    "dyn": ["-DSIMLOD_DYNAMIC_TILES"],                                   # tiles from a global cursor (written blind at the
    "dyn256": ["-DSIMLOD_DYNAMIC_TILES", "-DSIMLOD_TILE_POINTS=256"],   # end of round 1: never run, parity unknown)
}
EXP = os.path.join(ROOT, "tools", "exp")
os.makedirs(EXP, exist_ok=True)
import hashlib  # noqa: E402
_h = hashlib.sha256()
for _f in ("simlod_b200/csrc/construct.cu", "simlod_b200/csrc/fpmath.cuh", "include/simlod_abi.h"):
    _h.update(open(os.path.join(ROOT, _f), "rb").read())
SRC_TAG = _h.hexdigest()[:8]                  # variants are rebuilt when the kernel source changes
for name, flags in KNOBS.items():
    out = os.path.join(EXP, "knob_%s_%s.cubin" % (name, SRC_TAG))
    for stale in glob.glob(os.path.join(EXP, "knob_%s_*.cubin" % name)):
        if stale != out:
            os.remove(stale)
    if not os.path.exists(out):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-cubin"] + flags + \
              ["-o", out, os.path.join(ROOT, "simlod_b200", "csrc", "construct.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print("build of", name, "failed:", r.stderr[-400:], flush=True)

K, PRE = 36, 30
batches, mn, mx = bench.generate_batches(K, list(range(K)))
sim = SimLOD(1920, 1080, persistent_bytes=8 << 30)
sim.set_box(mn, mx)
dptr = sim.device_alloc(K * bench.BATCH * 16)
sim.memcpy_htod(dptr, np.concatenate(batches).view(np.uint8))
names = ["fused(alloc|count+sample|insert)", "split", "rewalk", "deferred", "final_alloc", "final_insert+stats", "split_rounds(count)", "prologue"]


def phases():
    return sim.memcpy_dtoh(sim.buffers().momentary + 96, 64).view(np.uint64).astype(np.float64) / 1e3


def full_build(module):
    sim.use_module(0, module)
    best = None
    for rep in range(3):
        sim.reset(); sim.flush_l2()
        kms, tms = sim.insert_device(dptr, K * bench.BATCH)
        best = kms if best is None else min(best, kms)
    st = sim.stats()
    canon = oracle.canon_from_image(*sim.download_octree())
    return best, st, canon


base_ms, base_stats, base_canon = full_build(None)
print("shipped kernel: %.3f ms for %d batches = %.0f Mpoints/s kernel-only" % (base_ms, K, K * bench.BATCH / base_ms / 1e3), flush=True)
for v in [None] + sorted(glob.glob(os.path.join(EXP, "*.cubin"))):
    label = os.path.basename(v) if v else "shipped"
    try:
        ms, st, canon = full_build(v)
        diffs = oracle.compare_canon(canon, base_canon, label) + oracle.compare_stats(st, base_stats)
        for rep in range(2):
            sim.use_module(0, None)
            sim.reset()
            sim.insert_device(dptr, PRE * bench.BATCH)
            p0 = phases()
            sim.use_module(0, v)
            kms, tms = sim.insert_device(dptr + PRE * bench.BATCH * 16, (K - PRE) * bench.BATCH)
            p1 = phases() - p0
        print(label.ljust(22), "full build %.3f ms (%+.1f %%)" % (ms, 100.0 * (ms - base_ms) / base_ms), "identical octree" if not diffs else "DIFFERS: %s" % diffs[:2],
              {n: round(float(x) / (K - PRE), 1) for n, x in zip(names, p1)}, flush=True)
    except Exception as e:
        print(label, "failed:", e, flush=True)
sim.use_module(0, None)
sim.close()
