"""The shipped cubins driven exactly like the reference host drives its kernels, bypassing simlod_b200/csrc/host.cpp:
tests/native/dropin_harness.cpp restates initCudaProgram's buffers (300 000 000-byte momentary buffer filled with
garbage, never cleared), the three argument arrays and the reference launch shapes (reset 1 x 1, construct numSMs x 256,
render occupancy x numSMs) on the plain driver API. Its octree and frame must equal the ones of the C-ABI path."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from simlod_b200 import SimLOD, camera, data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "native", "dropin_harness")


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="tests/native/dropin_harness not built (__graft_entry__.build())")
def test_cubins_under_the_reference_launch_sequence(tmp_path):
    pts, mn, mx = data.terrain(2_700_000)
    W, H = 1920, 1080
    sim = SimLOD(W, H, persistent_bytes=4 << 30)
    try:
        sim.set_box(mn, mx)
        view, proj = camera.autofocus(mx, W, H)
        sim.set_camera(view, proj)
        sim.reset()
        sim.insert_batches(list(data.batches(pts)))
        st = sim.stats()
        cn = oracle.canon_from_image(*sim.download_octree())
        sim.render()
        fb = sim.framebuffer()
        st_r = sim.stats()
        (tmp_path / "points.bin").write_bytes(pts.tobytes())
        (tmp_path / "uniforms.bin").write_bytes(sim.uniforms_bytes())
    finally:
        sim.close()
    out = str(tmp_path / "out")
    r = subprocess.run([HARNESS, os.path.join(ROOT, "simlod_b200", "cubin"), str(tmp_path / "points.bin"), str(tmp_path / "uniforms.bin"), out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    hs = np.fromfile(out + ".stats", dtype=np.uint8)
    from simlod_b200.api import Stats
    hst = Stats.from_buffer_copy(hs.tobytes())
    addrs = np.fromfile(out + ".addrs", dtype=np.uint64)
    hcn = oracle.canon_from_image(np.fromfile(out + ".nodes", dtype=np.uint8), np.fromfile(out + ".heap", dtype=np.uint8), int(addrs[0]), int(addrs[1]))
    assert hst.dbg == 0 and hst.numPoints == len(pts)
    diffs = oracle.compare_canon(hcn, cn, "reference launch sequence") + oracle.compare_stats(hst, st)
    assert not diffs, "\n".join(diffs)
    # same frame: visibility counters exactly, and the depth word of every pixel (which point or voxel colour wins a
    # depth tie depends on the racy order of samples inside chunks)
    for f in ("numVisibleNodes", "numVisibleInner", "numVisibleLeaves", "numVisiblePoints", "numVisibleVoxels"):
        assert getattr(hst, f) == getattr(st_r, f), f
    hfb = np.fromfile(out + ".fb", dtype=np.uint64).reshape(H, W)
    assert ((hfb >> np.uint64(32)) == (fb >> np.uint64(32))).all()
