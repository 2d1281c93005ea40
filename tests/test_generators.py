"""The device-side generators of the benchmark streams (simlod_b200/csrc/gen.cu, simlod_generate) against the
numpy generators they restate (simlod_b200/data.py). uniform_cube and terrain are bit-identical; shell uses the
device's sin/cos and may differ in the last place of a coordinate for a handful of points."""
import numpy as np
import pytest

from simlod_b200 import SimLOD, data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim():
    s = SimLOD(320, 176, persistent_bytes=1 << 30)
    yield s
    s.close()


def device_points(sim, kind, n_total, first, count, seed, size=0.0):
    dptr = sim.device_alloc(max(count, 1) * 16)
    try:
        sim.generate(kind, dptr, n_total, first, count, seed, size)
        return sim.memcpy_dtoh(dptr, count * 16).view(data.POINT_DTYPE)
    finally:
        sim.device_free(dptr)


def test_uniform_cube_is_bit_identical(sim):
    for n, size, seed, first in ((1_000_000, 1024.0, 42, 0), (120_000, 64.0, 5, 0), (300_000, 2048.0, 77, 1_000_000)):
        got = device_points(sim, sim.GEN_UNIFORM, 0, first, n, seed, size)
        want = data.uniform_cube(n, size=size, seed=seed, first=first)[0]
        assert got.tobytes() == want.tobytes()


def test_terrain_is_bit_identical(sim):
    # sub-ranges of differently sized scans, including strip boundaries (serpentine direction changes) and the last points
    for n_total, first, count in ((3_300_000, 0, 3_300_000), (36_000_000, 17_900_000, 400_000), (350_000_000, 349_700_000, 300_000),
                                  (350_000_000, 3_645_000, 10_000), (1_000, 0, 1_000)):
        got = device_points(sim, sim.GEN_TERRAIN, n_total, first, count, 7)
        want = data.terrain(n_total, first, count, seed=7)[0]
        neq = got.view(np.uint32).reshape(-1, 4) != want.view(np.uint32).reshape(-1, 4)
        assert not neq.any(), "%d of %d points differ (n_total %d, first %d)" % (int(neq.any(axis=1).sum()), count, n_total, first)
    got = device_points(sim, sim.GEN_TERRAIN, 2_000_000, 500_000, 100_000, 11)
    assert got.tobytes() == data.terrain(2_000_000, 500_000, 100_000, seed=11)[0].tobytes()


def test_shell_matches_numpy_to_the_last_place(sim):
    for n_total, first, count in ((2_400_000, 0, 2_400_000), (2_000_000_000, 1_999_000_000, 1_000_000)):
        got = device_points(sim, sim.GEN_SHELL, n_total, first, count, 1234)
        want = data.shell(n_total, first, count)[0]
        assert (got["color"] == want["color"]).all()
        for ax in "xyz":
            d = np.abs(got[ax].view(np.int32).astype(np.int64) - want[ax].view(np.int32).astype(np.int64))
            assert d.max() <= 1, (ax, int(d.max()))                       # neighbouring floats at worst
            assert (d != 0).mean() < 1e-4, (ax, float((d != 0).mean()))
