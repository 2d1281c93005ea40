"""The driver's round-end smoke check, kept under test: __graft_entry__.smoke() builds a small octree through the C ABI,
renders one frame and compares both with the CPU oracle."""
import pytest


@pytest.mark.gpu
def test_graft_entry_smoke():
    import __graft_entry__ as g
    g.smoke()
