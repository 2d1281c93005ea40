"""world_size-2 gloo tests of the multi-GPU host logic (sharding + reductions), CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from simlod_b200 import data
from simlod_b200 import dist as sdist


def test_round_robin_sharding_partitions_the_stream():
    for nb in (0, 1, 7, 36, 250):
        for g in (1, 2, 4, 8):
            owned = [sdist.shard_batches(nb, r, g) for r in range(g)]
            flat = sorted(b for o in owned for b in o)
            assert flat == list(range(nb))
            assert all(o == sorted(o) for o in owned)
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    for nb in (0, 1, 7, 36, 250):
        for g in (1, 2, 4, 8):
            owned = [sdist.shard_batches_blocks(nb, r, g) for r in range(g)]
            assert [b for o in owned for b in o] == list(range(nb))          # contiguous, in order, disjoint, complete
            assert max(len(o) for o in owned) <= -(-nb // g)
    ranges = sdist.shard_point_range(2_500_001, 1_000_000, 0, 2)
    assert ranges == [(0, 1_000_000), (2_000_000, 500_001)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, bs = 230_000, 50_000
        pts, mn, mx = data.uniform_cube(n, size=64.0, seed=3)
        # each rank sees only its shard; the cube comes from an all-reduce of the local extents
        mine = [pts[f:f + c] for f, c in sdist.shard_point_range(n, bs, rank, world)]
        cat = np.concatenate(mine)
        lmin = [float(cat[a].min()) for a in "xyz"]
        lmax = [float(cat[a].max()) for a in "xyz"]
        gmin, gmax = sdist.global_box(lmin, lmax)
        o = oracle.Oracle((0.0, 0.0, 0.0), (64.0, 64.0, 64.0))
        for b in mine:
            o.add_batch(b)
        total = sdist.reduce_stats(o.stats())
        tmax = sdist.max_over_ranks(1.0 + rank)
        # framebuffer compositing: u64 min across ranks
        fb = np.full((4, 4), (0x7f800000 << 32) | 0x00332211, dtype=np.uint64)
        fb[rank, rank] = (np.uint64(0x3f800000 + rank) << np.uint64(32)) | np.uint64(0xff000000 + rank)
        fb[3, 3] = (np.uint64(0x40000000 - rank) << np.uint64(32)) | np.uint64(rank)
        comp = sdist.composite_framebuffers(fb)
        if rank == 0:
            np.savez(out, gmin=gmin, gmax=gmax, numPoints=total["numPoints"], processed=total["numPointsProcessed"],
                     batches=total["batchletIndex"], tmax=tmax, comp=comp)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_insertion(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    pts, _, _ = data.uniform_cube(230_000, size=64.0, seed=3)
    assert np.allclose(r["gmin"], [pts[a].min() for a in "xyz"]) and np.allclose(r["gmax"], [pts[a].max() for a in "xyz"])
    assert int(r["numPoints"]) == 230_000 and int(r["processed"]) == 230_000 and int(r["batches"]) == 5
    assert float(r["tmax"]) == 2.0
    comp = r["comp"]
    assert comp[0, 0] == (np.uint64(0x3f800000) << np.uint64(32)) | np.uint64(0xff000000)
    assert comp[1, 1] == (np.uint64(0x3f800001) << np.uint64(32)) | np.uint64(0xff000001)
    assert comp[3, 3] == (np.uint64(0x3fffffff) << np.uint64(32)) | np.uint64(1)      # rank 1 is closer
    assert comp[2, 2] == (np.uint64(0x7f800000) << np.uint64(32)) | np.uint64(0x00332211)
