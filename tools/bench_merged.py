"""Developer tool: ONE octree over G GPUs — run under torchrun, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_merged.py [batches_per_gpu=8] [level=2] [out.json]

Every rank holds its round-robin shard of a K*G-batch terrain scan in HBM. Per step each rank partitions one
1 M-point batch by owner and the points travel to their owners ("p2p": the scatter kernel stores straight into the
owners' receive buffers over NVLink peer memory; "nccl": local scatter + all_to_all_single), then every rank inserts
what it received. Reported: aggregate Mpoints/s (host clock around synchronised regions, max over ranks), the split
between exchange and insertion, and a bit-exact check: each rank rebuilds, alone, the octree of the stream it should
have received (it regenerates all ranks' batches and partitions them locally) and compares canonical forms.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (checker only: canonical forms of the two device octrees)
from simlod_b200 import SimLOD, data  # noqa: E402
from simlod_b200 import dist as sdist  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LEVEL = int(sys.argv[2]) if len(sys.argv) > 2 else 2
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "merged.json")
DEPTH = int(sys.argv[4]) if len(sys.argv) > 4 else 8          # batches per exchange group

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 1_000_000
total_batches = K * world
mine = sdist.shard_batches(total_batches, rank, world)
batches, mn, mx = data.terrain_batches(total_batches, mine)
sim = SimLOD(640, 360, device=local, persistent_bytes=max(6 << 30, K * world * (260 << 20)))
sim.set_box(mn, mx)
src = sim.device_alloc(K * B * 16)
for i, b in enumerate(batches):
    sim.memcpy_htod(src + i * B * 16, b.view(np.uint8))

# plan: global per-cell histogram of everything -> LPT owners (same on every rank)
plan0 = sim.partition_plan(LEVEL, np.zeros(8 ** LEVEL, np.uint8), world)
hist = np.zeros(8 ** LEVEL, np.int64)
for i in range(K):
    hist += sim.partition_count(src + i * B * 16, B, plan0)[1].astype(np.int64)
t = torch.tensor(hist, device=dev)
dist.all_reduce(t)
owners = sdist.plan_owners(t.cpu().numpy(), world)
load = np.bincount(owners, weights=t.cpu().numpy(), minlength=world)
res = {"world": world, "batches_per_gpu": K, "level": LEVEL, "group_depth": DEPTH, "points_total": K * world * B, "owner_load": [int(v) for v in load]}


def sync_all():
    sim.synchronize()
    torch.cuda.synchronize()
    dist.barrier()


def run(mode):
    overlap = mode.endswith("_overlap")          # send group g+1 before inserting group g (dist.py: send_group / wait_group)
    ex = sdist.SpatialExchange(sim, LEVEL, owners, capacity_points=B, depth=DEPTH, mode=mode.replace("_overlap", ""), device=dev)
    best = None
    for rep in range(3):
        sim.reset()
        sync_all()
        t_ex = t_ins = 0.0
        kernel_ms = 0.0
        t0 = time.perf_counter()
        for w0 in range(0, K, 32):                   # plan a window of steps: their counts travel in one all_gather
            ex.prepare([(src + i * B * 16, B) for i in range(w0, min(K, w0 + 32))])
        t_prep = time.perf_counter() - t0
        groups = [[(src + i * B * 16, B) for i in range(g0, min(K, g0 + DEPTH))] for g0 in range(0, K, DEPTH)]
        if overlap:
            ex.send_group(groups[0])
        for gi, group in enumerate(groups):
            a = time.perf_counter()
            if overlap:
                ptr, n = ex.wait_group()
                if gi + 1 < len(groups):
                    ex.send_group(groups[gi + 1])
            else:
                ptr, n = ex.exchange_group(group)
            b = time.perf_counter()
            if n:
                kms, _ = sim.insert_device(ptr, n)
                kernel_ms += kms
            c = time.perf_counter()
            t_ex += b - a
            t_ins += c - b
        sync_all()
        dt = time.perf_counter() - t0
        vals = torch.tensor([dt, t_ex, t_ins, kernel_ms, t_prep], dtype=torch.float64, device=dev)
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dt, t_ex, t_ins, kernel_ms, t_prep = [float(v) for v in vals.cpu()]
        if best is None or dt < best["seconds"]:
            best = {"seconds": dt, "mpoints_per_s": K * world * B / dt / 1e6, "plan_window_s": t_prep, "exchange_s": t_ex, "insert_s": t_ins,
                    "construct_kernel_ms": kernel_ms}
    st = sim.stats()
    tot = sdist.reduce_stats(st, dev)
    best["numPoints_all_ranks"] = tot["numPoints"]
    best["numPoints_this_rank"] = int(st.numPoints)
    canon = oracle.canon_from_image(*sim.download_octree())
    return best, canon, st


# what this rank should have received, rebuilt locally: all ranks' batches of every step, partitioned here
def rebuild_locally():
    plan = sim.partition_plan(LEVEL, owners, world)
    everyone = [sdist.shard_batches(total_batches, r, world) for r in range(world)]
    tmp = sim.device_alloc(B * 16)
    land = sim.device_alloc(DEPTH * world * B * 16)
    scratch = [sim.device_alloc(B * 16) for _ in range(world)]          # the other ranks' buckets are thrown away
    sim.reset()
    pos = 0
    for i in range(K):
        for s in range(world):
            pts = data.terrain(total_batches * B, everyone[s][i] * B, B)[0]
            sim.memcpy_htod(tmp, pts.view(np.uint8))
            counts, _ = sim.partition_count(tmp, B, plan)
            ptrs = [land if d == rank else scratch[d] for d in range(world)]
            offs = [pos if d == rank else 0 for d in range(world)]
            sim.partition_scatter(tmp, B, plan, ptrs, offs)
            sim.synchronize()
            pos += int(counts[rank])
        if (i + 1) % DEPTH == 0 or i + 1 == K:      # the exchange delivers a group of DEPTH batches as one contiguous stream
            if pos:
                sim.insert_device(land, pos)
            pos = 0
    st = sim.stats()
    canon = oracle.canon_from_image(*sim.download_octree())
    for p in [tmp, land] + scratch:
        sim.device_free(p)
    return canon, st


want_canon, want_stats = rebuild_locally()
for mode in ("nccl", "p2p", "p2p_overlap"):
    try:
        r, canon, st = run(mode)
        diffs = oracle.compare_canon(canon, want_canon, mode) + oracle.compare_stats(st, want_stats)
        ok = torch.tensor([0 if diffs else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        r["octree_bit_exact_vs_local_rebuild_all_ranks"] = bool(int(ok.item()))
        if diffs:
            print("rank", rank, mode, "DIFFS", diffs[:5], flush=True)
        res[mode] = r
    except Exception as e:                      # tool only: report which transport is unavailable on this box
        res[mode] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(mode, json.dumps(res[mode]), flush=True)

# render the merged octree: every rank rasterises its part, the packed framebuffers are depth-composited over peer
# memory (one kernel per rank) and, beside it, with NCCL all_reduce(MIN) through the host path of dist.py
try:
    from simlod_b200 import camera
    W, H = 640, 360
    comp = sdist.FramebufferCompositor(sim, W, H, device=dev)
    sim.set_settings(useHighQualityShading=0)
    frames = []
    for name, cam in (("morro_bird", camera.orbit_camera(width=W, height=H, **camera.MORRO_BIRD)), ("autofocus", camera.autofocus(mx, W, H))):
        sim.set_camera(*cam)
        sim.render()
        sync_all()
        t0 = time.perf_counter()
        render_ms = sim.render()
        t1 = time.perf_counter()
        comp.composite()
        t2 = time.perf_counter()
        got = comp.read(W, H)
        want = sdist.composite_framebuffers(sim.framebuffer(), dev)
        ok = torch.tensor([1 if np.array_equal(got, want) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        tt = torch.tensor([t1 - t0, t2 - t1], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        covered = int((got >> np.uint64(32) != np.uint64(0x7f800000)).sum())
        frames.append({"camera": name, "render_kernel_ms": render_ms, "render_call_ms": float(tt[0]) * 1e3, "composite_ms": float(tt[1]) * 1e3,
                       "equals_nccl_allreduce_min_all_ranks": bool(int(ok.item())), "covered_pixels": covered})
    res["render_composite"] = {"width": W, "height": H, "frames": frames}
except Exception as e:
    res["render_composite"] = {"error": "%s: %s" % (type(e).__name__, e)}
if rank == 0:
    print("render", json.dumps(res["render_composite"]), flush=True)

# baseline beside it: the batch-sharded forest (no exchange), same data
sim.reset()
sync_all()
t0 = time.perf_counter()
kms, _ = sim.insert_device(src, K * B)
sync_all()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
res["batch_sharded_forest"] = {"seconds": float(dt.item()), "mpoints_per_s": K * world * B / float(dt.item()) / 1e6}
if rank == 0:
    print("sharded", json.dumps(res["batch_sharded_forest"]), flush=True)
    json.dump(res, open(out, "w"), indent=1)
sim.close()
dist.destroy_process_group()
