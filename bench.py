#!/usr/bin/env python
"""bench.py — octree insertion throughput (and render rate) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (oracle port)

A step is one pass of the hot path over one batch: 1 000 000 synthetic XYZRGBA points (16 B each)
inserted into the growing octree by kernel_construct. Workload = BASELINE.json configs[1]: the
Morro-Bay stand-in terrain (no dataset on the box) streamed in spatially coherent 1 M-point
batches from a freshly reset octree; K steps = a K-million-point scan (default K = 36 = the 36 M
configuration). With N > 1 every rank owns a builder and inserts K batches of a N*K-batch scan,
round-robin (b % N == rank): per-GPU work is fixed, so scaling is weak; there is no data-path
collective, only barrier / max-time / stats reductions over NCCL.

Numbers:
  value     Mpoints/s, all ranks' points / max-over-ranks device time, batches already resident in
            HBM (device-to-device copies into the 50-slot ring are inside the timed region)
  e2e       same metric through the public API from pinned HOST memory: per step one 16 MB
            host->device copy and one 112-byte Stats read-back per launch inside the timed region
  roofline  kernel_construct: algorithmic bytes (16 in + 16 out + 32*s + 16*v per point)
            / summed launch time (CUDA events on the launch stream) against the measured HBM copy peak
  cpu_baseline  the CPU oracle (port of the reference algorithm, 1 thread) on a bounded sample
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 1_000_000
METRIC = "Mpoints/sec octree insertion (16B XYZRGBA)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=36)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-batches", type=int, default=8)
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "blocks", "roundrobin"],
                    help="N > 1: every rank inserts its own K-batch scan of the N=1 extent and density (tiles: a survey of N tiles, "
                         "per-GPU work identical to N=1), or a share of ONE N*K-batch scan: batches [g*K, (g+1)*K) (blocks) / b %% N == g (roundrobin)")
    return ap.parse_args()


# ---- clocks sampled DURING the timed regions (B200_PROFILING.md) --------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not any(a <= t <= b + 0.15 for a, b in windows):
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:      # timed regions shorter than the sampling period: fall back to all samples
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx = max(mx, float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def generate_batches(total_batches, mine, threads=None, seed=7):
    """This rank's batches of the terrain scan of total_batches * 1 M points (counter-based generator)."""
    from concurrent.futures import ThreadPoolExecutor
    from simlod_b200 import data
    n_total = total_batches * BATCH
    threads = threads or min(16, os.cpu_count() or 4)
    with ThreadPoolExecutor(threads) as ex:
        out = list(ex.map(lambda b: data.terrain(n_total, b * BATCH, BATCH, seed)[0], mine))
    return out, (0.0, 0.0, 0.0), data.TERRAIN_EXTENT


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per kernel_construct launch from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "ncu_construct_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def bench_las(batches, mn, mx, device):
    """LAS format-2 records (26 B/point) -> 16-byte points: device decode with the records resident in HBM, end to end
    from pinned host memory, and the reference's own CPU loader (oracle/_ref/libref_las.so = LasLoader.cpp) beside it."""
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    from simlod_b200 import SimLOD, data
    nb = min(8, len(batches))
    scale, offset = (0.001, 0.001, 0.001), (0.0, 0.0, 0.0)
    recs = [data.las_records(b, 2, scale, offset) for b in batches[:nb]]
    bpp = recs[0].shape[1]
    n = nb * BATCH
    sim = SimLOD(320, 176, device=device, persistent_bytes=1 << 30)
    out = {"format": 2, "bytes_per_point": bpp, "points": n}
    try:
        sim.set_box(mn, mx)
        layout = sim.las_layout(bpp, 2, scale, offset)
        dptr = sim.device_alloc(n * bpp)
        hptr = sim.host_alloc(n * bpp)
        host = np.ctypeslib.as_array((ctypes.c_uint8 * (n * bpp)).from_address(hptr))
        host[:] = np.concatenate(recs).reshape(-1)
        sim.memcpy_htod(dptr, host)
        for mode in ("device", "host"):
            best = None
            for rep in range(3):
                sim.reset(); sim.flush_l2(); sim.synchronize()
                t0 = time.perf_counter()
                for k in range(nb):
                    if mode == "device":
                        sim.upload_batch_las_device(dptr + k * BATCH * bpp, BATCH, layout)
                    else:
                        sim._check(sim._lib.simlod_upload_batch_las(sim._ctx, hptr + k * BATCH * bpp, BATCH, ctypes.byref(layout)))
                sim.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            key = "value" if mode == "device" else "e2e"
            out[key] = {"value": round(n / best / 1e6, 1), "unit": "Mpoints/s",
                        "how": "records resident in HBM" if mode == "device" else "records in pinned host memory, H2D inside the timed region"}
        got = sim.ring_slot(nb - 1, 1000)
        want = oracle.decode_las(recs[nb - 1][:1000], 1000, bpp, 2, scale, offset)
        assert (got == want).all()
        out["roofline"] = {"bound": "hbm", "algorithmic_bytes_per_point": bpp + 16,
                           "achieved_gbs": round(out["value"]["value"] * 1e6 * (bpp + 16) / 1e9, 1)}
    finally:
        sim.close()
    if oracle.ref_las() is not None:
        d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        path = os.path.join(d, "simlod_bench_%d.las" % os.getpid())
        try:
            data.write_las(path, np.concatenate(batches[:nb]), 2, scale, offset)
            best = None
            for threads in sorted({1, min(8, os.cpu_count() or 1), os.cpu_count() or 1}):
                dt_t = oracle.ref_las_bench(path, n, 250_000, threads)
                if best is None or dt_t < best[0]:
                    best = (dt_t, threads)
            dt, threads = best
            out["cpu_baseline"] = {"value": round(n / dt / 1e6, 1), "unit": "Mpoints/s", "cores": threads, "kind": "reference",
                                   "sample": "loadLasNative (LasLoader.cpp compiled from /root/reference) on %d M points from tmpfs in 250k-point batches, long-lived loader threads; best of 1 / 8 / all cores" % (n // 1000000)}
        finally:
            if os.path.exists(path):
                os.remove(path)
    return out


def bench_stream_file(batches, mn, mx, device):
    """Second "next" row: .simlod file (tmpfs) -> loader threads -> pinned pool -> ring -> octree, with the reference's own
    loadFileNative (SimlodLoader.cpp compiled from /root/reference) timed beside it on the host cores."""
    import tempfile
    import oracle
    from simlod_b200 import SimLOD, data
    nb = min(16, len(batches))
    n = nb * BATCH
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "simlod_bench_%d.simlod" % os.getpid())
    out = {"points": n, "file_bytes": 24 + 16 * n}
    try:
        data.write_simlod(path, np.concatenate(batches[:nb]), mn, mx)
        sim = SimLOD(320, 176, device=device, persistent_bytes=max(4 << 30, nb * (220 << 20)))
        try:
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                got, kms, tms = sim.insert_simlod_file(path, loader_threads=min(16, os.cpu_count() or 8))
                dt = time.perf_counter() - t0
                assert got == n and sim.stats().numPoints == n
                if best is None or dt < best[0]:
                    best = (dt, kms, tms)
            out["e2e"] = {"value": round(n / best[0] / 1e6, 1), "unit": "Mpoints/s", "how": "wall clock incl. reset, %d loader threads, file in tmpfs" % min(16, os.cpu_count() or 8),
                          "device_ms": round(best[2], 3), "kernel_ms": round(best[1], 3)}
        finally:
            sim.close()
        if oracle.ref_simlod() is not None:
            best = None
            for threads in sorted({1, min(8, os.cpu_count() or 1), os.cpu_count() or 1}):
                dt_t = oracle.ref_simlod_bench(path, n, BATCH, threads)
                if best is None or dt_t < best[0]:
                    best = (dt_t, threads)
            out["cpu_baseline"] = {"value": round(n / best[0] / 1e6, 1), "unit": "Mpoints/s", "cores": best[1], "kind": "reference",
                                   "sample": "loadFileNative (SimlodLoader.cpp compiled from /root/reference), %d x 1M-point reads from tmpfs into host memory; best of 1 / 8 / all cores" % nb}
    finally:
        if os.path.exists(path):
            os.remove(path)
    return out


def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: oracle port (the reference has no CPU octree
    builder to compile; its kernels need a GPU). One step = one full 1 M-point batch."""
    if rank != 0:
        return
    import oracle
    W, K = args.warmup, args.steps
    total = world * K
    from simlod_b200 import dist as sdist
    if args.shard == "tiles":
        batches, mn, mx = generate_batches(K, list(range(K)), seed=7)            # rank 0's tile
    else:
        mine = sdist.shard_batches_blocks(total, 0, world) if args.shard == "blocks" else sdist.shard_batches(total, 0, world)
        batches, mn, mx = generate_batches(total, mine[:max(K, W)])
    o = oracle.Oracle(mn, mx)
    for b in batches[:W]:
        o.add_batch(b)
    o = oracle.Oracle(mn, mx)
    t0 = time.perf_counter()
    for b in batches[:K]:
        o.add_batch(b)
    dt = time.perf_counter() - t0
    s = o.stats()
    assert s.numPoints == K * BATCH
    val = K * BATCH / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpoints/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f32",
        "data": "synthetic",
        "config": {"workload": "terrain_synth_%dM (Morro Bay stand-in), %d x 1M-point batches, streamed from reset" % (K, K),
                   "note": "reference algorithm restated for the CPU (oracle/oracle.cpp); the reference itself has no CPU builder"},
        "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": 1, "kind": "port", "sample": "all %d batches of rank 0's stream" % K},
        "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from simlod_b200 import SimLOD, camera
    from simlod_b200 import dist as sdist

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = "cuda:%d" % local_rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 0), args.steps
    total_batches = world * K
    t_gen = time.time()
    if args.shard == "tiles":           # rank g: its own K-batch scan (tile g of the survey), same extent and density as N = 1
        batches, mn, mx = generate_batches(K, list(range(K)), seed=7 + rank)
    else:
        mine = sdist.shard_batches_blocks(total_batches, rank, world) if args.shard == "blocks" else sdist.shard_batches(total_batches, rank, world)
        batches, mn, mx = generate_batches(total_batches, mine)
    t_gen = time.time() - t_gen
    npts = K * BATCH

    sim = SimLOD(1920, 1080, device=local_rank, persistent_bytes=max(8 << 30, K * (220 << 20)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    windows = []
    try:
        sim.set_box(mn, mx)
        # pinned host copy (e2e source) and device-resident copy (value source) of this rank's stream
        host_ptr = sim.host_alloc(npts * 16)
        host = np.ctypeslib.as_array((ctypes.c_uint8 * (npts * 16)).from_address(host_ptr))
        for i, b in enumerate(batches):
            host[i * BATCH * 16:(i + 1) * BATCH * 16] = b.view(np.uint8).reshape(-1)
        dptr = sim.device_alloc(npts * 16)
        sim.memcpy_htod(dptr, host)

        # warm-up: W batches (code, clocks, allocator paths), then start over
        sim.reset()
        if W:
            sim.insert_device(dptr, W * BATCH)
        sim.reset()
        info0 = sim.launch_info()

        # ---- timed region 1: inputs resident in HBM --------------------------------------------
        sim.flush_l2()
        barrier()
        t0 = time.time()
        kernel_ms, total_ms = sim.insert_device(dptr, npts)
        barrier()
        windows.append((t0, time.time()))
        st = sim.stats()
        assert st.numPointsProcessed == npts and st.numPoints == npts and st.dbg == 0, (st.numPointsProcessed, st.numPoints, st.dbg)
        info1 = sim.launch_info()
        launches = info1["launches"] - info0["launches"] - 1          # minus the L2 flush fill
        b = sim.buffers()
        ctl = sim.memcpy_dtoh(b.momentary + 80, 16).view(np.uint64)    # Ctl::spilledTotal, voxelsTotal (construct.cu)
        spilled_total, voxels_total = int(ctl[0]), int(ctl[1])
        t_value = sdist.max_over_ranks(total_ms, dev)
        t_kernel = sdist.max_over_ranks(kernel_ms, dev)

        # ---- render on the built octree (rank-local; reported at rank 0) --------------------------
        render = None
        if not args.no_render:
            frames = []
            cams = [("autofocus+%d" % k, camera.autofocus(mx, 1920, 1080, yaw_offset=k * np.pi / 2)) for k in range(4)]
            cams += [("morro_bird", camera.orbit_camera(width=1920, height=1080, **camera.MORRO_BIRD)),
                     ("morro_close", camera.orbit_camera(width=1920, height=1080, **camera.MORRO_CLOSE))]
            t0 = time.time()
            for name, (view, proj) in cams:
                sim.set_camera(view, proj)
                sim.render()
                ms = min(sim.render() for _ in range(3))
                s = sim.stats()
                samples = s.numVisiblePoints + s.numVisibleVoxels
                frames.append({"camera": name, "ms": round(ms, 4), "visible_nodes": s.numVisibleNodes, "samples": samples,
                               "msamples_per_s": round(samples / ms / 1e3, 1) if ms > 0 else None})
            windows.append((t0, time.time()))
            tot_samples = sum(f["samples"] for f in frames)
            tot_ms = sum(f["ms"] for f in frames)
            render = {"metric": "render Msamples/s @1GPU (1920x1080, 64-bit atomicMin path)", "value": round(tot_samples / tot_ms / 1e3, 1),
                      "fps": round(1e3 * len(frames) / tot_ms, 1), "frames": frames}

        # ---- timed region 2: end to end from pinned host memory -------------------------------------
        sim.reset()
        sim.flush_l2()
        info2 = sim.launch_info()
        barrier()
        t0 = time.time()
        e_kernel_ms, e_total_ms = sim.insert_host_ptr(host_ptr, npts)
        st_e = sim.stats()                                  # device->host read of the result
        barrier()
        t1 = time.time()
        windows.append((t0, t1))
        assert st_e.numPointsProcessed == npts and st_e.numPoints == npts
        info3 = sim.launch_info()
        e_launches = info3["launches"] - info2["launches"]
        t_e2e = sdist.max_over_ranks(e_total_ms, dev)
        t_e2e_wall = sdist.max_over_ranks((t1 - t0) * 1e3, dev)
        totals = sdist.reduce_stats(st, dev)
    finally:
        sampler.stop()

    # ---- "next" row: LAS record decode, rank 0 at N = 1 ------------------------------
    las = None
    if rank == 0 and world == 1:
        try:
            las = bench_las(batches, mn, mx, local_rank)
        except Exception as e:          # the row is reported, never fatal for the headline
            las = {"error": repr(e)}

    stream = None
    if rank == 0 and world == 1:
        try:
            stream = bench_stream_file(batches, mn, mx, local_rank)
        except Exception as e:
            stream = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on a bounded sample -----------------------
    cpu = None
    if rank == 0 and world == 1:
        import oracle
        nb = min(args.cpu_sample_batches, K)
        o = oracle.Oracle(mn, mx)
        t0 = time.perf_counter()
        for bt in batches[:nb]:
            o.add_batch(bt)
        dt = time.perf_counter() - t0
        cpu = {"value": round(nb * BATCH / dt / 1e6, 4), "unit": "Mpoints/s", "cores": 1, "kind": "port",
               "sample": "first %d of the %d batches of the same stream, oracle/oracle.cpp, %.1f s" % (nb, K, dt)}

    if rank == 0:
        all_pts = world * npts
        value = all_pts / t_value / 1e3
        s_frac = spilled_total / npts
        v_frac = voxels_total / npts
        alg_bytes = (32.0 + 32.0 * s_frac + 16.0 * v_frac) * npts          # this rank, whole timed region
        peak, peak_src = peaks()
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        n_launch = max(launches, 1)
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mpoints/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(t_value / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32+f32", "data": "synthetic",
            "config": {"workload": "terrain_synth_%dM (Morro Bay 36M stand-in: 4800x4300x300 m fBm terrain in 50 m flight strips), "
                                   "%d x 1M-point batches per GPU streamed from a reset octree" % (K, K),
                       "batch_points": BATCH, "points_per_gpu": npts, "parallelism": "batch-sharded x%d (%s)" % (world, {"tiles": "one %d-batch scan tile per GPU" % K, "blocks": "contiguous blocks of one scan", "roundrobin": "round-robin over one scan"}[args.shard]),
                       "l2": "inputs %d MB > L2 (126 MB); L2 flushed before each timed region" % (npts * 16 // 1000000),
                       "kernel_only_mpoints_per_s": round(all_pts / t_kernel / 1e3, 2),
                       "octree": {k: totals[k] for k in ("numNodes", "numInner", "numLeaves", "numPoints", "numVoxels", "allocatedBytes_persistent")},
                       "construct_blocks": info1["construct_blocks"], "datagen_s": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "kernel": "kernel_construct", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5),
                         "traffic": round((traffic or {}).get("dram_bytes_per_batch", 0) * K / n_launch) if traffic else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_point": round(alg_bytes / npts, 3), "spilled_fraction_s": round(s_frac, 4),
                         "voxels_per_point_v": round(v_frac, 4), "launches": launches,
                         "avg_launch_ms": round(kernel_ms / n_launch, 4), "algorithmic_bytes_per_launch": round(alg_bytes / n_launch),
                         "note": "latency/atomic bound, not bandwidth bound: see DESIGN.md §7"},
            "e2e": {"value": round(all_pts / t_e2e / 1e3, 2), "unit": "Mpoints/s", "h2d_bytes_per_step": BATCH * 16,
                    "d2h_bytes_per_step": round(112.0 * e_launches / K, 1), "wall_clock_value": round(all_pts / t_e2e_wall / 1e3, 2),
                    "launches": e_launches},
            "gpu_launches": launches,
            "clocks": sampler.summary(windows),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if render:
            line["render"] = render
        if las:
            line["las_decode"] = las
        if stream:
            line["stream_file"] = stream
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
