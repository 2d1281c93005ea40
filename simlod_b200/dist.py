"""One process per GPU: sharding of the batch stream and the reductions around it.

Insertion shards by independent point batches: rank g of G owns a complete builder
instance over the same global cube and inserts batches b with b % G == g, in order. There is no
exchange step on the data path — the result is a forest of G octrees, each equal to what the
reference builds from that rank's batch sequence — so the only collectives are
  * all_reduce(MIN/MAX) of the 6 bounding-box floats when the cube is not known a priori,
  * all_reduce(SUM) of the Stats counters for reporting,
  * all_reduce(MAX) of the per-rank device time (throughput = total points / slowest rank),
  * optionally all_reduce(MIN) over the packed depth|colour framebuffers: u64 min is exactly the
    depth test the single-GPU atomicMin performs, so compositing is associative and exact.
torch.distributed (NCCL on the GPUs, gloo in the CPU tests) is plumbing only.
"""
import numpy as np

SUMMED_STATS = ["numNodes", "numInner", "numLeaves", "numNonemptyLeaves", "numPoints", "numVoxels", "numChunksPoints",
                "numChunksVoxels", "numPointsProcessed", "numAllocatedChunks", "chunkPoolSize", "allocatedBytes_persistent",
                "batchletIndex"]


def shard_batches(num_batches, rank, world_size):
    """Indices of the batches rank `rank` inserts, in insertion order (round-robin: b % G == g)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_batches, world_size))


def shard_point_range(num_points, batch_size, rank, world_size):
    """[(first, count)] point ranges of this rank's batches for a stream of num_points points."""
    num_batches = -(-num_points // batch_size)
    return [(b * batch_size, min(batch_size, num_points - b * batch_size)) for b in shard_batches(num_batches, rank, world_size)]


def _dist():
    import torch.distributed as dist
    return dist


def global_box(local_min, local_max, device="cpu"):
    """Bounding box over all ranks (two 3-float all-reduces)."""
    import torch
    dist = _dist()
    mn = torch.tensor(local_min, dtype=torch.float32, device=device)
    mx = torch.tensor(local_max, dtype=torch.float32, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return mn.cpu().numpy(), mx.cpu().numpy()


def reduce_stats(stats, device="cpu"):
    """Sum the additive Stats counters over ranks; returns a dict (same on every rank)."""
    import torch
    dist = _dist()
    vals = torch.tensor([int(getattr(stats, f)) for f in SUMMED_STATS], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
    return dict(zip(SUMMED_STATS, [int(v) for v in vals.cpu()]))


def max_over_ranks(value, device="cpu"):
    import torch
    dist = _dist()
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def composite_framebuffers(fb, device="cpu"):
    """Depth-composite per-rank packed framebuffers: element-wise unsigned 64-bit minimum.

    all_reduce has no uint64 MIN; the words are order-preservingly mapped to int64 (flip the top
    bit), reduced with MIN, and mapped back."""
    import torch
    dist = _dist()
    a = np.ascontiguousarray(fb, dtype=np.uint64)
    signed = (a ^ np.uint64(1 << 63)).view(np.int64)
    t = torch.from_numpy(signed.copy()).to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out = t.cpu().numpy().view(np.uint64) ^ np.uint64(1 << 63)
    return out.reshape(a.shape)
