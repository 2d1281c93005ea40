"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/liboracle.so (CPU restatement of the
reference, oracle.cpp) and around the canonicaliser. Imported only by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs; never by simlod_b200/."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_CUBINS = {0: os.path.join(REF_DIR, "ref_construct.cubin"), 1: os.path.join(REF_DIR, "ref_render.cubin"),
              2: os.path.join(REF_DIR, "ref_reset.cubin")}
REF_LAS_LIB = os.path.join(REF_DIR, "libref_las.so")
REF_SIMLOD_LIB = os.path.join(REF_DIR, "libref_simlod.so")
REF_MOMENTARY_BYTES = 420_000_000      # the reference carves 408 800 192 B out of its 300 MB buffer

POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("color", "<u4")])
RECORD_DTYPE = np.dtype([
    ("level", "<u4"), ("X", "<u4"), ("Y", "<u4"), ("Z", "<u4"), ("name", "S20"),
    ("counter", "<u4"), ("numPoints", "<u4"), ("numVoxels", "<u4"), ("numVoxelsStored", "<u4"),
    ("isLeaf", "<u4"), ("chunksPoints", "<u4"), ("chunksVoxels", "<u4"), ("nodeIndex", "<u4"), ("_pad", "<u4"),
    ("hashPoints", "<u8"), ("hashVoxelPos", "<u8")])
assert RECORD_DTYPE.itemsize == 88


class OStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("numNodes", "numInner", "numLeaves", "numNonemptyLeaves", "numPoints", "numVoxels",
                                           "numChunksPoints", "numChunksVoxels", "batchletIndex", "droppedSpilledPoints")] + \
               [(n, C.c_uint64) for n in ("numPointsProcessed", "numAllocatedChunks", "chunkPoolSize", "allocatedBytes_persistent")]


class RenderStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("numVisibleNodes", "numVisibleInner", "numVisibleLeaves", "numVisiblePoints", "numVisibleVoxels")]


# the Stats fields that are deterministic in the reference and therefore compared
STATS_FIELDS = ["numNodes", "numInner", "numLeaves", "numNonemptyLeaves", "numPoints", "numVoxels", "numChunksPoints",
                "numChunksVoxels", "batchletIndex", "numPointsProcessed", "numAllocatedChunks", "chunkPoolSize",
                "allocatedBytes_persistent"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_reset.argtypes = [vp]
        L.oracle_add_batch.argtypes = [vp, vp, C.c_uint32]
        L.oracle_get_stats.argtypes = [vp, C.POINTER(OStats)]
        L.canon_from_oracle.restype = vp
        L.canon_from_oracle.argtypes = [vp]
        L.canon_from_image.restype = vp
        L.canon_from_image.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64]
        L.canon_destroy.argtypes = [vp]
        L.canon_error.argtypes = [vp]
        L.canon_num_nodes.restype = C.c_uint32
        L.canon_num_nodes.argtypes = [vp]
        L.canon_records.argtypes = [vp, vp]
        L.canon_node_samples.restype = C.c_uint64
        L.canon_node_samples.argtypes = [vp, C.c_uint32, C.c_int, vp, C.c_uint64]
        L.oracle_check_voxel_colors.restype = C.c_int64
        L.oracle_check_voxel_colors.argtypes = [vp, vp]
        L.canon_render.argtypes = [vp, vp, vp, C.POINTER(RenderStats)]
        L.canon_flags.argtypes = [vp, vp]
        L.oracle_decode_las.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
        L.oracle_partition_cells.argtypes = [vp, C.c_uint64, vp, vp, C.c_float, C.c_uint32, vp]
        _lib = L
    return _lib


class Canon:
    """Canonical form of an octree: nodes keyed by (level, X, Y, Z), per-node counters, chunk counts,
    hashes of the sorted point multiset and of the sorted voxel positions."""

    def __init__(self, handle):
        self._h = handle
        err = lib().canon_error(handle)
        if err:
            raise RuntimeError("octree image is inconsistent (canon error %d)" % err)
        n = lib().canon_num_nodes(handle)
        self.records = np.zeros(n, dtype=RECORD_DTYPE)
        lib().canon_records(handle, self.records.ctypes.data)

    def samples(self, k, voxels=False):
        n = lib().canon_node_samples(self._h, k, 1 if voxels else 0, None, 0)
        out = np.empty(n, dtype=POINT_DTYPE)
        lib().canon_node_samples(self._h, k, 1 if voxels else 0, out.ctypes.data, n)
        return out

    def render(self, uniforms_bytes, width, height):
        fb = np.empty((height, width), dtype=np.uint64)
        rs = RenderStats()
        buf = C.create_string_buffer(uniforms_bytes, 480)
        lib().canon_render(self._h, buf, fb.ctypes.data, C.byref(rs))
        flags = np.zeros((len(self.records), 2), dtype=np.uint8)
        lib().canon_flags(self._h, flags.ctypes.data)
        return fb, rs, flags

    def __del__(self):
        try:
            lib().canon_destroy(self._h)
        except Exception:
            pass


def canon_from_image(nodes_bytes, heap_bytes, nodes_addr, heap_addr):
    nodes_bytes = np.ascontiguousarray(nodes_bytes, dtype=np.uint8)
    heap_bytes = np.ascontiguousarray(heap_bytes, dtype=np.uint8)
    h = lib().canon_from_image(nodes_bytes.ctypes.data, nodes_bytes.nbytes // 152, heap_bytes.ctypes.data, heap_bytes.nbytes,
                               nodes_addr, heap_addr)
    return Canon(h)


class Oracle:
    """Serial CPU octree builder with the reference's semantics."""

    def __init__(self, box_min, box_max, rcp_size=0.0):
        mn = (C.c_float * 3)(*[float(v) for v in box_min])
        mx = (C.c_float * 3)(*[float(v) for v in box_max])
        self._h = lib().oracle_create(mn, mx, float(rcp_size))

    def add_batch(self, points):
        pts = np.ascontiguousarray(points)
        assert pts.dtype.itemsize == 16 or pts.dtype == POINT_DTYPE
        lib().oracle_add_batch(self._h, pts.ctypes.data, pts.shape[0])

    def stats(self):
        s = OStats()
        lib().oracle_get_stats(self._h, C.byref(s))
        return s

    def canon(self):
        return Canon(lib().canon_from_oracle(self._h))

    def check_voxel_colors(self, canon):
        return lib().oracle_check_voxel_colors(self._h, canon._h)

    def __del__(self):
        try:
            lib().oracle_destroy(self._h)
        except Exception:
            pass


def compare_canon(a, b, what="octree"):
    """Field-by-field equality of two canonical forms; returns a list of human-readable differences."""
    diffs = []
    if len(a.records) != len(b.records):
        return ["%s: node count %d != %d" % (what, len(a.records), len(b.records))]
    for f in RECORD_DTYPE.names:
        if f in ("nodeIndex", "_pad"):
            continue
        neq = np.nonzero(a.records[f] != b.records[f])[0]
        if len(neq):
            k = int(neq[0])
            diffs.append("%s: %d nodes differ in %s (first: node %s: %r != %r)" % (
                what, len(neq), f, a.records["name"][k], a.records[f][k], b.records[f][k]))
    return diffs


def _level_cell(rec, level):
    """Level-`level` ancestor cell (child indices root first) of a canonical node record at level >= `level`."""
    sh = int(rec["level"]) - level
    x, y, z = int(rec["X"]) >> sh, int(rec["Y"]) >> sh, int(rec["Z"]) >> sh
    cell = 0
    for l in range(level):
        b = level - 1 - l
        cell = (cell << 3) | (((x >> b) & 1) << 2) | (((y >> b) & 1) << 1) | ((z >> b) & 1)
    return cell


def compare_merged(single, per_rank, level, owners):
    """The forest a spatial exchange builds (per_rank[r] = canonical octree of rank r, which owns the level-`level`
    cells c with owners[c] == r) against the single octree of the whole stream:
      * every node of the single octree at level >= `level` exists on the owner of its cell with an identical
        record (counters, chunk counts, hashes of the sorted points and voxel positions),
      * whatever else a rank holds at those levels is an empty leaf (the 7 siblings a split creates),
      * every node above `level` that is inner in the single octree: the ranks' voxel POSITION sets are disjoint
        and their union is the single node's set.
    Precondition (reported as a difference when violated): a node above `level` that is inner in the single octree
    is inner on every rank that holds points below it, i.e. every rank's share of it exceeds the leaf capacity —
    until then that rank still keeps the points in a leaf above `level`.
    Returns a list of differences."""
    diffs = []
    fields = [f for f in RECORD_DTYPE.names if f not in ("nodeIndex", "_pad")]
    key = lambda r: (int(r["level"]), int(r["X"]), int(r["Y"]), int(r["Z"]))
    maps = [{key(r): i for i, r in enumerate(c.records)} for c in per_rank]
    for i, rec in enumerate(single.records):
        k = key(rec)
        if k[0] >= level:
            own = int(owners[_level_cell(rec, level)])
            j = maps[own].get(k)
            if j is None:
                if int(rec["counter"]) or int(rec["numPoints"]):
                    diffs.append("node %s missing on its owner %d" % (rec["name"], own))
                continue
            other = per_rank[own].records[j]
            for f in fields:
                if rec[f] != other[f]:
                    diffs.append("node %s on rank %d differs in %s: %r != %r" % (rec["name"], own, f, other[f], rec[f]))
            for r, m in enumerate(maps):
                if r != own and k in m:
                    o = per_rank[r].records[m[k]]
                    if int(o["counter"]) or int(o["numPoints"]) or int(o["numVoxels"]) or not int(o["isLeaf"]):
                        diffs.append("rank %d holds samples in node %s owned by rank %d" % (r, rec["name"], own))
        elif not int(rec["isLeaf"]):
            want = np.unique(np.stack([single.samples(i, voxels=True)[a] for a in "xyz"], axis=1), axis=0)
            parts = []
            for r, m in enumerate(maps):
                if k in m and not int(per_rank[r].records[m[k]]["isLeaf"]):
                    v = per_rank[r].samples(m[k], voxels=True)
                    parts.append(np.unique(np.stack([v[a] for a in "xyz"], axis=1), axis=0))
                elif k in m and int(per_rank[r].records[m[k]]["numPoints"]):
                    diffs.append("precondition: node %s is inner in the single octree but still a leaf with %d points on rank %d"
                                 % (rec["name"], int(per_rank[r].records[m[k]]["numPoints"]), r))
            got = np.concatenate(parts) if parts else np.zeros((0, 3), np.float32)
            uniq = np.unique(got, axis=0)
            if len(uniq) != len(got):
                diffs.append("node %s: the ranks' voxel sets overlap" % rec["name"])
            if uniq.shape != want.shape or not np.array_equal(uniq, want):
                diffs.append("node %s: union of the ranks' voxels (%d) != single octree's (%d)" % (rec["name"], len(uniq), len(want)))
    # nothing but the single octree's nodes (or empty leaves) at level >= `level`
    have = {key(r) for r in single.records}
    for r, c in enumerate(per_rank):
        for rec in c.records:
            if int(rec["level"]) >= level and key(rec) not in have and (int(rec["counter"]) or int(rec["numPoints"])):
                diffs.append("rank %d has an extra non-empty node %s" % (r, rec["name"]))
    return diffs


def compare_stats(a, b, fields=STATS_FIELDS):
    return ["Stats.%s: %d != %d" % (f, getattr(a, f), getattr(b, f)) for f in fields if int(getattr(a, f)) != int(getattr(b, f))]


# ---- LAS decode -------------------------------------------------------------------
def decode_las(records, count, bytes_per_point, fmt, scale, offset, translation=(0.0, 0.0, 0.0)):
    """CPU restatement of the parse loop of loadLasNative (oracle.cpp:oracle_decode_las)."""
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    out = np.empty(count, dtype=POINT_DTYPE)
    sc, of, tr = (np.asarray(v, dtype=np.float64) for v in (scale, offset, translation))
    lib().oracle_decode_las(rec.ctypes.data, count, bytes_per_point, fmt, sc.ctypes.data, of.ctypes.data, tr.ctypes.data, out.ctypes.data)
    return out


# ---- spatial exchange ---------------------------------------------------------------
def partition_cells(points, box_min, box_max, level, rcp_size=None):
    """Level-`level` octree cell of every point (child indices root first), by the builder's quantisation."""
    pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    mn, mx = np.asarray(box_min, dtype=np.float32), np.asarray(box_max, dtype=np.float32)
    size = float(np.max(mx - mn))
    rcp = np.float32(1.0) / np.float32(size) if rcp_size is None else np.float32(rcp_size)
    out = np.empty(len(pts), dtype=np.uint32)
    lib().oracle_partition_cells(pts.ctypes.data, len(pts), mn.ctypes.data, mx.ctypes.data, C.c_float(float(rcp)), level, out.ctypes.data)
    return out


def partition_stable(points, box_min, box_max, level, owners, num_ranks, rcp_size=None):
    """[points owned by rank d, in input order] for d in range(num_ranks): what the scatter pass must produce."""
    pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    dst = np.asarray(owners, dtype=np.uint8)[partition_cells(pts, box_min, box_max, level, rcp_size)]
    return [pts[dst == d] for d in range(num_ranks)]


_ref_las = None


def ref_las():
    """The reference's own LasLoader.cpp compiled into oracle/_ref/libref_las.so (None if not built)."""
    global _ref_las
    if _ref_las is None and os.path.exists(REF_LAS_LIB):
        L = C.CDLL(REF_LAS_LIB)
        L.ref_las_load.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ref_las_header.argtypes = [C.c_char_p] + [C.c_void_p] * 8
        L.ref_las_bench.restype = C.c_double
        L.ref_las_bench.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        L.ref_las_load_parallel.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        _ref_las = L
    return _ref_las


def ref_las_load(path, first, count, translation=(0.0, 0.0, 0.0)):
    """loadLasNative of the reference (CPU): returns `count` 16-byte points."""
    out = np.zeros(count, dtype=POINT_DTYPE)
    tr = np.asarray(translation, dtype=np.float64)
    ref_las().ref_las_load(path.encode(), first, count, out.ctypes.data, tr.ctypes.data)
    return out


def ref_las_load_parallel(path, first, count, batch_points, threads, out=None, translation=(0.0, 0.0, 0.0)):
    """The reference loader on `threads` host threads, one loadLasNative call per batch (as spawnLoader does)."""
    out = np.zeros(count, dtype=POINT_DTYPE) if out is None else out
    tr = np.asarray(translation, dtype=np.float64)
    ref_las().ref_las_load_parallel(path.encode(), first, count, batch_points, out.ctypes.data, tr.ctypes.data, int(threads))
    return out


def ref_las_bench(path, count, batch_points, threads):
    """Seconds the reference loader needs for `count` points with `threads` long-lived loader threads."""
    out = np.zeros(count, dtype=POINT_DTYPE)
    return ref_las().ref_las_bench(path.encode(), count, batch_points, out.ctypes.data, int(threads))


# ---- .simlod loader ---------------------------------------------------------
_ref_simlod = None


def ref_simlod():
    """The reference's own SimlodLoader.cpp compiled into oracle/_ref/libref_simlod.so (None if not built)."""
    global _ref_simlod
    if _ref_simlod is None and os.path.exists(REF_SIMLOD_LIB):
        L = C.CDLL(REF_SIMLOD_LIB)
        L.ref_simlod_load.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ref_simlod_bench.restype = C.c_double
        L.ref_simlod_bench.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _ref_simlod = L
    return _ref_simlod


def ref_simlod_load(path, first_point, count):
    """loadFileNative of the reference: `count` points starting at point `first_point` of a .simlod file."""
    out = np.zeros(count, dtype=POINT_DTYPE)
    ref_simlod().ref_simlod_load(path.encode(), 24 + first_point * 16, count * 16, out.ctypes.data)
    return out


def ref_simlod_bench(path, count, batch_points, threads):
    out = np.zeros(count, dtype=POINT_DTYPE)
    return ref_simlod().ref_simlod_bench(path.encode(), count, batch_points, out.ctypes.data, int(threads))
