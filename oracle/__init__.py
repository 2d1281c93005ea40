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
