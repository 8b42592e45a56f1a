"""ctypes wrapper over oracle/_ref/libclodref.so (the UNMODIFIED reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under basicrenderer_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libclodref.so")


class ClodConfig(C.Structure):
    """Mirror of struct clodConfig (clusterlod.h:15-71)."""

    _fields_ = [
        ("max_vertices", C.c_size_t),
        ("min_triangles", C.c_size_t),
        ("max_triangles", C.c_size_t),
        ("partition_spatial", C.c_bool),
        ("partition_sort", C.c_bool),
        ("partition_size", C.c_size_t),
        ("partition_max_refined_groups", C.c_size_t),
        ("partition_refined_split_count", C.c_void_p),
        ("cluster_spatial", C.c_bool),
        ("cluster_fill_weight", C.c_float),
        ("cluster_split_factor", C.c_float),
        ("simplify_ratio", C.c_float),
        ("simplify_threshold", C.c_float),
        ("simplify_error_merge_previous", C.c_float),
        ("simplify_error_merge_additive", C.c_float),
        ("simplify_error_factor_sloppy", C.c_float),
        ("simplify_error_edge_limit", C.c_float),
        ("simplify_permissive", C.c_bool),
        ("simplify_fallback_permissive", C.c_bool),
        ("simplify_fallback_sloppy", C.c_bool),
        ("simplify_regularize", C.c_bool),
        ("optimize_bounds", C.c_bool),
        ("optimize_clusters", C.c_bool),
    ]


_lib = None


def build(full: bool = False) -> None:
    """(Re)build the oracle libraries; requires /root/reference for libclodref*.so."""
    subprocess.check_call(["make", "-s", "-j8", "-C", _HERE] + (["all", "full"] if full else ["all"]))


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.clodref_builder_config.restype = ClodConfig
        _lib.clodref_config_size.restype = C.c_size_t
        assert _lib.clodref_config_size() == C.sizeof(ClodConfig)
        for name in ("clodref_dag_build", "clodref_dag_build_dump"):
            fn = getattr(_lib, name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.POINTER(ClodConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
        _lib.clodref_blob_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        _lib.clodref_free.argtypes = [C.c_void_p]
        _lib.clodref_cluster_count.argtypes = [C.c_void_p]
        _lib.clodref_cluster_count.restype = C.c_size_t
        _lib.meshopt_generatePositionRemap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        _lib.clodref_clusterize.restype = C.c_size_t
        _lib.clodref_clusterize.argtypes = [C.POINTER(ClodConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.clodref_simplify.restype = C.c_size_t
        _lib.clodref_simplify.argtypes = [C.POINTER(ClodConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_float)]
        _lib.clodref_dag_build_mt.restype = C.c_size_t
        _lib.clodref_dag_build_mt.argtypes = [C.POINTER(ClodConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
        _lib.clodref_dag_build_stats.restype = C.c_void_p
        _lib.clodref_dag_build_stats.argtypes = [C.POINTER(ClodConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
        _lib.meshopt_computeClusterBounds.restype = None  # struct return handled by wrapper below
    return _lib


def builder_config() -> ClodConfig:
    return lib().clodref_builder_config()


_BLOB_DTYPES = {
    "remap": np.uint32, "protect_locks": np.uint8, "cluster_index_offsets": np.uint32, "cluster_indices": np.uint32,
    "cluster_refined": np.int32, "cluster_vertices": np.uint32, "cluster_depth": np.int32, "cluster_bounds": np.float32,
    "num_levels": np.int32, "pending": np.int32, "group_offsets": np.uint32, "group_clusters": np.int32, "locks": np.uint8,
    "group_terminal": np.uint8, "group_bounds": np.float32, "group_error": np.float32, "group_ids": np.int32,
    "simp_offsets": np.uint32, "simp_indices": np.uint32, "merged_offsets": np.uint32, "merged_indices": np.uint32,
    "out.group_depth": np.int32, "out.group_simplified": np.float32, "out.group_cluster_offsets": np.uint32,
    "out.cluster_refined": np.int32, "out.cluster_bounds": np.float32, "out.cluster_vertex_count": np.uint32,
    "out.cluster_indices": np.uint32, "out.cluster_index_offsets": np.uint32,
    "stats.level_groups": np.uint32, "stats.level_clusters": np.uint32, "stats.level_triangles": np.uint32, "stats.level_sloppy": np.uint32,
    "stats.level_max_error": np.float32, "stats.group_error": np.float32, "stats.group_depth": np.int32, "stats.group_clusters": np.uint32,
}


class Dag:
    """Result of a reference DAG build: named numpy arrays copied out of the native blob store."""

    def __init__(self, handle):
        self._h = handle
        self.cluster_count = lib().clodref_cluster_count(handle)
        self._cache = {}

    def get(self, name: str) -> np.ndarray:
        if name in self._cache:
            return self._cache[name]
        base = name.split(".", 1)[1] if name.startswith("L") and "." in name and name[1].isdigit() else name
        dtype = _BLOB_DTYPES[base]
        ptr = C.c_void_p()
        size = C.c_size_t()
        lib().clodref_blob_get(self._h, name.encode(), C.byref(ptr), C.byref(size))
        if size.value == 0:
            arr = np.zeros(0, dtype=dtype)
        else:
            arr = np.frombuffer(C.string_at(ptr, size.value), dtype=dtype).copy()
        if base.endswith("bounds") or base == "out.group_simplified":
            arr = arr.reshape(-1, 5)
        self._cache[name] = arr
        return arr

    def level(self, depth: int, name: str) -> np.ndarray:
        return self.get(f"L{depth}.{name}")

    @property
    def num_levels(self) -> int:
        return int(self.get("num_levels")[0])

    def close(self):
        if self._h:
            lib().clodref_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dag_build(positions, indices, attributes=None, attribute_weights=None, protect_mask=0, config=None, dump=True, positions_stride=None, attributes_stride=None, vertex_count=None) -> Dag:
    """Run the reference DAG build. positions: float32 [V,3] (or a strided view base with explicit stride)."""
    cfg = config or builder_config()
    positions = np.ascontiguousarray(positions, dtype=np.float32) if positions_stride is None else positions
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    V = vertex_count if vertex_count is not None else positions.shape[0]
    pstride = positions_stride or 12
    acount = 0
    astride = 0
    if attributes is not None:
        attributes = np.ascontiguousarray(attributes, dtype=np.float32) if attributes_stride is None else attributes
        attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
        acount = attribute_weights.size
        astride = attributes_stride or attributes.shape[1] * 4
    fn = lib().clodref_dag_build_dump if dump else lib().clodref_dag_build
    h = fn(C.byref(cfg), _ptr(indices), indices.size, _ptr(positions), V, pstride, _ptr(attributes), astride, _ptr(attribute_weights), acount, protect_mask, None)
    return Dag(h)


def dag_build_timed(positions, indices, attributes=None, attribute_weights=None, protect_mask=0, threads=1, config=None) -> int:
    """The reference's clodBuildEx with its per-group iteration tasks spread over `threads` host threads; output is
    discarded (timing runs). Returns the total cluster count."""
    cfg = config or builder_config()
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    acount = astride = 0
    if attributes is not None:
        attributes = np.ascontiguousarray(attributes, dtype=np.float32)
        attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
        acount, astride = attribute_weights.size, attributes.shape[1] * 4
    return lib().clodref_dag_build_mt(C.byref(cfg), _ptr(indices), indices.size, _ptr(positions), positions.shape[0], 12, _ptr(attributes), astride, _ptr(attribute_weights), acount, protect_mask, threads)


def dag_build_stats(positions, indices, attributes=None, attribute_weights=None, protect_mask=0, threads=None, config=None) -> dict:
    """The reference's clodBuildEx (iteration tasks on `threads` host threads, default all) reduced to per-depth statistics of
    its callback stream: groups, clusters, triangles, max finite group error and sloppy-fallback calls per depth."""
    cfg = config or builder_config()
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    acount = astride = 0
    if attributes is not None:
        attributes = np.ascontiguousarray(attributes, dtype=np.float32)
        attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
        acount, astride = attribute_weights.size, attributes.shape[1] * 4
    threads = threads or os.cpu_count() or 1
    h = lib().clodref_dag_build_stats(C.byref(cfg), _ptr(indices), indices.size, _ptr(positions), positions.shape[0], 12, _ptr(attributes), astride, _ptr(attribute_weights), acount, protect_mask, threads)
    d = Dag(h)
    out = {k.split(".", 1)[1]: d.get(k) for k in ("stats.level_groups", "stats.level_clusters", "stats.level_triangles", "stats.level_sloppy", "stats.level_max_error", "stats.group_error", "stats.group_depth", "stats.group_clusters")}
    out["total_clusters"] = d.cluster_count
    d.close()
    return out


def position_remap(positions: np.ndarray) -> np.ndarray:
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    remap = np.empty(positions.shape[0], dtype=np.uint32)
    lib().meshopt_generatePositionRemap(_ptr(remap), _ptr(positions), positions.shape[0], 12)
    return remap


def clusterize(positions: np.ndarray, indices: np.ndarray, config=None):
    """clod::clusterize -> (cluster_index_offsets[K+1], cluster_vertex_counts[K], indices[index_count])."""
    cfg = config or builder_config()
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    bound = indices.size // 3 + 1
    counts = np.zeros(bound, dtype=np.uint32)
    vcounts = np.zeros(bound, dtype=np.uint32)
    out = np.zeros(indices.size, dtype=np.uint32)
    k = lib().clodref_clusterize(C.byref(cfg), _ptr(indices), indices.size, _ptr(positions), positions.shape[0], 12, _ptr(counts), _ptr(vcounts), _ptr(out))
    offsets = np.zeros(k + 1, dtype=np.uint32)
    np.cumsum(counts[:k], out=offsets[1:])
    return offsets, vcounts[:k].copy(), out


def simplify(positions, indices, locks, target_count, attributes=None, attribute_weights=None, config=None):
    """clod::simplify on one group's merged index list -> (simplified indices, error)."""
    cfg = config or builder_config()
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    locks = np.ascontiguousarray(locks, dtype=np.uint8)
    acount = astride = 0
    if attributes is not None:
        attributes = np.ascontiguousarray(attributes, dtype=np.float32)
        attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
        acount, astride = attribute_weights.size, attributes.shape[1] * 4
    out = np.zeros(indices.size, dtype=np.uint32)
    err = C.c_float(0)
    n = lib().clodref_simplify(C.byref(cfg), _ptr(indices), indices.size, _ptr(positions), positions.shape[0], 12, _ptr(attributes), astride, _ptr(attribute_weights), acount, _ptr(locks), target_count, _ptr(out), C.byref(err))
    return out[:n].copy(), err.value


def local_indices(indices: np.ndarray):
    """The reference's clodLocalIndices (clusterlod.h:972-1023) -> (vertices[unique], triangles u8)"""
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    vertices = np.zeros(max(1, indices.size), dtype=np.uint32)
    triangles = np.zeros(indices.size, dtype=np.uint8)
    fn = lib().clodLocalIndices
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    n = fn(_ptr(vertices), _ptr(triangles), _ptr(indices), indices.size)
    return vertices[:n].copy(), triangles
