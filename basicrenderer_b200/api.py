"""ctypes binding of the clodb200 C ABI (include/clodb200.h).

`load()` returns the product library (basicrenderer_b200/libclodb200.so, CUDA sm_100a). It raises if the library is
missing or no CUDA device can be initialised: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libclodb200.so")


class ClodbError(RuntimeError):
    pass


class Config(C.Structure):
    """struct clodb200_config == struct clodConfig (clusterlod.h:15-71)."""

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


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ClodLib:
    """One loaded clodb200 library + initialised device."""

    def __init__(self, path: str, device: int = 0):
        if not os.path.exists(path):
            raise ClodbError(f"clodb200 native library not found at {path}; run `python -m basicrenderer_b200.build product`")
        self.path = path
        self._lib = C.CDLL(path)
        L = self._lib
        L.clodb200_last_error.restype = C.c_char_p
        L.clodb200_init.argtypes = [C.c_int]
        L.clodb200_launch_count.restype = C.c_uint64
        L.clodb200_builderConfig.restype = Config
        L.clodb200_generatePositionRemap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.clodb200_clusterize.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.clodb200_computeClusterBounds.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.clodb200_lockBoundary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.clodb200_simplifyGroups.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.clodb200_simplifyStats.argtypes = [C.c_void_p]
        self._check(L.clodb200_init(device))

    def _check(self, status: int):
        if status != 0:
            raise ClodbError(self._lib.clodb200_last_error().decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.clodb200_launch_count())

    def builder_config(self) -> Config:
        return self._lib.clodb200_builderConfig()

    def position_remap(self, positions: np.ndarray, stride: int | None = None, vertex_count: int | None = None) -> np.ndarray:
        if stride is None:
            positions = np.ascontiguousarray(positions, dtype=np.float32)
            vertex_count, stride = positions.shape[0], positions.shape[1] * 4
        remap = np.empty(vertex_count, dtype=np.uint32)
        self._check(self._lib.clodb200_generatePositionRemap(_ptr(remap), _ptr(positions), vertex_count, stride))
        return remap

    def clusterize(self, positions: np.ndarray, indices: np.ndarray, segment_offsets=None, config: Config | None = None):
        """-> (cluster_index_offsets[K+1], cluster_vertex_counts[K], cluster_segments[K], indices[index_count])"""
        cfg = config or self.builder_config()
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        cap = max(1, indices.size // 3)
        counts = np.zeros(cap, dtype=np.uint32)
        vcounts = np.zeros(cap, dtype=np.uint32)
        segs = np.zeros(cap, dtype=np.uint32)
        out = np.zeros(indices.size, dtype=np.uint32)
        k = C.c_size_t(0)
        so = None if segment_offsets is None else np.ascontiguousarray(segment_offsets, dtype=np.uint32)
        self._check(self._lib.clodb200_clusterize(C.byref(cfg), _ptr(indices), indices.size, _ptr(so), 0 if so is None else so.size - 1, _ptr(positions), positions.shape[0], positions.shape[1] * 4, _ptr(counts), _ptr(vcounts), _ptr(segs), _ptr(out), C.byref(k)))
        K = k.value
        offsets = np.zeros(K + 1, dtype=np.uint32)
        np.cumsum(counts[:K], out=offsets[1:])
        return offsets, vcounts[:K].copy(), segs[:K].copy(), out

    def cluster_bounds(self, positions: np.ndarray, indices: np.ndarray, cluster_index_counts: np.ndarray) -> np.ndarray:
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        counts = np.ascontiguousarray(cluster_index_counts, dtype=np.uint32)
        out = np.zeros((counts.size, 4), dtype=np.float32)
        self._check(self._lib.clodb200_computeClusterBounds(_ptr(indices), _ptr(counts), counts.size, _ptr(positions), positions.shape[0], positions.shape[1] * 4, _ptr(out)))
        return out


    def lock_boundary(self, locks: np.ndarray, indices: np.ndarray, group_index_offsets: np.ndarray, remap: np.ndarray, vertex_lock=None) -> np.ndarray:
        locks = np.ascontiguousarray(locks, dtype=np.uint8).copy()
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        offs = np.ascontiguousarray(group_index_offsets, dtype=np.uint32)
        remap = np.ascontiguousarray(remap, dtype=np.uint32)
        vl = None if vertex_lock is None else np.ascontiguousarray(vertex_lock, dtype=np.uint8)
        self._check(self._lib.clodb200_lockBoundary(_ptr(locks), _ptr(indices), _ptr(offs), offs.size - 1, _ptr(remap), _ptr(vl), locks.size))
        return locks

    def simplify_groups(self, positions, indices, group_index_offsets, locks, attributes=None, attribute_weights=None, config: Config | None = None):
        """-> (indices, group_index_offsets[G+1], group_errors[G])"""
        cfg = config or self.builder_config()
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        offs = np.ascontiguousarray(group_index_offsets, dtype=np.uint32)
        locks = None if locks is None else np.ascontiguousarray(locks, dtype=np.uint8)
        G = offs.size - 1
        acount = astride = 0
        if attributes is not None:
            attributes = np.ascontiguousarray(attributes, dtype=np.float32)
            attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
            acount, astride = attribute_weights.size, attributes.shape[1] * 4
        out = np.zeros(max(1, indices.size), dtype=np.uint32)
        counts = np.zeros(G, dtype=np.uint32)
        errors = np.zeros(G, dtype=np.float32)
        self._check(self._lib.clodb200_simplifyGroups(C.byref(cfg), _ptr(indices), _ptr(offs), G, _ptr(positions), positions.shape[0], positions.shape[1] * 4, _ptr(attributes), astride, _ptr(attribute_weights), acount, _ptr(locks), _ptr(out), _ptr(counts), _ptr(errors)))
        out_offs = np.zeros(G + 1, dtype=np.uint32)
        np.cumsum(counts, out=out_offs[1:])
        return out[: out_offs[-1]].copy(), out_offs, errors

    def simplify_stats(self):
        a = (C.c_uint * 3)()
        self._lib.clodb200_simplifyStats(a)
        return {"passes": a[0], "rounds": a[1], "max_rounds": a[2]}


_product = None


def load(device: int = 0) -> ClodLib:
    global _product
    if _product is None:
        _product = ClodLib(PRODUCT_LIB, device)
    return _product
