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


_product = None


def load(device: int = 0) -> ClodLib:
    global _product
    if _product is None:
        _product = ClodLib(PRODUCT_LIB, device)
    return _product
