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


class MeshDesc(C.Structure):
    """struct clodb200_mesh == struct clodMesh (clusterlod.h:73-99)."""

    _fields_ = [
        ("indices", C.c_void_p),
        ("index_count", C.c_size_t),
        ("vertex_count", C.c_size_t),
        ("vertex_positions", C.c_void_p),
        ("vertex_positions_stride", C.c_size_t),
        ("vertex_attributes", C.c_void_p),
        ("vertex_attributes_stride", C.c_size_t),
        ("vertex_lock", C.c_void_p),
        ("attribute_weights", C.c_void_p),
        ("attribute_count", C.c_size_t),
        ("attribute_protect_mask", C.c_uint),
    ]


class Bounds(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("error", C.c_float)]


class Cluster(C.Structure):
    _fields_ = [("refined", C.c_int), ("bounds", Bounds), ("indices", C.POINTER(C.c_uint)), ("index_count", C.c_size_t), ("vertex_count", C.c_size_t)]


class Group(C.Structure):
    _fields_ = [("depth", C.c_int), ("simplified", Bounds)]


OUTPUT_EX = C.CFUNCTYPE(C.c_int, C.c_void_p, Group, C.POINTER(Cluster), C.c_size_t, C.c_size_t, C.c_uint)

_RECORD_DTYPES = {
    "group_depth": np.int32, "group_simplified": np.float32, "group_cluster_offsets": np.uint32, "cluster_refined": np.int32,
    "cluster_bounds": np.float32, "cluster_vertex_count": np.uint32, "cluster_index_offsets": np.uint64, "cluster_indices": np.uint32,
    "level_triangles": np.uint32, "level_clusters": np.uint32, "level_groups": np.uint32, "level_passes": np.uint32, "level_sloppy": np.uint32, "stats": np.uint64,
}


class DagRecord:
    """The complete output callback stream of one DAG build, as numpy arrays."""

    def __init__(self, arrays):
        self.__dict__.update(arrays)
        self.group_simplified = self.group_simplified.reshape(-1, 5)
        self.cluster_bounds = self.cluster_bounds.reshape(-1, 5)
        st = self.stats
        self.total_clusters, self.levels, self.groups = int(st[0]), int(st[1]), int(st[2])
        self.simplify_passes, self.simplify_rounds, self.launches = int(st[5]), int(st[6]), int(st[7])


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
        L.clodb200_generateMikkTangents.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
        L.clodb200_clusterize.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.clodb200_computeClusterBounds.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.clodb200_lockBoundary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.clodb200_simplifyGroups.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.clodb200_simplifyStats.argtypes = [C.c_void_p]
        L.clodb200_primExclusiveScanU32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.clodb200_primExclusiveMaxScanU64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float)]
        L.clodb200_primSortPairsU32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.clodb200_buildRecorded.restype = C.c_void_p
        L.clodb200_buildRecorded.argtypes = [Config, MeshDesc]
        L.clodb200_meshUpload.restype = C.c_void_p
        L.clodb200_meshUpload.argtypes = [MeshDesc]
        L.clodb200_meshFree.argtypes = [C.c_void_p]
        L.clodb200_meshBuildRecorded.restype = C.c_void_p
        L.clodb200_meshBuildRecorded.argtypes = [Config, C.c_void_p, C.c_int]
        L.clodb200_recordGet.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.clodb200_recordFree.argtypes = [C.c_void_p]
        L.clodb200_buildEx.restype = C.c_size_t
        L.clodb200_buildEx.argtypes = [Config, MeshDesc, C.c_void_p, OUTPUT_EX, C.c_void_p]
        from . import artifacts as _art

        _art.bind(L)
        self._check(L.clodb200_init(device))

    def init(self, device: int = 0):
        self._check(self._lib.clodb200_init(device))

    def shutdown(self):
        self._lib.clodb200_shutdown()

    def _check(self, status: int):
        if status != 0:
            raise ClodbError(self._lib.clodb200_last_error().decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.clodb200_launch_count())

    def builder_config(self) -> Config:
        return self._lib.clodb200_builderConfig()

    # ---- device-wide primitives (parity tests / micro-benchmarks); each returns (result..., mean ms of the timed runs)
    def prim_exclusive_scan_u32(self, values: np.ndarray, repeat: int = 1):
        values = np.ascontiguousarray(values, dtype=np.uint32)
        out = np.empty_like(values)
        total = np.zeros(1, np.uint32)
        ms = C.c_float(0)
        self._check(self._lib.clodb200_primExclusiveScanU32(_ptr(values), _ptr(out), values.size, _ptr(total), repeat, C.byref(ms)))
        return out, int(total[0]), ms.value

    def prim_set_scan_epoch(self, epoch: int):
        self._check(self._lib.clodb200_primSetScanEpoch(C.c_uint(epoch)))

    def prim_exclusive_max_scan_u64(self, values: np.ndarray, repeat: int = 1):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.empty_like(values)
        ms = C.c_float(0)
        self._check(self._lib.clodb200_primExclusiveMaxScanU64(_ptr(values), _ptr(out), values.size, repeat, C.byref(ms)))
        return out, ms.value

    def prim_sort_pairs_u32(self, keys: np.ndarray, values: np.ndarray, bit_lo: int = 0, bit_hi: int = 32, repeat: int = 1):
        keys = np.array(keys, dtype=np.uint32)
        values = np.array(values, dtype=np.uint32)
        ms = C.c_float(0)
        self._check(self._lib.clodb200_primSortPairsU32(_ptr(keys), _ptr(values), keys.size, bit_lo, bit_hi, repeat, C.byref(ms)))
        return keys, values, ms.value

    def prim_acosf(self, values: np.ndarray) -> np.ndarray:
        values = np.ascontiguousarray(values, np.float32)
        out = np.zeros_like(values)
        self._lib.clodb200_primAcosf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self._check(self._lib.clodb200_primAcosf(_ptr(values), _ptr(out), values.size))
        return out

    def position_remap(self, positions: np.ndarray, stride: int | None = None, vertex_count: int | None = None) -> np.ndarray:
        if stride is None:
            positions = np.ascontiguousarray(positions, dtype=np.float32)
            vertex_count, stride = positions.shape[0], positions.shape[1] * 4
        remap = np.empty(vertex_count, dtype=np.uint32)
        self._check(self._lib.clodb200_generatePositionRemap(_ptr(remap), _ptr(positions), vertex_count, stride))
        return remap

    def mikk_tangents(self, vertices: np.ndarray, indices: np.ndarray, corners: bool = False):
        """GenerateMikkTangents (ClusterLODUtilities.cpp:655-737) on the interleaved [V, stride/4] float stream (pos, normal, uv, ...).
        Returns [V, 4] {tangent xyz, sign}, or None where the reference's generator refuses the input."""
        vertices = np.ascontiguousarray(vertices, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32)
        out = np.zeros((vertices.shape[0], 4), np.float32)
        ok = C.c_int(0)
        per_corner = np.zeros((indices.size, 4), np.float32) if corners else None
        self._check(self._lib.clodb200_generateMikkTangents(_ptr(vertices), vertices.shape[0], vertices.shape[1] * 4 if vertices.ndim == 2 else 0, _ptr(indices), indices.size, _ptr(out), C.byref(ok),
                                                            _ptr(per_corner) if corners else None))
        if not ok.value:
            return (None, None) if corners else None
        return (out, per_corner) if corners else out

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


    def local_indices(self, indices: np.ndarray):
        """clodLocalIndices for one cluster -> (vertices[unique], triangles u8[len(indices)])"""
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        vertices = np.zeros(max(1, indices.size), dtype=np.uint32)
        triangles = np.zeros(indices.size, dtype=np.uint8)
        self._lib.clodb200_localIndices.restype = C.c_size_t
        n = self._lib.clodb200_localIndices(_ptr(vertices), _ptr(triangles), _ptr(indices), C.c_size_t(indices.size))
        if n == 0 and indices.size:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200_localIndices failed")
        return vertices[:n].copy(), triangles

    def local_indices_batch(self, indices: np.ndarray, cluster_index_offsets: np.ndarray, vertex_capacity: int = 128):
        """-> (vertices[K, vertex_capacity], triangles u8[len(indices)], vertex_counts[K])"""
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        offs = np.ascontiguousarray(cluster_index_offsets, dtype=np.uint64)
        K = offs.size - 1
        vertices = np.zeros((K, vertex_capacity), dtype=np.uint32)
        triangles = np.zeros(indices.size, dtype=np.uint8)
        counts = np.zeros(K, dtype=np.uint32)
        self._check(self._lib.clodb200_localIndicesBatch(_ptr(indices), _ptr(offs), C.c_size_t(K), C.c_size_t(vertex_capacity), _ptr(vertices), _ptr(triangles), _ptr(counts)))
        return vertices, triangles, counts

    def protect_bits(self, attributes, protect_mask: int, remap, locks=None) -> np.ndarray:
        attributes = np.ascontiguousarray(attributes, dtype=np.float32)
        remap = np.ascontiguousarray(remap, dtype=np.uint32)
        out = np.zeros(remap.size, np.uint8) if locks is None else np.ascontiguousarray(locks, dtype=np.uint8).copy()
        self._lib.clodb200_protectBits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t]
        self._check(self._lib.clodb200_protectBits(_ptr(out), _ptr(attributes), attributes.shape[1] * 4, protect_mask, _ptr(remap), remap.size))
        return out

    def partition_finish(self, cluster_part, partition_count, cluster_refined, cluster_bounds5, config: Config | None = None):
        """-> (group_clusters[K], group_offsets[G+1]) for a given partition id per cluster (clusterlod.h:396-507)."""
        cfg = config or self.builder_config()
        part = np.ascontiguousarray(cluster_part, dtype=np.uint32)
        refined = np.ascontiguousarray(cluster_refined, dtype=np.int32)
        bounds = np.ascontiguousarray(cluster_bounds5, dtype=np.float32)
        K = part.size
        clusters = np.zeros(K, np.uint32)
        offsets = np.zeros(K + 1, np.uint32)
        count = C.c_size_t(0)
        self._check(self._lib.clodb200_partitionFinish(C.byref(cfg), _ptr(part), C.c_size_t(int(partition_count)), C.c_size_t(K), _ptr(refined), _ptr(bounds), _ptr(clusters), _ptr(offsets), C.byref(count)))
        return clusters, offsets[: count.value + 1].copy()

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

    # ---- DAG build ------------------------------------------------------------------------------------------------
    @staticmethod
    def mesh_desc(positions, indices, attributes=None, attribute_weights=None, protect_mask=0, vertex_lock=None, positions_stride=None, attributes_stride=None, vertex_count=None):
        """Builds a clodb200_mesh over numpy arrays; returns (desc, keepalive)."""
        if positions_stride is None:
            positions = np.ascontiguousarray(positions, dtype=np.float32)
            vertex_count, positions_stride = positions.shape[0], positions.shape[1] * 4
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        d = MeshDesc()
        d.indices, d.index_count, d.vertex_count = _ptr(indices), indices.size, vertex_count
        d.vertex_positions, d.vertex_positions_stride = _ptr(positions), positions_stride
        keep = [positions, indices]
        if attributes is not None:
            if attributes_stride is None:
                attributes = np.ascontiguousarray(attributes, dtype=np.float32)
                attributes_stride = attributes.shape[1] * 4
            attribute_weights = np.ascontiguousarray(attribute_weights, dtype=np.float32)
            d.vertex_attributes, d.vertex_attributes_stride = _ptr(attributes), attributes_stride
            d.attribute_weights, d.attribute_count = _ptr(attribute_weights), attribute_weights.size
            d.attribute_protect_mask = protect_mask
            keep += [attributes, attribute_weights]
        if vertex_lock is not None:
            vertex_lock = np.ascontiguousarray(vertex_lock, dtype=np.uint8)
            d.vertex_lock = _ptr(vertex_lock)
            keep.append(vertex_lock)
        return d, keep

    def _record(self, handle, views: bool = False) -> DagRecord:
        """views=True returns zero-copy views into the library's (recycled) record buffers: valid only until the next
        build call on this process (what a C caller of the callback ABI sees); the default copies."""
        if not handle:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200 build failed")
        arrays = {}
        for name, dtype in _RECORD_DTYPES.items():
            ptr, size = C.c_void_p(), C.c_size_t()
            self._lib.clodb200_recordGet(handle, name.encode(), C.byref(ptr), C.byref(size))
            if not size.value:
                arrays[name] = np.zeros(0, dtype)
                continue
            view = np.frombuffer((C.c_ubyte * size.value).from_address(ptr.value), dtype=dtype)
            arrays[name] = view if views else view.copy()
        self._lib.clodb200_recordFree(handle)
        return DagRecord(arrays)

    def build_dag(self, positions, indices, attributes=None, attribute_weights=None, protect_mask=0, config: Config | None = None, **kw) -> DagRecord:
        """clodBuildEx-equivalent on host arrays (upload + build + read-back), recording the callback stream."""
        views = bool(kw.pop("views", False))
        desc, keep = self.mesh_desc(positions, indices, attributes, attribute_weights, protect_mask, **kw)
        return self._record(self._lib.clodb200_buildRecorded(config or self.builder_config(), desc), views=views)

    def upload_mesh(self, positions, indices, attributes=None, attribute_weights=None, protect_mask=0, **kw):
        desc, keep = self.mesh_desc(positions, indices, attributes, attribute_weights, protect_mask, **kw)
        h = self._lib.clodb200_meshUpload(desc)
        if not h:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200 mesh upload failed")
        return h

    def free_mesh(self, handle):
        self._lib.clodb200_meshFree(handle)

    def build_dag_resident(self, mesh_handle, config: Config | None = None, keep_indices: bool = True) -> DagRecord:
        return self._record(self._lib.clodb200_meshBuildRecorded(config or self.builder_config(), mesh_handle, 1 if keep_indices else 0))

    def build_ex(self, positions, indices, callback, attributes=None, attribute_weights=None, protect_mask=0, config: Config | None = None) -> int:
        """clodBuildEx with a Python callback(group: Group, clusters: [Cluster], task_index) -> refined id."""
        desc, keep = self.mesh_desc(positions, indices, attributes, attribute_weights, protect_mask)

        def tramp(ctx, group, clusters, count, task_index, thread_index):
            return int(callback(group, [clusters[i] for i in range(count)], task_index))

        cb = OUTPUT_EX(tramp)
        n = self._lib.clodb200_buildEx(config or self.builder_config(), desc, None, cb, None)
        err = self._lib.clodb200_last_error().decode()
        if n == 0 and err:
            raise ClodbError(err)
        return n

    # ---- outer boundary: BuildClusterLODArtifactsFromGeometry (ClusterLODUtilities.h:5-13) ------------------------------
    def default_builder_settings(self):
        return self._lib.clodb200_defaultBuilderSettings()

    def _artifacts(self, handle, views: bool = False, keep_handle: bool = False):
        from . import artifacts as _art

        if not handle:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200 artifact build failed")
        a = _art.collect(self._lib, handle, views=views)
        if keep_handle:
            a.handle = handle
        else:
            self._lib.clodb200_artifactsFree(handle)
        return a

    def build_artifacts(self, vertices, indices, flags, uv_sets=None, tangents=None, settings=None, views: bool = False, keep_handle: bool = False, skinning=None):
        """Host arrays in, ClusterLODPrebuildArtifacts out (upload + DAG build + page encode + read-back)."""
        from . import artifacts as _art

        g, keep = _art.make_geometry(vertices, indices, flags, uv_sets, tangents, skinning)
        st = settings or self.default_builder_settings()
        return self._artifacts(self._lib.clodb200_buildArtifacts(C.byref(g), C.byref(st)), views=views, keep_handle=keep_handle)

    def upload_geometry(self, vertices, indices, flags, uv_sets=None, tangents=None, settings=None, skinning=None):
        from . import artifacts as _art

        g, keep = _art.make_geometry(vertices, indices, flags, uv_sets, tangents, skinning)
        st = settings or self.default_builder_settings()
        h = self._lib.clodb200_geometryUpload(C.byref(g), C.byref(st))
        if not h:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200 geometry upload failed")
        return h

    def free_geometry(self, handle):
        self._lib.clodb200_geometryFree(handle)

    def build_artifacts_resident(self, geometry_handle, views: bool = False, keep_handle: bool = False):
        return self._artifacts(self._lib.clodb200_geometryBuildArtifacts(geometry_handle), views=views, keep_handle=keep_handle)

    def free_artifacts(self, artifacts):
        self._lib.clodb200_artifactsFree(artifacts.handle)
        artifacts.handle = None

    def save_cache(self, artifacts, directory: str, container_file_name: str, metadata_file_name: str, source_identifier: str = "", prim_path: str = "", subset_name: str = "",
                   build_config_hash: int = 0):
        """CLodCache::Save without the .usdc wrapper: <directory>/<container> (.clodbin) + <directory>/<metadata> (clodBlob bytes)."""
        self._check(self._lib.clodb200_artifactsSaveCache(artifacts.handle, directory.encode(), container_file_name.encode(), metadata_file_name.encode(),
                                                          source_identifier.encode(), prim_path.encode(), subset_name.encode(), build_config_hash))

    def serialize_metadata(self, artifacts, container_file_name: str, source_identifier: str = "", prim_path: str = "", subset_name: str = "", build_config_hash: int = 0) -> bytes:
        args = (artifacts.handle, container_file_name.encode(), source_identifier.encode(), prim_path.encode(), subset_name.encode(), build_config_hash)
        n = self._lib.clodb200_artifactsSerializeMetadata(*args, None, 0)
        buf = C.create_string_buffer(n)
        self._lib.clodb200_artifactsSerializeMetadata(*args, buf, n)
        return buf.raw

    # ---- scene batches: the library's own NCCL gather of the metadata blobs (csrc/comm.cu) -------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self._lib.clodb200_commGetUniqueId(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes | None, world_size: int, rank: int):
        self._lib.clodb200_commInit.argtypes = [C.c_char_p, C.c_int, C.c_int]
        self._check(self._lib.clodb200_commInit(unique_id, world_size, rank))

    def comm_destroy(self):
        self._lib.clodb200_commDestroy()

    def gather_begin(self, payload: bytes):
        """Starts the all-gather of this rank's payload on the library's communication stream; returns a handle."""
        self._lib.clodb200_commGatherBegin.restype = C.c_void_p
        self._lib.clodb200_commGatherBegin.argtypes = [C.c_char_p, C.c_size_t]
        h = self._lib.clodb200_commGatherBegin(payload, len(payload))
        if not h:
            raise ClodbError(self._lib.clodb200_last_error().decode() or "clodb200 gather failed")
        return h

    def gather_end(self, handle):
        """Waits for a gather and returns the payloads of all ranks, by rank."""
        L = self._lib
        L.clodb200_commGatherWait.argtypes = [C.c_void_p]
        L.clodb200_commGatherGet.restype = C.c_void_p
        L.clodb200_commGatherGet.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.clodb200_commGatherFree.argtypes = [C.c_void_p]
        L.clodb200_commWorldSize.restype = C.c_int
        try:
            self._check(L.clodb200_commGatherWait(handle))
            out = []
            for r in range(L.clodb200_commWorldSize()):
                n = C.c_size_t()
                p = L.clodb200_commGatherGet(handle, r, C.byref(n))
                out.append(C.string_at(p, n.value) if n.value else b"")
            return out
        finally:
            L.clodb200_commGatherFree(handle)

    def cache_probe(self, directory: str, metadata_file_name: str, source_identifier: str = "", prim_path: str = "", subset_name: str = "", build_config_hash: int = 0) -> bool:
        self._lib.clodb200_cacheProbe.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
        return bool(self._lib.clodb200_cacheProbe(directory.encode(), metadata_file_name.encode(), source_identifier.encode(), prim_path.encode(), subset_name.encode(), build_config_hash))

    def timer_start(self):
        self._lib.clodb200_timerStart()

    def timer_stop_ms(self) -> float:
        self._lib.clodb200_timerStop.restype = C.c_float
        return float(self._lib.clodb200_timerStop())

    def profile_enable(self, enable: bool):
        self._lib.clodb200_profileEnable(1 if enable else 0)

    def profile_report(self):
        """-> list of (kernel, launches, total_ms, total_threads), slowest first"""
        self._lib.clodb200_profileReport.restype = C.c_size_t
        self._lib.clodb200_profileReport.argtypes = [C.c_char_p, C.c_size_t]
        buf = C.create_string_buffer(1 << 20)
        self._lib.clodb200_profileReport(buf, len(buf))
        rows = []
        for line in buf.value.decode().splitlines():
            name, count, ms, threads = line.rsplit(",", 3)
            rows.append((name, int(count), float(ms), int(threads)))
        return rows

    def simplify_stats(self):
        a = (C.c_uint * 3)()
        self._lib.clodb200_simplifyStats(a)
        return {"passes": a[0], "rounds": a[1], "max_rounds": a[2]}


_product = None


def load(device: int = 0) -> ClodLib:
    global _product
    if _product is None:
        _product = ClodLib(os.environ.get("CLODB200_LIB", PRODUCT_LIB), device)  # CLODB200_LIB: tuning variants (tools/)
    return _product
