"""ctypes wrapper over oracle/_ref/libclodref_full*.so: the reference's L3 builder BuildClusterLODArtifactsFromGeometry
(ClusterLODUtilities.cpp:5325) compiled unmodified by oracle/Makefile.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU legs).

  build(mesh_vertices, indices, flags)                          reference builder + reference clodBuildEx
  build(..., clodb200_lib="/path/to/libclodb200*.so")           reference builder + OUR clodBuildEx (INTEGRATION.md wiring)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VERTEX_COLORS, VERTEX_NORMALS, VERTEX_TEXCOORDS = 1, 2, 4

GROUP_DTYPE = np.dtype([
    ("bounds", np.float32, 5), ("firstMeshlet", np.uint32), ("meshletCount", np.uint32), ("depth", np.int32),
    ("firstGroupVertex", np.uint32), ("groupVertexCount", np.uint32), ("firstSegment", np.uint32), ("segmentCount", np.uint32),
    ("terminalSegmentCount", np.uint32), ("flags", np.uint32), ("pageMapBase", np.uint32), ("pageCount", np.uint32),
    ("parentGroupId", np.int32), ("maxParentError", np.float32), ("representationError", np.float32)])
SEGMENT_DTYPE = np.dtype([("refinedGroup", np.int32), ("firstMeshletInPage", np.uint32), ("meshletCount", np.uint32), ("pageIndex", np.uint32)])
LOCATOR_DTYPE = np.dtype([("blobOffset", np.uint64), ("blobSizeBytes", np.uint32), ("reserved", np.uint32)])
NODE_DTYPE = np.dtype([("isGroup", np.uint32), ("indexOrOffset", np.uint32), ("countMinusOne", np.uint32), ("ownerGroupId", np.uint32),
                       ("cullingSphere", np.float32, 4), ("lodBoundingSphere", np.float32, 4), ("maxQuadricError", np.float32), ("padding", np.float32, 3)])
CHUNK_DTYPE = np.dtype([("groupVertexCount", np.uint32), ("meshletCount", np.uint32), ("meshletTrianglesByteCount", np.uint32), ("compressedPositionQuantExp", np.uint32), ("compressedFlags", np.uint32)])
RANGE_DTYPE = np.dtype([("offset", np.uint32), ("count", np.uint32)])

_DTYPES = {
    "groups": GROUP_DTYPE, "segments": SEGMENT_DTYPE, "segmentBounds": np.dtype((np.float32, 4)), "groupChunks": CHUNK_DTYPE,
    "groupDiskLocators": LOCATOR_DTYPE, "pageDiskLocators": LOCATOR_DTYPE, "groupPageReferences": np.uint32, "groupPageReferenceOffsets": np.uint32,
    "nodes": NODE_DTYPE, "lodNodeRanges": RANGE_DTYPE, "lodLevelRoots": np.uint32, "objectBoundingSphere": np.float32, "counts": np.uint32,
    "meshPages": np.uint8, "meshPageOffsets": np.uint64,
}

_libs = {}


def available(ours: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libclodref_full_ours.so" if ours else "libclodref_full.so"))


def _lib(ours: bool):
    if ours not in _libs:
        lib = C.CDLL(os.path.join(_HERE, "_ref", "libclodref_full_ours.so" if ours else "libclodref_full.so"))
        lib.clodfull_build.restype = C.c_void_p
        lib.clodfull_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
        lib.clodfull_build_skinned.restype = C.c_void_p
        lib.clodfull_build_skinned.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_void_p, C.c_size_t, C.c_uint]
        lib.clodfull_error.restype = C.c_char_p
        lib.clodfull_error.argtypes = [C.c_void_p]
        lib.clodfull_seconds.restype = C.c_double
        lib.clodfull_seconds.argtypes = [C.c_void_p]
        lib.clodfull_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.clodfull_free.argtypes = [C.c_void_p]
        sizes = (C.c_uint * 8)()
        lib.clodfull_sizes(sizes)
        assert list(sizes) == [GROUP_DTYPE.itemsize, SEGMENT_DTYPE.itemsize, 16, CHUNK_DTYPE.itemsize, LOCATOR_DTYPE.itemsize, NODE_DTYPE.itemsize, RANGE_DTYPE.itemsize, 20], list(sizes)
        _libs[ours] = lib
    return _libs[ours]


def last_attributes() -> np.ndarray:
    """[V, A] simplification attribute stream the reference builder handed to clodBuildEx in the last build(..., clodb200_lib=...)
    call (normals, then its MikkTSpace tangents xyz + sign when the mesh has UVs)."""
    lib = _lib(True)
    lib.clodshim_last_attributes.restype = C.c_void_p
    lib.clodshim_last_attributes.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    v, a = C.c_size_t(), C.c_size_t()
    ptr = lib.clodshim_last_attributes(C.byref(v), C.byref(a))
    if not v.value or not a.value:
        return np.zeros((0, 0), np.float32)
    return np.frombuffer((C.c_float * (v.value * a.value)).from_address(ptr), np.float32).reshape(v.value, a.value).copy()


_mikk = None


def mikk_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libclodref_mikk.so"))


def mikk_tangents(vertices: np.ndarray, indices: np.ndarray, with_seconds: bool = False):
    """The reference's GenerateMikkTangents (ClusterLODUtilities.cpp:655-737, compiled unmodified into libclodref_mikk.so by
    oracle/ref_mikk_driver.cpp). Returns [V, 4] or None when the reference's generator returns false."""
    global _mikk
    if _mikk is None:
        _mikk = C.CDLL(os.path.join(_HERE, "_ref", "libclodref_mikk.so"))
        _mikk.clodref_mikk_tangents.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_double)]
    vertices = np.ascontiguousarray(vertices, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32)
    out = np.zeros((vertices.shape[0], 4), np.float32)
    sec = C.c_double(0)
    ok = _mikk.clodref_mikk_tangents(vertices.ctypes.data_as(C.c_void_p), vertices.shape[0], vertices.shape[1] * 4, indices.ctypes.data_as(C.c_void_p), indices.size,
                                     out.ctypes.data_as(C.c_void_p), C.byref(sec))
    res = out if ok else None
    return (res, sec.value) if with_seconds else res


class Artifacts:
    def __init__(self, arrays, seconds):
        self.__dict__.update(arrays)
        self.seconds = seconds

    def page(self, i: int) -> np.ndarray:
        return self.meshPages[int(self.meshPageOffsets[i]) : int(self.meshPageOffsets[i + 1])]


def interleave(positions, normals, uvs=None) -> np.ndarray:
    """VertexLayout.h: pos f32x3 @0, normal f32x3 @12[, uv f32x2 @24]."""
    cols = [np.asarray(positions, np.float32), np.asarray(normals, np.float32)]
    if uvs is not None:
        cols.append(np.asarray(uvs, np.float32))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def build(vertices: np.ndarray, indices: np.ndarray, flags: int = VERTEX_NORMALS, threads: int = 1, clodb200_lib: str | None = None, skinning: np.ndarray | None = None,
          recompute_normals: bool = False) -> Artifacts:
    ours = clodb200_lib is not None
    if ours:
        os.environ["CLODB200_LIB"] = clodb200_lib
    lib = _lib(ours)
    vertices = np.ascontiguousarray(vertices, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32)
    lib.clodfull_set_options(1 if recompute_normals else 0)
    if skinning is not None:
        skinning = np.ascontiguousarray(skinning, np.uint8)
        h = lib.clodfull_build_skinned(vertices.ctypes.data_as(C.c_void_p), vertices.shape[0], vertices.shape[1] * 4, indices.ctypes.data_as(C.c_void_p), indices.size, flags, threads,
                                       skinning.ctypes.data_as(C.c_void_p), skinning.size, skinning.shape[1])
    else:
        h = lib.clodfull_build(vertices.ctypes.data_as(C.c_void_p), vertices.shape[0], vertices.shape[1] * 4, indices.ctypes.data_as(C.c_void_p), indices.size, flags, threads)
    try:
        err = lib.clodfull_error(h).decode()
        if err:
            raise RuntimeError("reference builder failed: " + err)
        arrays = {}
        for name, dtype in _DTYPES.items():
            ptr, size = C.c_void_p(), C.c_size_t()
            lib.clodfull_get(h, name.encode(), C.byref(ptr), C.byref(size))
            arrays[name] = np.frombuffer((C.c_ubyte * size.value).from_address(ptr.value), dtype=dtype).copy() if size.value else np.zeros(0, dtype)
        return Artifacts(arrays, lib.clodfull_seconds(h))
    finally:
        lib.clodfull_free(h)
