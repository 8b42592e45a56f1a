"""Host-side mirror of the reference's outer builder boundary (BuildClusterLODArtifactsFromGeometry,
BasicRenderer/include/Mesh/ClusterLODUtilities.h:5-13) over the clodb200 C ABI: same inputs (interleaved vertex stream per
Mesh/VertexLayout.h, u32 indices, UV sets, VertexFlags, builder settings), same outputs (ClusterLODPrebuiltData arrays +
mesh page blobs) as numpy views of the reference's PODs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

VERTEX_COLORS, VERTEX_NORMALS, VERTEX_TEXCOORDS, VERTEX_SKINNED = 1, 2, 4, 8

# reference PODs (ClusterLODShaderTypes.h:97-139, ClusterLODTypes.h:36-77)
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
PAGE_HEADER_DTYPE = np.dtype([(n, np.uint32) for n in (
    "meshletCount", "compressedPositionQuantExp", "attributeMask", "uvSetCount", "descriptorOffset", "uvDescriptorOffset", "positionBitstreamOffset",
    "normalArrayOffset", "colorArrayOffset", "jointArrayOffset", "weightArrayOffset", "uvBitstreamDirectoryOffset", "triangleStreamOffset",
    "boneIndexStreamOffset", "reserved0", "reserved1")])
MESHLET_DESCRIPTOR_DTYPE = np.dtype([
    ("positionBitOffset", np.uint32), ("vertexAttributeOffset", np.uint32), ("triangleByteOffset", np.uint32), ("boneListOffset", np.uint32),
    ("minQ", np.int32, 3), ("bitsAndVertexCount", np.uint32), ("triangleCountAndRefinedGroup", np.uint32), ("boneCount", np.uint32),
    ("sourceGroupLocalIndex", np.uint32), ("reserved3", np.uint32), ("bounds", np.float32, 4)])
assert (GROUP_DTYPE.itemsize, SEGMENT_DTYPE.itemsize, CHUNK_DTYPE.itemsize, NODE_DTYPE.itemsize, LOCATOR_DTYPE.itemsize, PAGE_HEADER_DTYPE.itemsize, MESHLET_DESCRIPTOR_DTYPE.itemsize) == (76, 16, 20, 64, 16, 64, 64)

ARTIFACT_DTYPES = {
    "groups": GROUP_DTYPE, "segments": SEGMENT_DTYPE, "segmentBounds": np.dtype((np.float32, 4)), "groupChunks": CHUNK_DTYPE,
    "groupPageReferences": np.uint32, "groupPageReferenceOffsets": np.uint32, "nodes": NODE_DTYPE, "lodNodeRanges": RANGE_DTYPE,
    "lodLevelRoots": np.uint32, "objectBoundingSphere": np.float32, "counts": np.uint32, "meshPages": np.uint8, "meshPageOffsets": np.uint64,
    "stats": np.uint64,
}
STAT_NAMES = ("meshlets", "groups", "segments", "pages", "page_bytes", "meshlet_vertex_refs", "level_triangles", "group_vertices", "levels",
              "simplified_triangles", "d2h_bytes", "nodes", "launches")


class BuilderSettings(C.Structure):
    """clodb200_builder_settings == the mesh-mode fields of ClusterLODBuilderSettings (ClusterLODTypes.h:187-212)."""
    _fields_ = [("lodErrorMergePrevious", C.c_float), ("lodErrorMergeAdditive", C.c_float), ("partitionSizeFloor", C.c_uint32),
                ("preserveImportedNormals", C.c_int), ("enableNormalAttributeSimplification", C.c_int), ("normalAttributeWeight", C.c_float),
                ("simplifyTangentWeight", C.c_float), ("simplifyTangentSignWeight", C.c_float)]


class UvSet(C.Structure):
    _fields_ = [("values", C.c_void_p), ("count", C.c_size_t)]


class Geometry(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_count", C.c_size_t), ("vertex_stride", C.c_uint), ("vertex_flags", C.c_uint),
                ("indices", C.c_void_p), ("index_count", C.c_size_t), ("uv_sets", C.POINTER(UvSet)), ("uv_set_count", C.c_size_t), ("tangents", C.c_void_p),
                ("skinning_vertices", C.c_void_p), ("skinning_vertex_bytes", C.c_size_t), ("skinning_vertex_stride", C.c_uint)]


class Artifacts:
    """ClusterLODPrebuildArtifacts: prebuiltData arrays + cacheBuildData.meshPageBlobs (back to back, see page())."""

    def __init__(self, arrays):
        self.__dict__.update(arrays)
        self.stat = {n: int(self.stats[i]) for i, n in enumerate(STAT_NAMES)} if len(self.stats) else {}

    @property
    def page_count(self) -> int:
        return max(len(self.meshPageOffsets) - 1, 0)

    def page(self, i: int) -> np.ndarray:
        return self.meshPages[int(self.meshPageOffsets[i]): int(self.meshPageOffsets[i + 1])]

    def page_header(self, i: int):
        return np.frombuffer(self.page(i)[:64].tobytes(), PAGE_HEADER_DTYPE)[0]

    def page_descriptors(self, i: int) -> np.ndarray:
        p, h = self.page(i), self.page_header(i)
        o = int(h["descriptorOffset"])
        return np.frombuffer(p[o: o + 64 * int(h["meshletCount"])].tobytes(), MESHLET_DESCRIPTOR_DTYPE)


def interleave(positions, normals, uvs=None, colors=None) -> np.ndarray:
    """Mesh/VertexLayout.h: position f32x3 @0, normal f32x3 @12[, uv f32x2 @24][, colour f32x3]."""
    cols = [np.asarray(positions, np.float32), np.asarray(normals, np.float32)]
    if uvs is not None:
        cols.append(np.asarray(uvs, np.float32))
    if colors is not None:
        cols.append(np.asarray(colors, np.float32))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def bind(lib) -> None:
    L = lib
    L.clodb200_defaultBuilderSettings.restype = BuilderSettings
    L.clodb200_buildArtifacts.restype = C.c_void_p
    L.clodb200_buildArtifacts.argtypes = [C.POINTER(Geometry), C.POINTER(BuilderSettings)]
    L.clodb200_geometryUpload.restype = C.c_void_p
    L.clodb200_geometryUpload.argtypes = [C.POINTER(Geometry), C.POINTER(BuilderSettings)]
    L.clodb200_geometryFree.argtypes = [C.c_void_p]
    L.clodb200_geometryBuildArtifacts.restype = C.c_void_p
    L.clodb200_geometryBuildArtifacts.argtypes = [C.c_void_p]
    L.clodb200_artifactsGet.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.clodb200_artifactsFree.argtypes = [C.c_void_p]
    L.clodb200_artifactsSaveCache.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
    L.clodb200_artifactsSerializeMetadata.restype = C.c_size_t
    L.clodb200_artifactsSerializeMetadata.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_size_t]


def skinning_stream(positions, normals, joints, weights) -> np.ndarray:
    """The importer's skinning vertex stream as bytes [V, 88]: position f32x3, normal f32x3, joints u32x8, weights f32x8."""
    V = positions.shape[0]
    out = np.zeros((V, 88), np.uint8)
    out[:, 0:12] = np.ascontiguousarray(positions, np.float32).view(np.uint8).reshape(V, 12)
    out[:, 12:24] = np.ascontiguousarray(normals, np.float32).view(np.uint8).reshape(V, 12)
    out[:, 24:56] = np.ascontiguousarray(joints, np.uint32).view(np.uint8).reshape(V, 32)
    out[:, 56:88] = np.ascontiguousarray(weights, np.float32).view(np.uint8).reshape(V, 32)
    return out


def make_geometry(vertices, indices, flags, uv_sets=None, tangents=None, skinning=None):
    """-> (Geometry, keep-alive list). vertices: float32 [V, stride/4] interleaved; uv_sets: list of float32 [V, 2];
    skinning: uint8 [V', stride] skinning vertex stream (skinning_stream())."""
    vertices = np.ascontiguousarray(vertices, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    keep = [vertices, indices]
    g = Geometry()
    g.vertices = vertices.ctypes.data
    g.vertex_count = vertices.shape[0]
    g.vertex_stride = vertices.shape[1] * 4
    g.vertex_flags = flags
    g.indices = indices.ctypes.data
    g.index_count = indices.size
    if uv_sets:
        arr = (UvSet * len(uv_sets))()
        for i, u in enumerate(uv_sets):
            u = np.ascontiguousarray(u, np.float32)
            keep.append(u)
            arr[i].values = u.ctypes.data
            arr[i].count = u.shape[0]
        keep.append(arr)
        g.uv_sets = arr
        g.uv_set_count = len(uv_sets)
    if tangents is not None:
        tangents = np.ascontiguousarray(tangents, np.float32)
        keep.append(tangents)
        g.tangents = tangents.ctypes.data
    if skinning is not None:
        skinning = np.ascontiguousarray(skinning, np.uint8)
        keep.append(skinning)
        g.skinning_vertices = skinning.ctypes.data
        g.skinning_vertex_bytes = skinning.size
        g.skinning_vertex_stride = skinning.shape[1]
    return g, keep


def collect(lib, handle, views: bool = False) -> Artifacts:
    arrays = {}
    for name, dtype in ARTIFACT_DTYPES.items():
        ptr, size = C.c_void_p(), C.c_size_t()
        lib.clodb200_artifactsGet(handle, name.encode(), C.byref(ptr), C.byref(size))
        if not size.value:
            arrays[name] = np.zeros(0, dtype)
            continue
        view = np.frombuffer((C.c_ubyte * size.value).from_address(ptr.value), dtype=dtype)
        arrays[name] = view if views else view.copy()
    return Artifacts(arrays)
