"""Reader of the CLod cache files clodb200_artifactsSaveCache writes: DeserializeMetadata (BasicRenderer/src/Import/
CLodCache.cpp:209-250) and the .clodbin container as read back by ReadPageBlobDirect (:86-101, 252-259), restated.
No reference test pins these bytes ("parity unpinned"); the reader enforces the loader's acceptance rules
(:209-250 `offset == blob.size()`, :700-705 one disk locator per mesh page)."""
import struct

import numpy as np

from . import artifacts as art


class _Cursor:
    def __init__(self, data: bytes):
        self.data, self.off = data, 0

    def pod(self, fmt):
        v = struct.unpack_from("<" + fmt, self.data, self.off)
        self.off += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def vector(self, dtype):
        n = self.pod("Q")
        dtype = np.dtype(dtype)
        a = np.frombuffer(self.data, dtype, n, self.off).copy()
        self.off += n * dtype.itemsize
        return a

    def string(self):
        n = self.pod("Q")
        s = self.data[self.off: self.off + n].decode()
        self.off += n
        return s


def read_metadata(blob: bytes) -> dict:
    c = _Cursor(blob)
    out = {"schemaVersion": c.pod("I")}
    assert out["schemaVersion"] == 47
    out["buildConfigHash"] = c.pod("Q")
    out["groups"] = c.vector(art.GROUP_DTYPE)
    out["segments"] = c.vector(art.SEGMENT_DTYPE)
    out["segmentBounds"] = c.vector(np.dtype((np.float32, 4)))
    out["objectBoundingSphere"] = np.array(c.pod("4f"), np.float32)
    if c.pod("B"):
        out["groupChunks"] = c.vector(art.CHUNK_DTYPE)
    out["groupDiskLocators"] = c.vector(art.LOCATOR_DTYPE)
    out["pageDiskLocators"] = c.vector(art.LOCATOR_DTYPE)
    out["groupPageReferences"] = c.vector(np.uint32)
    out["groupPageReferenceOffsets"] = c.vector(np.uint32)
    out["trianglePageCount"], out["voxelPageBase"], out["voxelPageCount"] = c.pod("I"), c.pod("I"), c.pod("I")
    out["sourceIdentifier"], out["primPath"], out["subsetName"] = c.string(), c.string(), c.string()
    out["sourceBuildConfigHash"] = c.pod("Q")
    out["containerFileName"] = c.string()
    out["nodes"] = c.vector(art.NODE_DTYPE)
    out["lodNodeRanges"] = c.vector(art.RANGE_DTYPE)
    out["lodLevelRoots"] = c.vector(np.uint32)
    out["maxDepth"], out["maxTraversalDepth"] = c.pod("I"), c.pod("I")
    assert c.off == len(blob), "trailing bytes (DeserializeMetadata requires offset == size)"
    # TryLoad's acceptance rule (:700-705): one locator per mesh page
    assert len(out["pageDiskLocators"]) == out["voxelPageBase"] + out["voxelPageCount"] > 0
    return out


def read_container(path: str):
    data = open(path, "rb").read()
    magic, version, reserved, page_count = struct.unpack_from("<4I", data, 0)
    assert magic == 0x444F4C43 and version == 4
    loc = np.frombuffer(data, art.LOCATOR_DTYPE, page_count, 16)
    pages = []
    end = 16 + 16 * page_count
    for l in loc:
        assert int(l["blobOffset"]) == end  # blobs back to back, in page order
        pages.append(data[int(l["blobOffset"]): int(l["blobOffset"]) + int(l["blobSizeBytes"])])
        end += int(l["blobSizeBytes"])
    assert end == len(data)
    return pages


# ---- cache naming (CLodCache.cpp:62-80, 586-633) ------------------------------------------------------------------------------
# Independent Python restatement of the published Boost.ContainerHash (>= 1.82) algorithm the names are built from; the library's
# C entry points (csrc/cachenames.cu) must agree with it. PARITY UNPINNED: Boost is neither vendored by the reference nor
# installed here, and no reference test holds a known name (tests/cpp/boost_names_check.cpp is the check for a machine with Boost).
_M64 = (1 << 64) - 1
CONFIG_HASH_CONSTANTS = (47, 128, 32, 4, 4, 1, 1, 7, 1, 1, 1, 7, 27, 3)  # CLodCache.cpp:594-607
CONFIG_HASH_ENVIRONMENT = (
    "BASICRENDERER_CLOD_VOXEL_MODE", "BASICRENDERER_CLOD_VOXEL_GRID", "BASICRENDERER_CLOD_VOXEL_MIN_RES", "BASICRENDERER_CLOD_VOXEL_RAYS", "BASICRENDERER_CLOD_VOXEL_SCALE",
    "BASICRENDERER_CLOD_VOXEL_RETRIES", "BASICRENDERER_CLOD_VOXEL_GROWTH", "BASICRENDERER_CLOD_VOXEL_ACCEPTANCE_BIAS", "BASICRENDERER_CLOD_VOXEL_OPACITY_THRESHOLD",
    "BASICRENDERER_CLOD_VOXEL_CARRY_ZERO_COVERAGE", "BASICRENDERER_CLOD_VOXEL_PRUNING")


def _hash_mix(x: int) -> int:
    m = 0xE9846AF9B1A615D
    x ^= x >> 32
    x = (x * m) & _M64
    x ^= x >> 32
    x = (x * m) & _M64
    x ^= x >> 28
    return x


def _mulx(x: int, y: int) -> int:
    r = x * y
    return (r & _M64) ^ (r >> 64)


def boost_hash_string(data: bytes, seed: int = 0) -> int:
    """boost::hash_range over chars on a 64-bit target (mulxp1_hash)."""
    q = 0x9E3779B97F4A7C15
    k = (q * q) & _M64
    n = len(data)
    w = _mulx((seed + q) & _M64, k)
    h = w ^ n
    p = 0
    while n >= 8:
        v1 = int.from_bytes(data[p:p + 8], "little")
        w = (w + q) & _M64
        h ^= _mulx((v1 + w) & _M64, k)
        p += 8
        n -= 8
    v1 = 0
    if n >= 4:
        v1 = ((int.from_bytes(data[p + n - 4:p + n], "little") << ((n - 4) * 8)) | int.from_bytes(data[p:p + 4], "little")) & _M64
    elif n >= 1:
        x1, x2 = (n - 1) & 2, n >> 1
        v1 = (data[p + x1] << (x1 * 8)) | (data[p + x2] << (x2 * 8)) | data[p]
    w = (w + q) & _M64
    h ^= _mulx((v1 + w) & _M64, k)
    return _mulx((h + w) & _M64, k)


def boost_hash_combine(seed: int, hashed: int) -> int:
    return _hash_mix((seed + 0x9E3779B9 + hashed) & _M64)


def build_config_hash(environ=None) -> int:
    import os

    environ = os.environ if environ is None else environ
    seed = 0
    for c in CONFIG_HASH_CONSTANTS:
        seed = boost_hash_combine(seed, c)
    for name in CONFIG_HASH_ENVIRONMENT:
        seed = boost_hash_combine(seed, boost_hash_string(environ.get(name, "").encode()))
    return seed


def cache_file_name(source_identifier: str, prim_path: str, subset_name: str, config_hash: int) -> str:
    seed = 0
    for s in (source_identifier, prim_path, subset_name):
        seed = boost_hash_combine(seed, boost_hash_string(s.encode()))
    seed = boost_hash_combine(seed, config_hash)
    return f"clod_{seed:x}.usdc"


def cache_subdirectory(source_identifier: str) -> str:
    import re

    stem = "scene"
    if source_identifier:
        file = re.split(r"[/\\]", source_identifier)[-1]
        dot = file.rfind(".")
        s = file if (dot <= 0 or file == "..") else file[:dot]
        if s:
            stem = s
    clean = "".join(c if (c.isascii() and (c.isalnum() or c in "_-")) else "_" for c in stem) or "scene"
    return f"clod/{clean}_{boost_hash_combine(0, boost_hash_string(source_identifier.encode())):x}"
