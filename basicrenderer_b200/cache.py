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
