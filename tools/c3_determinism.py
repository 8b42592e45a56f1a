"""Debug driver: are the MikkTSpace tangents and the artifacts of a C3-shaped build reproducible run to run?
usage: python tools/c3_determinism.py [ico_f] [builds] [with_uv=1]"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import artifacts as art  # noqa: E402
from basicrenderer_b200 import load, meshgen  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 300
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with_uv = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
lib = load(0)
mesh, _ = meshgen.icosphere_seams_torch(f, seed=42)
if with_uv:
    t0 = lib.mikk_tangents(mesh.vertices, mesh.indices)
    t1 = lib.mikk_tangents(mesh.vertices, mesh.indices)
    print("mikk tangents reproducible:", np.array_equal(np.asarray(t0), np.asarray(t1)), flush=True)
    v, flags = mesh.vertices, art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS
else:
    v = np.ascontiguousarray(np.asarray(mesh.vertices).reshape(-1, 8)[:, :6])
    flags = art.VERTEX_NORMALS
h = lib.upload_geometry(v, mesh.indices, flags)
for it in range(builds):
    rec = lib.build_artifacts_resident(h, views=True)
    print(f"build {it}: pages {rec.stat['pages']} groups {rec.stat['groups']} meshlets {rec.stat['meshlets']} crc {zlib.crc32(np.asarray(rec.meshPages).tobytes()):08x}", flush=True)
lib.free_geometry(h)
