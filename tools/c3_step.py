"""Debug/profiling driver for the C3 shape (seamed icosphere, pos + normal + uv, MikkTSpace inside the build).
usage: python tools/c3_step.py [ico_f] [builds]   (numbers printed here are never bench values)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import artifacts as art  # noqa: E402
from basicrenderer_b200 import load, meshgen  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 300
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = load(0)
mesh, _ = meshgen.icosphere_seams_torch(f, seed=42)
h = lib.upload_geometry(mesh.vertices, mesh.indices, art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS)
for it in range(builds):
    t = time.time()
    rec = lib.build_artifacts_resident(h, views=True)
    print(f"build {it}: {mesh.triangle_count} tris {time.time() - t:.3f} s, pages {rec.stat['pages']} groups {rec.stat['groups']}", flush=True)
lib.free_geometry(h)
