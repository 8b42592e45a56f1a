"""Exactly N full builds of a bench workload (for ncu runs; numbers printed here are never bench values).
usage: python tools/one_build.py C2|C3 [builds] [grid_n or ico_f]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import artifacts as art  # noqa: E402
from basicrenderer_b200 import load, meshgen  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "C2"
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = load(0)
if what == "C3":
    f = int(sys.argv[3]) if len(sys.argv) > 3 else 2236
    mesh, _ = meshgen.icosphere_seams_torch(f, seed=42)
    v, i, flags = mesh.vertices, mesh.indices, art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS
else:
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2236
    m = meshgen.grid(n, seed=1234)
    v, i, flags = art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS
h = lib.upload_geometry(v, i, flags)
for it in range(builds):
    t = time.time()
    rec = lib.build_artifacts_resident(h, views=True)
    print(f"build {it}: {i.size // 3} tris {time.time() - t:.3f} s, pages {rec.stat['pages']} groups {rec.stat['groups']} launches {lib.launch_count}", flush=True)
lib.free_geometry(h)
