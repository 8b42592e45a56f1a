"""Profiling driver: W warm-up builds + K builds of one resident heightfield, for ncu / nvidia-smi sessions.
usage: python tools/prof_step.py [grid_n] [warmup] [steps]   (numbers printed here are never bench values)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from basicrenderer_b200 import load, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = load(0)
m = meshgen.grid(n, seed=1234)
from basicrenderer_b200 import artifacts as art  # noqa: E402

h = lib.upload_geometry(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS)
for it in range(warm + steps):
    l0 = lib.launch_count
    t = time.time()
    rec = lib.build_artifacts_resident(h, views=True)
    print(f"build {it}: {time.time() - t:.3f} s, launches {lib.launch_count - l0}, pages {rec.stat['pages']}", flush=True)
lib.free_geometry(h)
