"""Development check: the product CUDA library and the serial host emulation of the same kernel sources must produce the
same DAG (the partition's merge rounds are order-independent reductions, so the callback stream is bit-identical).
  python tools/cmp_gpu_emu.py [grid_n] [other_gpu_lib.so]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import build, load, meshgen  # noqa: E402
from basicrenderer_b200.api import ClodLib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
libs = {"gpu": load(0), "emu": ClodLib(build.EMU_LIB)}
if len(sys.argv) > 2:
    libs["gpu_other"] = ClodLib(sys.argv[2])
w = np.ones(3, np.float32)
KEYS = ("group_depth", "group_simplified", "group_cluster_offsets", "cluster_refined", "cluster_index_offsets", "cluster_indices", "cluster_bounds")
fail = False
for name, m in [(f"grid{n}", meshgen.grid(n, seed=3)), ("ico128", meshgen.icosphere(128)), ("torus", meshgen.torus(400, 200, seed=1))]:
    recs = {k: lib.build_dag(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7) for k, lib in libs.items()}
    again = libs["gpu"].build_dag(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7)
    recs["gpu_again"] = again
    ref = recs["gpu"]
    for k, r in recs.items():
        if k == "gpu":
            continue
        bad = [f for f in KEYS if not np.array_equal(np.asarray(getattr(ref, f)), np.asarray(getattr(r, f)))]
        print(name, m.triangle_count, "tris", len(ref.group_depth), "groups | gpu vs", k, ":", "IDENTICAL" if not bad else f"DIFFERENT {bad} groups {len(r.group_depth)}")
        fail |= bool(bad) and k != "gpu_other"
sys.exit(1 if fail else 0)
