"""How much does the per-level MAX simplification error of the UNMODIFIED reference move when only the order of its input changes?

The grouping stage (meshopt_partitionClusters) is a greedy heap agglomeration whose result depends on cluster numbering; the
north_star holds it to invariants, not to bit equality. Every level above the first inherits the grouping below it, so two valid
builds of the same surface end up simplifying different groups. This tool measures the spread of the I7 statistic (max group
error per DAG level) for the reference against ITSELF on (a) the same mesh with the triangle order reversed and (b) the same
surface mirrored (x <-> y, winding fixed). It is the yardstick tests/test_scale_parity.py uses for the per-level max.

  python tools/noise_floor.py [grid:N:seed | ico:F ...]
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("scale_parity", os.path.join(ROOT, "tools", "scale_parity.py"))
sp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sp)
from oracle import clodref  # noqa: E402


def variants(m):
    idx = m.indices.reshape(-1, 3)
    yield "reversed triangle order", m.positions, m.normals, idx[::-1].copy().reshape(-1)
    yield "mirrored x<->y", m.positions[:, [1, 0, 2]].copy(), m.normals[:, [1, 0, 2]].copy(), idx[:, [0, 2, 1]].copy().reshape(-1)


def main():
    w = np.ones(3, np.float32)
    for spec in sys.argv[1:] or ["grid:1300:5", "ico:224"]:
        m = sp.make_mesh(spec)
        base = clodref.dag_build_stats(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7)
        print(f"== {spec}: {m.triangle_count} triangles, reference vs reference")
        for name, pos, nrm, idx in variants(m):
            o = clodref.dag_build_stats(pos, idx, attributes=nrm, attribute_weights=w, protect_mask=7)
            n = min(len(base["level_max_error"]), len(o["level_max_error"]))
            ok = base["level_max_error"][:n] > 0
            ratio = np.where(ok, o["level_max_error"][:n] / np.maximum(base["level_max_error"][:n], 1e-30), 1.0)
            tri = (o["level_triangles"][:n].astype(float) - base["level_triangles"][:n]) / base["level_triangles"][:n] * 100
            print(f" {name}: groups per level {o['level_groups'][:6].tolist()} vs {base['level_groups'][:6].tolist()}")
            print("   max-error ratio per level:", [round(float(x), 3) for x in ratio])
            print(f"   worst max-error ratio {max(ratio.max(), (1 / ratio[ok]).max()):.3f}, worst triangle delta {np.abs(tri[base['level_triangles'][:n] > 2000]).max():.3f} %")


if __name__ == "__main__":
    main()
