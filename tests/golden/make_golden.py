"""Generates the committed golden fixtures in tests/golden/*.npz by running the compiled, UNMODIFIED reference
(oracle/_ref/libclodref.so, built from /root/reference by oracle/Makefile) on small seeded meshes.

    python tests/golden/make_golden.py

Each fixture holds the mesh (so the test does not depend on the generator staying bit-stable) and the reference's outputs
at every stage boundary of clodBuildEx: position remap, protect locks, depth-0 meshlets, and per level the merged group
index lists, lock bytes, simplified lists, errors and terminal flags, plus the whole callback stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from basicrenderer_b200 import meshgen  # noqa: E402
from oracle import clodref  # noqa: E402

CASES = {
    "grid40": lambda: meshgen.grid(40, seed=7),
    "ico12uv": lambda: meshgen.icosphere(12, True, True),
    "torus48": lambda: meshgen.torus(48, 24, seed=3),
}


def main():
    for name, make in CASES.items():
        m = make()
        attrs = np.ascontiguousarray(m.normals, dtype=np.float32)
        w = np.ones(3, np.float32)
        dag = clodref.dag_build(m.positions, m.indices, attributes=attrs, attribute_weights=w, protect_mask=7)
        out = {
            "positions": m.positions.astype(np.float32),
            "indices": m.indices.astype(np.uint32),
            "attributes": attrs,
            "attribute_weights": w,
            "protect_mask": np.array([7], np.uint32),
            "remap": dag.get("remap"),
            "protect_locks": dag.get("protect_locks"),
            "num_levels": np.array([dag.num_levels], np.uint32),
        }
        ro, rv, ri = clodref.clusterize(m.positions, m.indices)
        out["clusterize_offsets"], out["clusterize_vertices"], out["clusterize_indices"] = ro, rv, ri
        for level in range(dag.num_levels):
            for key in ("merged_indices", "merged_offsets", "locks", "simp_offsets", "simp_indices", "group_error", "group_terminal", "group_offsets"):
                out[f"L{level}.{key}"] = dag.level(level, key)
        for key in ("out.group_depth", "out.group_simplified", "out.group_cluster_offsets", "out.cluster_refined", "out.cluster_bounds", "out.cluster_vertex_count", "out.cluster_index_offsets", "out.cluster_indices",
                    "cluster_depth", "cluster_index_offsets", "cluster_indices", "cluster_vertices", "cluster_bounds"):
            out[key] = dag.get(key)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, m.triangle_count, "triangles,", dag.num_levels, "levels ->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
