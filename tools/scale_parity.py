"""Scale parity check: the DAG built by a clodb200 library (product CUDA library, or the development emulation with
`--emu`) next to the unmodified reference's clodBuildEx on the same mesh, per DAG level: triangles (I6, +-2 %), groups,
sloppy-fallback groups and max finite group error (I7, <= 1.05x).

  python tools/scale_parity.py [--emu] [--json out.json] mesh [mesh ...]
  mesh: grid:N[:seed]  ico:F  icouv:F  torus:A:B   (C2 = grid:2236:1234, C1 = ico:224)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import build, load, meshgen  # noqa: E402
from basicrenderer_b200.api import ClodLib  # noqa: E402
from oracle import clodref  # noqa: E402  (checker)

FLT_MAX = np.float32(3.4028234663852886e38)


def make_mesh(spec: str):
    p = spec.split(":")
    if p[0] == "grid":
        return meshgen.grid(int(p[1]), seed=int(p[2]) if len(p) > 2 else 1234)
    if p[0] == "ico":
        return meshgen.icosphere(int(p[1]))
    if p[0] == "icouv":
        return meshgen.icosphere(int(p[1]), True, True)
    if p[0] == "torus":
        return meshgen.torus(int(p[1]), int(p[2]), seed=1)
    raise SystemExit(f"unknown mesh spec {spec}")


def mesh_attributes(m):
    """Attribute stream + weights the way the tests and the bench feed the inner boundary (normals, weight 1)."""
    return m.normals, np.ones(3, np.float32)


def ours_stats(lib, m, keep_groups=False):
    attrs, w = mesh_attributes(m)
    t0 = time.time()
    h = lib.upload_mesh(m.positions, m.indices, attributes=attrs, attribute_weights=w, protect_mask=7)
    try:
        rec = lib.build_dag_resident(h, keep_indices=False)
    finally:
        lib.free_mesh(h)
    dt = time.time() - t0
    depth = rec.group_depth
    err = rec.group_simplified[:, 4]
    levels = int(depth.max()) + 1
    max_err = np.zeros(levels, np.float32)
    for d in range(levels):
        e = err[(depth == d) & (err < FLT_MAX)]
        max_err[d] = e.max() if e.size else 0.0
    extra = {"group_error": np.asarray(err).copy(), "group_depth": np.asarray(depth).copy()} if keep_groups else {}
    return {
        **extra,
        "level_triangles": rec.level_triangles.astype(np.int64), "level_groups": rec.level_groups.astype(np.int64),
        "level_sloppy": rec.level_sloppy.astype(np.int64), "level_passes": rec.level_passes.astype(np.int64),
        "level_max_error": max_err, "groups": rec.groups, "meshlets": rec.total_clusters, "seconds": dt,
    }


def ref_stats(m, threads=None):
    attrs, w = mesh_attributes(m)
    t0 = time.time()
    st = clodref.dag_build_stats(m.positions, m.indices, attributes=attrs, attribute_weights=w, protect_mask=7, threads=threads)
    st["seconds"] = time.time() - t0
    return st


def compare(ours, ref):
    """Returns (rows, summary): per-level comparison and the worst deviations."""
    n = max(len(ours["level_triangles"]), len(ref["level_triangles"]))
    rows = []
    worst_tri = 0.0
    worst_err = 0.0
    for d in range(n):
        ot = int(ours["level_triangles"][d]) if d < len(ours["level_triangles"]) else 0
        rt = int(ref["level_triangles"][d]) if d < len(ref["level_triangles"]) else 0
        oe = float(ours["level_max_error"][d]) if d < len(ours["level_max_error"]) else 0.0
        re_ = float(ref["level_max_error"][d]) if d < len(ref["level_max_error"]) else 0.0
        dt = (ot - rt) / rt * 100 if rt else 0.0
        er = oe / re_ if re_ > 0 else (1.0 if oe == 0 else float("inf"))
        rows.append({
            "depth": d, "triangles": ot, "ref_triangles": rt, "delta_pct": dt,
            "groups": int(ours["level_groups"][d]) if d < len(ours["level_groups"]) else 0, "ref_groups": int(ref["level_groups"][d]) if d < len(ref["level_groups"]) else 0,
            "sloppy": int(ours["level_sloppy"][d]) if d < len(ours["level_sloppy"]) else 0, "ref_sloppy": int(ref["level_sloppy"][d]) if d < len(ref["level_sloppy"]) else 0,
            "max_error": oe, "ref_max_error": re_, "error_ratio": er,
        })
        if rt > 2000:  # tiny top levels differ by a handful of triangles; the +2 absolute slack of the tests covers them
            worst_tri = max(worst_tri, abs(dt))
        if re_ > 0:
            worst_err = max(worst_err, er)
    return rows, {"levels": len(ours["level_triangles"]), "ref_levels": len(ref["level_triangles"]), "worst_triangle_delta_pct": worst_tri, "worst_error_ratio": worst_err,
                  "sloppy": int(np.sum(ours["level_sloppy"])), "ref_sloppy": int(np.sum(ref["level_sloppy"])), "groups": int(ours["groups"]), "ref_groups": int(np.sum(ref["level_groups"])),
                  "meshlets": int(ours["meshlets"]), "ref_meshlets": int(ref["total_clusters"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--json")
    ap.add_argument("--ref-threads", type=int, default=None)
    ap.add_argument("meshes", nargs="+")
    a = ap.parse_args()
    lib = ClodLib(build.build_emu()) if a.emu else load(0)
    report = {}
    for spec in a.meshes:
        m = make_mesh(spec)
        ours = ours_stats(lib, m)
        ref = ref_stats(m, a.ref_threads)
        rows, summary = compare(ours, ref)
        print(f"== {spec}: {m.triangle_count} triangles | ours {ours['seconds']:.2f} s, reference {ref['seconds']:.1f} s")
        print(" depth   triangles (ref)            d%   groups (ref)  sloppy (ref)  passes  max error (ref)           ratio")
        for r, p in zip(rows, list(ours["level_passes"]) + [0] * len(rows)):
            print(f" {r['depth']:>4} {r['triangles']:>10} ({r['ref_triangles']:>10}) {r['delta_pct']:>+7.2f} {r['groups']:>6} ({r['ref_groups']:>5}) {r['sloppy']:>6} ({r['ref_sloppy']:>4}) {int(p):>6}   {r['max_error']:.6g} ({r['ref_max_error']:.6g}) {r['error_ratio']:.3f}")
        print(" summary:", json.dumps(summary))
        report[spec] = {"summary": summary, "levels": rows}
    if a.json:
        with open(a.json, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
