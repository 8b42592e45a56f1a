"""Invariant checks for a cluster-LOD DAG, applied to the output callback stream of a build.

Used on this framework's builds and, unchanged, on the reference's own callback stream (the checker must accept the
reference). Pure numpy; no GPU.
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.finfo(np.float32).max


def _canon_triangles(idx: np.ndarray) -> np.ndarray:
    """Rotation-normalised triangle rows sorted lexicographically (multiset comparison)."""
    t = idx.reshape(-1, 3).astype(np.int64)
    k = np.argmin(t, axis=1)
    r = np.stack([t[np.arange(len(t)), k], t[np.arange(len(t)), (k + 1) % 3], t[np.arange(len(t)), (k + 2) % 3]], axis=1)
    order = np.lexsort((r[:, 2], r[:, 1], r[:, 0]))
    return r[order]


def check_dag(rec, positions: np.ndarray, indices: np.ndarray, max_triangles=128, max_vertices=128, max_group_clusters=512, max_refined=8, remap=None, check_indices=True):
    """Raises AssertionError on the first violated invariant; returns per-depth statistics.

    rec needs: group_depth, group_simplified[G,5], group_cluster_offsets[G+1], cluster_refined, cluster_bounds[K,5],
    cluster_vertex_count, cluster_index_offsets[K+1], cluster_indices.
    """
    gdepth = rec.group_depth
    G = len(gdepth)
    goff = rec.group_cluster_offsets.astype(np.int64)
    coff = rec.cluster_index_offsets.astype(np.int64)
    K = len(rec.cluster_refined)
    assert goff[-1] == K and len(goff) == G + 1
    tri_counts = np.diff(coff) // 3
    gerr = rec.group_simplified[:, 4]

    # I2 meshlet limits
    assert tri_counts.max() <= max_triangles, "I2: triangle limit"
    assert rec.cluster_vertex_count.max() <= max_vertices, "I2: vertex limit"
    assert tri_counts.min() >= 1

    cluster_group = np.repeat(np.arange(G), np.diff(goff))
    cluster_depth = gdepth[cluster_group]
    # groups are emitted depth by depth
    assert (np.diff(gdepth) >= 0).all(), "groups must be emitted in depth order"

    # I3 group limits
    assert np.diff(goff).max() <= max_group_clusters, "I3: clusters per group"
    for g in range(G):
        r = rec.cluster_refined[goff[g] : goff[g + 1]]
        assert len(np.unique(r)) <= max_refined, f"I3: group {g} has {len(np.unique(r))} refined ids"

    # refined ids point at groups of the previous depth; depth-0 clusters are original geometry
    ref = rec.cluster_refined
    assert (ref[cluster_depth == 0] == -1).all()
    later = cluster_depth > 0
    assert (ref[later] >= 0).all() and (ref[later] < G).all()
    assert (gdepth[ref[later]] == cluster_depth[later] - 1).all(), "refined must point one level down"

    # I4 monotone error along DAG edges: error(group containing cluster) > error(refined group) when both finite
    child_err = gerr[ref[later]]
    parent_err = gerr[cluster_group[later]]
    finite = (child_err < FLT_MAX) & (parent_err < FLT_MAX)
    assert (parent_err[finite] > child_err[finite]).all(), "I4: DAG error not strictly monotone"
    # a refined group that produced clusters cannot be terminal
    assert (child_err < FLT_MAX).all(), "clusters refer to a terminal group"
    # cluster error is the error of the group that produced it (0 for original geometry)
    assert (rec.cluster_bounds[cluster_depth == 0, 4] == 0).all()
    assert np.array_equal(rec.cluster_bounds[later, 4], child_err)

    # I8 single root / termination: deepest level holds exactly one terminal group
    dmax = gdepth.max()
    last = np.nonzero(gdepth == dmax)[0]
    assert (gerr[last] == FLT_MAX).all(), "I8: deepest groups must be terminal"

    stats = []
    for d in range(dmax + 1):
        sel = cluster_depth == d
        stats.append({"depth": d, "groups": int((gdepth == d).sum()), "clusters": int(sel.sum()), "triangles": int(tri_counts[sel].sum()),
                      "max_error": float(np.max(np.where(gerr[gdepth == d] < FLT_MAX, gerr[gdepth == d], 0)))})

    if check_indices:
        idx = rec.cluster_indices
        assert idx.size == coff[-1]
        # I1 coverage at depth 0: every input triangle exactly once
        sel0 = np.nonzero(cluster_depth == 0)[0]
        lo, hi = coff[sel0[0]], coff[sel0[-1] + 1]
        assert np.array_equal(_canon_triangles(idx[lo:hi]), _canon_triangles(indices)), "I1: depth-0 clusters must cover the input exactly"
        # unique vertex count per cluster matches vertex_count (sampled)
        for c in np.linspace(0, K - 1, num=min(K, 200), dtype=np.int64):
            assert len(np.unique(idx[coff[c] : coff[c + 1]])) == rec.cluster_vertex_count[c]
        # I5 crack-free: at every depth the set of boundary edges (by position) between groups... checked through locks:
        # every depth-(d+1) triangle only uses vertices that exist at depth d in the refined group
        if remap is not None:
            for g in range(G):
                kids = np.nonzero(ref == g)[0]
                if len(kids) == 0:
                    continue
                src = np.concatenate([idx[coff[c] : coff[c + 1]] for c in range(goff[g], goff[g + 1])])
                dst = np.concatenate([idx[coff[c] : coff[c + 1]] for c in kids])
                assert np.isin(remap[dst], remap[src]).all(), f"group {g}: simplified triangles use foreign vertices"
                assert dst.size <= src.size
    return stats
