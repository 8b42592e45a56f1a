"""Committed golden vectors (tests/golden/*.npz, produced by the compiled reference via tests/golden/make_golden.py):
the CUDA path (and the development emulation) must reproduce them without the reference being present."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.fixture(params=GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def gold(request):
    return np.load(request.param)


def test_golden_files_present():
    assert len(GOLDEN) >= 3


def test_remap_and_clusterize(lib, gold):
    assert np.array_equal(lib.position_remap(gold["positions"]), gold["remap"])
    go, gv, gs, gi = lib.clusterize(gold["positions"], gold["indices"])
    assert np.array_equal(go, gold["clusterize_offsets"])
    assert np.array_equal(gv, gold["clusterize_vertices"])
    assert np.array_equal(gi, gold["clusterize_indices"])


def test_locks_and_simplify_every_level(lib, gold):
    locks = gold["protect_locks"]
    for level in range(int(gold["num_levels"][0])):
        mi, mo = gold[f"L{level}.merged_indices"], gold[f"L{level}.merged_offsets"]
        got = lib.lock_boundary(locks, mi, mo, gold["remap"])
        locks = gold[f"L{level}.locks"]
        assert np.array_equal(got, locks), level
        si, so, se = lib.simplify_groups(gold["positions"], mi, mo, locks, attributes=gold["attributes"], attribute_weights=gold["attribute_weights"])
        rso, rsi = gold[f"L{level}.simp_offsets"], gold[f"L{level}.simp_indices"]
        term = gold[f"L{level}.group_terminal"].astype(bool)
        for g in range(len(mo) - 1):
            want = rsi[rso[g] : rso[g + 1]]
            if term[g]:
                continue  # terminal groups keep no simplified list (clusterlod.h:723-729)
            assert np.array_equal(si[so[g] : so[g + 1]], want), (level, g)
            assert se[g] == gold[f"L{level}.group_error"][g], (level, g)


def test_callback_stream(lib, gold):
    """These meshes have one group per level, so the whole clodBuildEx callback stream is determined: bit-exact."""
    if any(len(gold[f"L{l}.group_offsets"]) != 2 for l in range(int(gold["num_levels"][0]))):
        pytest.skip("several groups on some level")
    rec = lib.build_dag(gold["positions"], gold["indices"], attributes=gold["attributes"], attribute_weights=gold["attribute_weights"], protect_mask=int(gold["protect_mask"][0]))
    assert np.array_equal(rec.group_depth, gold["out.group_depth"])
    assert np.array_equal(rec.cluster_refined, gold["out.cluster_refined"])
    assert np.array_equal(rec.cluster_indices, gold["out.cluster_indices"])
    assert np.array_equal(rec.cluster_vertex_count, gold["out.cluster_vertex_count"])
    assert np.array_equal(rec.cluster_bounds, gold["out.cluster_bounds"])
    assert np.array_equal(rec.group_simplified, gold["out.group_simplified"])


def test_local_indices(lib, gold):
    coff = gold["out.cluster_index_offsets"].astype(np.uint64)
    idx = gold["out.cluster_indices"]
    verts, tris, counts = lib.local_indices_batch(idx, coff)
    assert np.array_equal(counts, gold["out.cluster_vertex_count"])
    for c in range(len(coff) - 1):
        seg = idx[coff[c] : coff[c + 1]]
        # defining property (clusterlod.h:180-182) + first-occurrence order
        assert np.array_equal(verts[c, : counts[c]][tris[coff[c] : coff[c + 1]]], seg)
        _, first = np.unique(seg, return_index=True)
        assert np.array_equal(verts[c, : counts[c]], seg[np.sort(first)])
