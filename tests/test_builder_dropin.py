"""The drop-in claim, tested with the reference's own code: the UNMODIFIED L3 builder BuildClusterLODArtifactsFromGeometry
(ClusterLODUtilities.cpp:5325) is linked against a clodBuildEx that forwards to libclodb200 (oracle/shims/
clusterlod_via_clodb200.cpp == the call-site change of INTEGRATION.md). The builder's throw-on-violation validation
(ClusterLODUtilities.cpp:4978-5253: monotone errors, fan-out <= 8, reachability, page/segment consistency) then judges our
DAG, and where the grouping is forced (one group per level) the pages, groups, segments and traversal nodes it produces
must be byte-identical to the all-reference build."""
import os

import numpy as np
import pytest

from basicrenderer_b200 import meshgen


def _lib_path(kind):
    from basicrenderer_b200 import api, build

    return build.build_emu() if kind == "emu" else api.PRODUCT_LIB


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def clodb_path(request):
    from oracle import clodfull

    if not clodfull.available(True):
        if not os.path.exists("/root/reference"):
            pytest.skip("oracle/_ref/libclodref_full_ours.so missing and /root/reference not mounted")
        from oracle import clodref

        clodref.build(full=True)
    return _lib_path(request.param)


@pytest.mark.parametrize("name", ["grid64", "ico24", "torus", "grid160"])
def test_reference_builder_on_our_dag_is_byte_identical_when_grouping_is_forced(clodb_path, meshes, name):
    from oracle import clodfull

    m = meshes[name]
    v = clodfull.interleave(m.positions, m.normals)
    ref = clodfull.build(v, m.indices)
    ours = clodfull.build(v, m.indices, clodb200_lib=clodb_path)
    if np.any(np.bincount(ref.groups["depth"]) != 1):
        pytest.skip("reference uses several groups on some level")
    for key in ("groups", "segments", "segmentBounds", "nodes", "lodNodeRanges", "lodLevelRoots", "pageDiskLocators", "groupPageReferences", "groupPageReferenceOffsets", "counts", "objectBoundingSphere",
                "meshPageOffsets", "meshPages"):
        assert np.array_equal(getattr(ref, key), getattr(ours, key)), key


def test_reference_builder_accepts_our_multi_group_dag(clodb_path):
    """Several groups per level: our grouping differs from the heap-serial reference, but the reference builder's own
    validation must accept the DAG (it throws otherwise) and the per-depth shape must stay within the stated bars."""
    from oracle import clodfull

    m = meshgen.grid(330, seed=11)
    v = clodfull.interleave(m.positions, m.normals)
    ref = clodfull.build(v, m.indices)
    ours = clodfull.build(v, m.indices, clodb200_lib=clodb_path)  # raises if the builder's validation fails
    assert ours.counts[3] == ref.counts[3]  # maxDepth
    for d in range(int(ref.counts[3]) + 1):
        r = ref.groups[ref.groups["depth"] == d]
        o = ours.groups[ours.groups["depth"] == d]
        assert abs(int(o["meshletCount"].sum()) - int(r["meshletCount"].sum())) <= 0.03 * r["meshletCount"].sum() + 2
        assert len(o) <= 2 * len(r) + 1
    # monotone DAG error (I4) as the renderer sees it: maxParentError > own error for every non-root group
    err = ours.groups["bounds"][:, 4]
    nonroot = ours.groups["parentGroupId"] >= 0
    finite = nonroot & (ours.groups["maxParentError"] < 3e38) & (err < 3e38)
    assert np.all(ours.groups["maxParentError"][finite] > err[finite])


def test_reference_builder_with_uv_seams_and_tangents(clodb_path, meshes):
    """normals + UV atlas => MikkTSpace tangents, 7 simplification attributes, protect bits on seams (config C3 shape)."""
    from oracle import clodfull

    m = meshes["ico16uv"]
    v = clodfull.interleave(m.positions, m.normals, m.vertices[:, 6:8])
    flags = clodfull.VERTEX_NORMALS | clodfull.VERTEX_TEXCOORDS
    ref = clodfull.build(v, m.indices, flags=flags)
    ours = clodfull.build(v, m.indices, flags=flags, clodb200_lib=clodb_path)
    assert ours.counts[3] == ref.counts[3]
    if np.all(np.bincount(ref.groups["depth"]) == 1):
        assert np.array_equal(ref.meshPages, ours.meshPages)
        assert np.array_equal(ref.groups, ours.groups)
        assert np.array_equal(ref.nodes, ours.nodes)
