"""Outer boundary: clodb200_buildArtifacts against the reference's own L3 builder.

Oracle: oracle/_ref/libclodref_full_ours.so = the UNMODIFIED BuildClusterLODArtifactsFromGeometry
(ClusterLODUtilities.cpp:5325) linked against a clodBuildEx that forwards to the library under test, so both sides see the
same clusters and groups. Every array of ClusterLODPrebuiltData and every byte of every mesh page must then be identical:
the group output tables (:856-1805), the traversal hierarchy (:4606-4963) and the mesh-wide page packing (:2313-2540)
are deterministic integer/float work given the cluster assignment (bit-exact bar)."""
import os

import numpy as np
import pytest

from basicrenderer_b200 import artifacts as art
from basicrenderer_b200 import meshgen

KEYS = ("groups", "segments", "segmentBounds", "groupChunks", "nodes", "lodNodeRanges", "lodLevelRoots", "groupPageReferences", "groupPageReferenceOffsets",
        "counts", "objectBoundingSphere", "meshPageOffsets", "meshPages")


@pytest.fixture
def full_oracle():
    from oracle import clodfull

    if not clodfull.available(True):
        if not os.path.exists("/root/reference"):
            pytest.skip("oracle/_ref/libclodref_full_ours.so missing and /root/reference not mounted")
        from oracle import clodref

        clodref.build(full=True)
    return clodfull


def _assert_identical(ref, ours):
    for key in KEYS:
        a, b = getattr(ref, key), getattr(ours, key)
        assert a.shape == b.shape, (key, a.shape, b.shape)
        if a.dtype.names:
            for f in a.dtype.names:
                assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), (key, f, np.flatnonzero((a[f].view(np.uint32) != b[f].view(np.uint32)).reshape(len(a), -1).any(axis=1))[:8])
        else:
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (key, np.flatnonzero(a.reshape(-1).view(np.uint8) != b.reshape(-1).view(np.uint8))[:8])


@pytest.mark.parametrize("name", ["grid64", "ico24", "torus", "grid160"])
def test_artifacts_identical_to_reference_builder_on_same_dag(lib, full_oracle, meshes, name):
    m = meshes[name]
    v = art.interleave(m.positions, m.normals)
    ref = full_oracle.build(v, m.indices, clodb200_lib=lib.path)
    ours = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS)
    _assert_identical(ref, ours)


def test_artifacts_multi_group_levels(lib, full_oracle):
    """several groups per level: multi-parent DAG edges, several pages per group, refined-id buckets"""
    m = meshgen.grid(330, seed=11)
    v = art.interleave(m.positions, m.normals)
    ref = full_oracle.build(v, m.indices, clodb200_lib=lib.path)
    ours = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS)
    assert np.bincount(ours.groups["depth"]).max() > 1
    _assert_identical(ref, ours)
    # the resident path gives the same bytes
    h = lib.upload_geometry(v, m.indices, art.VERTEX_NORMALS)
    again = lib.build_artifacts_resident(h)
    lib.free_geometry(h)
    _assert_identical(ours, again)


def test_artifacts_position_only_and_colors(lib, full_oracle, meshes):
    m = meshes["ico24"]
    rng = np.random.default_rng(5)
    colors = rng.random((m.vertex_count, 3), dtype=np.float32) * 1.2 - 0.1  # exercises the clamp
    v = art.interleave(m.positions, m.normals, colors=colors)
    flags = art.VERTEX_NORMALS | art.VERTEX_COLORS
    _assert_identical(full_oracle.build(v, m.indices, flags=flags, clodb200_lib=lib.path), lib.build_artifacts(v, m.indices, flags))
    # no normal flag: position-only pages, no simplification attributes
    _assert_identical(full_oracle.build(v, m.indices, flags=0, clodb200_lib=lib.path), lib.build_artifacts(v, m.indices, 0))


def test_artifacts_uv_seams_and_tangents(lib, full_oracle, meshes):
    """normals + UV atlas (config C3 shape): 7 simplification attributes, UV bitstreams in the pages. The MikkTSpace tangent
    stream is generated inside our call (csrc/mikk.cu) exactly as the reference does inside its own (:5359-5366): it must
    equal the attribute stream the reference hands to clodBuildEx bit for bit, and the artifacts must be byte-identical."""
    m = meshes["ico16uv"]
    v = art.interleave(m.positions, m.normals, m.vertices[:, 6:8])
    flags = art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS
    ref = full_oracle.build(v, m.indices, flags=flags, clodb200_lib=lib.path)
    attrs = full_oracle.last_attributes()
    assert attrs.shape == (m.vertex_count, 7)
    got = lib.mikk_tangents(v, m.indices)
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(attrs[:, 3:7]).view(np.uint32))
    ours = lib.build_artifacts(v, m.indices, flags)
    assert ours.page_header(0)["uvSetCount"] == 1
    _assert_identical(ref, ours)
    # a caller-supplied stream overrides the generator
    _assert_identical(ref, lib.build_artifacts(v, m.indices, flags, tangents=attrs[:, 3:7]))


def test_artifacts_empty_geometry(lib):
    v = np.zeros((0, 6), np.float32)
    a = lib.build_artifacts(v, np.zeros(0, np.uint32), art.VERTEX_NORMALS)
    assert len(a.groups) == 0 and a.page_count == 0


def test_cache_files_round_trip(lib, meshes, tmp_path):
    """.clodbin container + metadata blob (CLodCache.cpp:169-207, 252-259, 314-375): re-read with a restated
    DeserializeMetadata / loader acceptance check (:209-250, :635-713)."""
    from basicrenderer_b200.cache import read_container, read_metadata

    m = meshes["grid160"]
    v = art.interleave(m.positions, m.normals)
    a = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS, keep_handle=True)
    lib.save_cache(a, str(tmp_path), "clod_test.clodbin", "clod_test.clodblob", "scene.gltf", "/mesh0", "", 0x1234)
    blob = (tmp_path / "clod_test.clodblob").read_bytes()
    assert blob == lib.serialize_metadata(a, "clod_test.clodbin", "scene.gltf", "/mesh0", "", 0x1234)
    lib.free_artifacts(a)
    meta = read_metadata(blob)
    assert meta["schemaVersion"] == 47 and meta["buildConfigHash"] == 0x1234 and meta["containerFileName"] == "clod_test.clodbin"
    assert np.array_equal(meta["groups"], a.groups) and np.array_equal(meta["nodes"], a.nodes) and np.array_equal(meta["segments"], a.segments)
    pages = read_container(str(tmp_path / "clod_test.clodbin"))
    assert len(pages) == a.page_count == meta["trianglePageCount"] == len(meta["pageDiskLocators"])
    for i, p in enumerate(pages):
        assert np.array_equal(np.frombuffer(p, np.uint8), a.page(i))
        assert int(meta["pageDiskLocators"][i]["blobSizeBytes"]) == len(p)


def _skinning(m, bones=37, seed=9, zero_fraction=0.3):
    """Random skinning influences: up to 8 joints per vertex out of `bones`, a share of the weights exactly zero (those joints must
    not enter the meshlet bone lists), one vertex in ten without any positive weight."""
    rng = np.random.default_rng(seed)
    V = m.vertex_count
    joints = rng.integers(0, bones, (V, 8)).astype(np.uint32)
    weights = rng.random((V, 8), dtype=np.float32)
    weights[rng.random((V, 8)) < zero_fraction] = 0.0
    weights[rng.random(V) < 0.1] = 0.0
    weights[0, 0] = -0.5  # negative weights do not count either (weight > 0.0f)
    return art.skinning_stream(m.positions, m.normals, joints, weights)


@pytest.mark.parametrize("name", ["grid64", "torus"])
def test_skinned_artifacts_identical_to_reference_builder(lib, full_oracle, meshes, name):
    """Skinned meshes (ClusterLODUtilities.cpp:1091-1120, 1240-1265, 1677-1711): joint and weight arrays per meshlet vertex, sorted
    per-meshlet bone lists, boneListOffset / boneCount in the descriptors, joints | weights in the page attribute mask; the extra
    64 bytes per vertex also move the page splits."""
    m = meshes[name]
    v = art.interleave(m.positions, m.normals)
    skin = _skinning(m)
    ref = full_oracle.build(v, m.indices, clodb200_lib=lib.path, skinning=skin)
    ours = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS | art.VERTEX_SKINNED, skinning=skin)
    _assert_identical(ref, ours)
    header = ours.page(0)[:64].view(np.uint32)
    assert header[2] & 0b0110 == 0b0110 and header[9] != 0 and header[10] != 0  # attribute mask, joint / weight array offsets
    plain = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS)
    assert ours.meshPageOffsets[-1] > plain.meshPageOffsets[-1]


def test_skinned_multi_group_and_short_skinning_stream(lib, full_oracle):
    """Several groups and pages per level; a skinning stream shorter than the vertex stream (vertices beyond it read as zero
    influences, :1107-1110); the resident path gives the same bytes."""
    m = meshgen.grid(200, seed=21)
    v = art.interleave(m.positions, m.normals)
    skin = _skinning(m, bones=300, seed=4)[: m.vertex_count - 1234]
    flags = art.VERTEX_NORMALS | art.VERTEX_SKINNED
    ref = full_oracle.build(v, m.indices, flags=flags, clodb200_lib=lib.path, skinning=skin)
    ours = lib.build_artifacts(v, m.indices, flags, skinning=skin)
    assert np.bincount(ours.groups["depth"]).max() > 1
    _assert_identical(ref, ours)
    h = lib.upload_geometry(v, m.indices, flags, skinning=skin)
    again = lib.build_artifacts_resident(h)
    lib.free_geometry(h)
    _assert_identical(ours, again)


@pytest.mark.parametrize("name", ["grid64", "ico24"])
def test_recomputed_group_normals(lib, full_oracle, meshes, name):
    """preserveImportedNormals = false (RecalculateGroupNormals, ClusterLODUtilities.cpp:739-822): the page normals are the
    normalised sums of the group's face normals, accumulated in the reference's order (bit-exact), with the source normal as the
    fallback for degenerate sums."""
    m = meshes[name]
    normals = m.normals.copy()
    normals[::7] = 0.0  # fallback of the fallback: (0, 0, 1)
    v = art.interleave(m.positions, normals)
    st = lib.default_builder_settings()
    st.preserveImportedNormals = 0
    ours = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS, settings=st)
    ref = full_oracle.build(v, m.indices, clodb200_lib=lib.path, recompute_normals=True)
    _assert_identical(ref, ours)
    kept = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS)
    assert not np.array_equal(kept.meshPages, ours.meshPages)


def test_recomputed_group_normals_multi_group(lib, full_oracle):
    m = meshgen.grid(200, seed=3)
    v = art.interleave(m.positions, m.normals)
    st = lib.default_builder_settings()
    st.preserveImportedNormals = 0
    ours = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS, settings=st)
    assert np.bincount(ours.groups["depth"]).max() > 1
    _assert_identical(full_oracle.build(v, m.indices, clodb200_lib=lib.path, recompute_normals=True), ours)


def test_skip_if_cached_probe(lib, meshes, tmp_path):
    """clodb200_cacheProbe is the skip-if-cached test of a batch build (CLodCacheLoader::TryLoadPrebuilt,
    CLodCacheLoader.cpp:218-234): true only for the same identity and build configuration, false for a missing, truncated or
    foreign cache; the names come from the restated naming functions."""
    from basicrenderer_b200 import cache

    m = meshes["grid64"]
    v = art.interleave(m.positions, m.normals)
    config_hash = cache.build_config_hash({})
    usdc = cache.cache_file_name("scene.gltf", "/mesh0", "", config_hash)
    stem = usdc[:-5]
    d = str(tmp_path)
    assert not lib.cache_probe(d, stem + ".clodblob", "scene.gltf", "/mesh0", "", config_hash)
    a = lib.build_artifacts(v, m.indices, art.VERTEX_NORMALS, keep_handle=True)
    lib.save_cache(a, d, stem + ".clodbin", stem + ".clodblob", "scene.gltf", "/mesh0", "", config_hash)
    lib.free_artifacts(a)
    assert lib.cache_probe(d, stem + ".clodblob", "scene.gltf", "/mesh0", "", config_hash)
    assert not lib.cache_probe(d, stem + ".clodblob", "scene.gltf", "/mesh1", "", config_hash)       # another primitive
    assert not lib.cache_probe(d, stem + ".clodblob", "scene.gltf", "/mesh0", "", config_hash ^ 1)   # another build configuration
    blob = (tmp_path / (stem + ".clodblob")).read_bytes()
    (tmp_path / "cut.clodblob").write_bytes(blob[:-3])
    assert not lib.cache_probe(d, "cut.clodblob", "scene.gltf", "/mesh0", "", config_hash)            # truncated blob
    os.remove(tmp_path / (stem + ".clodbin"))
    assert not lib.cache_probe(d, stem + ".clodblob", "scene.gltf", "/mesh0", "", config_hash)         # container gone
