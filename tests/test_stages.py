"""Stage parity: each clodb200 stage against the compiled reference on the same inputs (bit-exact where stated)."""
import numpy as np
import pytest

MESHES = ["grid64", "ico24", "ico16uv", "torus", "grid160"]


@pytest.mark.parametrize("name", MESHES)
def test_position_remap_bit_exact(lib, oracle, meshes, name):
    m = meshes[name]
    assert np.array_equal(lib.position_remap(m.positions), oracle.position_remap(m.positions))


def test_position_remap_edge_cases(lib, oracle):
    rng = np.random.default_rng(1)
    p = rng.integers(-2, 3, size=(5000, 3)).astype(np.float32)  # many duplicates
    p[::7] *= -0.0  # signed zeros compare equal
    p[5] = [np.nan, 0, 0]
    p[6] = [np.nan, 0, 0]  # NaN never matches, not even itself
    got = lib.position_remap(p)
    assert np.array_equal(got, oracle.position_remap(p))
    assert got[5] == 5 and got[6] == 6
    # strided input (interleaved vertex buffer)
    inter = np.zeros((5000, 8), dtype=np.float32)
    inter[:, :3] = p
    assert np.array_equal(lib.position_remap(inter, stride=32, vertex_count=5000), got)
    assert lib.position_remap(np.zeros((0, 3), np.float32)).size == 0
    one = lib.position_remap(np.ones((1, 3), np.float32))
    assert one.tolist() == [0]


@pytest.mark.parametrize("name", MESHES)
def test_clusterize_bit_exact(lib, oracle, meshes, name):
    m = meshes[name]
    ro, rv, ri = oracle.clusterize(m.positions, m.indices)
    go, gv, gs, gi = lib.clusterize(m.positions, m.indices)
    assert np.array_equal(go, ro)
    assert np.array_equal(gv, rv)
    assert np.array_equal(gi, ri)
    assert (np.diff(go) // 3).max() <= 128 and gv.max() <= 128
    assert (gs == 0).all()


@pytest.mark.parametrize("name", ["grid160", "ico16uv", "torus"])
def test_clusterize_segments_match_reference_levels(lib, ref_dag, meshes, name):
    """Batched per-group re-clusterization == the reference's per-group clod::clusterize calls, every DAG level."""
    m = meshes[name]
    dag = ref_dag(name)
    cdepth = dag.get("cluster_depth")
    coff = dag.get("cluster_index_offsets")
    cidx = dag.get("cluster_indices")
    for level in range(dag.num_levels):
        so = dag.level(level, "simp_offsets")
        si = dag.level(level, "simp_indices")
        keep = np.diff(so) > 0
        if not keep.any():
            continue
        seg = np.concatenate([[0], np.cumsum(np.diff(so)[keep])]).astype(np.uint32) // 3
        go, gv, gs, gi = lib.clusterize(m.positions, si, segment_offsets=seg)
        sel = np.nonzero(cdepth == level + 1)[0]
        lo, hi = coff[sel[0]], coff[sel[-1] + 1]
        assert np.array_equal(gi, cidx[lo:hi]), f"level {level}"
        assert np.array_equal(go, coff[sel[0] : sel[-1] + 2] - lo), f"level {level}"
        assert np.array_equal(gv, dag.get("cluster_vertices")[sel])
        # clusters never straddle groups and come out group by group
        assert (np.diff(gs.astype(np.int64)) >= 0).all()


def test_clusterize_tiny_and_vertex_bound(lib, oracle):
    # unindexed soup: 128 triangles use 384 vertices => vertex-bound splits and tails
    rng = np.random.default_rng(3)
    for tcount in (1, 2, 41, 43, 127, 129, 300, 1000):
        pos = rng.random((tcount * 3, 3), dtype=np.float32)
        idx = np.arange(tcount * 3, dtype=np.uint32)
        ro, rv, ri = oracle.clusterize(pos, idx)
        go, gv, gs, gi = lib.clusterize(pos, idx)
        assert np.array_equal(go, ro) and np.array_equal(gv, rv) and np.array_equal(gi, ri), tcount


@pytest.mark.parametrize("name", MESHES)
def test_cluster_bounds_match_reference(lib, ref_dag, meshes, name):
    m = meshes[name]
    dag = ref_dag(name)
    sel = np.nonzero(dag.get("cluster_depth") == 0)[0]
    coff = dag.get("cluster_index_offsets")
    idx = dag.get("cluster_indices")[: coff[sel[-1] + 1]]
    got = lib.cluster_bounds(m.positions, idx, np.diff(coff[: sel[-1] + 2]))
    want = dag.get("cluster_bounds")[sel, :4]
    # north_star tolerance for float bounds: 1e-5 relative; the replayed sequential fit is in fact bit-exact
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    assert np.array_equal(got, want)
