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


@pytest.mark.parametrize("name", MESHES)
def test_lock_boundary_bit_exact(lib, ref_dag, meshes, name):
    dag = ref_dag(name)
    locks = dag.get("protect_locks")
    for level in range(dag.num_levels):
        got = lib.lock_boundary(locks, dag.level(level, "merged_indices"), dag.level(level, "merged_offsets"), dag.get("remap"))
        locks = dag.level(level, "locks")
        assert np.array_equal(got, locks), f"level {level}"


@pytest.mark.parametrize("name", MESHES)
def test_simplify_groups_match_reference(lib, ref_dag, meshes, name):
    """Batched group simplification vs the reference's per-group meshopt_simplifyWithAttributes, every DAG level.

    The dependency-wavefront collapse scheduling reproduces the serial greedy scan, and quadrics are accumulated in the
    reference's order, so the simplified index lists and errors are expected to be bit-identical (a stronger statement
    than the +-2 % triangle / 1.05x error invariants of the north star, which are asserted as well)."""
    m = meshes[name]
    dag = ref_dag(name)
    w = np.ones(3, np.float32)
    for level in range(dag.num_levels):
        mi, mo = dag.level(level, "merged_indices"), dag.level(level, "merged_offsets")
        si, so, se = lib.simplify_groups(m.positions, mi, mo, dag.level(level, "locks"), attributes=m.normals, attribute_weights=w)
        rso, rsi, rerr = dag.level(level, "simp_offsets"), dag.level(level, "simp_indices"), dag.level(level, "group_error")
        term = dag.level(level, "group_terminal").astype(bool)
        for g in range(len(mo) - 1):
            if term[g]:
                continue
            want = rsi[rso[g] : rso[g + 1]]
            got = si[so[g] : so[g + 1]]
            assert abs(got.size - want.size) <= 0.02 * want.size + 3, (level, g)
            assert se[g] <= rerr[g] * 1.05 + 1e-12, (level, g)
            assert np.array_equal(got, want), (level, g)
            assert se[g] == rerr[g], (level, g)


def _check_dag_simplify(lib, oracle, m, attrs, weights, protect):
    dag = oracle.dag_build(m.positions, m.indices, attributes=attrs, attribute_weights=weights, protect_mask=protect)
    exact = total = 0
    for level in range(dag.num_levels):
        mi, mo = dag.level(level, "merged_indices"), dag.level(level, "merged_offsets")
        locks = dag.level(level, "locks")
        si, so, se = lib.simplify_groups(m.positions, mi, mo, locks, attributes=attrs, attribute_weights=weights)
        rso, rsi, rerr = dag.level(level, "simp_offsets"), dag.level(level, "simp_indices"), dag.level(level, "group_error")
        term = dag.level(level, "group_terminal").astype(bool)
        for g in range(len(mo) - 1):
            want = rsi[rso[g] : rso[g + 1]]
            if term[g]:
                continue
            got = si[so[g] : so[g + 1]]
            total += 1
            assert abs(got.size - want.size) <= 0.02 * want.size + 3, (level, g)
            assert se[g] <= rerr[g] * 1.05 + 1e-12, (level, g)
            exact += int(np.array_equal(got, want) and se[g] == rerr[g])
    return exact, total


def test_simplify_protected_uv_seams(lib, oracle, meshes):
    """UV charts give different attributes on both sides of a position seam => protect bits, Seam/Locked kinds."""
    m = meshes["ico16uv"]
    attrs = np.ascontiguousarray(m.vertices[:, 3:8])
    weights = np.array([1, 1, 1, 0.5, 0.5], np.float32)
    exact, total = _check_dag_simplify(lib, oracle, m, attrs, weights, 31)
    assert total > 0 and exact == total


def test_simplify_many_groups(lib, oracle):
    from basicrenderer_b200 import meshgen

    m = meshgen.grid(330, seed=11)  # ~218k triangles => 5 groups at depth 0
    exact, total = _check_dag_simplify(lib, oracle, m, m.normals, np.ones(3, np.float32), 7)
    assert total >= 8 and exact == total


@pytest.mark.parametrize("name", ["grid64", "ico16uv", "torus"])
def test_sloppy_fallback_bit_exact(lib, oracle, meshes, name):
    """Groups whose edge collapse misses the target (here forced by locking random vertices) go through the sloppy
    fallback (clusterlod.h:567-599 -> meshopt_simplifySloppy): same triangles, same error as the reference, also when
    only some groups of a batched call need it."""
    m = meshes[name]
    T = m.indices.size // 3
    w = np.ones(3, np.float32)
    rng = np.random.default_rng(1)
    used_fallback = 0
    for p in (0.3, 0.6, 0.9, 1.0):
        locks = (rng.random(len(m.positions)) < p).astype(np.uint8)
        target = int(np.float32(T) * np.float32(0.5)) * 3
        want, werr = oracle.simplify(m.positions, m.indices, locks, target, attributes=m.normals, attribute_weights=w)
        cfg = oracle.builder_config()
        cfg.simplify_fallback_sloppy = False
        edge_only, _ = oracle.simplify(m.positions, m.indices, locks, target, attributes=m.normals, attribute_weights=w, config=cfg)
        used_fallback += int(edge_only.size > target)
        # batched call: group 0 = the whole mesh, group 1 = its first half (different target, usually no fallback at low p)
        half = (T // 2) * 3
        idx2 = np.concatenate([m.indices, m.indices[:half]])
        offs = np.array([0, m.indices.size, m.indices.size + half], np.uint32)
        si, so, se = lib.simplify_groups(m.positions, idx2, offs, locks, attributes=m.normals, attribute_weights=w)
        assert np.array_equal(si[so[0] : so[1]], want), p
        assert se[0] == np.float32(werr), p
        want1, werr1 = oracle.simplify(m.positions, m.indices[:half], locks, int(np.float32(T // 2) * np.float32(0.5)) * 3, attributes=m.normals, attribute_weights=w)
        assert np.array_equal(si[so[1] : so[2]], want1), p
        assert se[1] == np.float32(werr1), p
    assert used_fallback >= 2


@pytest.mark.parametrize("name", ["ico16uv", "grid64"])
def test_protect_bits_stage(lib, oracle, ref_dag, meshes, name):
    """a3 as a stage of its own (clusterlod.h:829-841): the seamed icosphere duplicates positions with different normals along the
    chart borders (bits set there); masks select attribute columns; -0 == +0 and NaN != NaN follow float comparison."""
    m = meshes[name]
    remap = oracle.position_remap(m.positions)
    assert np.array_equal(lib.protect_bits(m.normals, 7, remap), ref_dag(name).get("protect_locks"))
    # hand-made classes: vertex 1..4 share vertex 0's position class
    attrs = np.array([[0.0, 1.0, 2.0], [-0.0, 1.0, 2.0], [0.0, 1.5, 2.0], [0.0, 1.0, np.nan], [0.0, 1.0, 2.0], [9.0, 9.0, 9.0]], np.float32)
    remap = np.array([0, 0, 0, 0, 0, 5], np.uint32)
    assert lib.protect_bits(attrs, 7, remap).tolist() == [0, 0, 2, 2, 0, 0]
    assert lib.protect_bits(attrs, 1, remap).tolist() == [0, 0, 0, 0, 0, 0]  # column 0 only: -0 == +0
    assert lib.protect_bits(attrs, 2, remap).tolist() == [0, 0, 2, 0, 0, 0]
    assert lib.protect_bits(attrs, 4, remap, locks=np.array([1, 1, 1, 1, 1, 1], np.uint8)).tolist() == [1, 1, 1, 3, 1, 1]  # bits are OR-ed in
