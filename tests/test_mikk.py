"""MikkTSpace tangent stream (GenerateMikkTangents, ClusterLODUtilities.cpp:655-737) against the reference compiled
unmodified into oracle/_ref/libclodref_mikk.so: bit-exact per vertex, on smooth meshes and on the edge cases the
algorithm special-cases (uv seams, mirrored charts, degenerate triangles, degenerate uv mappings, non-manifold edges,
high-valence fans, duplicated vertices, unreferenced vertices)."""
import ctypes as C
import os

import numpy as np
import pytest

from basicrenderer_b200 import meshgen
from oracle import clodfull

pytestmark = pytest.mark.skipif(not clodfull.mikk_available(), reason="oracle/_ref/libclodref_mikk.so not built")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _check(lib, vertices, indices):
    vertices = np.ascontiguousarray(vertices, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32)
    ref = clodfull.mikk_tangents(vertices, indices)
    got = lib.mikk_tangents(vertices, indices)
    assert (ref is None) == (got is None)
    if ref is not None:
        bad = np.nonzero((_bits(got) != _bits(ref)).any(axis=1))[0]
        assert bad.size == 0, (bad[:8], got[bad[:4]], ref[bad[:4]])
    return got


def _grid_uv(n, uv_scale=(1.0, 1.0), seed=3):
    """(n+1)^2 displaced grid with pos, normal, uv = xy * uv_scale."""
    m = meshgen.grid(n, seed=seed)
    uv = m.positions[:, :2] * np.asarray(uv_scale, np.float32)
    return np.ascontiguousarray(np.concatenate([m.vertices[:, :6], uv.astype(np.float32)], axis=1)), m.indices.copy()


@pytest.mark.parametrize("f", [3, 12, 40])
def test_icosphere_with_uv_seams(lib, f):
    m = meshgen.icosphere(f, displace=True, uv_atlas=True)
    _check(lib, m.vertices, m.indices)


def test_grid_and_mirrored_charts(lib):
    v, i = _grid_uv(48)
    _check(lib, v, i)
    # mirror u on the right half: orientation-preserving flag flips across the seam (welded vertices differ there too)
    v2 = v.copy()
    right = v2[:, 0] > 0.5
    v2[right, 6] = 1.0 - v2[right, 6]
    _check(lib, v2, i)
    # every chart mirrored
    v3 = v.copy()
    v3[:, 6] = -v3[:, 6]
    got = _check(lib, v3, i)
    assert (got[:, 3] == -1.0).all()


def test_duplicated_vertices_weld(lib):
    """Unindexed triangle soup: every corner has its own vertex; the welding step must merge them."""
    v, i = _grid_uv(24)
    soup_v = v[i]
    soup_i = np.arange(i.size, dtype=np.uint32)
    got = _check(lib, soup_v, soup_i)
    base = _check(lib, v, i)
    np.testing.assert_allclose(got, base[i], atol=1e-6)  # same tangent spaces; only the per-vertex float sums differ


def test_degenerate_triangles_and_uvs(lib):
    v, i = _grid_uv(32)
    tri = i.reshape(-1, 3).copy()
    rng = np.random.default_rng(5)
    # position-degenerate triangles: repeated index, and distinct vertices at the same position
    deg = rng.choice(len(tri), 60, replace=False)
    tri[deg[:30], 2] = tri[deg[:30], 1]
    extra = v[tri[deg[30:], 1]].copy()
    extra[:, 3:8] += 0.25  # same position, other normal/uv: welded apart but still degenerate
    base = len(v)
    v = np.concatenate([v, extra])
    tri[deg[30:], 2] = np.arange(base, base + len(extra), dtype=np.uint32)
    # uv-degenerate ("group with anything") triangles: collapse the uv of isolated triangles' vertices via fresh vertices
    anyt = rng.choice(np.setdiff1d(np.arange(len(tri)), deg), 80, replace=False)
    fresh = v[tri[anyt].reshape(-1)].copy()
    fresh[:, 6:8] = 0.5
    base = len(v)
    v = np.concatenate([v, fresh])
    tri[anyt] = np.arange(base, base + len(fresh), dtype=np.uint32).reshape(-1, 3)
    _check(lib, v, tri.reshape(-1))


def test_uv_degenerate_patch_inside_welded_mesh(lib):
    """A block of triangles whose uv mapping has zero area but whose vertices stay welded to their neighbours: they join
    the neighbouring groups and take their orientation (mikktspace.cpp:1160-1172)."""
    v, i = _grid_uv(20)
    n = 21
    v = v.copy()
    for r in range(8, 12):
        v[r * n + 6 : r * n + 12, 6] = 0.25  # constant u along short runs => zero uv area for triangles inside the run
    _check(lib, v, i)


def test_all_uvs_zero(lib):
    v, i = _grid_uv(16, uv_scale=(0.0, 0.0))
    got = _check(lib, v, i)
    assert (got[:, 3] == -1.0).all()  # no group forms: initial space (1,0,0) with bOrient = 0 (mikktspace.cpp:340-346)


def test_non_manifold_and_high_valence(lib):
    # fan of 80 triangles around one vertex (> the in-register fan limit), plus three extra sheets on one spoke edge
    k = 80
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    ring = np.stack([np.cos(ang), np.sin(ang), 0.1 * np.sin(3 * ang)], axis=1)
    pos = np.concatenate([[[0, 0, 0.3]], ring, [[0.5, 0.0, 1.0], [0.5, 0.0, -1.0], [0.6, 0.1, 0.8]]]).astype(np.float32)
    nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    uv = pos[:, :2] * 0.5 + 0.5
    v = np.concatenate([pos, nrm, uv], axis=1).astype(np.float32)
    tris = [[0, 1 + j, 1 + (j + 1) % k] for j in range(k)]
    tris += [[0, 1, k + 1], [1, 0, k + 2], [0, 1, k + 3]]  # the edge (0,1) is now shared by five triangles
    _check(lib, v, np.asarray(tris, np.uint32).reshape(-1))


def test_unreferenced_vertices_and_signed_zero(lib):
    v, i = _grid_uv(12)
    v = np.concatenate([v, v[:5] * 2.0]).astype(np.float32)  # never referenced: fallback tangent from the normal
    v[-1, 3:6] = 0.0  # zero normal: fallback of the fallback
    v[-2, 3:6] = [0.0, 0.0, 1.0]  # |n.z| >= 0.999 picks the other axis
    _check(lib, v, i)


def test_generator_refuses_like_the_reference(lib):
    v, i = _grid_uv(4)
    assert lib.mikk_tangents(v[:, :6], i) is None  # stride 24 < 32 (ClusterLODUtilities.cpp:665)
    assert lib.mikk_tangents(v, i[:0]) is None
    bad = i.copy()
    bad[3] = len(v)
    assert lib.mikk_tangents(v, bad) is None  # index out of range (:677-683)


def test_corner_values_match_genTangSpace(lib):
    """Per-corner values against mikktspace's own callback stream (what m_setTSpaceBasic receives)."""
    m = meshgen.icosphere(10, displace=True, uv_atlas=True)
    got, corners = lib.mikk_tangents(m.vertices, m.indices, corners=True)
    # rebuild the per-vertex result from the corner stream exactly as the reference's callback does (:630-653, 706-734)
    acc = np.zeros((m.vertex_count, 3), np.float32)
    sign = np.zeros(m.vertex_count, np.float32)
    for c, vi in enumerate(m.indices):
        acc[vi] += corners[c, :3]
        sign[vi] += corners[c, 3]
    inv = np.float32(1.0) / np.sqrt((acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1] + acc[:, 2] * acc[:, 2]).astype(np.float32))
    want = np.concatenate([acc * inv[:, None], np.where(sign < 0, -1.0, 1.0)[:, None]], axis=1).astype(np.float32)
    assert np.array_equal(_bits(got), _bits(want))


def test_acosf_matches_libm(lib):
    """mk_acosf (csrc/mikk.cu) against the C library's acosf, which the reference calls through the acos(float) overload
    (mikktspace.cpp:1421): bit-exact on a dense sample of [-1, 1] plus the branch boundaries of the algorithm."""
    olib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(clodfull.__file__)), "_ref", "libclodoracle.so"))
    olib.clod_oracle_acosf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(11)
    edge = np.array([1.0, -1.0, 0.0, -0.0, 0.5, -0.5, 1e-9, -1e-9, 2.0 ** -26, 0.49999997, 0.50000006, 0.99999994, -0.99999994], np.float32)
    near = np.concatenate([e.view(np.uint32) + np.arange(-64, 65, dtype=np.int64) for e in edge[[0, 4, 5, 8]].reshape(-1, 1)]).astype(np.uint32).view(np.float32)
    near = near[np.abs(near) <= 1]
    n = 1 << (22 if lib.path.endswith("libclodb200.so") else 16)
    xs = np.concatenate([rng.uniform(-1, 1, n).astype(np.float32), (1 - rng.uniform(0, 1, n // 4) ** 4).astype(np.float32), edge, near])
    want = np.zeros_like(xs)
    olib.clod_oracle_acosf(xs.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), xs.size)
    got = lib.prim_acosf(xs)
    assert np.array_equal(_bits(got), _bits(want))
