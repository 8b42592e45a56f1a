"""Synthetic workload generators: the torch construction used for the full-size C3 mesh reproduces the
numpy one, and chart borders weld by position (seams, not cracks)."""
import numpy as np

from basicrenderer_b200 import meshgen


def test_seamed_icosphere_torch_matches_numpy():
    m, tangents = meshgen.icosphere_seams_torch(24, device="cpu", chunk=100)
    r = meshgen.icosphere(24, True, True)
    assert np.array_equal(m.indices, r.indices)
    assert np.array_equal(m.vertices, r.vertices)
    assert m.flags == r.flags == (meshgen.VERTEX_NORMALS | meshgen.VERTEX_TEXCOORDS)
    assert tangents.shape == (m.vertex_count, 4) and np.isfinite(tangents).all()
    assert np.allclose(np.linalg.norm(tangents[:, :3], axis=1), 1.0, atol=1e-5)
    # tangent lies in the tangent plane of the shading normal
    assert np.abs((tangents[:, :3] * m.normals).sum(axis=1)).max() < 1e-5


def test_seamed_icosphere_welds_by_position():
    m, _ = meshgen.icosphere_seams_torch(12, device="cpu")
    uniq = np.unique(m.positions, axis=0)
    # 10 f^2 + 2 distinct positions on a closed frequency-f icosphere; the rest are seam duplicates with other UVs
    assert len(uniq) == 10 * 12 * 12 + 2
    assert m.vertex_count == 20 * (13 * 14 // 2)


def test_scene_batch_sizes_total():
    t = meshgen.scene_batch_sizes(512, 1.0e9)
    assert abs(t.sum() / 1.0e9 - 1.0) < 0.01 and t.min() >= 1000.0
