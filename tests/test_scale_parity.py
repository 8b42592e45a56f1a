"""Whole-DAG parity at bench scale, through the C ABI, against the unmodified reference's clodBuildEx on the same mesh.

Round 1 shipped a CUDA-only divergence that no test could see (largest GPU test mesh: 218 k triangles; the bug needed >= 2 groups
that extend their candidate window in the same pass). These tests build C1 (1 M icosphere), a 1 M grid, a 3.4 M grid (> 60
groups per level, the large-scan code paths) and C2 itself (10 M, the benched mesh) on the GPU and assert, per DAG level:

  I6   triangles within +-2 % of the reference's (north_star bar; observed: <= 0.05 %)
  -    the number of groups that needed the sloppy fallback equals the reference's (0 on these meshes)
  I7   simplification error: median and 90th percentile of the depth-0 group errors (the only raw simplification errors the
       output carries) within 5 % of the reference's; the per-level MAX is reported and bounded by the reference's own
       sensitivity to its input order (tools/noise_floor.py: the unmodified reference against itself on the same surface with the
       triangle order reversed or x/y mirrored moves the per-level max by 1.2x - 1.9x), because the grouping stage is an
       invariant-bar stage (north_star) and every level above the first inherits the grouping of the levels below
  -    groups / meshlets / per-level triangles / per-level max error BITS equal to the development emulation's of the same
       kernel sources (tests/golden/scale_emu_stats.json): CUDA and emulation build the same DAG
"""
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("scale_parity", os.path.join(ROOT, "tools", "scale_parity.py"))
sp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sp)

GOLDEN = os.path.join(ROOT, "tests", "golden", "scale_emu_stats.json")

TRIANGLE_BAR = 0.02          # I6
QUANTILE_BAR = 1.05          # I7 on the median / p90 of a level's group errors
QUANTILE_MIN_GROUPS = 50     # below this the quantiles of a level are a handful of samples (C1: 19 groups spread over 15x in error)
MAX_ERROR_NOISE_BAR = 2.0    # the reference's own per-level max moves by up to 1.9x under input reordering


def _check(lib, oracle, spec, expect_golden):
    m = sp.make_mesh(spec)
    ours = sp.ours_stats(lib, m, keep_groups=True)
    ref = sp.ref_stats(m)
    rows, summary = sp.compare(ours, ref)
    assert summary["levels"] == summary["ref_levels"], summary
    for r in rows:
        assert abs(r["triangles"] - r["ref_triangles"]) <= TRIANGLE_BAR * r["ref_triangles"] + 2, r  # I6
        assert r["sloppy"] == r["ref_sloppy"], r
        assert r["groups"] <= 2 * r["ref_groups"] + 1, r
        if r["ref_max_error"] > 0:
            assert r["error_ratio"] <= MAX_ERROR_NOISE_BAR, r
    # I7 on distribution statistics of the raw simplification errors. Only depth 0 emits them: above it a group's error is
    # max(1.5 x inherited, own) (clusterlod.h:733), i.e. mostly the inherited maximum of the levels below.
    a = ours["group_error"][(ours["group_depth"] == 0) & (ours["group_error"] < sp.FLT_MAX)]
    b = ref["group_error"][(ref["group_depth"] == 0) & (ref["group_error"] < sp.FLT_MAX)]
    if min(a.size, b.size) >= QUANTILE_MIN_GROUPS:
        for q in (50, 90):
            assert np.percentile(a, q) <= QUANTILE_BAR * np.percentile(b, q), (spec, q, np.percentile(a, q), np.percentile(b, q))
    # terminal flags (FLT_MAX) exactly where simplified > 0.85 x input: same count per level as the reference
    for d in range(summary["levels"]):
        assert np.sum((ours["group_depth"] == d) & (ours["group_error"] >= sp.FLT_MAX)) == np.sum((ref["group_depth"] == d) & (ref["group_error"] >= sp.FLT_MAX)), (spec, d)
    if expect_golden:
        gold = json.load(open(GOLDEN))[spec]
        assert [int(x) for x in ours["level_triangles"]] == gold["level_triangles"]
        assert [int(x) for x in ours["level_groups"]] == gold["level_groups"]
        assert [int(x) for x in ours["level_sloppy"]] == gold["level_sloppy"]
        assert int(ours["groups"]) == gold["groups"] and int(ours["meshlets"]) == gold["meshlets"]
        assert [int(x) for x in ours["level_max_error"].view(np.uint32)] == gold["level_max_error_bits"]
    return summary


@pytest.mark.gpu
@pytest.mark.parametrize("spec", ["ico:224", "grid:707:7", "grid:1300:5", "grid:2236:1234"])
def test_gpu_dag_matches_reference_shape_and_emulation_at_scale(oracle, spec):
    from basicrenderer_b200 import load

    _check(load(0), oracle, spec, expect_golden=True)


def test_emulation_dag_matches_reference_shape(oracle):
    """The same assertions on the CPU suite's size (218 k triangles, 5 groups at depth 0)."""
    from basicrenderer_b200 import build
    from basicrenderer_b200.api import ClodLib

    _check(ClodLib(build.build_emu()), oracle, "grid:330:11", expect_golden=os.path.exists(GOLDEN))
