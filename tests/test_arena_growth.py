"""The device slabs are sized from the triangle count. Inputs that need more (e.g. triangle soups whose
vertices are shared by dozens of clusters) make the build entry points grow the exhausted slab and run the build again instead
of failing; the result does not depend on how often that happened. Exercised here by starting from a far too small temp slab
(CLODB200_TEMP_BYTES_PER_TRI / CLODB200_TEMP_BYTES_BASE) in a fresh process per entry point - so each entry point meets the
starved slab itself, the callback entry point included (its retry rule: restart only while nothing was delivered) - on the host
emulation (the retry is host code) and, with -m gpu, on the CUDA library (real cudaFree/cudaMalloc with async work in flight)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys, zlib
import numpy as np
sys.path.insert(0, "@ROOT@")
from basicrenderer_b200 import artifacts as art, build, meshgen, load
from basicrenderer_b200.api import ClodLib
kind, entry = sys.argv[1], sys.argv[2]
lib = ClodLib(build.build_emu()) if kind == "emu" else load(0)
m = meshgen.grid(300, seed=11)
w = np.ones(3, np.float32)
if entry == "build_ex":
    seen = []
    ids = set()
    def cb(g, c, t):
        seen.append(len(c))
        return len(seen)
    n = lib.build_ex(m.positions, m.indices, cb, attributes=m.normals, attribute_weights=w, protect_mask=7)
    print(n, sum(seen), len(seen))
elif entry == "build_dag":
    rec = lib.build_dag(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7)
    print(rec.total_clusters, zlib.crc32(np.asarray(rec.cluster_indices).tobytes()), rec.groups)
else:
    a = lib.build_artifacts(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS)
    print(len(a.groups), zlib.crc32(np.asarray(a.meshPages).tobytes()), len(a.nodes))
"""

STARVED = {"CLODB200_TEMP_BYTES_PER_TRI": "1", "CLODB200_TEMP_BYTES_BASE": "1000000"}  # ~4 MB for a build that needs ~60 MB


def _run(kind, entry, extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.check_output([sys.executable, "-c", WORKER.replace("@ROOT@", ROOT), kind, entry], env=env, text=True).split()


@pytest.mark.parametrize("kind", ["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("entry", ["build_ex", "build_dag", "build_artifacts"])
def test_builds_survive_a_too_small_slab_and_give_the_same_result(kind, entry):
    normal = _run(kind, entry, {})
    starved = _run(kind, entry, STARVED)
    assert starved == normal
    if entry == "build_ex":
        # every cluster delivered exactly once through the callbacks, no group delivered twice by a restarted build
        clusters, delivered, groups = (int(x) for x in normal)
        assert clusters == delivered
        dag = _run(kind, "build_dag", {})
        assert int(dag[0]) == clusters and int(dag[2]) == groups
