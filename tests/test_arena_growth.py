"""The device slabs are sized from the triangle count. Inputs that need more (e.g. triangle soups whose
vertices are shared by dozens of clusters) make the build entry points grow the exhausted slab and run the build again instead
of failing; the result does not depend on how often that happened. Exercised here by starting from a far too small temp slab
(CLODB200_TEMP_BYTES_PER_TRI / CLODB200_TEMP_BYTES_BASE) in a fresh process, on the host emulation of the kernel sources (the retry is host code)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys, zlib
import numpy as np
sys.path.insert(0, "@ROOT@")
from basicrenderer_b200 import artifacts as art, build, meshgen
from basicrenderer_b200.api import ClodLib
lib = ClodLib(build.build_emu())
m = meshgen.grid(300, seed=11)
w = np.ones(3, np.float32)
a = lib.build_artifacts(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS)
rec = lib.build_dag(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7)
seen = []
n = lib.build_ex(m.positions, m.indices, lambda g, c, t: seen.append(len(c)) or len(seen), attributes=m.normals, attribute_weights=w, protect_mask=7)
print(len(a.groups), zlib.crc32(np.asarray(a.meshPages).tobytes()), rec.total_clusters, zlib.crc32(np.asarray(rec.cluster_indices).tobytes()), n, sum(seen))
"""


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.check_output([sys.executable, "-c", WORKER.replace("@ROOT@", ROOT)], env=env, text=True).split()


def test_builds_survive_a_too_small_slab_and_give_the_same_result():
    normal = _run({})
    starved = _run({"CLODB200_TEMP_BYTES_PER_TRI": "1", "CLODB200_TEMP_BYTES_BASE": "1000000"})  # ~4 MB for a build that needs ~60 MB
    assert starved == normal
    assert int(normal[4]) == int(normal[5]) == int(normal[2])  # every cluster delivered exactly once through the callbacks
