"""Kernel tuning helper: times one resident C2 build with a tuning variant of the library and prints the per-kernel event
profile of the kernels named on the command line.
usage: python tools/tune_variant.py <lib.so> <grid_n> <kernel-substring> [...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from basicrenderer_b200 import meshgen  # noqa: E402
from basicrenderer_b200.api import ClodLib  # noqa: E402

lib = ClodLib(os.path.abspath(sys.argv[1]), 0)
n = int(sys.argv[2])
names = sys.argv[3:]
m = meshgen.grid(n, seed=1234)
h = lib.upload_mesh(m.positions, m.indices, attributes=m.normals, attribute_weights=np.ones(3, np.float32), protect_mask=7)
for _ in range(2):
    lib.build_dag_resident(h, keep_indices=False)
lib.timer_start()
for _ in range(3):
    lib.build_dag_resident(h, keep_indices=False)
ms = lib.timer_stop_ms() / 3
lib.profile_enable(True)
lib.build_dag_resident(h, keep_indices=False)
rows = lib.profile_report()
lib.profile_enable(False)
sel = "  ".join(f"{r[0]}={r[2]:.2f}ms/{r[1]}" for r in rows if any(s in r[0] for s in names))
print(f"{os.path.basename(sys.argv[1])}: {ms:.1f} ms/build  {sel}", flush=True)
