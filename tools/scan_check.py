"""Debug driver: exclusive scan primitive against numpy at large sizes.  usage: python tools/scan_check.py n [n ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import load  # noqa: E402

lib = load(0)
rng = np.random.default_rng(1)
for n in [int(float(x)) for x in sys.argv[1:]]:
    v = (rng.random(n) < 0.4).astype(np.uint32)
    ref = np.concatenate([[0], np.cumsum(v, dtype=np.uint64)[:-1]]).astype(np.uint32)
    for rep in range(3):
        out, total, ms = lib.prim_exclusive_scan_u32(v, repeat=5)
        bad = np.nonzero(out != ref)[0]
        print(f"n {n}: total {total} (expected {int(v.sum())}) mismatches {bad.size}" + (f" first at {bad[0]} (tile {bad[0] // 8192}): got {out[bad[0]]} want {ref[bad[0]]}" if bad.size else "") + f" {ms:.3f} ms", flush=True)
