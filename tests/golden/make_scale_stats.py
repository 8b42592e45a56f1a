"""Generates tests/golden/scale_emu_stats.json: per-level statistics of the DAG the development emulation (g++ build of the kernel
sources, tests/emu) produces for the large parity meshes. The CUDA library must reproduce them exactly (tests/test_scale_parity.py,
-m gpu): every stage is either bit-exact or an order-independent reduction, so CUDA and emulation build the same DAG. The
emulation needs minutes for the 10 M-triangle mesh, which is why the numbers are committed instead of recomputed by the test.

  python tests/golden/make_scale_stats.py            (about 6 minutes)
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("scale_parity", os.path.join(ROOT, "tools", "scale_parity.py"))
sp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sp)

from basicrenderer_b200 import build  # noqa: E402
from basicrenderer_b200.api import ClodLib  # noqa: E402

MESHES = ["grid:330:11", "ico:224", "grid:707:7", "grid:1300:5", "grid:2236:1234"]

if __name__ == "__main__":
    lib = ClodLib(build.build_emu())
    out = {}
    path = os.path.join(ROOT, "tests", "golden", "scale_emu_stats.json")
    for spec_ in (sys.argv[1:] or MESHES):
        st = sp.ours_stats(lib, sp.make_mesh(spec_))
        out[spec_] = {"level_triangles": [int(x) for x in st["level_triangles"]], "level_groups": [int(x) for x in st["level_groups"]],
                      "level_sloppy": [int(x) for x in st["level_sloppy"]], "level_passes": [int(x) for x in st["level_passes"]],
                      "groups": int(st["groups"]), "meshlets": int(st["meshlets"]), "level_max_error_bits": [int(x) for x in st["level_max_error"].view("uint32")]}
        print(spec_, out[spec_]["groups"], out[spec_]["meshlets"], f"{st['seconds']:.1f} s", flush=True)
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(out)
        with open(path, "w") as f:
            json.dump(old, f, indent=1)
