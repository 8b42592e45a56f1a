import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The compiled reference (oracle/_ref/libclodref.so). Built here when /root/reference is mounted; on the GPU box
    the prebuilt library travels with the snapshot."""
    from oracle import clodref

    if not clodref.available():
        if not os.path.exists("/root/reference"):
            pytest.skip("oracle/_ref/libclodref.so missing and /root/reference not mounted")
        clodref.build()
    return clodref


def _emu_lib():
    from basicrenderer_b200 import build
    from basicrenderer_b200.api import ClodLib

    return ClodLib(build.build_emu())


def _gpu_lib():
    from basicrenderer_b200 import load

    return load(0)


_cache = {}


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def lib(request):
    """Stage tests run twice: against the development-only host emulation of the kernel sources (tests/emu, no GPU
    needed, checks stage logic) and against the product CUDA library through the C ABI (-m gpu)."""
    kind = request.param
    if kind not in _cache:
        _cache[kind] = _emu_lib() if kind == "emu" else _gpu_lib()
    return _cache[kind]


@pytest.fixture(scope="session")
def meshes():
    from basicrenderer_b200 import meshgen

    return {
        "grid64": meshgen.grid(64),
        "ico24": meshgen.icosphere(24),
        "ico16uv": meshgen.icosphere(16, True, True),
        "torus": meshgen.torus(96, 40, seed=3),
        "grid160": meshgen.grid(160, seed=5),
    }


_dag_cache = {}


@pytest.fixture(scope="session")
def ref_dag(oracle, meshes):
    """Reference DAG dumps (per-level stage inputs/outputs) keyed by mesh name."""
    import numpy as np

    def get(name):
        if name not in _dag_cache:
            m = meshes[name]
            _dag_cache[name] = oracle.dag_build(m.positions, m.indices, attributes=m.normals, attribute_weights=np.ones(3, np.float32), protect_mask=7)
        return _dag_cache[name]

    return get
