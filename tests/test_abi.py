"""The C-ABI library loads and exports every symbol include/clodb200.h declares; struct layouts equal the reference's."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "clodb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clodb200_[A-Za-z0-9]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    names = _declared_symbols()
    for required in ("clodb200_buildEx", "clodb200_build", "clodb200_builderConfig", "clodb200_localIndices", "clodb200_generatePositionRemap", "clodb200_clusterize", "clodb200_lockBoundary", "clodb200_simplifyGroups",
                     "clodb200_computeClusterBounds"):
        assert required in names


def test_product_library_exports_every_declared_symbol():
    from basicrenderer_b200 import build

    path = build.build_product()
    lib = C.CDLL(path)  # loads without a GPU: CUDA is only touched by clodb200_init
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from basicrenderer_b200 import ClodbError, load

    with pytest.raises(ClodbError):
        load(0)


def test_struct_layouts_match_reference(oracle):
    from basicrenderer_b200 import api
    from oracle import clodref

    assert C.sizeof(api.Config) == C.sizeof(clodref.ClodConfig) == clodref.lib().clodref_config_size()
    for (n1, t1), (n2, t2) in zip(api.Config._fields_, clodref.ClodConfig._fields_):
        assert n1 == n2 and C.sizeof(t1) == C.sizeof(t2)
    assert C.sizeof(api.Bounds) == 20 and C.sizeof(api.Cluster) == 48 and C.sizeof(api.Group) == 24 and C.sizeof(api.MeshDesc) == 88
