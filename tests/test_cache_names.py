"""Cache naming (CLodCache.cpp:62-80, 586-633): the library's C entry points against an independent
Python restatement of the published Boost.ContainerHash algorithm. PARITY UNPINNED against Boost itself (not in this image, not
vendored by the reference): tests/cpp/boost_names_check.cpp prints the same quantities from the real header for a maintainer."""
import ctypes as C
import os

import pytest

from basicrenderer_b200 import build, cache


@pytest.fixture(scope="module")
def names():
    lib = C.CDLL(build.build_emu())  # host-only code: the same source is compiled into the product library
    lib.clodb200_cacheBuildConfigHash.restype = C.c_uint64
    lib.clodb200_cacheFileName.restype = C.c_size_t
    lib.clodb200_cacheFileName.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]
    lib.clodb200_cacheSubdirectory.restype = C.c_size_t
    lib.clodb200_cacheSubdirectory.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    return lib


def _name(fn, *args):
    n = fn(*args, None, 0)
    buf = C.create_string_buffer(n)
    assert fn(*args, buf, n) == n
    return buf.value.decode()


VOXEL_VARS = [v for v in cache.CONFIG_HASH_ENVIRONMENT]


def test_config_hash_follows_the_environment(names, monkeypatch):
    for v in VOXEL_VARS:
        monkeypatch.delenv(v, raising=False)
    base = names.clodb200_cacheBuildConfigHash()
    assert base == cache.build_config_hash({}) and base != 0
    monkeypatch.setenv("BASICRENDERER_CLOD_VOXEL_MODE", "mesh")
    changed = names.clodb200_cacheBuildConfigHash()
    assert changed != base and changed == cache.build_config_hash({"BASICRENDERER_CLOD_VOXEL_MODE": "mesh"})
    assert len(cache.CONFIG_HASH_CONSTANTS) == 14 and len(cache.CONFIG_HASH_ENVIRONMENT) == 11


@pytest.mark.parametrize("source,prim,subset", [
    ("models/zorah.usd", "/World/Mesh_0", ""),
    ("C:\\assets\\Bistro v5.2\\bistro.gltf", "mesh[12]/primitive[3]", "subset_a"),
    ("", "", ""),
    ("a", "ab", "abc"),
    ("abcdefg", "abcdefgh", "abcdefghi"),
    ("x" * 31, "y" * 32, "z" * 33),
])
def test_file_and_directory_names(names, source, prim, subset):
    h = 0x1234_5678_9ABC_DEF0
    assert _name(names.clodb200_cacheFileName, source.encode(), prim.encode(), subset.encode(), h) == cache.cache_file_name(source, prim, subset, h)
    got = _name(names.clodb200_cacheSubdirectory, source.encode())
    assert got == cache.cache_subdirectory(source)
    assert got.startswith("clod/") and all(c.isalnum() or c in "_-/" for c in got)


def test_stem_rules(names):
    assert _name(names.clodb200_cacheSubdirectory, b"").startswith("clod/scene_")
    assert _name(names.clodb200_cacheSubdirectory, b"scenes/My Scene (v2).usdc").startswith("clod/My_Scene__v2__")
    assert _name(names.clodb200_cacheSubdirectory, b"scenes/.hidden").startswith("clod/_hidden_")
    assert _name(names.clodb200_cacheSubdirectory, "t\u00e9st.glb".encode()).startswith("clod/t_st_")


def test_string_hash_tail_cases():
    """mulxp1_hash reads the tail through overlapping loads: every length 0..17 must give a distinct value, and equal prefixes of
    different length must not collide."""
    seen = {cache.boost_hash_string(b"abcdefghijklmnopq"[:n]) for n in range(18)}
    assert len(seen) == 18
    assert cache.boost_hash_string(b"\0") != cache.boost_hash_string(b"\0\0") != cache.boost_hash_string(b"")
