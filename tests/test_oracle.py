"""Pins the oracle: (1) the compiled reference passes meshoptimizer's own known-answer tests; (2) the plain-C restatement
(oracle/clod_oracle.c) agrees with the compiled reference and with the committed golden vectors. CPU only."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.fixture(scope="module")
def restatement():
    path = os.path.join(ROOT, "oracle", "_ref", "libclodoracle.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ROOT, "oracle", "clod_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restatement"])
    lib = C.CDLL(path)
    lib.clod_oracle_local_indices.restype = C.c_size_t
    lib.clod_oracle_refined_cap.restype = C.c_size_t
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _remap(lib, pos):
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.zeros(len(pos), np.uint32)
    assert lib.clod_oracle_position_remap(_p(out), _p(pos), C.c_size_t(len(pos)), C.c_size_t(pos.shape[1] * 4)) == 0
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="reference sources not mounted")
def test_reference_known_answer_tests(tmp_path):
    """meshoptimizer's demo/tests.cpp (runTests) built against the same sources the oracle is compiled from."""
    mo = os.path.join(REF, "ThirdParty", "meshoptimizer")
    main = tmp_path / "main.cpp"
    main.write_text("void runTests();\nint main() { runTests(); return 0; }\n")
    exe = tmp_path / "kat"
    srcs = sorted(glob.glob(os.path.join(mo, "src", "*.cpp")))
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(mo, "src"), str(main), os.path.join(mo, "demo", "tests.cpp")] + srcs + ["-o", str(exe)])
    subprocess.check_call([str(exe)])


def test_restated_remap_matches_reference(restatement, oracle, meshes):
    rng = np.random.default_rng(5)
    cases = [m.positions for m in meshes.values()]
    p = rng.integers(-2, 3, size=(4000, 3)).astype(np.float32)
    p[::5] *= -0.0
    p[7] = [np.nan, 1, 1]
    p[8] = [np.nan, 1, 1]
    cases.append(p)
    for pos in cases:
        assert np.array_equal(_remap(restatement, pos), oracle.position_remap(pos))


def test_restatement_matches_golden(restatement):
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        g = np.load(path)
        remap = _remap(restatement, g["positions"])
        assert np.array_equal(remap, g["remap"]), path
        locks = np.zeros(len(remap), np.uint8)
        attrs = np.ascontiguousarray(g["attributes"], np.float32)
        restatement.clod_oracle_protect_bits(_p(locks), _p(remap), _p(attrs), C.c_size_t(attrs.shape[1]), C.c_uint(int(g["protect_mask"][0])), C.c_size_t(len(remap)))
        assert np.array_equal(locks, g["protect_locks"]), path
        for level in range(int(g["num_levels"][0])):
            mi = np.ascontiguousarray(g[f"L{level}.merged_indices"], np.uint32)
            mo = np.ascontiguousarray(g[f"L{level}.merged_offsets"], np.uint32)
            assert restatement.clod_oracle_lock_boundary(_p(locks), _p(mi), _p(mo), C.c_size_t(len(mo) - 1), _p(remap), None, C.c_size_t(len(remap))) == 0
            assert np.array_equal(locks, g[f"L{level}.locks"]), (path, level)
        # local indices: vertex counts of the callback stream
        coff = g["out.cluster_index_offsets"]
        idx = np.ascontiguousarray(g["out.cluster_indices"], np.uint32)
        for c in range(0, len(coff) - 1, 7):
            seg = np.ascontiguousarray(idx[coff[c] : coff[c + 1]])
            v = np.zeros(len(seg), np.uint32)
            t = np.zeros(len(seg), np.uint8)
            n = restatement.clod_oracle_local_indices(_p(v), _p(t), _p(seg), C.c_size_t(len(seg)))
            assert n == g["out.cluster_vertex_count"][c]
            assert np.array_equal(v[:n][t], seg)


def test_restated_local_indices_matches_reference(restatement, oracle):
    rng = np.random.default_rng(9)
    for n in (3, 30, 384):
        seg = rng.integers(0, 90, size=n).astype(np.uint32)
        v = np.zeros(n, np.uint32)
        t = np.zeros(n, np.uint8)
        k = restatement.clod_oracle_local_indices(_p(v), _p(t), _p(seg), C.c_size_t(n))
        rv, rt = oracle.local_indices(seg)
        assert k == len(rv) and np.array_equal(v[:k], rv) and np.array_equal(t, rt)


def test_restated_error_rule(restatement):
    out = C.c_float()
    f = restatement.clod_oracle_error_rule
    f.argtypes = [C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
    assert f(300, 150, 0.25, 0.1, 0.85, 1.5, 0.0, C.byref(out)) == 0 and out.value == np.float32(0.25)
    assert f(300, 150, 0.05, 0.1, 0.85, 1.5, 0.0, C.byref(out)) == 0 and out.value == np.float32(np.float32(0.1) * np.float32(1.5))
    assert f(300, 258, 0.05, 0.1, 0.85, 1.5, 0.0, C.byref(out)) == 1 and out.value == np.finfo(np.float32).max
    assert f(300, 0, 0.05, 0.1, 0.85, 1.5, 0.0, C.byref(out)) == 1


def test_restated_refined_cap(restatement):
    refined = np.array([5, 5, 7, -1, 9, 7, 11, 13, 5, 15, 17, 19, 21, 23], np.int32)
    order = np.zeros(len(refined), np.uint32)
    group = np.zeros(len(refined), np.uint32)
    n = restatement.clod_oracle_refined_cap(_p(refined), C.c_size_t(len(refined)), C.c_size_t(8), _p(order), _p(group))
    # 11 distinct keys > 8 => two groups, bucket-major order (first-seen keys: 5,7,-1,9,11,13,15,17 | 19,21,23)
    assert n == 2
    assert refined[order].tolist() == [5, 5, 5, 7, 7, -1, 9, 11, 13, 15, 17, 19, 21, 23]
    assert group.tolist() == [0] * 11 + [1] * 3
    n = restatement.clod_oracle_refined_cap(_p(refined[:6]), C.c_size_t(6), C.c_size_t(8), _p(order), _p(group))
    assert n == 1 and order[:6].tolist() == list(range(6))
