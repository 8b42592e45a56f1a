"""Device-wide primitives under every stage (csrc/prims.cuh): single-pass chained exclusive scan (u32 add, u64 max) and
stable LSD radix sort, against numpy on the same seeded inputs. Sizes straddle the tile sizes (2048 / 8192 elements),
the small/large kernel switch (2^21) and the look-back window (128 tiles)."""
import numpy as np
import pytest

SIZES = [0, 1, 5, 2047, 2048, 2049, 8191, 8193, 70001, (1 << 21) - 3, (1 << 21) + 8192 * 130 + 17]


@pytest.mark.parametrize("n", SIZES)
def test_exclusive_scan_u32(lib, n):
    rng = np.random.default_rng(n + 1)
    v = rng.integers(0, 7, n, dtype=np.uint32)
    out, total, _ = lib.prim_exclusive_scan_u32(v)
    ref = np.concatenate([[0], np.cumsum(v, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(out, ref[:-1])
    assert total == int(ref[-1])


@pytest.mark.parametrize("n", [0, 3, 2049, 300001, (1 << 21) + 12345])
def test_exclusive_max_scan_u64(lib, n):
    rng = np.random.default_rng(n + 7)
    v = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    # long runs without a new maximum, as the tagged-error scan of the simplifier sees them
    v[rng.random(n) < 0.9] = 0
    out, _ = lib.prim_exclusive_max_scan_u64(v)
    ref = np.concatenate([np.zeros(1, np.uint64), np.maximum.accumulate(v)])[:-1] if n else v
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("n,bits", [(0, 32), (1, 32), (2049, 32), (100003, 30), (1 << 20, 12), ((1 << 21) + 5, 32)])
def test_sort_pairs_u32_is_stable(lib, n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)  # callers keep the bits above bit_hi clear
    if n:
        keys[rng.random(n) < 0.3] = keys[0]  # duplicates: stability is observable through the values
    vals = np.arange(n, dtype=np.uint32)
    k, v, _ = lib.prim_sort_pairs_u32(keys, vals, 0, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v, vals[order])
    assert np.array_equal(k, keys[order])


_alias_round = [0]


def test_scan_descriptor_formats_do_not_alias(lib):
    """Regression: the 8-byte max-scan leaves 16-byte {value, tag} tile descriptors; a 4-byte scan reading 8-byte
    {tag, value} words from the same memory would see the high half of a stale value as its tag. The simplifier's tagged
    errors are (group + 1) << 32 | bits, which equals (epoch << 2) | status for small epochs: the first build of a process on
    a mesh with many groups took garbage look-back prefixes. Here every stale inclusive prefix carries
    exactly the inclusive tag of the next call's epoch, and a 4-byte scan is that next call. The epoch only ever
    jumps forward (descriptors of the same format rely on epochs never repeating)."""
    n = 2048 * 900
    rng = np.random.default_rng(5)
    v = rng.integers(0, 3, n, dtype=np.uint32)
    ref = np.concatenate([[0], np.cumsum(v, dtype=np.uint64)]).astype(np.uint32)
    for _ in range(40):
        _alias_round[0] += 1
        base = 50_000_000 + 10 * _alias_round[0]
        lib.prim_set_scan_epoch(base)
        # the max-scan takes epoch base + 1; its stale descriptors carry the inclusive tag of the scan that follows it
        stale = np.full(n, ((4 * (base + 2) + 2) << 32) | 0x3F800000, dtype=np.uint64)  # tag_inc(base + 2) : bits of 1.0f
        lib.prim_exclusive_max_scan_u64(stale)
        out, total, _ = lib.prim_exclusive_scan_u32(v)  # epoch base + 2
        assert np.array_equal(out, ref[:-1]) and total == int(ref[-1])
