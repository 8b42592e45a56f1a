"""Micro-benchmark of the device-wide primitives (numbers are diagnostics, never bench values).
usage: python tools/prim_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from basicrenderer_b200 import load  # noqa: E402

lib = load(0)
rng = np.random.default_rng(1)
for n in (1 << 12, 1 << 16, 1 << 20, 1 << 22, 10_000_000, 30_000_000, 100_000_000):
    v = rng.integers(0, 3, n, dtype=np.uint32)
    out, total, ms = lib.prim_exclusive_scan_u32(v, repeat=11)
    assert total == int(v.sum(dtype=np.uint64) & 0xFFFFFFFF)
    print(f"scan u32 n={n:>10}: {ms * 1e3:8.1f} us  {n * 8 / ms / 1e6:8.1f} GB/s", flush=True)
for n in (1 << 20, 10_000_000, 30_000_000):
    v = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    out, ms = lib.prim_exclusive_max_scan_u64(v, repeat=11)
    print(f"max-scan u64 n={n:>10}: {ms * 1e3:8.1f} us  {n * 16 / ms / 1e6:8.1f} GB/s", flush=True)
for n in (1 << 16, 1 << 20, 10_000_000, 30_000_000):
    k = rng.integers(0, 1 << 30, n, dtype=np.uint64).astype(np.uint32)
    _, _, ms = lib.prim_sort_pairs_u32(k, np.arange(n, dtype=np.uint32), 0, 30, repeat=6)
    print(f"sort pairs u32 (30 bits) n={n:>10}: {ms * 1e3:8.1f} us  {n / ms / 1e3:8.1f} Mkeys/s  (incl. 2 restore copies)", flush=True)
