"""bench.py's reference arm runs without a GPU (it times the compiled reference on the host cores): check the JSON line
the driver parses. The CUDA arm needs a device and is exercised on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(oracle):
    from oracle import clodfull

    if not clodfull.available(True):
        import pytest

        pytest.skip("reference L3 builder not built")
    env = dict(os.environ, CLODB200_REF_ICO_F="16")  # bounded sample of the default (C3) workload: 5 120 triangles
    env.pop("CLODB200_BENCH_WORKLOAD", None)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], env=env, text=True)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench._config("C3", 1)  # both arms name the same configuration
    assert d["steps"] == 1 and d["warmup"] == 0
    assert d["impl"] == "reference" and d["unit"] == "Mtris/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_byte_model_covers_the_kernels_it_can_report():
    sys.path.insert(0, ROOT)
    import bench

    for k in ("k_sa_chained", "k_wave_rounds", "k_partition_chained", "k_pivot_large", "k_rs_scatter<K>"):
        assert bench.KERNEL_BYTES_PER_THREAD[k] > 0
    traffic = json.load(open(os.path.join(ROOT, "profiles", "kernel_dram_traffic.json")))
    assert set(traffic["kernels"]) <= set(bench.KERNEL_BYTES_PER_THREAD)


def test_kernel_table_on_the_committed_event_profile():
    sys.path.insert(0, ROOT)
    import bench

    rows = []
    for line in open(os.path.join(ROOT, "profiles", "r01_c2_kernel_events_v16.csv")).read().splitlines()[1:]:
        name, count, ms, threads = line.rsplit(",", 3)
        rows.append((name, int(count), float(ms), int(threads)))
    table = bench._kernel_table(rows, 6515.4)
    assert len(table) == 10 and table[0]["kernel"] == rows[0][0]
    by_name = {t["kernel"]: t for t in table}
    assert 0.3 < by_name["k_sa_chained"]["frac"] < 0.7  # 8 positions x 44 B per thread
    assert by_name["k_wave_rounds"]["frac"] < 0.05  # latency bound, reported as such
    json.dumps(table)


def test_reference_arm_fans_meshes_out_for_n_gpus(oracle):
    """--gpus N on the CPU arm: N independent meshes per step, built concurrently, as an importer fans primitives out."""
    from oracle import clodfull

    if not clodfull.available(True):
        import pytest

        pytest.skip("reference L3 builder not built")
    env = dict(os.environ, CLODB200_REF_GRID="60", CLODB200_BENCH_WORKLOAD="C2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env, text=True)
    d = json.loads([l for l in out.splitlines() if l.strip()][0])
    assert d["n_gpus"] == 2 and "2 such meshes built concurrently" in d["cpu_baseline"]["sample"] and d["value"] > 0
