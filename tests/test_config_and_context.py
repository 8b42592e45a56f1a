"""Boundary behaviour around the clodConfig struct and the per-thread build contexts.

* clodConfig fields whose reference behaviour is not built fail loudly (the struct is a layout-identical mirror of clodConfig,
  so a memcpy'd clodDefaultConfig() arrives with cluster_spatial = false, clusterlod.h:762);
* clodConfig::partition_refined_split_count is incremented like the reference's instrumentation counter (clusterlod.h:474-475);
* a host thread that builds and exits gives its context back; shutdown followed by init makes other threads rebuild theirs.
"""
import ctypes as C
import threading

import numpy as np
import pytest

from basicrenderer_b200 import ClodbError, meshgen


@pytest.mark.parametrize("field,value,needle", [
    ("cluster_spatial", False, "cluster_spatial"),
    ("simplify_regularize", True, "simplify_regularize"),
    ("simplify_error_edge_limit", 0.5, "simplify_error_edge_limit"),
    ("max_triangles", 0, "meshlet limits"),
])
def test_unsupported_config_fields_fail_loudly(lib, field, value, needle):
    m = meshgen.grid(24, seed=1)
    cfg = lib.builder_config()
    setattr(cfg, field, value)
    with pytest.raises(ClodbError) as e:
        lib.build_dag(m.positions, m.indices, config=cfg)
    assert needle in str(e.value)
    # the library stays usable
    assert lib.build_dag(m.positions, m.indices).total_clusters > 0


def test_fallback_permissive_only_rejected_without_permissive(lib):
    m = meshgen.grid(24, seed=1)
    cfg = lib.builder_config()
    cfg.simplify_fallback_permissive = True
    cfg.simplify_permissive = False
    with pytest.raises(ClodbError):
        lib.build_dag(m.positions, m.indices, config=cfg)
    cfg.simplify_permissive = True  # the fallback is then a no-op in the reference as well (clusterlod.h:616)
    assert lib.build_dag(m.positions, m.indices, config=cfg).total_clusters > 0


def test_refined_split_counter(lib, oracle):
    """cap = 1 refined id per group forces splits on every level above the first; cap = 0 disables them."""
    m = meshgen.grid(160, seed=5)
    counter = C.c_size_t(0)
    cfg = lib.builder_config()
    cfg.partition_size = 16
    cfg.partition_max_refined_groups = 1
    cfg.partition_refined_split_count = C.cast(C.pointer(counter), C.c_void_p)
    rec = lib.build_dag(m.positions, m.indices, config=cfg)
    assert counter.value > 0
    # the reference counts the same kind of event on the same input (different grouping, so only the order of magnitude)
    ref_counter = C.c_size_t(0)
    rcfg = oracle.builder_config()
    rcfg.partition_size = 16
    rcfg.partition_max_refined_groups = 1
    rcfg.partition_refined_split_count = C.cast(C.pointer(ref_counter), C.c_void_p)
    oracle.dag_build(m.positions, m.indices, config=rcfg, dump=False)
    assert ref_counter.value > 0 and 0.5 * ref_counter.value <= counter.value <= 2 * ref_counter.value
    first = counter.value
    cfg.partition_max_refined_groups = 0
    lib.build_dag(m.positions, m.indices, config=cfg)
    assert counter.value == first
    assert rec.total_clusters > 0


def test_threads_release_their_context_and_survive_reinit(lib):
    m = meshgen.grid(40, seed=2)
    want = lib.build_dag(m.positions, m.indices).cluster_indices.tobytes()
    got = []

    def work():
        got.append(lib.build_dag(m.positions, m.indices).cluster_indices.tobytes())

    for _ in range(3):  # each thread builds, exits, and its context (slabs, stream, staging) is torn down by its owner
        t = threading.Thread(target=work)
        t.start()
        t.join()
    assert got == [want] * 3

    # a long-lived worker keeps its context across a shutdown/init of the library and rebuilds it on the next call
    go, done = threading.Event(), threading.Event()
    out = []

    def worker():
        out.append(lib.build_dag(m.positions, m.indices).cluster_indices.tobytes())
        done.set()
        go.wait()
        out.append(lib.build_dag(m.positions, m.indices).cluster_indices.tobytes())

    t = threading.Thread(target=worker)
    t.start()
    done.wait()
    lib.shutdown()
    lib.init(0)
    go.set()
    t.join()
    assert out == [want, want]
    assert lib.build_dag(m.positions, m.indices).cluster_indices.tobytes() == want
