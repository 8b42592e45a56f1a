"""a8: meshopt_spatialSortRemap of the partitions + the refined-id cap split are deterministic integer
stages, "bit-exact given the same partitions". The reference's own meshopt_partitionClusters output for every multi-group level
of a reference DAG dump is fed to clodb200_partitionFinish; group order, cluster order inside the groups and the cap splits must
equal what clod::partition produced (clusterlod.h:396-507). Runs on the emulation and, with -m gpu, on the CUDA library (the
warp-cooperative cap kernel only exists there)."""
import ctypes as C

import numpy as np
import pytest

from basicrenderer_b200 import meshgen


def _reference_partition(oracle, dag, depth, partition_size):
    """meshopt_partitionClusters exactly as clod::partition calls it (clusterlod.h:387-399) for the level's pending clusters."""
    lib = oracle.lib()
    pending = dag.level(depth, "pending")
    remap = dag.get("remap")
    offs = dag.get("cluster_index_offsets")
    idx = dag.get("cluster_indices")
    counts = np.array([offs[c + 1] - offs[c] for c in pending], np.uint32)
    flat = np.concatenate([remap[idx[offs[c]:offs[c + 1]]] for c in pending]).astype(np.uint32)
    part = np.zeros(len(pending), np.uint32)
    fn = lib.meshopt_partitionClusters
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    return part, fn, flat, counts, pending, remap


@pytest.mark.parametrize("partition_size,cap", [(384, 8), (16, 8), (16, 2), (24, 1), (16, 0)])
def test_group_order_and_cap_split_given_reference_partitions(lib, oracle, partition_size, cap):
    m = meshgen.grid(200, seed=13) if partition_size < 100 else meshgen.grid(330, seed=11)
    w = np.ones(3, np.float32)
    rcfg = oracle.builder_config()
    rcfg.partition_size = partition_size
    rcfg.partition_max_refined_groups = cap
    dag = oracle.dag_build(m.positions, m.indices, attributes=m.normals, attribute_weights=w, protect_mask=7, config=rcfg)
    cfg = lib.builder_config()
    cfg.partition_size = partition_size
    cfg.partition_max_refined_groups = cap
    positions = np.ascontiguousarray(m.positions, np.float32)
    refined_all = dag.get("cluster_refined")
    bounds_all = dag.get("cluster_bounds")
    checked = splits = 0
    for depth in range(dag.num_levels):
        want_offsets = dag.level(depth, "group_offsets")
        want_clusters = dag.level(depth, "group_clusters")
        part, fn, flat, counts, pending, remap = _reference_partition(oracle, dag, depth, partition_size)
        if len(want_offsets) == 2 and len(pending) <= partition_size:
            continue  # small pending set: clod::partition returns it as one group without partitioning (clusterlod.h:352-385)
        P = fn(part.ctypes.data, flat.ctypes.data, flat.size, counts.ctypes.data, counts.size, positions.ctypes.data, remap.size, 12, partition_size)
        clusters, offsets = lib.partition_finish(part, P, refined_all[pending], bounds_all[pending], config=cfg)
        assert np.array_equal(offsets, want_offsets), (depth, P, len(offsets) - 1, len(want_offsets) - 1)
        assert np.array_equal(np.asarray(pending)[clusters], want_clusters), depth
        checked += 1
        splits += (len(want_offsets) - 1) - P
    assert checked >= 2
    if cap in (1, 2):
        assert splits > 0  # the cap really split partitions on this input
