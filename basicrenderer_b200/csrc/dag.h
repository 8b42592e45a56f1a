// DAG driver interface (dag.cu).
#pragma once

#include "clodb.h"

namespace clodb
{

// Layout-compatible with clodCluster / clodGroup (clusterlod.h:105-141) so the C ABI can hand them to the reference's
// own callback type without repacking.
struct DagCluster
{
	int refined;
	float bounds[5]; // center xyz, radius, error
	const unsigned int* indices;
	size_t index_count;
	size_t vertex_count;
};

struct DagGroup
{
	int depth;
	float simplified[5];
};

// One whole DAG level for sinks that can take it at once (host arrays; cluster ids are level-local).
struct LevelBulk
{
	int depth;
	u32 cluster_count, group_count;
	const u32* group_cluster_offset; // group_count + 1
	const u32* group_clusters;       // cluster ids, group-major (callback order)
	const int* refined;              // per cluster
	const float* bounds5;            // per cluster: inherited sphere + error
	const float* precise4;           // per cluster: sphere of the cluster's own geometry
	bool use_precise;                // clusterlod.h:689 (optimize_bounds): precise sphere for clusters with refined != -1
	const float* group_bounds5;      // per group: simplified sphere + error (after the error rule)
	const u32* cluster_tri_offset;   // cluster_count + 1
	const u32* cluster_vertex_count; // per cluster
};

struct DagSink
{
	virtual ~DagSink() {}
	// Optional fast path: take the whole level in one call (the sink may use host threads) and return true after filling
	// group_ids[g] with what group() would have returned for group g, in order. Sinks that forward to a user callback keep the
	// serial per-group contract (clusterlod.h:895-926) and leave this alone.
	virtual bool emit_level_bulk(const LevelBulk& /*level*/, std::vector<int>& /*group_ids*/)
	{
		return false;
	}
	// returns the id stored as `refined` for clusters produced from this group
	virtual int group(const DagGroup& group, const DagCluster* clusters, size_t cluster_count, size_t task_index) = 0;
	// false: the level's index lists stay on the device (DagCluster::indices is null); the sink reads them from the
	// ClusterSet handed to begin_level, whose arrays live in the persist arena until the build ends
	virtual bool wants_indices() const
	{
		return true;
	}
	virtual void begin_level(const ClusterSet& /*level*/, int /*depth*/)
	{
	}
	// set by the driver before every group(): ids (inside the level) of the clusters passed, and the level's
	// cluster -> first triangle table (host copy)
	const u32* cluster_ids = nullptr;
	const u32* level_cluster_tri_offset = nullptr;
};

struct BuildStats
{
	u32 levels = 0;
	u32 groups = 0;
	size_t total_clusters = 0;
	size_t simplified_triangles = 0;
	size_t d2h_bytes = 0;
	u32 simplify_passes = 0;
	u32 simplify_rounds = 0;
	std::vector<u32> level_triangles, level_clusters, level_groups;
	size_t refined_splits = 0; // partitions split by the refined-id cap (clodConfig::partition_refined_split_count)
	std::vector<u32> level_passes, level_sloppy; // edge-collapse passes / groups sent through the sloppy fallback, per level
};

size_t build_dag(const Config& config, const DeviceMesh& mesh, const u32* indices_dev, size_t index_count, Workspace& ws, DagSink& sink, BuildStats& stats);

} // namespace clodb
