// clodb200 internal stage API (device pointers unless stated otherwise). One translation unit per stage; the C ABI in
// capi.cu and the DAG driver in dag.cu are built on these.
#pragma once

#include "prims.cuh"

namespace clodb
{

// Builder configuration actually consumed by the stages; mirrors the reference's clodConfig fields
// (clusterlod.h:15-71) that are live on the BasicRenderer path (ClusterLODUtilities.cpp:5426-5460).
struct Config
{
	u32 max_vertices = 128;
	u32 min_triangles = 64;
	u32 max_triangles = 128;
	float cluster_fill_weight = 0.5f;
	u32 partition_size = 384;
	u32 partition_max_refined_groups = 8;
	bool partition_sort = true;
	bool partition_spatial = true;
	float simplify_ratio = 0.5f;
	float simplify_threshold = 0.85f;
	float simplify_error_merge_previous = 1.5f;
	float simplify_error_merge_additive = 0.0f;
	float simplify_error_factor_sloppy = 100.f;
	bool simplify_permissive = true;
	bool simplify_fallback_sloppy = true;
	bool optimize_clusters = true;
	bool optimize_bounds = true;
};

// Device-resident input mesh (positions/attributes are tightly packed copies made at upload).
struct DeviceMesh
{
	const float* positions = nullptr; // 3 floats per vertex
	size_t vertex_count = 0;
	const float* attributes = nullptr; // attribute_stride floats per vertex (may be null)
	u32 attribute_stride = 0;          // floats per vertex in `attributes`
	u32 attribute_count = 0;           // leading floats used for simplification
	float attribute_weights[32] = {};
	u32 attribute_protect_mask = 0;
	const u8* vertex_lock = nullptr;
};

struct Workspace
{
	Arena persist; // level outputs, kept until the build finishes
	Arena temp;    // stage temporaries, stack discipline
	HostStage stage; // pinned host buffer receiving each level's cluster tables for the output callbacks
};

// ---- S1: position remap + protect bits (remap.cu) ----------------------------------------------------------------
// remap[i] = lowest j with bit-equal-under-== position (meshopt_generatePositionRemap, indexgenerator.cpp:442-465)
void position_remap(const float* positions, size_t vertex_count, u32* remap, Arena& temp);
// locks[i] |= 2 where a protected attribute differs from the canonical vertex (clusterlod.h:829-841)
void protect_bits(const float* attributes, u32 attribute_stride, u32 protect_mask, const u32* remap, size_t vertex_count, u8* locks);

// ---- a1: MikkTSpace tangent stream (mikk.cu) -------------------------------------------------------------------------
// GenerateMikkTangents (ClusterLODUtilities.cpp:655-737) over the interleaved vertex buffer (pos @0, normal @12, uv @24).
// Writes float4 {tangent xyz, sign} per vertex; false where the reference's generator would refuse the input.
// corner_tangents4 (optional, parity tests): the per-corner values handed to m_setTSpaceBasic, zeros for degenerate triangles.
size_t mikk_temp_bytes(size_t vertex_count, size_t index_count);
void mikk_acosf(const float* in, float* out, size_t n); // the corner-angle acosf on its own, for its parity test
bool mikk_tangents(const u8* vertices, u32 vertex_stride, size_t vertex_count, const u32* indices, size_t index_count, float* tangents4, Arena& temp, float* corner_tangents4 = nullptr);

// ---- S3: spatial clusterization (clusterize.cu) ---------------------------------------------------------------------
// Splits every segment [seg_offsets[s], seg_offsets[s+1]) of the triangle list independently into meshlets, exactly as
// clod::clusterize -> meshopt_buildMeshletsSpatial + meshopt_optimizeMeshlet would for that segment's index list
// (clusterlod.h:305-348; clusterizer.cpp:1380-1477, 1051-1124, 1682-1770).
struct ClusterSet
{
	u32 cluster_count = 0;
	u32 triangle_count = 0;
	u32* tri = nullptr;             // 3 u32 per triangle, cluster-major, optimized order
	u32* cluster_tri_offset = nullptr; // cluster_count + 1, in triangles
	u32* cluster_vertex_count = nullptr;
	u32* cluster_segment = nullptr; // source segment of each cluster
};
ClusterSet clusterize(const u32* tri, u32 triangle_count, const u32* seg_offsets_host, u32 segment_count, const float* positions, const Config& config, Workspace& ws);

// ---- S7: bounds (bounds.cu) -----------------------------------------------------------------------------------------
// per-cluster sphere of meshopt_computeClusterBounds (clusterizer.cpp:1479-1632); bounds = {cx,cy,cz,r} per cluster
void cluster_bounds(const u32* tri, const u32* cluster_tri_offset, u32 cluster_count, const float* positions, float* bounds4);
// sphere-of-spheres per group as clod::boundsMerge (clusterlod.h:283-303): out5 = {c, r, max error}
void group_bounds_merge(const float* cluster_bounds5, const u32* group_cluster_offset, const u32* group_clusters, u32 group_count, float* out5);

// ---- S4: grouping (partition.cu) ------------------------------------------------------------------------------------
struct GroupSet
{
	u32 group_count = 0;
	u32 cluster_count = 0;
	u32* group_clusters = nullptr;       // cluster ids, group-major (ascending cluster id inside a group unless cap-split)
	u32* group_cluster_offset = nullptr; // group_count + 1
	std::vector<u32> group_cluster_offset_host;
	u32 merge_rounds = 0;
	u32 refined_splits = 0; // partitions the refined-id cap had to split (clusterlod.h:474-475)
};
// clod::partition (clusterlod.h:350-510) for the pending clusters of one level (all K clusters of the ClusterSet)
GroupSet partition_clusters(const u32* tri, const u32* cluster_tri_offset, u32 cluster_count, const int* cluster_refined, const float* cluster_bounds5, const u32* remap, const float* positions, size_t vertex_count, const Config& config, Workspace& ws);

// The deterministic second half of clod::partition (clusterlod.h:396-507) for a GIVEN partition id per cluster (device array):
// spatial order of the partitions, cluster order inside them, refined-id cap split. "Bit-exact given the same partitions."
GroupSet partition_finish(const u32* cluster_part, u32 partition_count, u32 cluster_count, const int* cluster_refined, const float* cluster_bounds5, const Config& config, Workspace& ws);

// ---- S5: group assembly + boundary locks (groups.cu) --------------------------------------------------------------
// Gathers the triangles of each group (clusters listed group-major in group_clusters) into one contiguous run per group,
// as runIterationTask's merge (clusterlod.h:708-711). Returns the per-group triangle offsets (host) in out_offsets.
void gather_group_triangles(const u32* tri, const u32* cluster_tri_offset, const u32* group_clusters, u32 cluster_count, u32* gtri_out, u32* gc_tri_offset /* cluster_count + 1 */, Arena& temp);
// clod::lockBoundary (clusterlod.h:512-559): bit0 = position class touched by >= 2 groups, keeps bit1 (protect), ORs vertex_lock
void lock_boundary(const u32* gtri, const u32* tri_group_offsets_dev, u32 group_count, u32 triangle_count, const u32* remap, const u8* vertex_lock, size_t vertex_count, u8* locks, Arena& temp);

// ---- S8: group output (output.cu) -----------------------------------------------------------------------------------
// clodLocalIndices (clusterlod.h:972-1023) for a batch of clusters; vertices holds vertex_capacity slots per cluster
void local_indices(const u32* indices, const u64* cluster_index_offset, u32 cluster_count, u32 vertex_capacity, u32* vertices, u8* triangles, u32* vertex_count);

// ---- S6: simplification (simplify.cu) -------------------------------------------------------------------------------
struct SimplifyOutput
{
	u32 group_count = 0;
	u32 triangle_count = 0;
	u32* tri = nullptr;              // simplified triangles (global vertex ids), group-major
	u32* group_tri_offset = nullptr; // group_count + 1
	float* group_error = nullptr;    // absolute error per group (meshopt_SimplifyErrorAbsolute)
};
struct SimplifyStats
{
	u32 passes = 0;
	u32 rounds = 0;
	u32 max_rounds = 0;
	u32 window_extensions = 0;
	u32 sloppy_groups = 0;
};
extern thread_local SimplifyStats g_simplify_stats;
// One meshopt_simplifyWithAttributes(Sparse|ErrorAbsolute|Permissive) per group, all groups batched. gtri holds each
// group's merged index list back to back; locks is the per-vertex lock byte array of the level.
SimplifyOutput simplify_groups(const u32* gtri, const u32* group_tri_offset_host, u32 group_count, const DeviceMesh& mesh, const u32* global_remap, const u8* locks, const Config& config, Workspace& ws);

} // namespace clodb
