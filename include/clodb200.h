/* clodb200 — B200-native cluster-LOD DAG builder: C ABI.
 *
 * Drop-in boundary for the cluster-LOD build path of panthuncia/BasicRenderer. Every entry point takes plain pointers
 * and sizes (no C++/torch types) and returns 0 on success or a negative status; clodb200_last_error() returns the
 * message of the last failure on the calling thread. Pointers are HOST pointers unless the name says _device.
 * The library has no CPU code path: every call fails with CLODB200_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   clodb200_generatePositionRemap  <- meshopt_generatePositionRemap, ThirdParty/meshoptimizer/src/meshoptimizer.h
 *                                      (called at BasicRenderer/include/ThirdParty/meshoptimizer/clusterlod.h:826)
 *   clodb200_clusterize             <- clod::clusterize, clusterlod.h:305-348 (meshopt_buildMeshletsSpatial +
 *                                      meshopt_optimizeMeshlet), batched over independent index segments
 *   clodb200_computeClusterBounds   <- meshopt_computeClusterBounds via clod::boundsCompute, clusterlod.h:270-281
 *   clodb200_build / clodb200_buildEx <- clodBuild / clodBuildEx, clusterlod.h:159-184 (same structs, same callback)
 *   clodb200_localIndices           <- clodLocalIndices, clusterlod.h:184 (implementation :972-1023)
 *   clodb200_lockBoundary           <- clod::lockBoundary, clusterlod.h:512-559
 *   clodb200_simplifyGroups         <- clod::simplify, clusterlod.h:601-659 (meshopt_simplifyWithAttributes per group)
 */
#ifndef CLODB200_H
#define CLODB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define CLODB200_OK 0
#define CLODB200_ERR_NO_DEVICE (-1)
#define CLODB200_ERR_INVALID (-2)
#define CLODB200_ERR_RUNTIME (-3)

/* Same layout as struct clodConfig, clusterlod.h:15-71. */
typedef struct clodb200_config
{
	size_t max_vertices;
	size_t min_triangles;
	size_t max_triangles;
	bool partition_spatial;
	bool partition_sort;
	size_t partition_size;
	size_t partition_max_refined_groups;
	size_t* partition_refined_split_count;
	bool cluster_spatial;
	float cluster_fill_weight;
	float cluster_split_factor;
	float simplify_ratio;
	float simplify_threshold;
	float simplify_error_merge_previous;
	float simplify_error_merge_additive;
	float simplify_error_factor_sloppy;
	float simplify_error_edge_limit;
	bool simplify_permissive;
	bool simplify_fallback_permissive;
	bool simplify_fallback_sloppy;
	bool simplify_regularize;
	bool optimize_bounds;
	bool optimize_clusters;
} clodb200_config;

const char* clodb200_last_error(void);
/* Selects the CUDA device for this process (one process per GPU) and creates the build stream. */
int clodb200_init(int device);
void clodb200_shutdown(void);
/* Kernel launches issued by this library since clodb200_init (bench.py's gpu_launches). */
uint64_t clodb200_launch_count(void);

/* clodDefaultConfig(max_triangles) followed by the BasicRenderer overrides (ClusterLODUtilities.cpp:5426-5460). */
clodb200_config clodb200_builderConfig(void);

/* remap[i] = lowest index with the same position (IEEE ==). positions_stride in bytes (>= 12, multiple of 4). */
int clodb200_generatePositionRemap(unsigned int* remap, const float* positions, size_t vertex_count, size_t positions_stride);

/* Splits each segment [segment_offsets[s], segment_offsets[s+1]) (in triangles) of `indices` into meshlets.
 * Outputs: cluster_index_counts/cluster_vertex_counts/cluster_segments hold one entry per cluster (capacity
 * index_count / 3 entries each), out_indices receives index_count cluster-major indices. Returns the cluster count in
 * *out_cluster_count. segment_offsets == NULL means one segment covering everything. */
int clodb200_clusterize(const clodb200_config* config, const unsigned int* indices, size_t index_count, const unsigned int* segment_offsets, size_t segment_count,
    const float* positions, size_t vertex_count, size_t positions_stride,
    unsigned int* cluster_index_counts, unsigned int* cluster_vertex_counts, unsigned int* cluster_segments, unsigned int* out_indices, size_t* out_cluster_count);

/* Bounding sphere {cx, cy, cz, r} per cluster; clusters are given as cluster-major indices + per-cluster index counts. */
int clodb200_computeClusterBounds(const unsigned int* indices, const unsigned int* cluster_index_counts, size_t cluster_count,
    const float* positions, size_t vertex_count, size_t positions_stride, float* out_bounds4);

/* clodLocalIndices (clusterlod.h:972-1023): vertices[] = distinct indices in first-occurrence order, triangles[i] = position
 * of indices[i] in vertices[]; returns the number of unique vertices (0 on failure). index_count <= 768, at most 256
 * distinct vertices, exactly as the reference's unsigned char triangle ids imply. */
size_t clodb200_localIndices(unsigned int* vertices, unsigned char* triangles, const unsigned int* indices, size_t index_count);
/* The same for many clusters in one call: cluster c owns indices [cluster_index_offsets[c], cluster_index_offsets[c+1]);
 * out_vertices holds vertex_capacity slots per cluster, out_triangles one byte per index, out_vertex_counts one per cluster. */
int clodb200_localIndicesBatch(const unsigned int* indices, const uint64_t* cluster_index_offsets, size_t cluster_count, size_t vertex_capacity,
    unsigned int* out_vertices, unsigned char* out_triangles, unsigned int* out_vertex_counts);

/* clod::lockBoundary (clusterlod.h:512-559) for one DAG level. indices holds the merged index lists of all groups back
 * to back, group_index_offsets[group_count + 1] delimits them. locks[vertex_count] is updated in place: bit0 = position
 * shared by >= 2 groups, bit1 (protect) kept, vertex_lock (optional) ORed in. remap = position remap of the mesh. */
int clodb200_lockBoundary(unsigned char* locks, const unsigned int* indices, const unsigned int* group_index_offsets, size_t group_count,
    const unsigned int* remap, const unsigned char* vertex_lock, size_t vertex_count);

/* clod::simplify (clusterlod.h:601-659: meshopt_simplifyWithAttributes with Sparse|ErrorAbsolute|Permissive) for every
 * group of a DAG level in one batched call. Each group g is simplified towards size_t(T_g * simplify_ratio) triangles.
 * out_indices (capacity = total index count) receives the simplified lists back to back, out_group_index_counts[g] and
 * out_group_errors[g] the per-group result size and absolute error. attributes may be NULL (attribute_count 0). */
int clodb200_simplifyGroups(const clodb200_config* config, const unsigned int* indices, const unsigned int* group_index_offsets, size_t group_count,
    const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count,
    const unsigned char* locks, unsigned int* out_indices, unsigned int* out_group_index_counts, float* out_group_errors);

/* ---- DAG build: drop-in for clodBuild / clodBuildEx (clusterlod.h:159-184) -------------------------------------- */

/* Same layout as struct clodMesh, clusterlod.h:73-99. Pointers are host pointers, borrowed for the duration of the call. */
typedef struct clodb200_mesh
{
	const unsigned int* indices;
	size_t index_count;
	size_t vertex_count;
	const float* vertex_positions;
	size_t vertex_positions_stride;
	const float* vertex_attributes;
	size_t vertex_attributes_stride;
	const unsigned char* vertex_lock;
	const float* attribute_weights;
	size_t attribute_count;
	unsigned int attribute_protect_mask;
} clodb200_mesh;

/* Same layouts as clodBounds / clodCluster / clodGroup, clusterlod.h:105-141. */
typedef struct clodb200_bounds
{
	float center[3];
	float radius;
	float error;
} clodb200_bounds;

typedef struct clodb200_cluster
{
	int refined;
	clodb200_bounds bounds;
	const unsigned int* indices; /* valid only during the callback, like the reference's */
	size_t index_count;
	size_t vertex_count;
} clodb200_cluster;

typedef struct clodb200_group
{
	int depth;
	clodb200_bounds simplified;
} clodb200_group;

/* Same signatures as clodOutput / clodOutputEx (clusterlod.h:146-148). Called serially on the calling thread, depth by
 * depth, groups in partition order; the return value becomes clodb200_cluster::refined of the clusters simplified from
 * that group. thread_index is always 0 (the per-group iteration tasks run on the GPU). */
typedef int (*clodb200_output)(void* output_context, clodb200_group group, const clodb200_cluster* clusters, size_t cluster_count);
typedef int (*clodb200_outputEx)(void* output_context, clodb200_group group, const clodb200_cluster* clusters, size_t cluster_count, size_t task_index, unsigned int thread_index);

/* Returns the total number of clusters produced; 0 for empty/invalid geometry (one line on stderr, as clusterlod.h:803-816)
 * or on failure (clodb200_last_error() is then non-empty). parallel_config of the reference is accepted and ignored. */
size_t clodb200_build(clodb200_config config, clodb200_mesh mesh, void* output_context, clodb200_output output_callback);
size_t clodb200_buildEx(clodb200_config config, clodb200_mesh mesh, void* output_context, clodb200_outputEx output_callback, const void* parallel_config);

/* Device-resident variant: upload once, build many times (benchmarks; scene batches). */
typedef struct clodb200_device_mesh clodb200_device_mesh;
clodb200_device_mesh* clodb200_meshUpload(clodb200_mesh mesh);
void clodb200_meshFree(clodb200_device_mesh* mesh);
size_t clodb200_meshBuildEx(clodb200_config config, const clodb200_device_mesh* mesh, void* output_context, clodb200_outputEx output_callback);

/* Recording build: runs clodb200_meshBuildEx / clodb200_buildEx with an internal callback that stores the whole output
 * stream, returned as named arrays ("group_depth" i32, "group_simplified" f32x5, "group_cluster_offsets" u32,
 * "cluster_refined" i32, "cluster_bounds" f32x5, "cluster_vertex_count" u32, "cluster_index_offsets" u64,
 * "cluster_indices" u32, "level_triangles"/"level_clusters"/"level_groups" u32, "stats" u64). */
typedef struct clodb200_record clodb200_record;
clodb200_record* clodb200_buildRecorded(clodb200_config config, clodb200_mesh mesh);
clodb200_record* clodb200_meshBuildRecorded(clodb200_config config, const clodb200_device_mesh* mesh, int keep_indices);
int clodb200_recordGet(const clodb200_record* record, const char* name, const void** out_ptr, size_t* out_bytes);
void clodb200_recordFree(clodb200_record* record);

/* CUDA-event stopwatch on the build stream: Start records an event, Stop records another, synchronises and returns the
 * elapsed device time in milliseconds (bench.py brackets its timed steps with these). */
void clodb200_timerStart(void);
float clodb200_timerStop(void);

/* Per-kernel CUDA-event timing on the build stream (off by default; adds two event records per launch).
 * clodb200_profileReport drains the recorded spans into "kernel,launches,total_ms,total_threads" lines, slowest first; returns the
 * number of bytes needed (including the terminator); at most `capacity` bytes are written. */
void clodb200_profileEnable(int enable);
size_t clodb200_profileReport(char* buffer, size_t capacity);

/* Diagnostics of the last clodb200_simplifyGroups / build on this process: {passes, wavefront rounds, max rounds in a pass}. */
void clodb200_simplifyStats(unsigned int out3[3]);

#ifdef __cplusplus
}
#endif
#endif
