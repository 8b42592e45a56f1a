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
 *   clodb200_buildArtifacts         <- BuildClusterLODArtifactsFromGeometry, BasicRenderer/include/Mesh/ClusterLODUtilities.h:5-13
 *   clodb200_artifactsSaveCache     <- CLodCache::Save (container + metadata blob), BasicRenderer/src/Import/CLodCache.cpp:512-583
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

/* Same layout as struct clodConfig, clusterlod.h:15-71 (a memcpy of the reference's struct is valid).
 * Supported subset: every field the reference's builder sets (ClusterLODUtilities.cpp:5426-5460) - max_vertices / max_triangles /
 * min_triangles (1..256), partition_spatial, partition_sort, partition_size, partition_max_refined_groups,
 * partition_refined_split_count (incremented once per split partition), cluster_spatial = true, cluster_fill_weight, simplify_ratio,
 * simplify_threshold, simplify_error_merge_previous / _additive, simplify_error_factor_sloppy, simplify_permissive,
 * simplify_fallback_sloppy, optimize_bounds, optimize_clusters. Values whose reference behaviour is not built make the build fail with
 * an error instead of being ignored: cluster_spatial = false (meshopt_buildMeshletsFlex; note that clodDefaultConfig() leaves it
 * false, clusterlod.h:762), simplify_regularize, simplify_fallback_permissive without simplify_permissive,
 * simplify_error_edge_limit > 0. cluster_split_factor only matters for the flex clusterizer and is ignored. */
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
/* Selects the CUDA device for this process (one process per GPU). Threading: every host thread that calls the library gets
 * its own build context on first use (CUDA stream, device arenas, pinned staging), so independent meshes may be built
 * concurrently from several threads — the reference builds one primitive per worker thread the same way
 * (Import/GlTFGeometryExtractor.cpp:1349, Import/USDGeometryExtractor.cpp:945-961). A handle (uploaded mesh/geometry,
 * record, artifacts) may be used by any thread, by one thread at a time; an artifacts/record handle should be freed by the
 * thread that built it (it is recycled for that thread's next build). clodb200_launch_count and the timer/profile calls
 * refer to the calling thread's context. */
int clodb200_init(int device);
void clodb200_shutdown(void);
/* Kernel launches issued by the calling thread since it first used the library (bench.py's gpu_launches). */
uint64_t clodb200_launch_count(void);

/* clodDefaultConfig(max_triangles) followed by the BasicRenderer overrides (ClusterLODUtilities.cpp:5426-5460). */
clodb200_config clodb200_builderConfig(void);

/* remap[i] = lowest index with the same position (IEEE ==). positions_stride in bytes (>= 12, multiple of 4). */
int clodb200_generatePositionRemap(unsigned int* remap, const float* positions, size_t vertex_count, size_t positions_stride);

/* GenerateMikkTangents (ClusterLODUtilities.cpp:655-737; genTangSpaceDefault of Utilities/mikktspace.cpp over the
 * indexed triangle list, then the per-vertex sum in (face, corner) order, normalisation and the normal-derived fallback).
 * `vertices` is the interleaved MeshVertexLayout stream with normals and texcoords (position @0, normal @12, uv @24;
 * vertex_stride >= 32). out_tangents4 receives {tangent xyz, sign} per vertex. *out_generated = 0 where the reference's
 * generator returns false (the builder then continues without tangent attributes, :5361-5365). out_corner_tangents4
 * (optional, index_count float4) receives what genTangSpace hands to m_setTSpaceBasic per (face, corner); corners of
 * degenerate triangles, which copy from another corner (mikktspace.cpp:1820-1861), read as zeros there. */
int clodb200_generateMikkTangents(const void* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count,
    float* out_tangents4, int* out_generated, float* out_corner_tangents4);

/* Splits each segment [segment_offsets[s], segment_offsets[s+1]) (in triangles) of `indices` into meshlets.
 * Outputs: cluster_index_counts/cluster_vertex_counts/cluster_segments hold one entry per cluster (capacity
 * index_count / 3 entries each), out_indices receives index_count cluster-major indices. Returns the cluster count in
 * *out_cluster_count. segment_offsets == NULL means one segment covering everything. */
/* a3: locks[i] |= 2 (meshopt_SimplifyVertex_Protect) where vertex i is not its position class's canonical vertex and one of the
 * attribute columns selected by protect_mask differs from the canonical vertex's (float !=), clusterlod.h:829-841. */
int clodb200_protectBits(unsigned char* locks, const float* attributes, size_t attributes_stride, unsigned int protect_mask, const unsigned int* remap, size_t vertex_count);

/* a8: the deterministic second half of clod::partition (clusterlod.h:396-507) for a given partition id per cluster (what
 * meshopt_partitionClusters returned): partitions ordered by meshopt_spatialSortRemap of their last cluster's centre
 * (spatialorder.cpp:218-251), clusters in ascending order inside a partition, then the refined-id cap split
 * (partition_max_refined_groups). cluster_bounds5 = {centre xyz, radius, error} per cluster. out_group_clusters has cluster_count
 * entries, out_group_offsets cluster_count + 1 (worst case); bit-exact given the same partitions. */
int clodb200_partitionFinish(const clodb200_config* config, const unsigned int* cluster_part, size_t partition_count, size_t cluster_count, const int* cluster_refined,
    const float* cluster_bounds5, unsigned int* out_group_clusters, unsigned int* out_group_offsets, size_t* out_group_count);

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

/* ---- outer boundary: drop-in for BuildClusterLODArtifactsFromGeometry ------------------------------------------------
 * Reference: BasicRenderer/include/Mesh/ClusterLODUtilities.h:5-13 (implementation src/Mesh/ClusterLODUtilities.cpp:5325-5766),
 * i.e. MeshIngestBuilder::BuildClusterLODArtifacts() (ClusterLODTypes.h:415). The DAG build, the group output tables
 * (:856-1805), the traversal hierarchy (:4606-4963) and the mesh-wide page packing (:2313-2540) run inside this call; the
 * page blobs are written by the GPU and read back once. */

/* The ClusterLODBuilderSettings fields that are live in mesh mode (ClusterLODTypes.h:187-212). The voxel fallback is not
 * built: the call behaves as enableVoxelFallback = false / voxelFallbackMode = MeshOnly (CLodCacheTool --clod-voxel-mode=mesh).
 * preserveImportedNormals = 0 recomputes the page normals per group from the group's triangles (RecalculateGroupNormals,
 * ClusterLODUtilities.cpp:739-822; one GPU thread per group, not tuned: the reference itself never clears the flag). */
typedef struct clodb200_builder_settings
{
	float lodErrorMergePrevious;
	float lodErrorMergeAdditive;
	uint32_t partitionSizeFloor;
	int preserveImportedNormals;
	int enableNormalAttributeSimplification;
	float normalAttributeWeight;
	float simplifyTangentWeight;
	float simplifyTangentSignWeight;
} clodb200_builder_settings;
/* GetDefaultBuilderSettings(), BasicRenderer/src/Import/DefaultCLodSettings.cpp:3-30 */
clodb200_builder_settings clodb200_defaultBuilderSettings(void);

/* MeshUvSetData (Import/MeshData.h:14-17): `count` float2 values; a set whose count differs from the vertex count reads
 * as zeros, as the reference's (ClusterLODUtilities.cpp:1046). */
typedef struct clodb200_uv_set
{
	const float* values;
	size_t count;
} clodb200_uv_set;

/* The arguments of BuildClusterLODArtifactsFromGeometry as plain pointers. `vertices` is the interleaved stream of
 * MeshVertexLayout (Mesh/VertexLayout.h: position f32x3 @0, normal f32x3 @12, uv f32x2 @24 if VERTEX_TEXCOORDS, then
 * colour f32x3 if VERTEX_COLORS), vertex_flags the VertexFlags bits (Mesh/VertexFlags.h). `tangents` is normally NULL:
 * when the stream has normals and texcoords and normal-attribute simplification is on, the MikkTSpace tangent stream is
 * generated on the device as the reference's GenerateMikkTangents does inside its call (:655-737, :5359-5366). A
 * non-NULL `tangents` (float4 per vertex) overrides the generator.
 * Skinned meshes: `skinning_vertices` is the importer's second stream (MeshIngestBuilder::AppendSkinningVertexBytes; position f32x3,
 * normal f32x3, joints u32x4 x2, weights f32x4 x2 = PackedSkinningInfluences at byte 24, skinning_vertex_stride >= 88), NULL when the
 * mesh is not skinned. It only feeds the pages: joint and weight arrays per meshlet vertex and the sorted per-meshlet bone lists
 * (ClusterLODUtilities.cpp:1091-1120, 1240-1265, 1677-1711); skinning_vertex_bytes is the size of the stream (vertices beyond it
 * read as zero influences, :1107-1110). */
typedef struct clodb200_geometry
{
	const void* vertices;
	size_t vertex_count;
	unsigned int vertex_stride;
	unsigned int vertex_flags;
	const unsigned int* indices;
	size_t index_count;
	const clodb200_uv_set* uv_sets;
	size_t uv_set_count;
	const float* tangents;
	const void* skinning_vertices;
	size_t skinning_vertex_bytes;
	unsigned int skinning_vertex_stride;
} clodb200_geometry;

#define CLODB200_VERTEX_COLORS 1u
#define CLODB200_VERTEX_NORMALS 2u
#define CLODB200_VERTEX_TEXCOORDS 4u
#define CLODB200_VERTEX_SKINNED 8u

/* ClusterLODPrebuildArtifacts (ClusterLODTypes.h:165-169) as named arrays of the reference's PODs; see
 * clodb200_artifactsGet. Empty / invalid geometry gives artifacts with empty arrays, as the reference does. */
typedef struct clodb200_artifacts clodb200_artifacts;
clodb200_artifacts* clodb200_buildArtifacts(const clodb200_geometry* geometry, const clodb200_builder_settings* settings);

/* Geometry kept in HBM: upload once, build many times (benchmarks, scene batches). */
typedef struct clodb200_device_geometry clodb200_device_geometry;
clodb200_device_geometry* clodb200_geometryUpload(const clodb200_geometry* geometry, const clodb200_builder_settings* settings);
void clodb200_geometryFree(clodb200_device_geometry* geometry);
clodb200_artifacts* clodb200_geometryBuildArtifacts(const clodb200_device_geometry* geometry);

/* Names: "groups" (ClusterLODGroup, 76 B), "segments" (ClusterLODGroupSegment, 16 B), "segmentBounds" (float4),
 * "groupChunks" (ClusterLODGroupChunk, 20 B), "groupPageReferences" / "groupPageReferenceOffsets" (u32), "nodes"
 * (ClusterLODNode, 64 B), "lodNodeRanges" ({u32 offset, u32 count}), "lodLevelRoots" (u32), "objectBoundingSphere"
 * (float4), "counts" (u32 {trianglePageCount, voxelPageBase, voxelPageCount, maxDepth, maxTraversalDepth}),
 * "meshPages" (cacheBuildData.meshPageBlobs back to back; pinned host memory), "meshPageOffsets" (u64[pages + 1]),
 * "stats" (u64[16]: meshlets, groups, segments, pages, page bytes, meshlet-vertex refs, triangles over all levels,
 * group-unique vertices, levels, simplified triangles, device-to-host bytes, nodes). Returns 1 if the name exists. */
int clodb200_artifactsGet(const clodb200_artifacts* artifacts, const char* name, const void** out_ptr, size_t* out_bytes);
void clodb200_artifactsFree(clodb200_artifacts* artifacts);

/* CLod cache files (BasicRenderer/src/Import/CLodCache.cpp): writes `<directory>/<container_file_name>` — the .clodbin
 * container (ContainerHeader + page directory + page blobs, :252-259, 314-375) — and `<directory>/<metadata_file_name>`,
 * the bytes of SerializeMetadata (:169-207) that the reference stores as `uchar[] clodBlob` on prim /CLodCache of its
 * .usdc stage (:528-574; the OpenUSD crate wrapper itself is not written here). The file names are
 * the caller's: the reference derives them with boost::hash_combine (:586-633), which is not restated. */
int clodb200_artifactsSaveCache(const clodb200_artifacts* artifacts, const char* directory, const char* container_file_name, const char* metadata_file_name,
    const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash);
/* Skip-if-cached (CLodCacheLoader::TryLoadPrebuilt, CLodCacheLoader.cpp:218-234): 1 when `directory` holds a cache for this identity
 * and build configuration that the loader's acceptance rules take (CLodCache.cpp:209-250, 635-713: schema 47, build hash and
 * identity equal, blob consumed exactly, container file present with one locator per mesh page), else 0. Needs no GPU: a batch
 * tool probes before it uploads a mesh, which is how an interrupted scene build resumes. */
int clodb200_cacheProbe(const char* directory, const char* metadata_file_name, const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash);
/* The metadata blob alone: returns the byte count; writes at most `capacity` bytes. */
size_t clodb200_artifactsSerializeMetadata(const clodb200_artifacts* artifacts, const char* container_file_name, const char* source_identifier, const char* prim_path,
    const char* subset_name, uint64_t build_config_hash, void* buffer, size_t capacity);

/* Cache naming: how the renderer finds a cache (CLodCache.cpp:62-80, 586-633). boost::hash_combine chains restated from the
 * published Boost.ContainerHash (>= 1.82) algorithm; Boost is neither vendored nor installed, so these names are "parity unpinned"
 * (csrc/cachenames.cu). Each name function returns the byte count incl. the terminator and writes at most `capacity` bytes.
 *   clodb200_cacheBuildConfigHash  == CLodCache::ComputeBuildConfigHash()                        (reads BASICRENDERER_CLOD_VOXEL_*)
 *   clodb200_cacheFileName         == CLodCache::BuildCacheFileName(key, hash)                   "clod_<hex>.usdc"
 *   clodb200_cacheSubdirectory     == BuildSceneCacheSubdirectory(sourceIdentifier)              "clod/<stem>_<hex>" */
uint64_t clodb200_cacheBuildConfigHash(void);
size_t clodb200_cacheFileName(const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash, char* out, size_t capacity);
size_t clodb200_cacheSubdirectory(const char* source_identifier, char* out, size_t capacity);

/* ---- scene batches across the GPUs of one box -------------------------------------------------------------
 * Meshes are independent: a batch is sharded by mesh (one process per GPU, each calling the build entry points above for its own
 * meshes; the reference fans primitives out over worker threads the same way, GlTFGeometryExtractor.cpp:1349). The only exchange
 * is the gather of the per-mesh metadata blobs (clodb200_artifactsSerializeMetadata bytes, CLodCache.cpp:169-207) to every rank:
 * an NCCL all-gather over NVLink on the library's own communication stream, asynchronous to the caller's builds.
 *   rank 0: clodb200_commGetUniqueId -> ship the 128 bytes to the other ranks by any means (file, socket, MPI, torch.distributed)
 *   all   : clodb200_init(device); clodb200_commInit(id, world, rank)        (collective)
 *   all   : h = clodb200_commGatherBegin(blob, bytes)  ... keep building ...  clodb200_commGatherWait(h);
 *           clodb200_commGatherGet(h, r, &n) for r in 0..world-1; clodb200_commGatherFree(h)     (collective, same order on all ranks)
 * Issue the gather after the rank's last build of the batch: a pending NCCL kernel waits on the GPU for the slowest peer and keeps
 * this rank's cooperative kernels (which need every SM) from becoming co-resident until it has finished.
 * World size 1 needs no NCCL (loopback). NCCL is loaded at run time (libnccl.so.2; CLODB200_NCCL_LIB overrides the name). */
#define CLODB200_COMM_ID_BYTES 128
typedef struct clodb200_gather clodb200_gather;
int clodb200_commGetUniqueId(void* out_id);
int clodb200_commInit(const void* id, int world_size, int rank);
int clodb200_commWorldSize(void);
int clodb200_commRank(void);
void clodb200_commDestroy(void);
clodb200_gather* clodb200_commGatherBegin(const void* payload, size_t bytes);
int clodb200_commGatherWait(clodb200_gather* gather);
const void* clodb200_commGatherGet(const clodb200_gather* gather, int rank, size_t* out_bytes);
void clodb200_commGatherFree(clodb200_gather* gather);

/* CUDA-event stopwatch on the build stream: Start records an event, Stop records another, synchronises and returns the
 * elapsed device time in milliseconds (bench.py brackets its timed steps with these). */
void clodb200_timerStart(void);
float clodb200_timerStop(void);

/* Per-kernel CUDA-event timing on the build stream (off by default; adds two event records per launch).
 * clodb200_profileReport drains the recorded spans into "kernel,launches,total_ms,total_threads" lines, slowest first; returns the
 * number of bytes needed (including the terminator); at most `capacity` bytes are written. */
void clodb200_profileEnable(int enable);
size_t clodb200_profileReport(char* buffer, size_t capacity);

/* Device-wide primitives under the stages (hand-written single-pass chained scan and LSD radix sort, csrc/prims.cuh),
 * exposed for their own parity tests and micro-benchmarks. Host pointers; the primitive runs `repeat` times and *ms
 * (optional) receives the mean device time of the runs after the first. Scans are exclusive; the sort is stable on key
 * bits [bit_lo, bit_hi) rounded up to whole 8-bit digits (callers keep the bits above bit_hi clear) and reorders `values`
 * with the keys. */
int clodb200_primExclusiveScanU32(const unsigned int* in, unsigned int* out, size_t n, unsigned int* total, int repeat, float* ms);
int clodb200_primExclusiveMaxScanU64(const uint64_t* in, uint64_t* out, size_t n, int repeat, float* ms);
int clodb200_primSortPairsU32(unsigned int* keys, unsigned int* values, size_t n, int bit_lo, int bit_hi, int repeat, float* ms);
/* Test hook: sets the call epoch of the chained scans (every scan call takes the next epoch and tags its tile descriptors
 * with it). Lets a test put the epoch where stale descriptors of another scan format would alias it (tests/test_prims.py). */
int clodb200_primSetScanEpoch(unsigned int epoch);
/* The corner-angle arccosine of the tangent generator: the reference's `acos(float)` (Utilities/mikktspace.cpp:1421)
 * resolves to the C library's acosf; out[i] must equal it bit for bit on [-1, 1] (NaN outside). */
int clodb200_primAcosf(const float* in, float* out, size_t n);

/* Diagnostics of the last clodb200_simplifyGroups / build on this process: {passes, wavefront rounds, max rounds in a pass}. */
void clodb200_simplifyStats(unsigned int out3[3]);

#ifdef __cplusplus
}
#endif
#endif
