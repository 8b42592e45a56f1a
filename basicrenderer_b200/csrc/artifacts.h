// Outer boundary of the cluster-LOD build: the work of BuildClusterLODArtifactsFromGeometry
// (BasicRenderer/src/Mesh/ClusterLODUtilities.cpp:5325-5766) around clodBuildEx — group output tables, traversal
// hierarchy, mesh-wide page packing and the SoA page blobs — with the DAG and the page bytes produced on the device.
#pragma once

#include "dag.h"

#include <map>
#include <string>

namespace clodb
{

static const u32 kMaxUvSets = 4;

// BasicRenderer/include/Mesh/VertexFlags.h
static const u32 kVertexColors = 1u << 0;
static const u32 kVertexNormals = 1u << 1;
static const u32 kVertexTexcoords = 1u << 2;
static const u32 kVertexSkinned = 1u << 3;

// The ClusterLODBuilderSettings fields that are live in mesh mode (ClusterLODTypes.h:187-212; the voxel fields are not,
// SURVEY.md §8a a20 / §8f rank 4).
struct BuilderSettings
{
	float lod_error_merge_previous = 1.5f;
	float lod_error_merge_additive = 0.0f;
	u32 partition_size_floor = 8;
	bool preserve_imported_normals = true;
	bool enable_normal_attribute_simplification = true;
	float normal_attribute_weight = 1.0f;
	float simplify_tangent_weight = 0.01f;
	float simplify_tangent_sign_weight = 0.5f;
};

// Importer geometry resident in HBM: the interleaved vertex stream of MeshVertexLayout (VertexLayout.h: position f32x3 @0,
// normal f32x3 @12, then uv f32x2, then color f32x3 as the flags say), u32 indices, optional separate UV sets, and the
// streams the DAG build consumes (tight positions, simplification attributes) split off on the device.
struct DeviceGeometry
{
	const u8* vertices = nullptr;
	u32 vertex_stride = 0; // bytes
	u32 vertex_flags = 0;
	size_t vertex_count = 0;
	const u32* indices = nullptr;
	size_t index_count = 0;
	u32 uv_set_count = 0;
	const float* uv_values[kMaxUvSets] = {}; // u at [i * uv_stride], v at [i * uv_stride + 1]
	u32 uv_stride[kMaxUvSets] = {};          // in floats
	// skinned meshes: the importer's second vertex stream (position f32x3, normal f32x3, joints u32x4 x2, weights f32x4 x2;
	// ClusterLODUtilities.cpp:50-56, :1091-1120); only the page writer reads it
	const u8* skinning_vertices = nullptr;
	u32 skinning_stride = 0;
	size_t skinning_vertex_count = 0;
	DeviceMesh mesh;
	// MikkTSpace tangents are generated inside every build call, as the reference does (ClusterLODUtilities.cpp:5359-5366):
	// tangents4 = float4 per vertex scratch, attribute columns [tangent_column, +4) of mesh.attributes receive them
	float* generated_tangents4 = nullptr;
	float* attributes_rw = nullptr;
	u32 tangent_column = 0;
};

// positions3[i] = vertex position; attributes[i] = {normal xyz}{tangent xyzw} as selected (either may be skipped)
void write_tangent_columns(const float* tangents4, size_t vertex_count, float* attributes, u32 attribute_stride, u32 tangent_column);
void split_vertex_streams(const u8* vertices, u32 vertex_stride, size_t vertex_count, float* positions3, float* attributes, u32 attribute_stride, bool with_normals,
    const float* tangents4);

struct Artifacts
{
	// ClusterLODPrebuiltData (ClusterLODTypes.h:124-145) as named byte blobs of the reference's PODs:
	// "groups" ClusterLODGroup[76 B], "segments" ClusterLODGroupSegment[16 B], "segmentBounds" float4, "groupChunks"
	// ClusterLODGroupChunk[20 B], "groupPageReferences" u32, "groupPageReferenceOffsets" u32, "nodes" ClusterLODNode[64 B],
	// "lodNodeRanges" {offset,count}, "lodLevelRoots" u32, "objectBoundingSphere" float4,
	// "counts" u32{trianglePageCount, voxelPageBase, voxelPageCount, maxDepth, maxTraversalDepth},
	// "meshPageOffsets" u64[pageCount + 1]; the page bytes themselves sit in `pages` ("meshPages").
	std::map<std::string, std::vector<u8>> blobs;
	HostStage pages; // pinned
	size_t page_bytes = 0;
	// what the build moved and produced (bench.py's roofline terms)
	u64 stats[16] = {};
};

void build_artifacts(const DeviceGeometry& geometry, const BuilderSettings& settings, Workspace& ws, Artifacts& out, BuildStats& stats);

// ---- CLod cache files (BasicRenderer/src/Mesh/CLodCache.cpp) -----------------------------------------------------------
struct CacheIdentity
{
	std::string source_identifier, prim_path, subset_name;
	u64 build_config_hash = 0;
};
// SerializeMetadata (CLodCache.cpp:169-207): the blob stored as `clodBlob`; page locators are filled from the artifacts
std::vector<u8> serialize_cache_metadata(const Artifacts& artifacts, const CacheIdentity& identity, const std::string& container_file_name);
// .clodbin container (CLodCache.cpp:252-259, 314-375): header, page table, blobs back to back
void write_cache_container(const Artifacts& artifacts, const std::string& path);

} // namespace clodb
