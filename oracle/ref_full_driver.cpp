// ORACLE DRIVER (test infrastructure): calls the reference's L3 builder BuildClusterLODArtifactsFromGeometry
// (BasicRenderer/src/Mesh/ClusterLODUtilities.cpp:5325, compiled unmodified from /root/reference by oracle/Makefile) and
// exposes the resulting ClusterLODPrebuiltData + page blobs as named byte blobs through a small C interface.
// Two libraries are built from this driver:
//   libclodref_full.so       reference builder + reference clodBuildEx (ClusterLOD.cpp)
//   libclodref_full_ours.so  reference builder + clodBuildEx forwarded to libclodb200 (shims/clusterlod_via_clodb200.cpp),
//                            i.e. exactly the integration of INTEGRATION.md: the reference's own validation then judges
//                            our DAG, and its page/hierarchy output on OUR clusters is the oracle for the L3 rows.
#include "Mesh/ClusterLODUtilities.h"
#include "Mesh/VertexFlags.h"
#include "Managers/Singletons/TaskSchedulerManager.h"

#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace
{
struct Handle
{
	std::map<std::string, std::vector<unsigned char>> blobs;
	std::string error;
	double seconds = 0;
};

template <typename T>
void put(Handle* h, const char* name, const std::vector<T>& v)
{
	std::vector<unsigned char>& b = h->blobs[name];
	b.resize(v.size() * sizeof(T));
	if (!v.empty())
		memcpy(b.data(), v.data(), b.size());
}

ClusterLODBuilderSettings mesh_mode_settings()
{
	// GetDefaultBuilderSettings() (BasicRenderer/src/Import/DefaultCLodSettings.cpp:3-30; not compilable here because of its
	// backslash include) with the voxel fallback switched off, the CLI's --clod-voxel-mode=mesh (CLodCacheTool/main.cpp:77,97)
	ClusterLODBuilderSettings s;
	s.disableSloppyFallback = false;
	s.lodErrorMergePrevious = 1.5f;
	s.lodErrorMergeAdditive = 0.0f;
	s.partitionSizeFloor = 8u;
	s.preserveImportedNormals = true;
	s.enableNormalAttributeSimplification = true;
	s.normalAttributeWeight = 1.0f;
	s.simplifyTangentWeight = 0.01f;
	s.simplifyTangentSignWeight = 0.5f;
	s.enableVoxelFallback = false;
	s.voxelFallbackMode = ClusterLODVoxelFallbackMode::MeshOnly;
	s.voxelGridBaseResolution = 32u;
	s.voxelMinResolution = 0u;
	s.voxelRaysPerCell = 64u;
	s.voxelFallbackScalingFactor = 1.0f;
	s.voxelFallbackMaxRetryCount = 10u;
	s.voxelFallbackGrowthFactor = 1.1f;
	s.voxelFallbackAcceptanceBias = 1.0f;
	s.voxelFallbackOpacityThreshold = 0.0f;
	s.voxelFallbackCarryZeroCoverage = false;
	s.voxelFallbackPruningMode = ClusterLODVoxelPruningMode::Coverage;
	return s;
}
} // namespace

extern "C"
{

// vertices: interleaved per VertexLayout.h (pos f32x3 @0, normal f32x3 @12[, uv f32x2 @24][, color f32x3]); flags = VertexFlags.
// threads: 0/1 = serial (scheduler left uninitialised, ClusterLODUtilities.cpp:5619), n > 1 = n worker threads.
void* clodfull_build_skinned(const unsigned char* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count, unsigned int flags, unsigned int threads,
    const unsigned char* skinning_vertices, size_t skinning_bytes, unsigned int skinning_stride);

// options for the next clodfull_build* call on this process: bit 0 = preserveImportedNormals false (RecalculateGroupNormals)
static unsigned int g_build_options = 0;
void clodfull_set_options(unsigned int options)
{
	g_build_options = options;
}

void* clodfull_build(const unsigned char* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count, unsigned int flags, unsigned int threads)
{
	return clodfull_build_skinned(vertices, vertex_count, vertex_stride, indices, index_count, flags, threads, nullptr, 0, 0);
}

// skinning_vertices: the importer's second vertex stream (MeshIngestBuilder::AppendSkinningVertexBytes), or null
void* clodfull_build_skinned(const unsigned char* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count, unsigned int flags, unsigned int threads,
    const unsigned char* skinning_vertices, size_t skinning_bytes, unsigned int skinning_stride)
{
	Handle* h = new Handle();
	try
	{
		br::TaskSchedulerManager& tsm = br::TaskSchedulerManager::GetInstance();
		tsm.Cleanup();
		if (threads > 1)
		{
			setenv("CLODREF_THREADS", std::to_string(threads).c_str(), 1);
			tsm.Initialize();
		}
		std::vector<std::byte> v(reinterpret_cast<const std::byte*>(vertices), reinterpret_cast<const std::byte*>(vertices) + vertex_count * vertex_stride);
		std::vector<uint32_t> idx(indices, indices + index_count);
		std::vector<MeshUvSetData> uvSets;
		auto t0 = std::chrono::steady_clock::now();
		std::vector<std::byte> skin;
		if (skinning_vertices && skinning_bytes)
			skin.assign(reinterpret_cast<const std::byte*>(skinning_vertices), reinterpret_cast<const std::byte*>(skinning_vertices) + skinning_bytes);
		ClusterLODBuilderSettings build_settings = mesh_mode_settings();
		if (g_build_options & 1u)
			build_settings.preserveImportedNormals = false;
		ClusterLODPrebuildArtifacts a = BuildClusterLODArtifactsFromGeometry(v, vertex_stride, skin.empty() ? nullptr : &skin, skinning_stride, idx, uvSets, flags, build_settings);
		h->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		tsm.Cleanup();

		const ClusterLODPrebuiltData& d = a.prebuiltData;
		put(h, "groups", d.groups);
		put(h, "segments", d.segments);
		put(h, "segmentBounds", d.segmentBounds);
		put(h, "groupChunks", d.groupChunks);
		put(h, "groupDiskLocators", d.groupDiskLocators);
		put(h, "pageDiskLocators", d.pageDiskLocators);
		put(h, "groupPageReferences", d.groupPageReferences);
		put(h, "groupPageReferenceOffsets", d.groupPageReferenceOffsets);
		put(h, "nodes", d.nodes);
		put(h, "lodNodeRanges", d.lodNodeRanges);
		put(h, "lodLevelRoots", d.lodLevelRoots);
		std::vector<float> sphere = {d.objectBoundingSphere.sphere.x, d.objectBoundingSphere.sphere.y, d.objectBoundingSphere.sphere.z, d.objectBoundingSphere.sphere.w};
		put(h, "objectBoundingSphere", sphere);
		std::vector<uint32_t> counts = {d.trianglePageCount, d.voxelPageBase, d.voxelPageCount, d.maxDepth, d.maxTraversalDepth};
		put(h, "counts", counts);
		// mesh pages back to back + offsets
		std::vector<uint64_t> offsets(1, 0);
		std::vector<unsigned char>& pages = h->blobs["meshPages"];
		for (const std::vector<std::byte>& p : a.cacheBuildData.meshPageBlobs)
		{
			size_t old = pages.size();
			pages.resize(old + p.size());
			if (!p.empty())
				memcpy(pages.data() + old, p.data(), p.size());
			offsets.push_back(pages.size());
		}
		put(h, "meshPageOffsets", offsets);
	}
	catch (const std::exception& e)
	{
		h->error = e.what();
	}
	return h;
}

const char* clodfull_error(const void* handle)
{
	return static_cast<const Handle*>(handle)->error.c_str();
}

double clodfull_seconds(const void* handle)
{
	return static_cast<const Handle*>(handle)->seconds;
}

int clodfull_get(const void* handle, const char* name, const void** out_ptr, size_t* out_bytes)
{
	const Handle* h = static_cast<const Handle*>(handle);
	auto it = h->blobs.find(name);
	*out_ptr = nullptr;
	*out_bytes = 0;
	if (it == h->blobs.end())
		return 0;
	*out_ptr = it->second.data();
	*out_bytes = it->second.size();
	return 1;
}

void clodfull_free(void* handle)
{
	delete static_cast<Handle*>(handle);
}

// struct sizes, so the python side can check its dtypes
void clodfull_sizes(unsigned int out[8])
{
	out[0] = sizeof(ClusterLODGroup);
	out[1] = sizeof(ClusterLODGroupSegment);
	out[2] = sizeof(BoundingSphere);
	out[3] = sizeof(ClusterLODGroupChunk);
	out[4] = sizeof(ClusterLODGroupDiskLocator);
	out[5] = sizeof(ClusterLODNode);
	out[6] = sizeof(ClusterLODNodeRangeAlloc);
	out[7] = sizeof(clodBounds);
}

}
