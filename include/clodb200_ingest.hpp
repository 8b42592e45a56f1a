// clodb200 — C++ mirror of the reference's ingest-side builder interface, above the C ABI (include/clodb200.h).
//
// Reference: class MeshIngestBuilder, BasicRenderer/include/Mesh/ClusterLODTypes.h:354-434 (BuildClusterLODArtifacts() is
// implemented in src/Mesh/MeshIngestBuilder.cpp:7-18 as a call of BuildClusterLODArtifactsFromGeometry). Same method
// names, argument meaning and error behaviour (std::runtime_error with the reference's messages), so an importer
// (Import/GlTFGeometryExtractor.cpp:1025-1298, Import/USDGeometryExtractor.cpp:743-859) switches by changing a namespace.
// Differences: BuildClusterLODArtifacts() returns an owning handle to the artifacts held by the library (named arrays of
// the reference's PODs, see clodb200_artifactsGet) instead of a ClusterLODPrebuildArtifacts value; Build() (the
// renderer-side GPU Mesh object) is out of scope.
#pragma once

#include "clodb200.h"

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace clodb200
{

// MeshUvSetData (Import/MeshData.h:14-17)
struct MeshUvSetData
{
	std::vector<float> values; // float2 per vertex
};

// Owning view of the result of one build; movable, frees the library-side artifacts on destruction.
class ClusterLODPrebuildArtifacts
{
public:
	ClusterLODPrebuildArtifacts() = default;
	explicit ClusterLODPrebuildArtifacts(clodb200_artifacts* handle)
	    : m_handle(handle)
	{
	}
	ClusterLODPrebuildArtifacts(ClusterLODPrebuildArtifacts&& other) noexcept
	    : m_handle(other.m_handle)
	{
		other.m_handle = nullptr;
	}
	ClusterLODPrebuildArtifacts& operator=(ClusterLODPrebuildArtifacts&& other) noexcept
	{
		if (this != &other)
		{
			reset();
			m_handle = other.m_handle;
			other.m_handle = nullptr;
		}
		return *this;
	}
	ClusterLODPrebuildArtifacts(const ClusterLODPrebuildArtifacts&) = delete;
	ClusterLODPrebuildArtifacts& operator=(const ClusterLODPrebuildArtifacts&) = delete;
	~ClusterLODPrebuildArtifacts()
	{
		reset();
	}

	// Raw bytes of a named array ("groups", "segments", "nodes", "meshPages", "meshPageOffsets", "stats", ...).
	std::pair<const void*, size_t> Get(const char* name) const
	{
		const void* p = nullptr;
		size_t bytes = 0;
		if (!m_handle || !clodb200_artifactsGet(m_handle, name, &p, &bytes))
			return std::make_pair(static_cast<const void*>(nullptr), size_t(0));
		return std::make_pair(p, bytes);
	}
	size_t Count(const char* name, size_t element_size) const
	{
		return Get(name).second / element_size;
	}
	// CLodCacheLoader::SavePrebuiltLocked equivalent without the .usdc wrapper (clodb200_artifactsSaveCache)
	void SaveCache(const std::string& directory, const std::string& containerFileName, const std::string& metadataFileName, const std::string& sourceIdentifier, const std::string& primPath,
	    const std::string& subsetName, uint64_t buildConfigHash) const
	{
		if (clodb200_artifactsSaveCache(m_handle, directory.c_str(), containerFileName.c_str(), metadataFileName.c_str(), sourceIdentifier.c_str(), primPath.c_str(), subsetName.c_str(), buildConfigHash) != CLODB200_OK)
			throw std::runtime_error(clodb200_last_error());
	}
	clodb200_artifacts* handle() const
	{
		return m_handle;
	}

private:
	void reset()
	{
		if (m_handle)
			clodb200_artifactsFree(m_handle);
		m_handle = nullptr;
	}
	clodb200_artifacts* m_handle = nullptr;
};

typedef clodb200_builder_settings ClusterLODBuilderSettings;

class MeshIngestBuilder
{
public:
	MeshIngestBuilder(unsigned int vertexSize, unsigned int skinningVertexSize, unsigned int flags)
	    : MeshIngestBuilder(vertexSize, skinningVertexSize, flags, clodb200_defaultBuilderSettings())
	{
	}
	MeshIngestBuilder(unsigned int vertexSize, unsigned int skinningVertexSize, unsigned int flags, ClusterLODBuilderSettings clusterLODBuilderSettings)
	    : m_vertexSize(vertexSize), m_skinningVertexSize(skinningVertexSize), m_flags(flags), m_clusterLODBuilderSettings(clusterLODBuilderSettings)
	{
	}

	void ReserveVertices(size_t vertexCount)
	{
		m_vertices.reserve(vertexCount * static_cast<size_t>(m_vertexSize));
	}
	void ReserveIndices(size_t indexCount)
	{
		m_indices.reserve(indexCount);
	}
	void AppendVertexBytes(const std::byte* data, size_t byteCount)
	{
		if (byteCount != m_vertexSize)
			throw std::runtime_error("MeshIngestBuilder vertex byte size mismatch");
		m_vertices.insert(m_vertices.end(), data, data + byteCount);
	}
	void AppendSkinningVertexBytes(const std::byte* data, size_t byteCount)
	{
		if (m_skinningVertexSize == 0)
			throw std::runtime_error("MeshIngestBuilder has no skinning vertex format");
		if (byteCount != m_skinningVertexSize)
			throw std::runtime_error("MeshIngestBuilder skinning vertex byte size mismatch");
		m_skinningVertices.insert(m_skinningVertices.end(), data, data + byteCount);
	}
	void AppendIndex(uint32_t index)
	{
		m_indices.push_back(index);
	}
	void AppendIndices(const uint32_t* data, size_t count)
	{
		m_indices.insert(m_indices.end(), data, data + count);
	}
	void SetUvSets(std::vector<MeshUvSetData> uvSets)
	{
		m_uvSets = std::move(uvSets);
	}
	const std::vector<MeshUvSetData>& GetUvSets() const
	{
		return m_uvSets;
	}
	void SetClusterLODBuilderSettings(const ClusterLODBuilderSettings& settings)
	{
		m_clusterLODBuilderSettings = settings;
	}
	const ClusterLODBuilderSettings& GetClusterLODBuilderSettings() const
	{
		return m_clusterLODBuilderSettings;
	}

	// Runs the full cluster-LOD build on the GPU (mesh mode). Throws std::runtime_error with the library's message on
	// failure, as the reference throws from its validation (ClusterLODUtilities.cpp:4637, 5252, 5726).
	ClusterLODPrebuildArtifacts BuildClusterLODArtifacts() const
	{
		std::vector<clodb200_uv_set> sets(m_uvSets.size());
		for (size_t i = 0; i < m_uvSets.size(); ++i)
		{
			sets[i].values = m_uvSets[i].values.data();
			sets[i].count = m_uvSets[i].values.size() / 2;
		}
		clodb200_geometry g;
		g.vertices = m_vertices.data();
		g.vertex_count = m_vertexSize ? m_vertices.size() / m_vertexSize : 0;
		g.vertex_stride = m_vertexSize;
		g.vertex_flags = m_flags;
		g.indices = m_indices.data();
		g.index_count = m_indices.size();
		g.uv_sets = sets.empty() ? nullptr : sets.data();
		g.uv_set_count = sets.size();
		g.tangents = nullptr;
		g.skinning_vertices = m_skinningVertices.empty() ? nullptr : m_skinningVertices.data();
		g.skinning_vertex_bytes = m_skinningVertices.size();
		g.skinning_vertex_stride = m_skinningVertexSize;
		clodb200_artifacts* a = clodb200_buildArtifacts(&g, &m_clusterLODBuilderSettings);
		if (!a)
			throw std::runtime_error(clodb200_last_error());
		return ClusterLODPrebuildArtifacts(a);
	}

private:
	unsigned int m_vertexSize = 0;
	unsigned int m_skinningVertexSize = 0;
	unsigned int m_flags = 0;
	std::vector<std::byte> m_vertices;
	std::vector<std::byte> m_skinningVertices;
	std::vector<uint32_t> m_indices;
	std::vector<MeshUvSetData> m_uvSets;
	ClusterLODBuilderSettings m_clusterLODBuilderSettings;
};

} // namespace clodb200
