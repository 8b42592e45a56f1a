// Drives include/clodb200_ingest.hpp the way the reference importers drive MeshIngestBuilder
// (BasicRenderer/src/Import/GlTFGeometryExtractor.cpp:1025-1298): one AppendVertexBytes per vertex, AppendIndices, build.
// usage: ingest_test <mesh.bin>   (u32 V, u32 I, V * 6 f32 {position, normal}, I u32) -> one line of counts and checksums
#include "clodb200_ingest.hpp"

#include <cstdio>
#include <cstring>

static uint64_t fnv1a(const void* data, size_t bytes)
{
	const unsigned char* p = static_cast<const unsigned char*>(data);
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < bytes; ++i)
		h = (h ^ p[i]) * 1099511628211ull;
	return h;
}

int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	FILE* f = fopen(argv[1], "rb");
	if (!f)
		return 2;
	uint32_t V = 0, I = 0;
	if (fread(&V, 4, 1, f) != 1 || fread(&I, 4, 1, f) != 1)
		return 2;
	std::vector<float> vertices(size_t(V) * 6);
	std::vector<uint32_t> indices(I);
	if (fread(vertices.data(), 4, vertices.size(), f) != vertices.size() || fread(indices.data(), 4, I, f) != I)
		return 2;
	fclose(f);

	if (clodb200_init(0) != CLODB200_OK)
	{
		fprintf(stderr, "init failed: %s\n", clodb200_last_error());
		return 3;
	}
	const unsigned int vertexSize = 24;
	clodb200::MeshIngestBuilder builder(vertexSize, 0, CLODB200_VERTEX_NORMALS);

	// error behaviour of the reference (ClusterLODTypes.h:374-389)
	int thrown = 0;
	try
	{
		builder.AppendVertexBytes(reinterpret_cast<const std::byte*>(vertices.data()), 20);
	}
	catch (const std::runtime_error& e)
	{
		thrown += strcmp(e.what(), "MeshIngestBuilder vertex byte size mismatch") == 0;
	}
	try
	{
		builder.AppendSkinningVertexBytes(reinterpret_cast<const std::byte*>(vertices.data()), 16);
	}
	catch (const std::runtime_error& e)
	{
		thrown += strcmp(e.what(), "MeshIngestBuilder has no skinning vertex format") == 0;
	}

	builder.ReserveVertices(V);
	builder.ReserveIndices(I);
	for (uint32_t v = 0; v < V; ++v)
		builder.AppendVertexBytes(reinterpret_cast<const std::byte*>(&vertices[size_t(v) * 6]), vertexSize);
	builder.AppendIndices(indices.data(), I / 2);
	for (uint32_t i = I / 2; i < I; ++i)
		builder.AppendIndex(indices[i]);

	clodb200::ClusterLODPrebuildArtifacts artifacts = builder.BuildClusterLODArtifacts();
	auto pages = artifacts.Get("meshPages");
	auto groups = artifacts.Get("groups");
	auto nodes = artifacts.Get("nodes");
	printf("thrown %d groups %zu segments %zu nodes %zu pages_fnv %016llx groups_fnv %016llx nodes_fnv %016llx\n", thrown, artifacts.Count("groups", 76), artifacts.Count("segments", 16),
	    artifacts.Count("nodes", 64), (unsigned long long)fnv1a(pages.first, pages.second), (unsigned long long)fnv1a(groups.first, groups.second), (unsigned long long)fnv1a(nodes.first, nodes.second));

	// empty geometry: empty artifacts, as the reference returns (no throw)
	clodb200::MeshIngestBuilder empty(vertexSize, 0, CLODB200_VERTEX_NORMALS);
	clodb200::ClusterLODPrebuildArtifacts none = empty.BuildClusterLODArtifacts();
	printf("empty groups %zu\n", none.Count("groups", 76));
	clodb200_shutdown();
	return 0;
}
