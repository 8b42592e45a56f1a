// ORACLE DRIVER (test infrastructure): exposes the reference's GenerateMikkTangents
// (BasicRenderer/src/Mesh/ClusterLODUtilities.cpp:655-737, an internal-linkage function) by compiling the UNMODIFIED
// reference translation unit into this one where it lies under /root/reference (oracle/Makefile passes
// -I$(REF)/BasicRenderer/src). It runs genTangSpaceDefault of BasicRenderer/src/Utilities/mikktspace.cpp through the
// reference's own callbacks, per-vertex accumulation and fallback, i.e. exactly the tangent stream the builder feeds to
// the simplifier. Built into oracle/_ref/libclodref_mikk.so; only tests/, smoke() and bench.py's CPU legs load it.
#include "Mesh/ClusterLODUtilities.cpp"

#include <chrono>

extern "C" int clodref_mikk_tangents(const void* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count,
    float* out_tangents4, double* out_seconds)
{
	std::vector<std::byte> v(static_cast<const std::byte*>(vertices), static_cast<const std::byte*>(vertices) + vertex_count * vertex_stride);
	std::vector<uint32_t> idx(indices, indices + index_count);
	std::vector<DirectX::XMFLOAT4> tangents;
	auto t0 = std::chrono::steady_clock::now();
	bool ok = GenerateMikkTangents(v, vertex_stride, idx, tangents);
	auto t1 = std::chrono::steady_clock::now();
	if (out_seconds)
		*out_seconds = std::chrono::duration<double>(t1 - t0).count();
	if (!ok)
		return 0;
	memcpy(out_tangents4, tangents.data(), tangents.size() * sizeof(DirectX::XMFLOAT4));
	return 1;
}
