// ORACLE BUILD SHIM: replaces BasicRenderer/src/Mesh/ClusterLOD.cpp when building libclodref_full_ours.so.
// The reference's clusterlod.h implementation is still compiled (clodDefaultConfig, clodLocalIndices, ... stay reference
// code) but its clodBuildEx / clodBuild are renamed away and re-defined here to forward to libclodb200's C ABI — the
// call-site change INTEGRATION.md describes, done at link level so ClusterLODUtilities.cpp stays unmodified.
// The library to drive is taken from $CLODB200_LIB (product .so on a GPU box, the tests/emu build on CPU).
#include <meshoptimizer.h>
#define clodBuildEx clodBuildEx_reference
#define clodBuild clodBuild_reference
#define CLUSTERLOD_IMPLEMENTATION
#include <ThirdParty/meshoptimizer/clusterlod.h>
#undef clodBuildEx
#undef clodBuild

#include "../../include/clodb200.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace
{
struct Clodb
{
	void* so = nullptr;
	int (*init)(int) = nullptr;
	const char* (*last_error)(void) = nullptr;
	size_t (*build_ex)(clodb200_config, clodb200_mesh, void*, clodb200_outputEx, const void*) = nullptr;
};

Clodb& lib()
{
	static Clodb l;
	if (!l.so)
	{
		const char* path = getenv("CLODB200_LIB");
		if (!path)
			throw std::runtime_error("CLODB200_LIB is not set");
		l.so = dlopen(path, RTLD_NOW | RTLD_LOCAL);
		if (!l.so)
			throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
		l.init = reinterpret_cast<int (*)(int)>(dlsym(l.so, "clodb200_init"));
		l.last_error = reinterpret_cast<const char* (*)(void)>(dlsym(l.so, "clodb200_last_error"));
		l.build_ex = reinterpret_cast<size_t (*)(clodb200_config, clodb200_mesh, void*, clodb200_outputEx, const void*)>(dlsym(l.so, "clodb200_buildEx"));
		if (!l.init || !l.last_error || !l.build_ex)
			throw std::runtime_error("libclodb200 is missing entry points");
		if (l.init(0) != CLODB200_OK)
			throw std::runtime_error(l.last_error());
	}
	return l;
}
} // namespace

static_assert(sizeof(clodConfig) == sizeof(clodb200_config) && sizeof(clodMesh) == sizeof(clodb200_mesh) && sizeof(clodCluster) == sizeof(clodb200_cluster) && sizeof(clodGroup) == sizeof(clodb200_group),
    "clodb200 structs must be layout-identical to the reference's");

// The simplification attribute stream of the last build (normals + the reference's MikkTSpace tangents): test input for
// clodb200_buildArtifacts, whose caller supplies the tangents the reference generates internally.
static std::vector<float> g_last_attributes;
static size_t g_last_attribute_count = 0;

extern "C" const float* clodshim_last_attributes(size_t* vertex_count, size_t* attribute_count)
{
	*attribute_count = g_last_attribute_count;
	*vertex_count = g_last_attribute_count ? g_last_attributes.size() / g_last_attribute_count : 0;
	return g_last_attributes.data();
}

extern "C" size_t clodBuildEx(clodConfig config, clodMesh mesh, void* output_context, clodOutputEx output_callback, const clodBuildParallelConfig*)
{
	g_last_attribute_count = mesh.vertex_attributes ? mesh.attribute_count : 0;
	g_last_attributes.clear();
	for (size_t i = 0; i < mesh.vertex_count && g_last_attribute_count; ++i)
		g_last_attributes.insert(g_last_attributes.end(), mesh.vertex_attributes + i * (mesh.vertex_attributes_stride / sizeof(float)),
		    mesh.vertex_attributes + i * (mesh.vertex_attributes_stride / sizeof(float)) + g_last_attribute_count);
	clodb200_config cfg;
	memcpy(&cfg, &config, sizeof(cfg));
	clodb200_mesh m;
	memcpy(&m, &mesh, sizeof(m));
	Clodb& l = lib();
	size_t clusters = l.build_ex(cfg, m, output_context, reinterpret_cast<clodb200_outputEx>(output_callback), nullptr);
	if (clusters == 0 && *l.last_error())
		throw std::runtime_error(l.last_error());
	return clusters;
}

extern "C" size_t clodBuild(clodConfig config, clodMesh mesh, void* output_context, clodOutput output_callback)
{
	(void)config, (void)mesh, (void)output_context, (void)output_callback;
	throw std::runtime_error("clodBuild is not used by the builder");
}
