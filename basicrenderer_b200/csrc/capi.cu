// clodb200 C ABI (include/clodb200.h): argument checking, host<->device staging, error translation.
#include "../../include/clodb200.h"
#include "clodb.h"
#include "dag.h"
#include "artifacts.h"

#include <map>
#include <algorithm>

#include <atomic>
#include <mutex>
#include <functional>

using namespace clodb;

namespace
{
thread_local std::string t_last_error;
std::mutex g_api_mutex; // guards process-wide state only (initialisation); builds run on per-thread contexts without it
bool g_initialized = false;
int g_device = 0;
// the workspace and the block cache live on the heap behind trivially destructible thread_local pointers: the context owner
// below releases them from its destructor at thread exit, when other thread_local objects may already have been destroyed
thread_local Workspace* t_ws = nullptr;
Workspace& ws_ref()
{
	if (!t_ws)
		t_ws = new Workspace();
	return *t_ws;
}
#define g_ws (ws_ref())
thread_local bool t_context_ready = false;
// Bumped by clodb200_shutdown: a thread whose context was created before the last shutdown drops it (streams and slabs may
// belong to another device) and builds a new one on its next call.
std::atomic<uint64_t> g_generation{1};
thread_local uint64_t t_generation = 0;
void release_thread_context() noexcept;
// RAII owner of the calling thread's build context: a host thread that builds and then exits gives back its device slabs,
// pinned staging, streams and cached blocks instead of leaking them until process exit.
struct ThreadContextOwner
{
	~ThreadContextOwner()
	{
		release_thread_context();
	}
};
thread_local ThreadContextOwner t_context_owner;

int fail(int code, const std::string& message)
{
	t_last_error = message;
	return code;
}

void ensure_workspace(size_t persist_bytes, size_t temp_bytes)
{
	if (g_ws.persist.capacity < persist_bytes)
		g_ws.persist.init(persist_bytes);
	if (g_ws.temp.capacity < temp_bytes)
		g_ws.temp.init(temp_bytes);
	g_ws.persist.release(0);
	g_ws.temp.release(0);
}

// Runs `attempt` again with a larger slab while it fails with ArenaExhausted and `may_retry()` allows it (nothing has been
// handed to the caller yet). ensure_workspace never shrinks a slab, so the growth sticks for the retry and later builds.
template <typename Attempt, typename MayRetry>
auto with_arena_growth(Attempt&& attempt, MayRetry&& may_retry) -> decltype(attempt())
{
	for (int tries = 0;; ++tries)
	{
		try
		{
			return attempt();
		}
		catch (const ArenaExhausted& e)
		{
			if (tries >= 16 || !may_retry())
				throw;
			Arena& slab = e.arena == &g_ws.persist ? g_ws.persist : g_ws.temp;
			if (e.arena != &g_ws.persist && e.arena != &g_ws.temp)
				throw;
			size_t want = std::max(slab.capacity * 2, e.need * 2);
			// everything in flight (build stream and the asynchronous level downloads) must be done with the slab before it goes
			dev_sync();
			dev_d2h_async_wait();
#ifndef CLODB_EMU
			{
				// never ask for more than the device can give: the old slab is released first, so its bytes count as free
				size_t free_bytes = 0, total_bytes = 0;
				if (cudaMemGetInfo(&free_bytes, &total_bytes) == cudaSuccess)
				{
					const size_t available = free_bytes + slab.capacity;
					const size_t margin = size_t(256) << 20;
					if (e.need + margin > available)
						throw Error("clodb200: out of device memory (the build needs a " + std::to_string(e.need >> 20) + " MB slab, " + std::to_string(available >> 20) + " MB are available)");
					want = std::min(want, available - margin);
				}
			}
#endif
			try
			{
				slab.init(want);
			}
			catch (const Error&)
			{
#ifndef CLODB_EMU
				cudaGetLastError(); // a failed cudaMalloc leaves its error behind; the next launch check must not trip over it
#endif
				throw Error("clodb200: out of device memory while growing a workspace slab to " + std::to_string(want >> 20) + " MB");
			}
		}
	}
}

// clodConfig fields that change the reference's output and are not built here fail loudly instead of being dropped
// (the struct is layout-identical to clodConfig, so a memcpy'd clodDefaultConfig() arrives with cluster_spatial = false).
struct UnsupportedConfig : Error
{
	explicit UnsupportedConfig(const std::string& what)
	    : Error(what)
	{
	}
};

Config to_config(const clodb200_config* c)
{
	Config r;
	if (!c)
		return r;
	if (!c->cluster_spatial)
		throw UnsupportedConfig("clodb200: clodConfig::cluster_spatial = false (meshopt_buildMeshletsFlex) is not supported; the reference builder sets cluster_spatial = true (ClusterLODUtilities.cpp:5430)");
	if (c->simplify_regularize)
		throw UnsupportedConfig("clodb200: clodConfig::simplify_regularize is not supported");
	if (c->simplify_fallback_permissive && !c->simplify_permissive)
		throw UnsupportedConfig("clodb200: clodConfig::simplify_fallback_permissive is not supported (use simplify_permissive)");
	if (c->simplify_error_edge_limit > 0)
		throw UnsupportedConfig("clodb200: clodConfig::simplify_error_edge_limit is not supported");
	if (c->max_vertices == 0 || c->max_vertices > 256 || c->max_triangles == 0 || c->max_triangles > 256 || c->min_triangles > c->max_triangles)
		throw UnsupportedConfig("clodb200: clodConfig meshlet limits out of range (max_vertices and max_triangles in 1..256, min_triangles <= max_triangles)");
	r.max_vertices = u32(c->max_vertices);
	r.min_triangles = u32(c->min_triangles);
	r.max_triangles = u32(c->max_triangles);
	r.cluster_fill_weight = c->cluster_fill_weight;
	r.partition_size = u32(c->partition_size);
	r.partition_max_refined_groups = u32(c->partition_max_refined_groups);
	r.partition_sort = c->partition_sort;
	r.partition_spatial = c->partition_spatial;
	r.simplify_ratio = c->simplify_ratio;
	r.simplify_threshold = c->simplify_threshold;
	r.simplify_error_merge_previous = c->simplify_error_merge_previous;
	r.simplify_error_merge_additive = c->simplify_error_merge_additive;
	r.simplify_error_factor_sloppy = c->simplify_error_factor_sloppy;
	r.simplify_permissive = c->simplify_permissive;
	r.simplify_fallback_sloppy = c->simplify_fallback_sloppy;
	r.optimize_clusters = c->optimize_clusters;
	r.optimize_bounds = c->optimize_bounds;
	return r;
}

// tightly packed device copy of strided host positions
float* upload_positions(const float* positions, size_t vertex_count, size_t stride_bytes, Arena& arena)
{
	float* dev = arena.alloc<float>(vertex_count * 3);
	if (stride_bytes == 12)
	{
		dev_h2d(dev, positions, vertex_count * 12);
	}
	else
	{
		std::vector<float> packed(vertex_count * 3);
		size_t stride = stride_bytes / 4;
		for (size_t i = 0; i < vertex_count; ++i)
		{
			packed[i * 3 + 0] = positions[i * stride + 0];
			packed[i * 3 + 1] = positions[i * stride + 1];
			packed[i * 3 + 2] = positions[i * stride + 2];
		}
		dev_h2d(dev, packed.data(), vertex_count * 12);
	}
	return dev;
}

// The calling thread's build context: device selection and its own non-blocking stream (the arenas, scan-chain
// descriptors and pinned staging of the thread grow on demand).
void ensure_thread_context()
{
	if (t_context_ready && t_generation == g_generation.load())
		return;
	if (t_context_ready)
		release_thread_context(); // left over from before a shutdown
	(void)&t_context_owner;       // odr-use: constructs the owner so its destructor runs at thread exit
	t_generation = g_generation.load();
#ifndef CLODB_EMU
	CUDA_CHECK(cudaSetDevice(g_device));
	cudaStream_t s;
	CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	g_stream = s;
#endif
	t_context_ready = true;
}

} // namespace
namespace clodb
{
// comm.cu reports through the same per-thread error slot
void capi_set_error(const std::string& message)
{
	t_last_error = message;
}
bool capi_initialized()
{
	return g_initialized;
}
} // namespace clodb
namespace
{
template <typename F>
int guarded(F&& body)
{
	if (!g_initialized)
		return fail(CLODB200_ERR_NO_DEVICE, "clodb200: not initialised (call clodb200_init; a CUDA device is required, there is no CPU path)");
	try
	{
		ensure_thread_context();
		return body();
	}
	catch (const std::exception& e)
	{
		return fail(CLODB200_ERR_RUNTIME, e.what());
	}
}
} // namespace

struct clodb200_record
{
	std::map<std::string, std::vector<unsigned char> > blobs;

	template <typename T>
	void append(const char* name, const T* data, size_t count)
	{
		std::vector<unsigned char>& b = blobs[name];
		size_t old = b.size();
		b.resize(old + count * sizeof(T));
		if (count)
			memcpy(b.data() + old, data, count * sizeof(T));
	}
	template <typename T>
	void push(const char* name, T v)
	{
		append(name, &v, 1);
	}
};

struct RecordSink : DagSink
{
	clodb200_record* rec;
	bool keep_indices = true;
	int next_group = 0;
	u32 cluster_total = 0;
	u64 index_total = 0;

	typedef std::vector<unsigned char> Blob;
	Blob *b_depth = nullptr, *b_simplified, *b_refined, *b_bounds, *b_vcount, *b_indices, *b_ioffsets, *b_goffsets;

	template <typename T>
	static void put(Blob* b, const T* data, size_t count)
	{
		const unsigned char* p = reinterpret_cast<const unsigned char*>(data);
		b->insert(b->end(), p, p + count * sizeof(T));
	}

	// index_hint: expected total index count of the stream (about 2x the input), reserved up front
	void bind(size_t index_hint)
	{
		b_depth = &rec->blobs["group_depth"];
		b_simplified = &rec->blobs["group_simplified"];
		b_refined = &rec->blobs["cluster_refined"];
		b_bounds = &rec->blobs["cluster_bounds"];
		b_vcount = &rec->blobs["cluster_vertex_count"];
		b_indices = &rec->blobs["cluster_indices"];
		b_ioffsets = &rec->blobs["cluster_index_offsets"];
		b_goffsets = &rec->blobs["group_cluster_offsets"];
		if (keep_indices)
			b_indices->reserve(index_hint * 4);
	}

	int group(const DagGroup& group, const DagCluster* clusters, size_t cluster_count, size_t) override
	{
		put(b_depth, &group.depth, 1);
		put(b_simplified, group.simplified, 5);
		for (size_t i = 0; i < cluster_count; ++i)
		{
			const DagCluster& c = clusters[i];
			put(b_refined, &c.refined, 1);
			put(b_bounds, c.bounds, 5);
			u32 vc = u32(c.vertex_count);
			put(b_vcount, &vc, 1);
			if (keep_indices)
				put(b_indices, c.indices, c.index_count);
			index_total += c.index_count;
			put(b_ioffsets, &index_total, 1);
		}
		cluster_total += u32(cluster_count);
		put(b_goffsets, &cluster_total, 1);
		return next_group++;
	}
};

extern "C"
{

const char* clodb200_last_error(void)
{
	return t_last_error.c_str();
}

int clodb200_init(int device)
{
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (g_initialized)
		return CLODB200_OK;
#ifndef CLODB_EMU
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(CLODB200_ERR_NO_DEVICE, std::string("clodb200: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU path");
	if (device < 0 || device >= count)
		return fail(CLODB200_ERR_INVALID, "clodb200: invalid device ordinal");
	e = cudaSetDevice(device);
	if (e != cudaSuccess)
		return fail(CLODB200_ERR_NO_DEVICE, std::string("clodb200: cudaSetDevice failed: ") + cudaGetErrorString(e));
	g_device = device;
	try
	{
		ensure_thread_context();
	}
	catch (const std::exception& ex)
	{
		return fail(CLODB200_ERR_NO_DEVICE, std::string("clodb200: stream creation failed: ") + ex.what());
	}
	if (const char* dbg = getenv("CLODB200_SYNC_DEBUG"))
		g_sync_debug = atoi(dbg);
#else
	(void)device;
	if (const char* rev = getenv("CLODB200_EMU_REVERSE"))
		emu_reverse = atoi(rev);
#endif
	g_initialized = true;
	return CLODB200_OK;
}

void clodb200_shutdown(void)
{
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (!g_initialized)
		return;
	release_thread_context();
	// the contexts of other threads are released by those threads: at their next call (stale generation) or when they exit
	g_generation.fetch_add(1);
	g_initialized = false;
}

uint64_t clodb200_launch_count(void)
{
	return g_launches;
}

clodb200_config clodb200_builderConfig(void)
{
	clodb200_config c;
	memset(&c, 0, sizeof(c));
	c.max_vertices = 128;
	c.min_triangles = 64;
	c.max_triangles = 128;
	c.partition_spatial = true;
	c.partition_sort = true;
	c.partition_size = 384;
	c.partition_max_refined_groups = 8;
	c.cluster_spatial = true;
	c.cluster_fill_weight = 0.5f;
	c.cluster_split_factor = 2.0f;
	c.simplify_ratio = 0.5f;
	c.simplify_threshold = 0.85f;
	c.simplify_error_merge_previous = 1.5f;
	c.simplify_error_merge_additive = 0.0f;
	c.simplify_error_factor_sloppy = 100.f;
	c.simplify_permissive = true;
	c.simplify_fallback_permissive = false;
	c.simplify_fallback_sloppy = true;
	c.simplify_regularize = false;
	c.optimize_bounds = true;
	c.optimize_clusters = true;
	return c;
}

int clodb200_generatePositionRemap(unsigned int* remap, const float* positions, size_t vertex_count, size_t positions_stride)
{
	return guarded([&]() -> int {
		if (vertex_count == 0)
			return CLODB200_OK;
		if (!remap || !positions || positions_stride < 12 || positions_stride % 4)
			return fail(CLODB200_ERR_INVALID, "clodb200_generatePositionRemap: invalid arguments");
		ensure_workspace(vertex_count * 16 + (1 << 20), vertex_count * 40 + (1 << 20));
		float* dpos = upload_positions(positions, vertex_count, positions_stride, g_ws.persist);
		u32* dremap = g_ws.persist.alloc<u32>(vertex_count);
		position_remap(dpos, vertex_count, dremap, g_ws.temp);
		dev_d2h(remap, dremap, vertex_count * sizeof(u32));
		return CLODB200_OK;
	});
}

int clodb200_protectBits(unsigned char* locks, const float* attributes, size_t attributes_stride, unsigned int protect_mask, const unsigned int* remap, size_t vertex_count)
{
	return guarded([&]() -> int {
		if (vertex_count == 0)
			return CLODB200_OK;
		if (!locks || !remap || (protect_mask && attributes && (attributes_stride < 4 || attributes_stride % 4)))
			return fail(CLODB200_ERR_INVALID, "clodb200_protectBits: invalid arguments");
		for (size_t i = 0; i < vertex_count; ++i)
			if (remap[i] >= vertex_count)
				return fail(CLODB200_ERR_INVALID, "clodb200_protectBits: remap entry out of range");
		ensure_workspace(vertex_count * (attributes_stride + 8) + (1 << 20), 1 << 20);
		u8* dlocks = g_ws.persist.alloc<u8>(vertex_count);
		u32* dremap = g_ws.persist.alloc<u32>(vertex_count);
		dev_h2d(dlocks, locks, vertex_count);
		dev_h2d(dremap, remap, vertex_count * 4);
		if (attributes && protect_mask)
		{
			float* dattr = g_ws.persist.alloc<float>(vertex_count * (attributes_stride / 4));
			dev_h2d(dattr, attributes, vertex_count * attributes_stride);
			protect_bits(dattr, u32(attributes_stride / 4), protect_mask, dremap, vertex_count, dlocks);
		}
		dev_d2h(locks, dlocks, vertex_count);
		return CLODB200_OK;
	});
}

int clodb200_generateMikkTangents(const void* vertices, size_t vertex_count, unsigned int vertex_stride, const unsigned int* indices, size_t index_count,
    float* out_tangents4, int* out_generated, float* out_corner_tangents4)
{
	return guarded([&]() -> int {
		if (out_generated)
			*out_generated = 0;
		if (!vertices || !indices || !out_tangents4 || !out_generated)
			return fail(CLODB200_ERR_INVALID, "clodb200_generateMikkTangents: invalid arguments");
		if (vertex_count == 0 || index_count == 0 || vertex_stride < 32 || vertex_stride % 4)
			return CLODB200_OK; // GenerateMikkTangents returns false (ClusterLODUtilities.cpp:665-675)
		for (size_t i = 0; i < index_count; ++i)
			if (indices[i] >= vertex_count)
				return CLODB200_OK; // :677-683
		ensure_workspace(vertex_count * (size_t(vertex_stride) + 16) + index_count * 20 + (1 << 20), mikk_temp_bytes(vertex_count, index_count));
		u8* dv = g_ws.persist.alloc<u8>(vertex_count * vertex_stride);
		u32* di = g_ws.persist.alloc<u32>(index_count);
		float* dt = g_ws.persist.alloc<float>(vertex_count * 4);
		dev_h2d(dv, vertices, vertex_count * vertex_stride);
		dev_h2d(di, indices, index_count * 4);
		float* dc = out_corner_tangents4 ? g_ws.persist.alloc<float>(index_count * 4) : nullptr;
		if (mikk_tangents(dv, vertex_stride, vertex_count, di, index_count, dt, g_ws.temp, dc))
		{
			dev_d2h(out_tangents4, dt, vertex_count * 16);
			if (dc)
				dev_d2h(out_corner_tangents4, dc, index_count * 16);
			*out_generated = 1;
		}
		return CLODB200_OK;
	});
}

int clodb200_partitionFinish(const clodb200_config* config, const unsigned int* cluster_part, size_t partition_count, size_t cluster_count, const int* cluster_refined,
    const float* cluster_bounds5, unsigned int* out_group_clusters, unsigned int* out_group_offsets, size_t* out_group_count)
{
	return guarded([&]() -> int {
		if (out_group_count)
			*out_group_count = 0;
		if (cluster_count == 0 || partition_count == 0)
			return CLODB200_OK;
		if (!cluster_part || !cluster_refined || !cluster_bounds5 || !out_group_clusters || !out_group_offsets || !out_group_count)
			return fail(CLODB200_ERR_INVALID, "clodb200_partitionFinish: invalid arguments");
		for (size_t c = 0; c < cluster_count; ++c)
			if (cluster_part[c] >= partition_count)
				return fail(CLODB200_ERR_INVALID, "clodb200_partitionFinish: partition id out of range");
		const u32 K = u32(cluster_count);
		ensure_workspace(size_t(K) * 64 + (16 << 20), size_t(K) * 256 + (16 << 20));
		u32* d_part = g_ws.persist.alloc<u32>(K);
		int* d_refined = g_ws.persist.alloc<int>(K);
		float* d_bounds = g_ws.persist.alloc<float>(size_t(K) * 5);
		dev_h2d(d_part, cluster_part, size_t(K) * 4);
		dev_h2d(d_refined, cluster_refined, size_t(K) * 4);
		dev_h2d(d_bounds, cluster_bounds5, size_t(K) * 20);
		GroupSet gs = partition_finish(d_part, u32(partition_count), K, d_refined, d_bounds, to_config(config), g_ws);
		dev_d2h(out_group_clusters, gs.group_clusters, size_t(K) * 4);
		for (u32 g = 0; g <= gs.group_count; ++g)
			out_group_offsets[g] = gs.group_cluster_offset_host[g];
		*out_group_count = gs.group_count;
		return CLODB200_OK;
	});
}

int clodb200_clusterize(const clodb200_config* config, const unsigned int* indices, size_t index_count, const unsigned int* segment_offsets, size_t segment_count,
    const float* positions, size_t vertex_count, size_t positions_stride,
    unsigned int* cluster_index_counts, unsigned int* cluster_vertex_counts, unsigned int* cluster_segments, unsigned int* out_indices, size_t* out_cluster_count)
{
	return guarded([&]() -> int {
		if (out_cluster_count)
			*out_cluster_count = 0;
		if (index_count == 0)
			return CLODB200_OK;
		if (!indices || !positions || index_count % 3 || positions_stride < 12 || positions_stride % 4 || !out_indices || !out_cluster_count)
			return fail(CLODB200_ERR_INVALID, "clodb200_clusterize: invalid arguments");
		for (size_t i = 0; i < index_count; ++i)
			if (indices[i] >= vertex_count)
				return fail(CLODB200_ERR_INVALID, "clodb200_clusterize: index out of range");
		u32 T = u32(index_count / 3);
		std::vector<u32> segs;
		if (segment_offsets && segment_count)
			segs.assign(segment_offsets, segment_offsets + segment_count + 1);
		else
			segs = {0u, T};
		if (segs.front() != 0 || segs.back() != T)
			return fail(CLODB200_ERR_INVALID, "clodb200_clusterize: segment offsets must cover [0, triangle_count]");

		ensure_workspace(vertex_count * 12 + index_count * 12 + (16 << 20), size_t(T) * 160 + (16 << 20));
		float* dpos = upload_positions(positions, vertex_count, positions_stride, g_ws.persist);
		u32* dtri = g_ws.persist.alloc<u32>(index_count);
		dev_h2d(dtri, indices, index_count * sizeof(u32));

		ClusterSet cs = clusterize(dtri, T, segs.data(), u32(segs.size() - 1), dpos, to_config(config), g_ws);

		std::vector<u32> offsets = dev_download(cs.cluster_tri_offset, size_t(cs.cluster_count) + 1);
		for (u32 c = 0; c < cs.cluster_count; ++c)
			if (cluster_index_counts)
				cluster_index_counts[c] = (offsets[c + 1] - offsets[c]) * 3;
		if (cluster_vertex_counts)
			dev_d2h(cluster_vertex_counts, cs.cluster_vertex_count, size_t(cs.cluster_count) * sizeof(u32));
		if (cluster_segments)
			dev_d2h(cluster_segments, cs.cluster_segment, size_t(cs.cluster_count) * sizeof(u32));
		dev_d2h(out_indices, cs.tri, index_count * sizeof(u32));
		*out_cluster_count = cs.cluster_count;
		return CLODB200_OK;
	});
}

int clodb200_computeClusterBounds(const unsigned int* indices, const unsigned int* cluster_index_counts, size_t cluster_count,
    const float* positions, size_t vertex_count, size_t positions_stride, float* out_bounds4)
{
	return guarded([&]() -> int {
		if (cluster_count == 0)
			return CLODB200_OK;
		if (!indices || !cluster_index_counts || !positions || !out_bounds4 || positions_stride < 12 || positions_stride % 4)
			return fail(CLODB200_ERR_INVALID, "clodb200_computeClusterBounds: invalid arguments");
		std::vector<u32> offsets(cluster_count + 1, 0);
		for (size_t c = 0; c < cluster_count; ++c)
		{
			if (cluster_index_counts[c] % 3 || cluster_index_counts[c] > 128 * 3)
				return fail(CLODB200_ERR_INVALID, "clodb200_computeClusterBounds: clusters must hold at most 128 triangles");
			offsets[c + 1] = offsets[c] + cluster_index_counts[c] / 3;
		}
		size_t index_count = size_t(offsets.back()) * 3;
		ensure_workspace(vertex_count * 12 + index_count * 4 + cluster_count * 24 + (16 << 20), 1 << 20);
		float* dpos = upload_positions(positions, vertex_count, positions_stride, g_ws.persist);
		u32* dtri = g_ws.persist.alloc<u32>(index_count);
		dev_h2d(dtri, indices, index_count * sizeof(u32));
		u32* doff = g_ws.persist.alloc<u32>(cluster_count + 1);
		dev_h2d(doff, offsets.data(), offsets.size() * sizeof(u32));
		float* dbounds = g_ws.persist.alloc<float>(cluster_count * 4);
		cluster_bounds(dtri, doff, u32(cluster_count), dpos, dbounds);
		dev_d2h(out_bounds4, dbounds, cluster_count * 4 * sizeof(float));
		return CLODB200_OK;
	});
}

int clodb200_lockBoundary(unsigned char* locks, const unsigned int* indices, const unsigned int* group_index_offsets, size_t group_count,
    const unsigned int* remap, const unsigned char* vertex_lock, size_t vertex_count)
{
	return guarded([&]() -> int {
		if (!locks || !indices || !group_index_offsets || !remap)
			return fail(CLODB200_ERR_INVALID, "clodb200_lockBoundary: invalid arguments");
		size_t index_count = group_index_offsets[group_count];
		ensure_workspace(index_count * 4 + vertex_count * 8 + group_count * 4 + (16 << 20), vertex_count * 8 + (16 << 20));
		std::vector<u32> tri_offsets(group_count + 1);
		for (size_t g = 0; g <= group_count; ++g)
			tri_offsets[g] = group_index_offsets[g] / 3;
		u32* dtri = g_ws.persist.alloc<u32>(index_count);
		u32* doff = g_ws.persist.alloc<u32>(group_count + 1);
		u32* dremap = g_ws.persist.alloc<u32>(vertex_count);
		u8* dlocks = g_ws.persist.alloc<u8>(vertex_count);
		u8* dvlock = vertex_lock ? g_ws.persist.alloc<u8>(vertex_count) : nullptr;
		dev_h2d(dtri, indices, index_count * 4);
		dev_h2d(doff, tri_offsets.data(), tri_offsets.size() * 4);
		dev_h2d(dremap, remap, vertex_count * 4);
		dev_h2d(dlocks, locks, vertex_count);
		if (vertex_lock)
			dev_h2d(dvlock, vertex_lock, vertex_count);
		lock_boundary(dtri, doff, u32(group_count), u32(index_count / 3), dremap, dvlock, vertex_count, dlocks, g_ws.temp);
		dev_d2h(locks, dlocks, vertex_count);
		return CLODB200_OK;
	});
}

int clodb200_localIndicesBatch(const unsigned int* indices, const uint64_t* cluster_index_offsets, size_t cluster_count, size_t vertex_capacity,
    unsigned int* out_vertices, unsigned char* out_triangles, unsigned int* out_vertex_counts)
{
	return guarded([&]() -> int {
		if (cluster_count == 0)
			return CLODB200_OK;
		if (!indices || !cluster_index_offsets || !out_vertices || !out_triangles || !out_vertex_counts || vertex_capacity == 0 || vertex_capacity > 256)
			return fail(CLODB200_ERR_INVALID, "clodb200_localIndicesBatch: invalid arguments");
		size_t index_count = size_t(cluster_index_offsets[cluster_count]);
		for (size_t c = 0; c < cluster_count; ++c)
			if (cluster_index_offsets[c + 1] < cluster_index_offsets[c] || cluster_index_offsets[c + 1] - cluster_index_offsets[c] > 768)
				return fail(CLODB200_ERR_INVALID, "clodb200_localIndicesBatch: clusters hold at most 256 triangles");
		ensure_workspace(index_count * 5 + cluster_count * (vertex_capacity * 4 + 16) + (16 << 20), 16 << 20);
		u32* dind = g_ws.persist.alloc<u32>(index_count);
		u64* doff = g_ws.persist.alloc<u64>(cluster_count + 1);
		u32* dvert = g_ws.persist.alloc<u32>(cluster_count * vertex_capacity);
		u8* dtri = g_ws.persist.alloc<u8>(index_count);
		u32* dcount = g_ws.persist.alloc<u32>(cluster_count);
		dev_h2d(dind, indices, index_count * 4);
		dev_h2d(doff, cluster_index_offsets, (cluster_count + 1) * 8);
		dev_memset(dvert, 0, cluster_count * vertex_capacity * 4);
		local_indices(dind, doff, u32(cluster_count), u32(vertex_capacity), dvert, dtri, dcount);
		dev_d2h(out_vertices, dvert, cluster_count * vertex_capacity * 4);
		dev_d2h(out_triangles, dtri, index_count);
		dev_d2h(out_vertex_counts, dcount, cluster_count * 4);
		return CLODB200_OK;
	});
}

size_t clodb200_localIndices(unsigned int* vertices, unsigned char* triangles, const unsigned int* indices, size_t index_count)
{
	if (index_count == 0)
		return 0;
	uint64_t offsets[2] = {0, index_count};
	unsigned int tmp[256];
	unsigned int count = 0;
	if (clodb200_localIndicesBatch(indices, offsets, 1, 256, tmp, triangles, &count) != CLODB200_OK)
		return 0;
	memcpy(vertices, tmp, size_t(count) * sizeof(unsigned int));
	return count;
}

int clodb200_simplifyGroups(const clodb200_config* config, const unsigned int* indices, const unsigned int* group_index_offsets, size_t group_count,
    const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count,
    const unsigned char* locks, unsigned int* out_indices, unsigned int* out_group_index_counts, float* out_group_errors)
{
	return guarded([&]() -> int {
		if (group_count == 0)
			return CLODB200_OK;
		if (!indices || !group_index_offsets || !positions || positions_stride < 12 || positions_stride % 4 || !out_indices || !out_group_index_counts || !out_group_errors)
			return fail(CLODB200_ERR_INVALID, "clodb200_simplifyGroups: invalid arguments");
		if (attribute_count && (!attributes || !attribute_weights || attributes_stride < attribute_count * 4 || attributes_stride % 4 || attribute_count > 32))
			return fail(CLODB200_ERR_INVALID, "clodb200_simplifyGroups: invalid attribute arguments");
		size_t index_count = group_index_offsets[group_count];
		std::vector<u32> tri_offsets(group_count + 1);
		for (size_t g = 0; g <= group_count; ++g)
		{
			if (group_index_offsets[g] % 3)
				return fail(CLODB200_ERR_INVALID, "clodb200_simplifyGroups: group offsets must be multiples of 3");
			tri_offsets[g] = group_index_offsets[g] / 3;
		}
		size_t astride = attribute_count ? attributes_stride / 4 : 0;
		ensure_workspace(vertex_count * (24 + astride * 4) + index_count * 12 + (32 << 20), index_count * 200 + vertex_count * 16 + (32 << 20));

		DeviceMesh mesh;
		float* dpos = upload_positions(positions, vertex_count, positions_stride, g_ws.persist);
		mesh.positions = dpos;
		mesh.vertex_count = vertex_count;
		if (attribute_count)
		{
			float* dattr = g_ws.persist.alloc<float>(vertex_count * astride);
			dev_h2d(dattr, attributes, vertex_count * astride * 4);
			mesh.attributes = dattr;
			mesh.attribute_stride = u32(astride);
			mesh.attribute_count = u32(attribute_count);
			for (size_t i = 0; i < attribute_count; ++i)
				mesh.attribute_weights[i] = attribute_weights[i];
		}
		u8* dlocks = nullptr;
		if (locks)
		{
			dlocks = g_ws.persist.alloc<u8>(vertex_count);
			dev_h2d(dlocks, locks, vertex_count);
		}
		u32* dremap = g_ws.persist.alloc<u32>(vertex_count);
		position_remap(dpos, vertex_count, dremap, g_ws.temp);
		u32* dtri = g_ws.persist.alloc<u32>(index_count);
		dev_h2d(dtri, indices, index_count * 4);

		SimplifyOutput so = simplify_groups(dtri, tri_offsets.data(), u32(group_count), mesh, dremap, dlocks, to_config(config), g_ws);

		std::vector<u32> offs = dev_download(so.group_tri_offset, group_count + 1);
		for (size_t g = 0; g < group_count; ++g)
			out_group_index_counts[g] = (offs[g + 1] - offs[g]) * 3;
		dev_d2h(out_indices, so.tri, size_t(so.triangle_count) * 12);
		dev_d2h(out_group_errors, so.group_error, group_count * 4);
		return CLODB200_OK;
	});
}

// Small cache of device blocks for mesh uploads: cudaMalloc / cudaFree per call cost milliseconds and cudaFree synchronises the
// device. Freed blocks are kept (up to 8) and handed out again to requests they can hold without wasting more than 2x.
struct DeviceBlock
{
	void* ptr;
	size_t bytes;
};
static thread_local std::vector<DeviceBlock>* t_block_cache = nullptr;
static std::vector<DeviceBlock>& block_cache_ref()
{
	if (!t_block_cache)
		t_block_cache = new std::vector<DeviceBlock>();
	return *t_block_cache;
}
#define g_block_cache (block_cache_ref())

static void* block_alloc(size_t bytes, std::vector<DeviceBlock>& owned)
{
	bytes = bytes ? bytes : 1;
	size_t best = g_block_cache.size();
	for (size_t i = 0; i < g_block_cache.size(); ++i)
		if (g_block_cache[i].bytes >= bytes && g_block_cache[i].bytes <= bytes * 2 + 4096 && (best == g_block_cache.size() || g_block_cache[i].bytes < g_block_cache[best].bytes))
			best = i;
	DeviceBlock b;
	if (best != g_block_cache.size())
	{
		b = g_block_cache[best];
		g_block_cache.erase(g_block_cache.begin() + best);
	}
	else
	{
		b.ptr = dev_malloc(bytes);
		b.bytes = bytes;
	}
	owned.push_back(b);
	return b.ptr;
}

static void block_release(std::vector<DeviceBlock>& owned)
{
	for (const DeviceBlock& b : owned)
		g_block_cache.push_back(b);
	owned.clear();
	while (g_block_cache.size() > 8)
	{
		size_t smallest = 0;
		for (size_t i = 1; i < g_block_cache.size(); ++i)
			if (g_block_cache[i].bytes < g_block_cache[smallest].bytes)
				smallest = i;
		dev_free(g_block_cache[smallest].ptr);
		g_block_cache.erase(g_block_cache.begin() + smallest);
	}
}

KERNEL k_index_range(const u32* __restrict__ indices, size_t n, u32 vertex_count, u32* bad)
{
	size_t i = GTID;
	if (i < n && indices[i] >= vertex_count)
		atomicOr(bad, 1u);
}

struct clodb200_device_mesh
{
	DeviceMesh mesh;
	u32* indices = nullptr;
	size_t index_count = 0;
	std::vector<DeviceBlock> allocations;
};

static bool validate_mesh(const clodb200_mesh& mesh)
{
	// clusterlod.h:796-816
	const bool missingIndices = mesh.index_count > 0 && mesh.indices == NULL;
	const bool missingPositions = mesh.vertex_count > 0 && mesh.vertex_positions == NULL;
	const bool missingAttributes = mesh.attribute_count > 0 && mesh.vertex_attributes == NULL;
	const bool emptyGeometry = mesh.index_count == 0 || mesh.vertex_count == 0;
	const bool invalidPositionStride = mesh.vertex_positions_stride < sizeof(float) * 3;
	const bool invalidAttributeStride = mesh.attribute_count > 0 && mesh.vertex_attributes_stride < mesh.attribute_count * sizeof(float);
	if (emptyGeometry || missingIndices || missingPositions || missingAttributes || invalidPositionStride || invalidAttributeStride)
	{
		fprintf(stderr,
		    "clusterlod: skipping mesh with invalid or empty geometry (index_count=%zu, vertex_count=%zu, indices=%p, positions=%p, position_stride=%zu, attribute_count=%zu, attributes=%p, attribute_stride=%zu)\n",
		    mesh.index_count, mesh.vertex_count, static_cast<const void*>(mesh.indices), static_cast<const void*>(mesh.vertex_positions), mesh.vertex_positions_stride,
		    mesh.attribute_count, static_cast<const void*>(mesh.vertex_attributes), mesh.vertex_attributes_stride);
		return false;
	}
	return true;
}

static clodb200_device_mesh* upload_mesh_locked(const clodb200_mesh& mesh)
{
	if (mesh.index_count % 3 || mesh.vertex_positions_stride % 4 || mesh.vertex_attributes_stride % 4 || mesh.attribute_count > 32)
		throw Error("clodb200: index count must be a multiple of 3, strides multiples of 4, at most 32 attributes");
	clodb200_device_mesh* dm = new clodb200_device_mesh();
	try
	{
		size_t V = mesh.vertex_count;
		float* dpos = static_cast<float*>(block_alloc(V * 12, dm->allocations));
		if (mesh.vertex_positions_stride == 12)
			dev_h2d(dpos, mesh.vertex_positions, V * 12);
		else
		{
			std::vector<float> packed(V * 3);
			size_t stride = mesh.vertex_positions_stride / 4;
			for (size_t i = 0; i < V; ++i)
				for (int k = 0; k < 3; ++k)
					packed[i * 3 + k] = mesh.vertex_positions[i * stride + k];
			dev_h2d(dpos, packed.data(), V * 12);
		}
		dm->mesh.positions = dpos;
		dm->mesh.vertex_count = V;
		size_t astride = mesh.vertex_attributes_stride / 4;
		if (mesh.vertex_attributes && astride)
		{
			float* dattr = static_cast<float*>(block_alloc(V * astride * 4, dm->allocations));
			dev_h2d(dattr, mesh.vertex_attributes, V * astride * 4);
			dm->mesh.attributes = dattr;
			dm->mesh.attribute_stride = u32(astride);
			dm->mesh.attribute_count = u32(mesh.attribute_count);
			for (size_t i = 0; i < mesh.attribute_count; ++i)
				dm->mesh.attribute_weights[i] = mesh.attribute_weights[i];
			dm->mesh.attribute_protect_mask = mesh.attribute_protect_mask;
		}
		if (mesh.vertex_lock)
		{
			u8* dl = static_cast<u8*>(block_alloc(V, dm->allocations));
			dev_h2d(dl, mesh.vertex_lock, V);
			dm->mesh.vertex_lock = dl;
		}
		dm->indices = static_cast<u32*>(block_alloc(mesh.index_count * 4, dm->allocations));
		dev_h2d(dm->indices, mesh.indices, mesh.index_count * 4);
		dm->index_count = mesh.index_count;
		// range check on the device copy (the reference asserts; a bad index would read out of bounds in every stage)
		u32* bad = static_cast<u32*>(block_alloc(sizeof(u32), dm->allocations));
		dev_memset(bad, 0, sizeof(u32));
		LAUNCH(k_index_range, mesh.index_count, dm->indices, mesh.index_count, u32(V), bad);
		if (dev_read(bad))
			throw Error("clodb200: index out of range");
	}
	catch (...)
	{
		block_release(dm->allocations);
		delete dm;
		throw;
	}
	return dm;
}

static void free_mesh_locked(clodb200_device_mesh* dm)
{
	if (!dm)
		return;
	block_release(dm->allocations);
	delete dm;
}

struct CallbackSink : DagSink
{
	void* context;
	clodb200_outputEx callback_ex;
	clodb200_output callback;
	size_t delivered = 0; // groups handed to the caller so far (a build that has delivered anything is not restarted)

	int group(const DagGroup& group, const DagCluster* clusters, size_t cluster_count, size_t task_index) override
	{
		static_assert(sizeof(DagCluster) == sizeof(clodb200_cluster) && sizeof(DagGroup) == sizeof(clodb200_group), "layout");
		clodb200_group g;
		memcpy(&g, &group, sizeof(g));
		const clodb200_cluster* c = reinterpret_cast<const clodb200_cluster*>(clusters);
		delivered++;
		if (callback_ex)
			return callback_ex(context, g, c, cluster_count, task_index, 0);
		if (callback)
			return callback(context, g, c, cluster_count);
		return -1;
	}
};

static thread_local BuildStats g_last_build_stats;

static size_t build_locked(const clodb200_config& config, const clodb200_device_mesh* dm, DagSink& sink)
{
	size_t T = dm->index_count / 3;
	size_t V = dm->mesh.vertex_count;
	size_t scale_temp = 640, scale_persist = 96, base_temp = 64u << 20;
	if (const char* e = getenv("CLODB200_TEMP_BYTES_PER_TRI"))
		scale_temp = size_t(atoll(e));
	if (const char* e = getenv("CLODB200_TEMP_BYTES_BASE")) // tests start from a starved slab to exercise the growth path
		base_temp = size_t(atoll(e));
	ensure_workspace(T * scale_persist + V * 32 + (64u << 20), T * scale_temp + V * 16 + base_temp);
	size_t clusters = build_dag(to_config(&config), dm->mesh, dm->indices, dm->index_count, g_ws, sink, g_last_build_stats);
	// optional instrumentation counter of the reference (clusterlod.h:33-35, incremented at :474-475)
	if (config.partition_refined_split_count)
		*config.partition_refined_split_count += g_last_build_stats.refined_splits;
	return clusters;
}

size_t clodb200_meshBuildEx(clodb200_config config, const clodb200_device_mesh* mesh, void* output_context, clodb200_outputEx output_callback)
{
	size_t result = 0;
	int status = guarded([&]() -> int {
		if (!mesh)
			return fail(CLODB200_ERR_INVALID, "clodb200_meshBuildEx: null mesh");
		CallbackSink sink;
		sink.context = output_context;
		sink.callback_ex = output_callback;
		sink.callback = nullptr;
		t_last_error.clear();
		result = with_arena_growth([&]() { return build_locked(config, mesh, sink); }, [&]() { return sink.delivered == 0; });
		return CLODB200_OK;
	});
	return status == CLODB200_OK ? result : 0;
}

clodb200_device_mesh* clodb200_meshUpload(clodb200_mesh mesh)
{
	clodb200_device_mesh* dm = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!validate_mesh(mesh))
			return fail(CLODB200_ERR_INVALID, "clodb200: invalid or empty geometry");
		dm = upload_mesh_locked(mesh);
		return CLODB200_OK;
	});
	return dm;
}

void clodb200_meshFree(clodb200_device_mesh* mesh)
{
	guarded([&]() -> int {
		free_mesh_locked(mesh);
		return CLODB200_OK;
	});
}

static size_t build_host(clodb200_config config, clodb200_mesh mesh, void* output_context, clodb200_outputEx cb_ex, clodb200_output cb)
{
	size_t result = 0;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!validate_mesh(mesh))
			return CLODB200_OK; // the reference returns 0 without failing
		clodb200_device_mesh* dm = upload_mesh_locked(mesh);
		try
		{
			CallbackSink sink;
			sink.context = output_context;
			sink.callback_ex = cb_ex;
			sink.callback = cb;
			result = with_arena_growth([&]() { return build_locked(config, dm, sink); }, [&]() { return sink.delivered == 0; });
		}
		catch (...)
		{
			free_mesh_locked(dm);
			throw;
		}
		free_mesh_locked(dm);
		return CLODB200_OK;
	});
	return result;
}

size_t clodb200_build(clodb200_config config, clodb200_mesh mesh, void* output_context, clodb200_output output_callback)
{
	return build_host(config, mesh, output_context, nullptr, output_callback);
}

size_t clodb200_buildEx(clodb200_config config, clodb200_mesh mesh, void* output_context, clodb200_outputEx output_callback, const void*)
{
	return build_host(config, mesh, output_context, output_callback, nullptr);
}

static thread_local clodb200_record* g_record_pool = nullptr; // one recycled record (buffers keep their capacity between builds)

static clodb200_record* record_build_once(const clodb200_config& config, const clodb200_device_mesh* dm, bool keep_indices)
{
	clodb200_record* rec = g_record_pool ? g_record_pool : new clodb200_record();
	g_record_pool = nullptr;
	try
	{
		RecordSink sink;
		sink.rec = rec;
		sink.keep_indices = keep_indices;
		rec->push<u32>("group_cluster_offsets", 0);
		rec->push<u64>("cluster_index_offsets", 0);
		sink.bind(dm->index_count * 2 + dm->index_count / 8);
		size_t clusters = build_locked(config, dm, sink);
		const BuildStats& st = g_last_build_stats;
		rec->append<u32>("level_triangles", st.level_triangles.data(), st.level_triangles.size());
		rec->append<u32>("level_clusters", st.level_clusters.data(), st.level_clusters.size());
		rec->append<u32>("level_groups", st.level_groups.data(), st.level_groups.size());
		rec->append<u32>("level_passes", st.level_passes.data(), st.level_passes.size());
		rec->append<u32>("level_sloppy", st.level_sloppy.data(), st.level_sloppy.size());
		u64 stats[8] = {u64(clusters), st.levels, st.groups, st.simplified_triangles, st.d2h_bytes, st.simplify_passes, st.simplify_rounds, g_launches};
		rec->append<u64>("stats", stats, 8);
	}
	catch (...)
	{
		delete rec;
		throw;
	}
	return rec;
}

static clodb200_record* record_build(const clodb200_config& config, const clodb200_device_mesh* dm, bool keep_indices)
{
	return with_arena_growth([&]() { return record_build_once(config, dm, keep_indices); }, []() { return true; });
}

clodb200_record* clodb200_meshBuildRecorded(clodb200_config config, const clodb200_device_mesh* mesh, int keep_indices)
{
	clodb200_record* rec = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!mesh)
			return fail(CLODB200_ERR_INVALID, "clodb200_meshBuildRecorded: null mesh");
		rec = record_build(config, mesh, keep_indices != 0);
		return CLODB200_OK;
	});
	return rec;
}

clodb200_record* clodb200_buildRecorded(clodb200_config config, clodb200_mesh mesh)
{
	clodb200_record* rec = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!validate_mesh(mesh))
			return fail(CLODB200_ERR_INVALID, "clodb200: invalid or empty geometry");
		clodb200_device_mesh* dm = upload_mesh_locked(mesh);
		try
		{
			rec = record_build(config, dm, true);
		}
		catch (...)
		{
			free_mesh_locked(dm);
			throw;
		}
		free_mesh_locked(dm);
		return CLODB200_OK;
	});
	return rec;
}

int clodb200_recordGet(const clodb200_record* record, const char* name, const void** out_ptr, size_t* out_bytes)
{
	*out_ptr = nullptr;
	*out_bytes = 0;
	if (!record)
		return 0;
	auto it = record->blobs.find(name);
	if (it == record->blobs.end())
		return 0;
	*out_ptr = it->second.data();
	*out_bytes = it->second.size();
	return 1;
}

void clodb200_recordFree(clodb200_record* record)
{
	if (!record)
		return;
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (!g_record_pool)
	{
		for (auto& kv : record->blobs)
			kv.second.clear(); // keeps the capacity
		g_record_pool = record;
	}
	else
		delete record;
}

// ---- outer boundary: BuildClusterLODArtifactsFromGeometry --------------------------------------------------------------
struct clodb200_artifacts
{
	Artifacts data;
};

struct clodb200_device_geometry
{
	DeviceGeometry geometry;
	BuilderSettings settings;
	std::vector<DeviceBlock> allocations;
};

static thread_local clodb200_artifacts* g_artifacts_pool = nullptr; // one recycled result (keeps its pinned page buffer)

static BuilderSettings to_settings(const clodb200_builder_settings* s)
{
	BuilderSettings r;
	if (!s)
		return r;
	r.lod_error_merge_previous = s->lodErrorMergePrevious;
	r.lod_error_merge_additive = s->lodErrorMergeAdditive;
	r.partition_size_floor = s->partitionSizeFloor;
	r.preserve_imported_normals = s->preserveImportedNormals != 0;
	r.enable_normal_attribute_simplification = s->enableNormalAttributeSimplification != 0;
	r.normal_attribute_weight = s->normalAttributeWeight;
	r.simplify_tangent_weight = s->simplifyTangentWeight;
	r.simplify_tangent_sign_weight = s->simplifyTangentSignWeight;
	return r;
}

static bool geometry_is_buildable(const clodb200_geometry& g)
{
	// the conditions under which clodBuildEx returns 0 (clusterlod.h:796-816); the builder then returns empty artifacts
	return g.vertices && g.indices && g.vertex_count > 0 && g.index_count > 0 && g.vertex_stride >= 12;
}

static clodb200_device_geometry* upload_geometry_locked(const clodb200_geometry& g, const clodb200_builder_settings* settings)
{
	if (g.index_count % 3 || g.vertex_stride % 4 || g.uv_set_count > kMaxUvSets || g.vertex_count >= 0xffffffffull || g.index_count / 3 >= 0xffffffffull)
		throw Error("clodb200: index count must be a multiple of 3, the vertex stride a multiple of 4, at most 4 UV sets");
	clodb200_device_geometry* dg = new clodb200_device_geometry();
	try
	{
		dg->settings = to_settings(settings);
		const BuilderSettings& st = dg->settings;
		DeviceGeometry& geo = dg->geometry;
		const size_t V = g.vertex_count;
		u8* dv = static_cast<u8*>(block_alloc(V * g.vertex_stride, dg->allocations));
		dev_h2d(dv, g.vertices, V * g.vertex_stride);
		geo.vertices = dv;
		geo.vertex_stride = g.vertex_stride;
		geo.vertex_flags = g.vertex_flags;
		geo.vertex_count = V;
		u32* di = static_cast<u32*>(block_alloc(g.index_count * 4, dg->allocations));
		dev_h2d(di, g.indices, g.index_count * 4);
		geo.indices = di;
		geo.index_count = g.index_count;
		u32* bad = static_cast<u32*>(block_alloc(sizeof(u32), dg->allocations));
		dev_memset(bad, 0, sizeof(u32));
		LAUNCH(k_index_range, g.index_count, di, g.index_count, u32(V), bad);
		if (dev_read(bad))
			throw Error("clodb200: index out of range");

		geo.uv_set_count = u32(g.uv_set_count);
		for (size_t s = 0; s < g.uv_set_count; ++s)
		{
			float* duv = static_cast<float*>(block_alloc(V * 8, dg->allocations));
			if (g.uv_sets[s].values && g.uv_sets[s].count == V)
				dev_h2d(duv, g.uv_sets[s].values, V * 8);
			else
				dev_memset(duv, 0, V * 8);
			geo.uv_values[s] = duv;
			geo.uv_stride[s] = 2;
		}

		// skinned meshes: the second vertex stream only feeds the page writer (joints, weights, bone lists)
		if (g.skinning_vertices && g.skinning_vertex_bytes && g.skinning_vertex_stride)
		{
			if (g.skinning_vertex_stride % 4)
				throw Error("clodb200: the skinning vertex stride must be a multiple of 4");
			u8* dsk = static_cast<u8*>(block_alloc(g.skinning_vertex_bytes, dg->allocations));
			dev_h2d(dsk, g.skinning_vertices, g.skinning_vertex_bytes);
			geo.skinning_vertices = dsk;
			geo.skinning_stride = g.skinning_vertex_stride;
			geo.skinning_vertex_count = g.skinning_vertex_bytes / g.skinning_vertex_stride;
		}

		// simplification attribute stream (ClusterLODUtilities.cpp:5359-5410): normals x3, then tangent xyz + sign
		const bool has_normals = (g.vertex_flags & kVertexNormals) != 0 && g.vertex_stride >= 24;
		const bool has_texcoords = (g.vertex_flags & kVertexTexcoords) != 0 && g.vertex_stride >= 32;
		const bool use_normals = st.enable_normal_attribute_simplification && has_normals;
		bool wants_tangents = use_normals && has_texcoords;
		DeviceMesh& mesh = geo.mesh;
		float* dpos = static_cast<float*>(block_alloc(V * 12, dg->allocations));
		float* dtan = nullptr;
		if (wants_tangents)
		{
			dtan = static_cast<float*>(block_alloc(V * 16, dg->allocations));
			if (g.tangents)
				dev_h2d(dtan, g.tangents, V * 16);
			else if (g.vertex_stride % 4 == 0 && g.index_count % 3 == 0)
			{
				// GenerateMikkTangents runs inside every build call, as in the reference (:5359-5366); see build_artifacts_locked
				dev_memset(dtan, 0, V * 16);
				geo.generated_tangents4 = dtan;
			}
			else
			{
				// the generator refuses the input (:665-675): the build goes on without tangent attributes
				wants_tangents = false;
				dtan = nullptr;
			}
		}
		u32 acount = (use_normals ? 3u : 0u) + (wants_tangents ? 4u : 0u);
		float* dattr = acount ? static_cast<float*>(block_alloc(V * acount * 4, dg->allocations)) : nullptr;
		split_vertex_streams(dv, g.vertex_stride, V, dpos, dattr, acount, use_normals, dtan);
		mesh.positions = dpos;
		mesh.vertex_count = V;
		if (acount)
		{
			mesh.attributes = dattr;
			geo.attributes_rw = dattr;
			geo.tangent_column = use_normals ? 3u : 0u;
			mesh.attribute_stride = acount;
			mesh.attribute_count = acount;
			u32 k = 0;
			if (use_normals)
			{
				float w = std::max(0.0f, st.normal_attribute_weight);
				mesh.attribute_weights[0] = mesh.attribute_weights[1] = mesh.attribute_weights[2] = w;
				mesh.attribute_protect_mask |= 7u << k;
				k += 3;
			}
			if (wants_tangents)
			{
				float w = std::max(0.0f, st.simplify_tangent_weight);
				mesh.attribute_weights[k] = mesh.attribute_weights[k + 1] = mesh.attribute_weights[k + 2] = w;
				mesh.attribute_weights[k + 3] = std::max(0.0f, st.simplify_tangent_sign_weight);
				mesh.attribute_protect_mask |= 15u << k;
			}
		}
	}
	catch (...)
	{
		block_release(dg->allocations);
		delete dg;
		throw;
	}
	return dg;
}

static void free_geometry_locked(clodb200_device_geometry* dg)
{
	if (!dg)
		return;
	block_release(dg->allocations);
	delete dg;
}

static clodb200_artifacts* take_artifacts()
{
	clodb200_artifacts* a = g_artifacts_pool ? g_artifacts_pool : new clodb200_artifacts();
	g_artifacts_pool = nullptr;
	a->data.blobs.clear();
	a->data.page_bytes = 0;
	return a;
}

static clodb200_artifacts* build_artifacts_once(const clodb200_device_geometry* dg)
{
	clodb200_artifacts* a = take_artifacts();
	try
	{
		size_t T = dg->geometry.index_count / 3, V = dg->geometry.vertex_count;
		size_t scale_temp = 640, scale_persist = 96, base_temp = 64u << 20;
		if (const char* e = getenv("CLODB200_TEMP_BYTES_PER_TRI"))
			scale_temp = size_t(atoll(e));
		if (const char* e = getenv("CLODB200_TEMP_BYTES_BASE"))
			base_temp = size_t(atoll(e));
		size_t temp_bytes = T * scale_temp + V * 16 + base_temp;
		const DeviceGeometry& geo = dg->geometry;
		if (geo.generated_tangents4)
			temp_bytes = std::max(temp_bytes, mikk_temp_bytes(V, geo.index_count));
		ensure_workspace(T * scale_persist + V * 32 + (64u << 20), temp_bytes);
		if (geo.generated_tangents4)
		{
			if (!mikk_tangents(geo.vertices, geo.vertex_stride, V, geo.indices, geo.index_count, geo.generated_tangents4, g_ws.temp))
				throw Error("clodb200: tangent generation refused an input that passed the upload checks");
			write_tangent_columns(geo.generated_tangents4, V, geo.attributes_rw, geo.mesh.attribute_stride, geo.tangent_column);
		}
		build_artifacts(dg->geometry, dg->settings, g_ws, a->data, g_last_build_stats);
		a->data.stats[12] = g_launches;
	}
	catch (...)
	{
		a->data.pages.destroy();
		delete a;
		throw;
	}
	return a;
}

static clodb200_artifacts* build_artifacts_locked(const clodb200_device_geometry* dg)
{
	return with_arena_growth([&]() { return build_artifacts_once(dg); }, []() { return true; });
}

clodb200_builder_settings clodb200_defaultBuilderSettings(void)
{
	clodb200_builder_settings s;
	s.lodErrorMergePrevious = 1.5f;
	s.lodErrorMergeAdditive = 0.0f;
	s.partitionSizeFloor = 8u;
	s.preserveImportedNormals = 1;
	s.enableNormalAttributeSimplification = 1;
	s.normalAttributeWeight = 1.0f;
	s.simplifyTangentWeight = 0.01f;
	s.simplifyTangentSignWeight = 0.5f;
	return s;
}

clodb200_device_geometry* clodb200_geometryUpload(const clodb200_geometry* geometry, const clodb200_builder_settings* settings)
{
	clodb200_device_geometry* dg = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!geometry || !geometry_is_buildable(*geometry))
			return fail(CLODB200_ERR_INVALID, "clodb200: invalid or empty geometry");
		dg = upload_geometry_locked(*geometry, settings);
		return CLODB200_OK;
	});
	return dg;
}

void clodb200_geometryFree(clodb200_device_geometry* geometry)
{
	guarded([&]() -> int {
		free_geometry_locked(geometry);
		return CLODB200_OK;
	});
}

clodb200_artifacts* clodb200_geometryBuildArtifacts(const clodb200_device_geometry* geometry)
{
	clodb200_artifacts* a = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!geometry)
			return fail(CLODB200_ERR_INVALID, "clodb200_geometryBuildArtifacts: null geometry");
		a = build_artifacts_locked(geometry);
		return CLODB200_OK;
	});
	return a;
}

clodb200_artifacts* clodb200_buildArtifacts(const clodb200_geometry* geometry, const clodb200_builder_settings* settings)
{
	clodb200_artifacts* a = nullptr;
	guarded([&]() -> int {
		t_last_error.clear();
		if (!geometry)
			return fail(CLODB200_ERR_INVALID, "clodb200_buildArtifacts: null geometry");
		if (!geometry_is_buildable(*geometry))
		{
			// clodBuildEx returns 0 and the builder hands back empty artifacts (ClusterLODUtilities.cpp:4608-4609)
			a = take_artifacts();
			return CLODB200_OK;
		}
		clodb200_device_geometry* dg = upload_geometry_locked(*geometry, settings);
		try
		{
			a = build_artifacts_locked(dg);
		}
		catch (...)
		{
			free_geometry_locked(dg);
			throw;
		}
		free_geometry_locked(dg);
		return CLODB200_OK;
	});
	return a;
}

int clodb200_artifactsGet(const clodb200_artifacts* artifacts, const char* name, const void** out_ptr, size_t* out_bytes)
{
	*out_ptr = nullptr;
	*out_bytes = 0;
	if (!artifacts || !name)
		return 0;
	if (!strcmp(name, "meshPages"))
	{
		*out_ptr = artifacts->data.pages.base;
		*out_bytes = artifacts->data.page_bytes;
		return 1;
	}
	if (!strcmp(name, "stats"))
	{
		*out_ptr = artifacts->data.stats;
		*out_bytes = sizeof(artifacts->data.stats);
		return 1;
	}
	auto it = artifacts->data.blobs.find(name);
	if (it == artifacts->data.blobs.end())
		return 0;
	*out_ptr = it->second.data();
	*out_bytes = it->second.size();
	return 1;
}

void clodb200_artifactsFree(clodb200_artifacts* artifacts)
{
	if (!artifacts)
		return;
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (!g_artifacts_pool)
		g_artifacts_pool = artifacts;
	else
	{
		artifacts->data.pages.destroy();
		delete artifacts;
	}
}

size_t clodb200_artifactsSerializeMetadata(const clodb200_artifacts* artifacts, const char* container_file_name, const char* source_identifier, const char* prim_path,
    const char* subset_name, uint64_t build_config_hash, void* buffer, size_t capacity)
{
	if (!artifacts || artifacts->data.blobs.find("counts") == artifacts->data.blobs.end())
		return 0;
	CacheIdentity id;
	id.source_identifier = source_identifier ? source_identifier : "";
	id.prim_path = prim_path ? prim_path : "";
	id.subset_name = subset_name ? subset_name : "";
	id.build_config_hash = build_config_hash;
	std::vector<u8> blob = serialize_cache_metadata(artifacts->data, id, container_file_name ? container_file_name : "");
	if (buffer && capacity)
		memcpy(buffer, blob.data(), std::min(capacity, blob.size()));
	return blob.size();
}

int clodb200_artifactsSaveCache(const clodb200_artifacts* artifacts, const char* directory, const char* container_file_name, const char* metadata_file_name,
    const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash)
{
	try
	{
		if (!artifacts || !directory || !container_file_name || !metadata_file_name || artifacts->data.blobs.find("counts") == artifacts->data.blobs.end())
			return fail(CLODB200_ERR_INVALID, "clodb200_artifactsSaveCache: null argument or empty artifacts");
		CacheIdentity id;
		id.source_identifier = source_identifier ? source_identifier : "";
		id.prim_path = prim_path ? prim_path : "";
		id.subset_name = subset_name ? subset_name : "";
		id.build_config_hash = build_config_hash;
		std::string dir(directory);
		write_cache_container(artifacts->data, dir + "/" + container_file_name);
		std::vector<u8> blob = serialize_cache_metadata(artifacts->data, id, container_file_name);
		FILE* f = fopen((dir + "/" + metadata_file_name).c_str(), "wb");
		if (!f)
			return fail(CLODB200_ERR_RUNTIME, "clodb200: cannot open the metadata file");
		size_t written = fwrite(blob.data(), 1, blob.size(), f);
		fclose(f);
		if (written != blob.size())
			return fail(CLODB200_ERR_RUNTIME, "clodb200: short write of the metadata file");
		return CLODB200_OK;
	}
	catch (const std::exception& e)
	{
		return fail(CLODB200_ERR_RUNTIME, e.what());
	}
}

// Skip-if-cached (CLodCacheLoader::TryLoadPrebuilt -> CLodCache::TryLoad, CLodCacheLoader.cpp:218-234, CLodCache.cpp:635-713):
// is there a cache for this identity and build configuration that the loader would accept? Walks the metadata blob the way
// DeserializeMetadata does (CLodCache.cpp:209-250) and applies its acceptance rules: schema version, build hash and identity equal,
// blob consumed exactly, container present with the container magic/version and one locator per mesh page. No GPU involved.
int clodb200_cacheProbe(const char* directory, const char* metadata_file_name, const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash)
{
	if (!directory || !metadata_file_name)
		return 0;
	const std::string dir(directory);
	FILE* f = fopen((dir + "/" + metadata_file_name).c_str(), "rb");
	if (!f)
		return 0;
	std::vector<u8> blob;
	u8 chunk[65536];
	for (size_t n; (n = fread(chunk, 1, sizeof(chunk), f)) > 0;)
		blob.insert(blob.end(), chunk, chunk + n);
	fclose(f);
	size_t off = 0;
	bool ok = true;
	auto pod = [&](void* out, size_t bytes) {
		if (!ok || off + bytes > blob.size())
		{
			ok = false;
			return;
		}
		memcpy(out, blob.data() + off, bytes);
		off += bytes;
	};
	auto skip_vector = [&](size_t element) -> u64 {
		u64 count = 0;
		pod(&count, 8);
		if (!ok || count > (blob.size() - off) / (element ? element : 1))
		{
			ok = false;
			return 0;
		}
		off += size_t(count) * element;
		return count;
	};
	auto string = [&]() -> std::string {
		u64 n = 0;
		pod(&n, 8);
		if (!ok || n > blob.size() - off)
		{
			ok = false;
			return std::string();
		}
		std::string s(reinterpret_cast<const char*>(blob.data() + off), size_t(n));
		off += size_t(n);
		return s;
	};
	u32 schema = 0;
	u64 hash = 0, hash2 = 0;
	pod(&schema, 4);
	pod(&hash, 8);
	skip_vector(76); // groups
	skip_vector(16); // segments
	skip_vector(16); // segmentBounds
	off += 16;       // objectBoundingSphere
	u8 has_chunks = 0;
	pod(&has_chunks, 1);
	if (has_chunks)
		skip_vector(20);
	skip_vector(16); // groupDiskLocators
	const u64 page_locators = skip_vector(16);
	skip_vector(4); // groupPageReferences
	skip_vector(4); // groupPageReferenceOffsets
	u32 page_counts[3] = {0, 0, 0};
	pod(page_counts, 12);
	const std::string src = string(), prim = string(), subset = string();
	pod(&hash2, 8);
	const std::string container = string();
	skip_vector(64); // nodes
	skip_vector(8);  // lodNodeRanges
	skip_vector(4);  // lodLevelRoots
	off += 8;        // maxDepth, maxTraversalDepth
	if (!ok || off != blob.size() || schema != 47 || hash != build_config_hash || hash2 != build_config_hash)
		return 0;
	if (src != (source_identifier ? source_identifier : "") || prim != (prim_path ? prim_path : "") || subset != (subset_name ? subset_name : ""))
		return 0;
	if (page_locators != u64(page_counts[0]) + page_counts[2])
		return 0;
	FILE* c = fopen((dir + "/" + container).c_str(), "rb");
	if (!c)
		return 0;
	u32 header[4] = {0, 0, 0, 0};
	size_t got = fread(header, 1, sizeof(header), c);
	fclose(c);
	return got == sizeof(header) && header[0] == 0x444F4C43u && header[1] == 4u && header[3] == u32(page_locators) ? 1 : 0;
}

#ifndef CLODB_EMU
static thread_local cudaEvent_t g_timer_start = nullptr, g_timer_stop = nullptr;
#endif

extern "C++"
{
namespace
{
void release_thread_context() noexcept
{
	if (!t_context_ready)
		return;
	try
	{
#ifndef CLODB_EMU
		if (g_stream)
			cudaStreamSynchronize(g_stream);
#endif
		dev_d2h_async_wait();
	}
	catch (...)
	{
	}
	try
	{
		if (t_ws)
		{
			t_ws->persist.destroy();
			t_ws->temp.destroy();
			t_ws->stage.destroy();
			delete t_ws;
			t_ws = nullptr;
		}
		if (t_block_cache)
		{
			for (const DeviceBlock& b : *t_block_cache)
				dev_free(b.ptr);
			delete t_block_cache;
			t_block_cache = nullptr;
		}
		delete g_record_pool;
		g_record_pool = nullptr;
		if (g_artifacts_pool)
		{
			g_artifacts_pool->data.pages.destroy();
			delete g_artifacts_pool;
			g_artifacts_pool = nullptr;
		}
#ifndef CLODB_EMU
		if (g_timer_start)
		{
			cudaEventDestroy(g_timer_start);
			cudaEventDestroy(g_timer_stop);
			g_timer_start = g_timer_stop = nullptr;
		}
#endif
	}
	catch (...)
	{
	}
	rt_thread_release();
	t_context_ready = false;
}
} // namespace
} // extern "C++"

void clodb200_timerStart(void)
{
#ifndef CLODB_EMU
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (!g_timer_start)
	{
		cudaEventCreate(&g_timer_start);
		cudaEventCreate(&g_timer_stop);
	}
	cudaStreamSynchronize(g_stream);
	cudaEventRecord(g_timer_start, g_stream);
#endif
}

float clodb200_timerStop(void)
{
	float ms = 0.f;
#ifndef CLODB_EMU
	std::lock_guard<std::mutex> lock(g_api_mutex);
	if (!g_timer_start)
		return 0.f;
	cudaEventRecord(g_timer_stop, g_stream);
	cudaEventSynchronize(g_timer_stop);
	cudaEventElapsedTime(&ms, g_timer_start, g_timer_stop);
#endif
	return ms;
}

void clodb200_profileEnable(int enable)
{
	std::lock_guard<std::mutex> lock(g_api_mutex);
	g_profile = enable;
}

size_t clodb200_profileReport(char* buffer, size_t capacity)
{
	std::lock_guard<std::mutex> lock(g_api_mutex);
	std::string report;
	try
	{
		report = profile_report();
	}
	catch (const std::exception& e)
	{
		report = std::string("error,0,0 ") + e.what() + "\n";
	}
	if (buffer && capacity)
	{
		size_t n = std::min(capacity - 1, report.size());
		memcpy(buffer, report.data(), n);
		buffer[n] = 0;
	}
	return report.size() + 1;
}

// ---- device-wide primitives, exposed for their own parity tests and micro-benchmarks (tests/test_prims.py) ------------
// Each call uploads the input, runs the primitive `repeat` times on the build stream (first run untimed when repeat > 1),
// downloads the result of the last run and returns the mean device time of the timed runs in *ms (may be NULL).
static float timed_runs(int repeat, const std::function<void()>& run)
{
	float ms = 0.f;
	run();
#ifndef CLODB_EMU
	if (repeat > 1)
	{
		cudaEvent_t a, b;
		cudaEventCreate(&a);
		cudaEventCreate(&b);
		cudaEventRecord(a, g_stream);
		for (int i = 1; i < repeat; ++i)
			run();
		cudaEventRecord(b, g_stream);
		cudaEventSynchronize(b);
		cudaEventElapsedTime(&ms, a, b);
		ms /= float(repeat - 1);
		cudaEventDestroy(a);
		cudaEventDestroy(b);
	}
#else
	for (int i = 1; i < repeat; ++i)
		run();
#endif
	return ms;
}

int clodb200_primExclusiveScanU32(const unsigned int* in, unsigned int* out, size_t n, unsigned int* total, int repeat, float* ms)
{
	return guarded([&]() -> int {
		if (n && (!in || !out))
			return fail(CLODB200_ERR_INVALID, "clodb200_primExclusiveScanU32: invalid arguments");
		ensure_workspace(n * 8 + (1 << 20), 1 << 20);
		u32* din = g_ws.persist.alloc<u32>(n + 1);
		u32* dout = g_ws.persist.alloc<u32>(n + 1);
		u32* dtotal = g_ws.persist.alloc<u32>(4);
		dev_h2d(din, in, n * sizeof(u32));
		float t = timed_runs(repeat, [&]() { exclusive_scan_u32(din, dout, n, dtotal, g_ws.temp); });
		dev_d2h(out, dout, n * sizeof(u32));
		if (total)
			dev_d2h(total, dtotal, sizeof(u32));
		if (ms)
			*ms = t;
		return CLODB200_OK;
	});
}

int clodb200_primAcosf(const float* in, float* out, size_t n)
{
	return guarded([&]() -> int {
		if (n && (!in || !out))
			return fail(CLODB200_ERR_INVALID, "clodb200_primAcosf: invalid arguments");
		ensure_workspace(n * 8 + (1 << 20), 1 << 20);
		float* din = g_ws.persist.alloc<float>(n + 1);
		float* dout = g_ws.persist.alloc<float>(n + 1);
		dev_h2d(din, in, n * sizeof(float));
		mikk_acosf(din, dout, n);
		dev_d2h(out, dout, n * sizeof(float));
		return CLODB200_OK;
	});
}

int clodb200_primExclusiveMaxScanU64(const uint64_t* in, uint64_t* out, size_t n, int repeat, float* ms)
{
	return guarded([&]() -> int {
		if (n && (!in || !out))
			return fail(CLODB200_ERR_INVALID, "clodb200_primExclusiveMaxScanU64: invalid arguments");
		ensure_workspace(n * 16 + (1 << 20), 1 << 20);
		u64* din = g_ws.persist.alloc<u64>(n + 1);
		u64* dout = g_ws.persist.alloc<u64>(n + 1);
		dev_h2d(din, in, n * sizeof(u64));
		float t = timed_runs(repeat, [&]() { exclusive_scan<u64, OpMaxU64>(din, dout, n, nullptr, g_ws.temp); });
		dev_d2h(out, dout, n * sizeof(u64));
		if (ms)
			*ms = t;
		return CLODB200_OK;
	});
}

int clodb200_primSetScanEpoch(unsigned int epoch)
{
	return guarded([&]() -> int {
#ifndef CLODB_EMU
		g_scan_chain.epoch = epoch & 0x3fffffffu;
#else
		(void)epoch;
#endif
		return CLODB200_OK;
	});
}

int clodb200_primSortPairsU32(unsigned int* keys, unsigned int* values, size_t n, int bit_lo, int bit_hi, int repeat, float* ms)
{
	return guarded([&]() -> int {
		if (n && (!keys || !values))
			return fail(CLODB200_ERR_INVALID, "clodb200_primSortPairsU32: invalid arguments");
		if (bit_lo < 0 || bit_hi > 32 || bit_hi < bit_lo)
			return fail(CLODB200_ERR_INVALID, "clodb200_primSortPairsU32: invalid bit range");
		ensure_workspace(n * 24 + (1 << 20), n * 8 + (16 << 20));
		u32* k0 = g_ws.persist.alloc<u32>(n + 1);
		u32* v0 = g_ws.persist.alloc<u32>(n + 1);
		u32* k = g_ws.persist.alloc<u32>(n + 1);
		u32* v = g_ws.persist.alloc<u32>(n + 1);
		u32* kt = g_ws.persist.alloc<u32>(n + 1);
		u32* vt = g_ws.persist.alloc<u32>(n + 1);
		dev_h2d(k0, keys, n * sizeof(u32));
		dev_h2d(v0, values, n * sizeof(u32));
		// the copies that restore the unsorted input are part of every run (and of the reported time)
		float t = timed_runs(repeat, [&]() {
			dev_d2d(k, k0, n * sizeof(u32));
			dev_d2d(v, v0, n * sizeof(u32));
			radix_sort_pairs<u32>(k, kt, v, vt, n, bit_lo, bit_hi, g_ws.temp);
		});
		dev_d2h(keys, k, n * sizeof(u32));
		dev_d2h(values, v, n * sizeof(u32));
		if (ms)
			*ms = t;
		return CLODB200_OK;
	});
}

void clodb200_simplifyStats(unsigned int out3[3])
{
	out3[0] = g_simplify_stats.passes;
	out3[1] = g_simplify_stats.rounds;
	out3[2] = g_simplify_stats.max_rounds;
}

} // extern "C"
