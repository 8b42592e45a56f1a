// ORACLE (test infrastructure, NOT product code; nothing under basicrenderer_b200/ links this).
//
// Thin C-ABI driver around the UNMODIFIED reference sources, compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libclodref.so:
//   - ThirdParty/meshoptimizer/src/*.cpp            (meshopt_* kernels, extern "C", exported as-is)
//   - BasicRenderer/include/ThirdParty/meshoptimizer/clusterlod.h (clodBuildEx + clod:: internals)
//
// Besides re-exporting the reference's own C ABI (meshopt_*, clodBuildEx, clodLocalIndices ...), this file adds
//   clodref_dag_build      : runs the reference's real clodBuildEx and records the callback stream
//   clodref_dag_build_dump : walks the same loop as clodBuildEx (clusterlod.h:792-943) by calling the reference's
//                            own clod::clusterize / partition / lockBoundary / boundsMerge / simplify, and records
//                            every stage's inputs and outputs per DAG level, so each CUDA stage can be checked
//                            against the reference "given the same cluster/group assignment" (SURVEY.md §8a).
//                            The recorded callback stream must equal clodref_dag_build's (tests/test_oracle.py).
// Results are returned as named raw blobs (clodref_blob_get).
#include <meshoptimizer.h>

// Observation hook, not a modification: inside the reference's clusterlod.h the one call of meshopt_simplifySloppy (the
// fallback of clod::simplify, clusterlod.h:591) is routed through a forwarding wrapper that counts calls, so tests can
// compare how many groups of a level needed the fallback. The wrapper calls the real, unmodified function.
extern "C" size_t clodref_counted_simplifySloppy(unsigned int* destination, const unsigned int* indices, size_t index_count, const float* vertex_positions, size_t vertex_count, size_t vertex_positions_stride,
    const unsigned char* vertex_lock, size_t target_index_count, float target_error, float* result_error);
#define meshopt_simplifySloppy clodref_counted_simplifySloppy
#define CLUSTERLOD_IMPLEMENTATION
#include <ThirdParty/meshoptimizer/clusterlod.h>
#undef meshopt_simplifySloppy

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

namespace
{

struct BlobStore
{
	std::map<std::string, std::vector<unsigned char> > blobs;

	template <typename T>
	void put(const std::string& name, const std::vector<T>& v)
	{
		std::vector<unsigned char>& b = blobs[name];
		b.resize(v.size() * sizeof(T));
		if (!v.empty())
			memcpy(b.data(), v.data(), b.size());
	}

	template <typename T>
	void append(const std::string& name, const T* data, size_t count)
	{
		std::vector<unsigned char>& b = blobs[name];
		size_t old = b.size();
		b.resize(old + count * sizeof(T));
		if (count)
			memcpy(b.data() + old, data, count * sizeof(T));
	}

	template <typename T>
	void push(const std::string& name, T value)
	{
		append(name, &value, 1);
	}
};

struct OutputRecorder
{
	BlobStore* store;
	int next_group;
	uint32_t cluster_total;
	uint32_t index_total;

	explicit OutputRecorder(BlobStore* s)
	    : store(s), next_group(0), cluster_total(0), index_total(0)
	{
		store->push<uint32_t>("out.group_cluster_offsets", 0);
		store->push<uint32_t>("out.cluster_index_offsets", 0);
	}

	int record(clodGroup group, const clodCluster* clusters, size_t cluster_count)
	{
		store->push<int32_t>("out.group_depth", group.depth);
		store->append<float>("out.group_simplified", group.simplified.center, 3);
		store->push<float>("out.group_simplified", group.simplified.radius);
		store->push<float>("out.group_simplified", group.simplified.error);

		for (size_t i = 0; i < cluster_count; ++i)
		{
			const clodCluster& c = clusters[i];
			store->push<int32_t>("out.cluster_refined", c.refined);
			store->append<float>("out.cluster_bounds", c.bounds.center, 3);
			store->push<float>("out.cluster_bounds", c.bounds.radius);
			store->push<float>("out.cluster_bounds", c.bounds.error);
			store->push<uint32_t>("out.cluster_vertex_count", uint32_t(c.vertex_count));
			store->append<unsigned int>("out.cluster_indices", c.indices, c.index_count);
			index_total += uint32_t(c.index_count);
			store->push<uint32_t>("out.cluster_index_offsets", index_total);
		}

		cluster_total += uint32_t(cluster_count);
		store->push<uint32_t>("out.group_cluster_offsets", cluster_total);
		return next_group++;
	}

	static int callback(void* ctx, clodGroup group, const clodCluster* clusters, size_t cluster_count, size_t, unsigned int)
	{
		return static_cast<OutputRecorder*>(ctx)->record(group, clusters, cluster_count);
	}
};

std::string levelKey(int depth, const char* name)
{
	return "L" + std::to_string(depth) + "." + name;
}

} // namespace

extern "C"
{

struct clodref_handle
{
	BlobStore store;
	size_t cluster_count;
};

// The effective builder configuration of the reference (ClusterLODUtilities.cpp:5426-5460 on top of
// clodDefaultConfig(128), clusterlod.h:742-774); SURVEY.md §2.4.
clodConfig clodref_builder_config(void)
{
	clodConfig c = clodDefaultConfig(128);
	c.max_vertices = 128;
	c.max_triangles = 128;
	c.min_triangles = 64;
	c.cluster_spatial = true;
	c.cluster_fill_weight = 0.5f;
	c.cluster_split_factor = 2.0f;
	c.partition_spatial = true;
	c.partition_sort = true;
	c.optimize_clusters = true;
	c.optimize_bounds = true;
	c.simplify_fallback_permissive = false;
	c.simplify_error_factor_sloppy = 100.f;
	c.simplify_fallback_sloppy = true;
	c.simplify_regularize = false;
	c.simplify_error_merge_previous = 1.5f;
	c.simplify_error_merge_additive = 0.0f;
	c.partition_size = 384;
	c.partition_max_refined_groups = 8;
	return c;
}

size_t clodref_config_size(void)
{
	return sizeof(clodConfig);
}

static clodMesh makeMesh(const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, unsigned int protect_mask, const unsigned char* vertex_lock)
{
	clodMesh mesh = {};
	mesh.indices = indices;
	mesh.index_count = index_count;
	mesh.vertex_count = vertex_count;
	mesh.vertex_positions = positions;
	mesh.vertex_positions_stride = positions_stride;
	mesh.vertex_attributes = attributes;
	mesh.vertex_attributes_stride = attributes_stride;
	mesh.vertex_lock = vertex_lock;
	mesh.attribute_weights = attribute_weights;
	mesh.attribute_count = attribute_count;
	mesh.attribute_protect_mask = protect_mask;
	return mesh;
}

// Runs the reference's real clodBuildEx (serial) and records the output callback stream.
clodref_handle* clodref_dag_build(const clodConfig* config, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, unsigned int protect_mask, const unsigned char* vertex_lock)
{
	clodref_handle* h = new clodref_handle();
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, attributes, attributes_stride, attribute_weights, attribute_count, protect_mask, vertex_lock);

	OutputRecorder recorder(&h->store);
	h->cluster_count = clodBuildEx(*config, mesh, &recorder, &OutputRecorder::callback, NULL);
	return h;
}

// Multi-threaded variant: the reference's own inner-threaded path. clodBuildEx calls the iteration callback once per DAG
// level; the callback fans the per-group tasks out over `threads` host threads, which is exactly how the renderer drives
// it through TaskSchedulerManager::ParallelFor (ClusterLODUtilities.cpp:5587-5593). No output is recorded unless
// `record` is non-zero (timing runs).
struct MtContext
{
	unsigned int threads;
};

static void mtIterate(void* iteration_context, void* output_context, int, size_t task_count)
{
	// output_context is the recorder/discard context; the thread count travels in a static
	(void)output_context;
	extern unsigned int g_clodref_threads;
	unsigned int n = g_clodref_threads ? g_clodref_threads : 1;
	if (n <= 1 || task_count <= 1)
	{
		for (size_t i = 0; i < task_count; ++i)
			clodBuild_iterationTask(iteration_context, i, 0);
		return;
	}
	std::atomic<size_t> next(0);
	std::vector<std::thread> pool;
	auto worker = [&](unsigned int tid) {
		for (;;)
		{
			size_t i = next.fetch_add(1);
			if (i >= task_count)
				break;
			clodBuild_iterationTask(iteration_context, i, tid);
		}
	};
	unsigned int count = unsigned(std::min<size_t>(n, task_count));
	for (unsigned int t = 1; t < count; ++t)
		pool.emplace_back(worker, t);
	worker(0);
	for (std::thread& t : pool)
		t.join();
}

unsigned int g_clodref_threads = 1;
static std::atomic<unsigned int> g_sloppy_calls(0);

size_t clodref_counted_simplifySloppy(unsigned int* destination, const unsigned int* indices, size_t index_count, const float* vertex_positions, size_t vertex_count, size_t vertex_positions_stride,
    const unsigned char* vertex_lock, size_t target_index_count, float target_error, float* result_error)
{
	g_sloppy_calls.fetch_add(1);
	return meshopt_simplifySloppy(destination, indices, index_count, vertex_positions, vertex_count, vertex_positions_stride, vertex_lock, target_index_count, target_error, result_error);
}

unsigned int clodref_sloppy_calls(void)
{
	return g_sloppy_calls.load();
}

static int discardCallback(void* ctx, clodGroup, const clodCluster*, size_t, size_t, unsigned int)
{
	int* next = static_cast<int*>(ctx);
	return (*next)++;
}

size_t clodref_dag_build_mt(const clodConfig* config, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, unsigned int protect_mask, unsigned int threads)
{
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, attributes, attributes_stride, attribute_weights, attribute_count, protect_mask, NULL);
	g_clodref_threads = threads;
	clodBuildParallelConfig parallel = {};
	parallel.iteration_callback = &mtIterate;
	int next = 0;
	return clodBuildEx(*config, mesh, &next, &discardCallback, &parallel);
}

// The reference's real clodBuildEx on `threads` host threads, keeping only per-depth statistics of the callback stream
// (no index lists): what the scale-parity tests compare at sizes where a full dump would take minutes.
//   stats.level_groups / level_clusters / level_triangles (u32 per depth), stats.level_max_error (f32, finite errors only),
//   stats.level_sloppy (u32: fallback calls while that depth's groups were simplified)
//   stats.group_error (f32) / group_depth (i32) / group_clusters (u32): one entry per emitted group, callback order
struct StatsRecorder
{
	std::vector<uint32_t> groups, clusters, triangles, sloppy;
	std::vector<float> max_error;
	std::vector<float> group_error; // every group's emitted error, callback order
	std::vector<int32_t> group_depth;
	std::vector<uint32_t> group_clusters;
	int next_group = 0;
	unsigned int sloppy_seen = 0;

	void touch(size_t depth)
	{
		if (groups.size() <= depth)
		{
			groups.resize(depth + 1);
			clusters.resize(depth + 1);
			triangles.resize(depth + 1);
			sloppy.resize(depth + 1);
			max_error.resize(depth + 1);
		}
	}

	static int callback(void* ctx, clodGroup group, const clodCluster* cl, size_t cluster_count, size_t, unsigned int)
	{
		StatsRecorder* r = static_cast<StatsRecorder*>(ctx);
		size_t d = size_t(group.depth);
		r->touch(d);
		// all iteration tasks of a depth finish before its first callback (clusterlod.h:884-926)
		unsigned int now = g_sloppy_calls.load();
		r->sloppy[d] += now - r->sloppy_seen;
		r->sloppy_seen = now;
		r->groups[d]++;
		r->group_error.push_back(group.simplified.error);
		r->group_depth.push_back(group.depth);
		r->group_clusters.push_back(uint32_t(cluster_count));
		r->clusters[d] += uint32_t(cluster_count);
		for (size_t i = 0; i < cluster_count; ++i)
			r->triangles[d] += uint32_t(cl[i].index_count / 3);
		if (group.simplified.error < FLT_MAX && group.simplified.error > r->max_error[d])
			r->max_error[d] = group.simplified.error;
		return r->next_group++;
	}
};

clodref_handle* clodref_dag_build_stats(const clodConfig* config, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, unsigned int protect_mask, unsigned int threads)
{
	clodref_handle* h = new clodref_handle();
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, attributes, attributes_stride, attribute_weights, attribute_count, protect_mask, NULL);
	g_clodref_threads = threads;
	clodBuildParallelConfig parallel = {};
	parallel.iteration_callback = &mtIterate;
	StatsRecorder rec;
	rec.sloppy_seen = g_sloppy_calls.load();
	h->cluster_count = clodBuildEx(*config, mesh, &rec, &StatsRecorder::callback, &parallel);
	h->store.put("stats.level_groups", rec.groups);
	h->store.put("stats.level_clusters", rec.clusters);
	h->store.put("stats.level_triangles", rec.triangles);
	h->store.put("stats.level_sloppy", rec.sloppy);
	h->store.put("stats.level_max_error", rec.max_error);
	h->store.put("stats.group_error", rec.group_error);
	h->store.put("stats.group_depth", rec.group_depth);
	h->store.put("stats.group_clusters", rec.group_clusters);
	return h;
}

// Same loop as clodBuildEx (clusterlod.h:792-943) expressed with the reference's own clod:: functions, recording
// the inputs/outputs of each stage per DAG level.
clodref_handle* clodref_dag_build_dump(const clodConfig* config_ptr, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, unsigned int protect_mask, const unsigned char* vertex_lock)
{
	using namespace clod;

	clodref_handle* h = new clodref_handle();
	BlobStore& store = h->store;
	clodConfig config = *config_ptr;
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, attributes, attributes_stride, attribute_weights, attribute_count, protect_mask, vertex_lock);

	OutputRecorder recorder(&store);

	std::vector<unsigned char> locks(mesh.vertex_count);

	// clusterlod.h:825-826
	std::vector<unsigned int> remap(mesh.vertex_count);
	meshopt_generatePositionRemap(&remap[0], mesh.vertex_positions, mesh.vertex_count, mesh.vertex_positions_stride);
	store.put("remap", remap);

	// clusterlod.h:829-841
	if (mesh.attribute_protect_mask)
	{
		size_t max_attributes = mesh.vertex_attributes_stride / sizeof(float);

		for (size_t i = 0; i < mesh.vertex_count; ++i)
		{
			unsigned int r = remap[i];

			for (size_t j = 0; j < max_attributes; ++j)
				if (r != i && (mesh.attribute_protect_mask & (1u << j)) && mesh.vertex_attributes[i * max_attributes + j] != mesh.vertex_attributes[r * max_attributes + j])
					locks[i] |= meshopt_SimplifyVertex_Protect;
		}
	}
	store.put("protect_locks", locks);

	// clusterlod.h:844-848
	std::vector<Cluster> clusters = clusterize(config, mesh, mesh.indices, mesh.index_count);
	for (Cluster& cluster : clusters)
		cluster.bounds = boundsCompute(mesh, cluster.indices, 0.f);

	uint32_t cluster_index_total = 0;
	store.push<uint32_t>("cluster_index_offsets", 0);
	auto recordCluster = [&](const Cluster& c, int depth) {
		store.append<unsigned int>("cluster_indices", c.indices.data(), c.indices.size());
		cluster_index_total += uint32_t(c.indices.size());
		store.push<uint32_t>("cluster_index_offsets", cluster_index_total);
		store.push<int32_t>("cluster_refined", c.refined);
		store.push<uint32_t>("cluster_vertices", uint32_t(c.vertices));
		store.push<int32_t>("cluster_depth", depth);
		store.append<float>("cluster_bounds", c.bounds.center, 3);
		store.push<float>("cluster_bounds", c.bounds.radius);
		store.push<float>("cluster_bounds", c.bounds.error);
	};

	for (const Cluster& c : clusters)
		recordCluster(c, 0);

	std::vector<int> pending(clusters.size());
	for (size_t i = 0; i < clusters.size(); ++i)
		pending[i] = int(i);

	int depth = 0;

	while (pending.size() > 1)
	{
		store.put(levelKey(depth, "pending"), pending);

		std::vector<std::vector<int> > groups = partition(config, mesh, clusters, pending, remap);

		std::vector<uint32_t> group_offsets(1, 0);
		std::vector<int> group_clusters;
		for (const std::vector<int>& g : groups)
		{
			group_clusters.insert(group_clusters.end(), g.begin(), g.end());
			group_offsets.push_back(uint32_t(group_clusters.size()));
		}
		store.put(levelKey(depth, "group_offsets"), group_offsets);
		store.put(levelKey(depth, "group_clusters"), group_clusters);

		pending.clear();

		lockBoundary(locks, groups, clusters, remap, mesh.vertex_lock);
		store.put(levelKey(depth, "locks"), locks);

		std::vector<unsigned char> group_terminal(groups.size());
		std::vector<float> group_bounds;
		std::vector<float> group_error(groups.size());
		std::vector<uint32_t> simp_offsets(1, 0);
		std::vector<unsigned int> simp_indices;
		std::vector<uint32_t> merged_offsets(1, 0);
		std::vector<unsigned int> merged_indices;
		std::vector<int> group_ids(groups.size());

		for (size_t i = 0; i < groups.size(); ++i)
		{
			const std::vector<int>& group = groups[i];

			// runIterationTask, clusterlod.h:699-738
			std::vector<unsigned int> merged;
			for (size_t j = 0; j < group.size(); ++j)
				merged.insert(merged.end(), clusters[group[j]].indices.begin(), clusters[group[j]].indices.end());

			merged_indices.insert(merged_indices.end(), merged.begin(), merged.end());
			merged_offsets.push_back(uint32_t(merged_indices.size()));

			size_t target_size = size_t((merged.size() / 3) * config.simplify_ratio) * 3;
			if (!merged.empty())
				target_size = std::max<size_t>(3, target_size);

			clodBounds bounds = boundsMerge(clusters, group);

			float error = 0.f;
			std::vector<unsigned int> simplified = simplify(config, mesh, merged, locks, target_size, &error);
			group_error[i] = error;

			const bool invalidSimplifiedTopology = !simplified.empty() && (simplified.size() % 3) != 0;
			const bool emptyOrDegenerateSimplified = !merged.empty() && simplified.size() < 3;
			bool terminal = false;
			if (simplified.size() > merged.size() * config.simplify_threshold || invalidSimplifiedTopology || emptyOrDegenerateSimplified)
			{
				terminal = true;
				bounds.error = FLT_MAX;
				simplified.clear();
			}
			else
			{
				bounds.error = std::max(bounds.error * config.simplify_error_merge_previous, error) + error * config.simplify_error_merge_additive;
			}

			group_terminal[i] = terminal;
			group_bounds.insert(group_bounds.end(), bounds.center, bounds.center + 3);
			group_bounds.push_back(bounds.radius);
			group_bounds.push_back(bounds.error);
			simp_indices.insert(simp_indices.end(), simplified.begin(), simplified.end());
			simp_offsets.push_back(uint32_t(simp_indices.size()));

			// clusterlod.h:895-926
			if (terminal)
			{
				group_ids[i] = outputGroupEx(config, mesh, clusters, group, bounds, depth, &recorder, &OutputRecorder::callback, i, 0);
				continue;
			}

			const int refined = outputGroupEx(config, mesh, clusters, group, bounds, depth, &recorder, &OutputRecorder::callback, i, 0);
			group_ids[i] = refined;

			std::vector<Cluster> split = clusterize(config, mesh, simplified.data(), simplified.size());
			if (split.empty())
			{
				clodBounds terminalBounds = bounds;
				terminalBounds.error = FLT_MAX;
				outputGroupEx(config, mesh, clusters, group, terminalBounds, depth, &recorder, &OutputRecorder::callback, i, 0);
				continue;
			}

			for (size_t j = 0; j < group.size(); ++j)
				clusters[group[j]].indices = std::vector<unsigned int>();

			for (Cluster& cluster : split)
			{
				cluster.refined = refined;
				cluster.bounds = bounds;
				recordCluster(cluster, depth + 1);
				clusters.push_back(std::move(cluster));
				pending.push_back(int(clusters.size()) - 1);
			}
		}

		store.put(levelKey(depth, "group_terminal"), group_terminal);
		store.put(levelKey(depth, "group_bounds"), group_bounds);
		store.put(levelKey(depth, "group_error"), group_error);
		store.put(levelKey(depth, "group_ids"), group_ids);
		store.put(levelKey(depth, "simp_offsets"), simp_offsets);
		store.put(levelKey(depth, "simp_indices"), simp_indices);
		store.put(levelKey(depth, "merged_offsets"), merged_offsets);
		store.put(levelKey(depth, "merged_indices"), merged_indices);

		depth++;
	}

	if (pending.size())
	{
		const Cluster& cluster = clusters[pending[0]];
		clodBounds bounds = cluster.bounds;
		bounds.error = FLT_MAX;
		store.put(levelKey(depth, "pending"), pending);
		outputGroupEx(config, mesh, clusters, pending, bounds, depth, &recorder, &OutputRecorder::callback, 0, 0);
	}

	store.push<int32_t>("num_levels", depth);
	h->cluster_count = clusters.size();
	return h;
}

size_t clodref_cluster_count(const clodref_handle* h)
{
	return h->cluster_count;
}

// returns 1 and fills ptr/bytes if the blob exists (an absent blob reads as empty)
int clodref_blob_get(const clodref_handle* h, const char* name, const void** out_ptr, size_t* out_bytes)
{
	std::map<std::string, std::vector<unsigned char> >::const_iterator it = h->store.blobs.find(name);
	if (it == h->store.blobs.end())
	{
		*out_ptr = NULL;
		*out_bytes = 0;
		return 0;
	}
	*out_ptr = it->second.data();
	*out_bytes = it->second.size();
	return 1;
}

void clodref_free(clodref_handle* h)
{
	delete h;
}

// Stage-level entry points onto the reference's file-static clod:: helpers (same TU), for stage parity tests.

// clod::clusterize (clusterlod.h:305-348): returns cluster count; fills per-cluster index counts/vertex counts and the
// concatenated (optimized) cluster index lists (index_count entries).
size_t clodref_clusterize(const clodConfig* config, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    unsigned int* out_cluster_index_counts, unsigned int* out_cluster_vertex_counts, unsigned int* out_indices)
{
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, NULL, 0, NULL, 0, 0, NULL);
	std::vector<clod::Cluster> clusters = clod::clusterize(*config, mesh, indices, index_count);

	size_t offset = 0;
	for (size_t i = 0; i < clusters.size(); ++i)
	{
		out_cluster_index_counts[i] = unsigned(clusters[i].indices.size());
		out_cluster_vertex_counts[i] = unsigned(clusters[i].vertices);
		memcpy(out_indices + offset, clusters[i].indices.data(), clusters[i].indices.size() * sizeof(unsigned int));
		offset += clusters[i].indices.size();
	}
	return clusters.size();
}

// clod::simplify (clusterlod.h:601-659) on one group's merged index list; returns the simplified index count.
size_t clodref_simplify(const clodConfig* config, const unsigned int* indices, size_t index_count, const float* positions, size_t vertex_count, size_t positions_stride,
    const float* attributes, size_t attributes_stride, const float* attribute_weights, size_t attribute_count, const unsigned char* locks, size_t target_count,
    unsigned int* out_indices, float* out_error)
{
	clodMesh mesh = makeMesh(indices, index_count, positions, vertex_count, positions_stride, attributes, attributes_stride, attribute_weights, attribute_count, 0, NULL);
	std::vector<unsigned int> merged(indices, indices + index_count);
	std::vector<unsigned char> lockv(locks, locks + vertex_count);
	float error = 0.f;
	std::vector<unsigned int> lod = clod::simplify(*config, mesh, merged, lockv, target_count, &error);
	memcpy(out_indices, lod.data(), lod.size() * sizeof(unsigned int));
	*out_error = error;
	return lod.size();
}

} // extern "C"
