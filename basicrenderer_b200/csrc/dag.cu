// DAG driver: the loop of clodBuildEx (BasicRenderer/include/ThirdParty/meshoptimizer/clusterlod.h:792-943) with every
// stage running as level-wide CUDA kernels and the output callback stream produced on the host in the reference's order.
//
// Per level: partition -> merged group index lists -> boundary locks -> merged bounds -> batched simplification ->
// (host) terminal test + error rule -> output callbacks -> batched per-group re-clusterization of the simplified lists.
#include "clodb.h"
#include "dag.h"

#include <cfloat>
#include <algorithm>
#include <chrono>

namespace clodb
{

KERNEL k_expand_bounds(const float* __restrict__ bounds4, float error, float* bounds5, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	bounds5[c * 5 + 0] = bounds4[c * 4 + 0];
	bounds5[c * 5 + 1] = bounds4[c * 4 + 1];
	bounds5[c * 5 + 2] = bounds4[c * 4 + 2];
	bounds5[c * 5 + 3] = bounds4[c * 4 + 3];
	bounds5[c * 5 + 4] = error;
}

KERNEL k_group_tri_offsets(const u32* __restrict__ gc_tri_offset, const u32* __restrict__ group_cluster_offset, u32 G, u32* group_tri_offset)
{
	size_t g = GTID;
	if (g > G)
		return;
	group_tri_offset[g] = gc_tri_offset[group_cluster_offset[g]];
}

KERNEL k_assign_cluster_parent(const u32* __restrict__ cluster_segment, const int* __restrict__ segment_refined, const float* __restrict__ segment_bounds5, int* cluster_refined, float* cluster_bounds5, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 s = cluster_segment[c];
	cluster_refined[c] = segment_refined[s];
	for (int k = 0; k < 5; ++k)
		cluster_bounds5[c * 5 + k] = segment_bounds5[s * 5 + k];
}

KERNEL k_copy_segments(const u32* __restrict__ src_tri, const u32* __restrict__ seg_src_offset, const u32* __restrict__ seg_dst_offset, u32 S, u32* dst_tri, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 lo = 0, hi = S;
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (seg_dst_offset[mid] <= u32(t))
			lo = mid;
		else
			hi = mid;
	}
	size_t src = size_t(seg_src_offset[lo]) + (t - seg_dst_offset[lo]);
	dst_tri[t * 3 + 0] = src_tri[src * 3 + 0];
	dst_tri[t * 3 + 1] = src_tri[src * 3 + 1];
	dst_tri[t * 3 + 2] = src_tri[src * 3 + 2];
}

// Host view of one level's cluster tables inside the pinned staging buffer (filled by an asynchronous copy that is
// enqueued right after the level is clusterized, so it overlaps the level's partition / simplification kernels).
struct LevelHost
{
	const u32* tri = nullptr;
	const u32* off = nullptr;
	const u32* vcount = nullptr;
};

static LevelHost download_level_async(const ClusterSet& cs, Workspace& ws, BuildStats& stats, bool with_indices)
{
	size_t tri_bytes = with_indices ? size_t(cs.triangle_count) * 12 : 0, off_bytes = (size_t(cs.cluster_count) + 1) * 4, vc_bytes = size_t(cs.cluster_count) * 4;
	size_t off_at = (tri_bytes + 255) & ~size_t(255), vc_at = (off_at + off_bytes + 255) & ~size_t(255);
	ws.stage.reserve(vc_at + vc_bytes + 256);
	char* base = ws.stage.base;
	if (tri_bytes)
		dev_d2h_async(base, cs.tri, tri_bytes);
	dev_d2h_async(base + off_at, cs.cluster_tri_offset, off_bytes);
	dev_d2h_async(base + vc_at, cs.cluster_vertex_count, vc_bytes);
	stats.d2h_bytes += tri_bytes + off_bytes + vc_bytes;
	LevelHost h;
	h.tri = with_indices ? reinterpret_cast<const u32*>(base) : nullptr;
	h.off = reinterpret_cast<const u32*>(base + off_at);
	h.vcount = reinterpret_cast<const u32*>(base + vc_at);
	return h;
}

static size_t emit_level(const ClusterSet& cs, const LevelHost& host, const std::vector<int>& refined, const std::vector<float>& bounds5, const std::vector<float>& precise4, const Config& config,
    const GroupSet& groups, const std::vector<u32>& group_clusters, const std::vector<float>& group_bounds5, int depth, DagSink& sink, std::vector<int>& group_ids, BuildStats& stats)
{
	// the level's cluster tables were sent to the pinned staging buffer when the level was built
	dev_d2h_async_wait();
	const u32* tri = host.tri;
	const u32* off = host.off;
	const u32* vcount = host.vcount;

	std::vector<DagCluster> out;
	sink.begin_level(cs, depth);
	sink.level_cluster_tri_offset = off;
	group_ids.assign(groups.group_count, -1);
	if (!tri)
	{
		LevelBulk bulk;
		bulk.depth = depth;
		bulk.cluster_count = cs.cluster_count;
		bulk.group_count = groups.group_count;
		bulk.group_cluster_offset = groups.group_cluster_offset_host.data();
		bulk.group_clusters = group_clusters.data();
		bulk.refined = refined.data();
		bulk.bounds5 = bounds5.data();
		bulk.precise4 = precise4.data();
		bulk.use_precise = config.optimize_bounds;
		bulk.group_bounds5 = group_bounds5.data();
		bulk.cluster_tri_offset = off;
		bulk.cluster_vertex_count = vcount;
		if (sink.emit_level_bulk(bulk, group_ids))
		{
			stats.groups += groups.group_count;
			return cs.cluster_count;
		}
	}
	for (u32 g = 0; g < groups.group_count; ++g)
	{
		u32 b = groups.group_cluster_offset_host[g], e = groups.group_cluster_offset_host[g + 1];
		out.resize(e - b);
		for (u32 j = b; j < e; ++j)
		{
			u32 c = group_clusters[j];
			DagCluster& dc = out[j - b];
			dc.refined = refined[c];
			// clusterlod.h:689: precise bounds for simplified clusters when optimize_bounds is set, error is inherited
			bool precise = config.optimize_bounds && refined[c] != -1;
			const float* src = precise ? &precise4[size_t(c) * 4] : &bounds5[size_t(c) * 5];
			dc.bounds[0] = src[0];
			dc.bounds[1] = src[1];
			dc.bounds[2] = src[2];
			dc.bounds[3] = src[3];
			dc.bounds[4] = bounds5[size_t(c) * 5 + 4];
			dc.indices = tri ? tri + size_t(off[c]) * 3 : nullptr;
			dc.index_count = size_t(off[c + 1] - off[c]) * 3;
			dc.vertex_count = vcount[c];
		}
		DagGroup dg;
		dg.depth = depth;
		for (int k = 0; k < 5; ++k)
			dg.simplified[k] = group_bounds5[size_t(g) * 5 + k];
		sink.cluster_ids = &group_clusters[b];
		group_ids[g] = sink.group(dg, out.data(), out.size(), g);
		stats.groups++;
	}
	return cs.cluster_count;
}

// CLODB200_LEVEL_TIMES=1: host wall time per stage and level on stderr (each mark synchronises the build stream, so the
// sum is slower than an untraced build; diagnostics only)
struct StageClock
{
	bool on;
	std::chrono::steady_clock::time_point t0;
	StageClock()
	    : on(getenv("CLODB200_LEVEL_TIMES") != nullptr)
	{
		restart();
	}
	void restart()
	{
		if (on)
		{
			dev_sync();
			t0 = std::chrono::steady_clock::now();
		}
	}
	double lap()
	{
		if (!on)
			return 0;
		dev_sync();
		auto t1 = std::chrono::steady_clock::now();
		double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		t0 = t1;
		return ms;
	}
};

size_t build_dag(const Config& config, const DeviceMesh& mesh, const u32* indices_dev, size_t index_count, Workspace& ws, DagSink& sink, BuildStats& stats)
{
	stats = BuildStats();
	StageClock clock;
	u32 T0 = u32(index_count / 3);
	size_t V = mesh.vertex_count;
	Arena& persist = ws.persist;
	Arena& temp = ws.temp;

	u32* remap = persist.alloc<u32>(V);
	u8* locks = persist.alloc<u8>(V);
	position_remap(mesh.positions, V, remap, temp);
	dev_memset(locks, 0, V);
	protect_bits(mesh.attributes, mesh.attribute_stride, mesh.attribute_protect_mask, remap, V, locks);

	// initial clusterization + precise bounds (clusterlod.h:844-848)
	u32 seg0[2] = {0, T0};
	ClusterSet level = clusterize(indices_dev, T0, seg0, 1, mesh.positions, config, ws);
	LevelHost level_host = download_level_async(level, ws, stats, sink.wants_indices());
	size_t total_clusters = level.cluster_count;

	float* bounds4 = persist.alloc<float>(size_t(level.cluster_count) * 4);
	float* bounds5 = persist.alloc<float>(size_t(level.cluster_count) * 5);
	int* refined_dev = persist.alloc<int>(level.cluster_count);
	cluster_bounds(level.tri, level.cluster_tri_offset, level.cluster_count, mesh.positions, bounds4);
	LAUNCH(k_expand_bounds, level.cluster_count, bounds4, 0.f, bounds5, level.cluster_count);
	dev_memset(refined_dev, 0xff, size_t(level.cluster_count) * sizeof(int));

	std::vector<int> refined_host(level.cluster_count, -1);
	int depth = 0;
	if (clock.on)
		fprintf(stderr, "level -1: remap + first clusterize + bounds %.3f ms (T %u K %u)\n", clock.lap(), T0, level.cluster_count);

	while (level.cluster_count > 1)
	{
		u32 K = level.cluster_count;
		stats.levels++;
		stats.level_triangles.push_back(level.triangle_count);
		stats.level_clusters.push_back(K);

		GroupSet groups = partition_clusters(level.tri, level.cluster_tri_offset, K, refined_dev, bounds5, remap, mesh.positions, V, config, ws);
		u32 G = groups.group_count;
		stats.level_groups.push_back(G);
		stats.refined_splits += groups.refined_splits;
		double t_partition = clock.lap();
		u64 launches0 = g_launches;

		ArenaScope level_scope(temp);
		u32* gtri = temp.alloc<u32>(size_t(level.triangle_count) * 3);
		u32* gc_tri_offset = temp.alloc<u32>(size_t(K) + 1);
		u32* group_tri_offset = temp.alloc<u32>(size_t(G) + 1);
		gather_group_triangles(level.tri, level.cluster_tri_offset, groups.group_clusters, K, gtri, gc_tri_offset, temp);
		LAUNCH(k_group_tri_offsets, size_t(G) + 1, gc_tri_offset, groups.group_cluster_offset, G, group_tri_offset);
		std::vector<u32> group_tri_offset_host = dev_download(group_tri_offset, size_t(G) + 1);

		lock_boundary(gtri, group_tri_offset, G, level.triangle_count, remap, mesh.vertex_lock, V, locks, temp);

		float* group_bounds5_dev = temp.alloc<float>(size_t(G) * 5);
		group_bounds_merge(bounds5, groups.group_cluster_offset, groups.group_clusters, G, group_bounds5_dev);

		double t_locks = clock.lap();
		SimplifyOutput simp = simplify_groups(gtri, group_tri_offset_host.data(), G, mesh, remap, locks, config, ws);
		double t_simplify = clock.lap();
		stats.simplify_passes += g_simplify_stats.passes;
		stats.level_passes.push_back(g_simplify_stats.passes);
		stats.level_sloppy.push_back(g_simplify_stats.sloppy_groups);
		stats.simplify_rounds += g_simplify_stats.rounds;

		std::vector<u32> simp_offset = dev_download(simp.group_tri_offset, size_t(G) + 1);
		std::vector<float> simp_error = dev_download(simp.group_error, G);
		std::vector<float> group_bounds5 = dev_download(group_bounds5_dev, size_t(G) * 5);
		std::vector<u32> group_clusters_host = dev_download(groups.group_clusters, K);

		// terminal test + error rule (clusterlod.h:719-735), same float expressions
		std::vector<unsigned char> terminal(G);
		for (u32 g = 0; g < G; ++g)
		{
			size_t merged = size_t(group_tri_offset_host[g + 1] - group_tri_offset_host[g]) * 3;
			size_t simplified = size_t(simp_offset[g + 1] - simp_offset[g]) * 3;
			bool empty_or_degenerate = merged != 0 && simplified < 3;
			float error = simp_error[g];
			float& berr = group_bounds5[size_t(g) * 5 + 4];
			if (simplified > merged * config.simplify_threshold || empty_or_degenerate)
			{
				terminal[g] = 1;
				berr = FLT_MAX;
			}
			else
			{
				terminal[g] = 0;
				berr = std::max(berr * config.simplify_error_merge_previous, error) + error * config.simplify_error_merge_additive;
			}
		}

		// precise bounds of this level's clusters for output (depth 0 already has them in bounds5)
		std::vector<float> bounds5_host = dev_download(bounds5, size_t(K) * 5);
		std::vector<float> precise4;
		if (depth > 0 && config.optimize_bounds)
		{
			float* p4 = temp.alloc<float>(size_t(K) * 4);
			cluster_bounds(level.tri, level.cluster_tri_offset, K, mesh.positions, p4);
			precise4 = dev_download(p4, size_t(K) * 4);
		}
		else
			precise4.assign(size_t(K) * 4, 0.f);

		std::vector<int> group_ids;
		emit_level(level, level_host, refined_host, bounds5_host, precise4, config, groups, group_clusters_host, group_bounds5, depth, sink, group_ids, stats);
		double t_emit = clock.lap();

		// segments for re-clusterization: simplified lists of non-terminal groups
		std::vector<u32> seg_src, seg_dst(1, 0);
		std::vector<int> seg_refined;
		std::vector<float> seg_bounds5;
		for (u32 g = 0; g < G; ++g)
		{
			u32 count = simp_offset[g + 1] - simp_offset[g];
			if (terminal[g] || count == 0)
				continue;
			seg_src.push_back(simp_offset[g]);
			seg_dst.push_back(seg_dst.back() + count);
			seg_refined.push_back(group_ids[g]);
			seg_bounds5.insert(seg_bounds5.end(), group_bounds5.begin() + size_t(g) * 5, group_bounds5.begin() + size_t(g) * 5 + 5);
			stats.simplified_triangles += count;
		}
		u32 S = u32(seg_src.size());
		depth++;
		if (S == 0)
		{
			level = ClusterSet();
			break;
		}
		u32 T_next = seg_dst.back();
		const u32* next_tri = simp.tri;
		if (T_next != simp.triangle_count)
		{
			// drop terminal groups' triangles
			u32* packed = temp.alloc<u32>(size_t(T_next) * 3);
			u32* d_src = temp.alloc<u32>(S);
			u32* d_dst = temp.alloc<u32>(size_t(S) + 1);
			dev_h2d(d_src, seg_src.data(), size_t(S) * 4);
			dev_h2d(d_dst, seg_dst.data(), (size_t(S) + 1) * 4);
			LAUNCH(k_copy_segments, T_next, simp.tri, d_src, d_dst, S, packed, T_next);
			next_tri = packed;
		}

		ClusterSet next = clusterize(next_tri, T_next, seg_dst.data(), S, mesh.positions, config, ws);
		level_host = download_level_async(next, ws, stats, sink.wants_indices());
		total_clusters += next.cluster_count;

		// clusters inherit the refined id and the bounds of the group they came from (clusterlod.h:919-925)
		bounds5 = persist.alloc<float>(size_t(next.cluster_count) * 5);
		refined_dev = persist.alloc<int>(next.cluster_count);
		{
			int* d_ref = temp.alloc<int>(S);
			float* d_b5 = temp.alloc<float>(size_t(S) * 5);
			dev_h2d(d_ref, seg_refined.data(), size_t(S) * 4);
			dev_h2d(d_b5, seg_bounds5.data(), size_t(S) * 20);
			LAUNCH(k_assign_cluster_parent, next.cluster_count, next.cluster_segment, d_ref, d_b5, refined_dev, bounds5, next.cluster_count);
		}
		refined_host = dev_download(refined_dev, next.cluster_count);
		if (clock.on)
			fprintf(stderr, "level %d: T %u K %u G %u | partition %.3f locks %.3f simplify %.3f (passes %u sloppy %u) emit %.3f clusterize %.3f ms | launches %llu\n", depth - 1, level.triangle_count, K, G,
			    t_partition, t_locks, t_simplify, g_simplify_stats.passes, g_simplify_stats.sloppy_groups, t_emit, clock.lap(), (unsigned long long)(g_launches - launches0));
		level = next;
	}

	if (level.cluster_count == 1)
	{
		// final terminal group (clusterlod.h:931-941)
		stats.levels++;
		stats.level_triangles.push_back(level.triangle_count);
		stats.level_clusters.push_back(1);
		stats.level_groups.push_back(1);
		stats.level_passes.push_back(0);
		stats.level_sloppy.push_back(0);
		GroupSet last;
		last.group_count = 1;
		last.cluster_count = 1;
		last.group_cluster_offset_host = {0u, 1u};
		std::vector<u32> gc = {0u};
		std::vector<float> b5 = dev_download(bounds5, 5);
		std::vector<float> gb5 = b5;
		gb5[4] = FLT_MAX;
		std::vector<float> precise4(4, 0.f);
		if (depth > 0 && config.optimize_bounds)
		{
			ArenaScope s(temp);
			float* p4 = temp.alloc<float>(4);
			cluster_bounds(level.tri, level.cluster_tri_offset, 1, mesh.positions, p4);
			precise4 = dev_download(p4, 4);
		}
		std::vector<int> ids;
		emit_level(level, level_host, refined_host, b5, precise4, config, last, gc, gb5, depth, sink, ids, stats);
	}

	stats.total_clusters = total_clusters;
	return total_clusters;
}

} // namespace clodb
