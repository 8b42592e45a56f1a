// S4: cluster grouping (partitioning) for one DAG level.
//
// Reference: clod::partition (clusterlod.h:350-510) -> meshopt_partitionClusters
// (ThirdParty/meshoptimizer/src/partition.cpp:484-624): per-cluster unique remapped vertices (filterClusterIndices :23-56),
// cluster adjacency weighted by shared vertex count (buildClusterAdjacency :98-204), agglomerative merging of groups by
// score = shared * (1/sqrt(v1) + 1/sqrt(v2)) * (1 + 0.4 * boundsScore) up to target/max sizes (:326-366, :546-596), then
// meshopt_spatialSortRemap of the partitions (spatialorder.cpp:218-251) and the refined-id cap split (clusterlod.h:432-507).
//
// B200 formulation. The reference merges one pair at a time off a heap (serial, and its hottest function). Here the same
// objective is optimised bulk-synchronously: every round each open group picks its best admissible neighbour under the
// same score, locally-dominant (mutually best) pairs merge, and the group graph is contracted with a sort + run-length
// pass. Group sizes follow the same target/max rules, so the result satisfies the same invariants (<= max size, groups
// closed once they reach the target) but is not the same partition bit for bit. Everything downstream of the merge
// (Morton ordering of groups, cluster order inside groups, refined-id cap) follows the reference exactly.
#include "clodb.h"

#include <cfloat>
#include <algorithm>

namespace clodb
{

static const u32 NONE = 0xffffffffu;

// ---------------------------------------------------------------------------------------------- cluster vertex sets
KERNEL k_cluster_unique(const u32* __restrict__ tri, const u32* __restrict__ cluster_tri_offset, const u32* __restrict__ remap, const float* __restrict__ positions, u32 K, u32 stride,
    u32* cv, u32* cv_count, float* center_radius)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 begin = cluster_tri_offset[c] * 3, end = cluster_tri_offset[c + 1] * 3;
	u32 slots[512];
	for (int i = 0; i < 512; ++i)
		slots[i] = NONE;
	u32 count = 0;
	float cx = 0, cy = 0, cz = 0;
	for (u32 j = begin; j < end; ++j)
	{
		u32 v = remap[tri[j]];
		u32 h = (v * 0x9E3779B1u) >> 23;
		bool fresh = false;
		for (;;)
		{
			if (slots[h] == v)
				break;
			if (slots[h] == NONE)
			{
				slots[h] = v;
				fresh = true;
				break;
			}
			h = (h + 1) & 511;
		}
		if (fresh && count < stride)
		{
			cv[size_t(c) * stride + count++] = v;
			cx += positions[size_t(v) * 3 + 0];
			cy += positions[size_t(v) * 3 + 1];
			cz += positions[size_t(v) * 3 + 2];
		}
	}
	cv_count[c] = count;
	if (count)
	{
		cx /= float(count);
		cy /= float(count);
		cz /= float(count);
	}
	float r2 = 0;
	for (u32 j = 0; j < count; ++j)
	{
		const float* p = positions + size_t(cv[size_t(c) * stride + j]) * 3;
		float d2 = (p[0] - cx) * (p[0] - cx) + (p[1] - cy) * (p[1] - cy) + (p[2] - cz) * (p[2] - cz);
		r2 = r2 < d2 ? d2 : r2;
	}
	center_radius[c * 4 + 0] = cx;
	center_radius[c * 4 + 1] = cy;
	center_radius[c * 4 + 2] = cz;
	center_radius[c * 4 + 3] = sqrtf(r2);
}

KERNEL k_emit_vertex_pairs(const u32* __restrict__ cv, const u32* __restrict__ cv_count, const u32* __restrict__ cv_offset, u32 K, u32 stride, u32* pair_vertex, u32* pair_cluster)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 base = cv_offset[c];
	for (u32 j = 0; j < cv_count[c]; ++j)
	{
		pair_vertex[base + j] = cv[size_t(c) * stride + j];
		pair_cluster[base + j] = u32(c);
	}
}

// for every (vertex, cluster) entry: number of other clusters sharing that vertex
KERNEL k_shared_counts(const u32* __restrict__ pair_vertex, u32 P, u32* counts)
{
	size_t p = GTID;
	if (p >= P)
		return;
	u32 v = pair_vertex[p];
	u32 m = 0;
	for (size_t q = p; q-- > 0 && pair_vertex[q] == v;)
		++m;
	for (size_t q = p + 1; q < P && pair_vertex[q] == v; ++q)
		++m;
	counts[p] = m;
}

KERNEL k_emit_cluster_edges(const u32* __restrict__ pair_vertex, const u32* __restrict__ pair_cluster, const u32* __restrict__ edge_offset, u32 P, u64 K, u64* edge_key)
{
	size_t p = GTID;
	if (p >= P)
		return;
	u32 v = pair_vertex[p];
	u64 src = pair_cluster[p];
	u32 out = edge_offset[p];
	size_t s = p;
	while (s > 0 && pair_vertex[s - 1] == v)
		--s;
	for (size_t q = s; q < P && pair_vertex[q] == v; ++q)
		if (q != p)
			edge_key[out++] = src * K + pair_cluster[q];
}

// ------------------------------------------------------------------------------------------------ edge list contraction
KERNEL k_edge_run_heads(const u64* __restrict__ edge_key, const u32* __restrict__ edge_w_in, u32 E, u32* head_flag)
{
	size_t e = GTID;
	if (e >= E)
		return;
	head_flag[e] = (e == 0 || edge_key[e] != edge_key[e - 1]) ? 1u : 0u;
}

KERNEL k_edge_combine(const u64* __restrict__ edge_key, const u32* __restrict__ edge_w_in, const u32* __restrict__ head_scanned, u32 total, u32 E, u64 K, u32* out_src, u32* out_dst, u32* out_w)
{
	size_t e = GTID;
	if (e >= E)
		return;
	u32 pos = head_scanned[e];
	u32 next = e + 1 < E ? head_scanned[e + 1] : total;
	if (next == pos)
		return; // not a run head
	u64 key = edge_key[e];
	u32 w = 0;
	for (size_t q = e; q < E && edge_key[q] == key; ++q)
		w += edge_w_in ? edge_w_in[q] : 1u;
	out_src[pos] = u32(key / K);
	out_dst[pos] = u32(key % K);
	out_w[pos] = w;
}

KERNEL k_edge_src_count(const u32* __restrict__ src, u32 E, u32* counts)
{
	size_t e = GTID;
	if (e >= E)
		return;
	atomicAdd(&counts[src[e]], 1u);
}

// ---------------------------------------------------------------------------------------------------- agglomeration
struct GroupInfo
{
	float cx, cy, cz, radius;
	u32 size;     // clusters in the group (0 = merged away)
	u32 vertices; // running estimate as in partition.cpp:575-577
};

KERNEL k_init_group_info(const float* __restrict__ center_radius, const u32* __restrict__ cv_count, GroupInfo* info, u32* label, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	GroupInfo g;
	g.cx = center_radius[c * 4 + 0];
	g.cy = center_radius[c * 4 + 1];
	g.cz = center_radius[c * 4 + 2];
	g.radius = center_radius[c * 4 + 3];
	g.size = 1;
	g.vertices = cv_count[c];
	info[c] = g;
	label[c] = u32(c);
}

// symmetric variant of boundsScore (partition.cpp:314-324): larger radius over merged radius, operands ordered by id so
// both endpoints of an edge compute identical bits
DEVFN float merge_score(const GroupInfo& lo, const GroupInfo& hi, u32 shared, bool use_bounds)
{
	float score = float(int(shared)) * (1.f / sqrtf(float(int(lo.vertices))) + 1.f / sqrtf(float(int(hi.vertices))));
	if (use_bounds)
	{
		float r1 = lo.radius, r2 = hi.radius;
		float dx = hi.cx - lo.cx, dy = hi.cy - lo.cy, dz = hi.cz - lo.cz;
		float d = sqrtf(dx * dx + dy * dy + dz * dz);
		float mr = d + r1 < r2 ? r2 : (d + r2 < r1 ? r1 : (d + r2 + r1) / 2);
		float rmax = r1 > r2 ? r1 : r2;
		score *= 1.f + 0.4f * (mr > 0 ? rmax / mr : 0.f);
	}
	return score;
}

// best admissible neighbour of every open group; edges ordered by (score desc, min id asc, max id asc)
KERNEL k_pick_neighbor(const GroupInfo* __restrict__ info, const u32* __restrict__ edge_off, const u32* __restrict__ edge_dst, const u32* __restrict__ edge_w, u32 K, u32 target, u32 max_size, int use_bounds, u32* best, float* best_score_out, u32* min_size)
{
	size_t gg = GTID;
	if (gg >= K)
		return;
	u32 g = u32(gg);
	best[g] = NONE;
	const GroupInfo& me = info[g];
	if (me.size == 0 || me.size >= target)
		return;
	float best_score = 0.f;
	u32 best_h = NONE;
	for (u32 e = edge_off[g]; e < edge_off[g + 1]; ++e)
	{
		u32 h = edge_dst[e];
		const GroupInfo& other = info[h];
		if (other.size == 0 || other.size >= target)
			continue;
		if (me.size + other.size > max_size)
			continue;
		float score = g < h ? merge_score(me, other, edge_w[e], use_bounds != 0) : merge_score(other, me, edge_w[e], use_bounds != 0);
		if (!(score > 0.f))
			continue;
		bool better = score > best_score;
		if (!better && score == best_score && best_h != NONE)
		{
			u32 a0 = g < h ? g : h, a1 = g < h ? h : g;
			u32 b0 = g < best_h ? g : best_h, b1 = g < best_h ? best_h : g;
			better = a0 < b0 || (a0 == b0 && a1 < b1);
		}
		if (better)
		{
			best_score = score;
			best_h = h;
		}
	}
	best[g] = best_h;
	best_score_out[g] = best_score;
	if (best_h != NONE)
		atomicMin(min_size, me.size);
}

// mergeBounds (partition.cpp:289-312)
DEVFN void merge_bounds(GroupInfo& target, const GroupInfo& source)
{
	float r1 = target.radius, r2 = source.radius;
	float dx = source.cx - target.cx, dy = source.cy - target.cy, dz = source.cz - target.cz;
	float d = sqrtf(dx * dx + dy * dy + dz * dz);
	if (d + r1 < r2)
	{
		target.cx = source.cx;
		target.cy = source.cy;
		target.cz = source.cz;
		target.radius = source.radius;
		return;
	}
	if (d + r2 > r1)
	{
		float k = d > 0 ? (d + r2 - r1) / (2 * d) : 0.f;
		target.cx += dx * k;
		target.cy += dy * k;
		target.cz += dz * k;
		target.radius = (d + r2 + r1) / 2;
	}
}

// Smallest-first agglomeration, bulk-synchronous: groups within 2x of the smallest open size are "movers" this round.
// A mover merges with its pick when the pick is mutual (both movers; the lower id survives), or when the pick is a
// larger, non-moving group that selected it as its best proposer (one proposer per group per round).
KERNEL k_propose(const GroupInfo* __restrict__ info, const u32* __restrict__ best, const float* __restrict__ best_score, u32 K, const u32* __restrict__ min_size, u64* claim)
{
	size_t gg = GTID;
	if (gg >= K)
		return;
	u32 g = u32(gg);
	u32 h = best[g];
	if (h == NONE)
		return;
	u32 limit = *min_size * 2;
	if (info[g].size > limit || info[h].size <= limit)
		return; // only movers propose, and only to non-movers
	u64 key = (u64(__float_as_uint(best_score[g])) << 32) | u64(~g);
	atomicMax(reinterpret_cast<unsigned long long*>(&claim[h]), (unsigned long long)key);
}

KERNEL k_merge_pairs(GroupInfo* info, const u32* __restrict__ best, const u32* __restrict__ edge_off, const u32* __restrict__ edge_dst, const u32* __restrict__ edge_w, u32 K,
    const u32* __restrict__ min_size, const u64* __restrict__ claim, u32* parent, u32* merge_count)
{
	size_t gg = GTID;
	if (gg >= K)
		return;
	u32 g = u32(gg);
	u32 h = best[g];
	if (h == NONE)
		return;
	u32 limit = *min_size * 2;
	if (info[g].size > limit)
		return;
	u32 dst, src;
	if (info[h].size <= limit)
	{
		// mover-mover: mutual picks only, handled once by the lower id
		if (best[h] != g || h < g)
			return;
		dst = g;
		src = h;
	}
	else
	{
		if (u32(~u32(claim[h])) != g)
			return;
		dst = h;
		src = g;
	}
	u32 shared = 0;
	for (u32 e = edge_off[dst]; e < edge_off[dst + 1]; ++e)
		if (edge_dst[e] == src)
			shared = edge_w[e];
	GroupInfo a = info[dst], b = info[src];
	a.size += b.size;
	a.vertices += b.vertices;
	a.vertices = a.vertices > shared ? a.vertices - shared : 1;
	merge_bounds(a, b);
	b.size = 0;
	b.vertices = 0;
	b.radius = 0;
	info[dst] = a;
	info[src] = b;
	parent[src] = dst;
	atomicAdd(merge_count, 1u);
}

KERNEL k_relabel_clusters(u32* label, const u32* __restrict__ parent, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	label[c] = parent[label[c]];
}

// ---- group-graph contraction by hashing (one merge round) ------------------------------------------------------------
// Relabels every edge to the merged groups, drops self loops and combines parallel edges by inserting (src, dst) into an
// open-addressed table that accumulates the shared-vertex weight. The first inserter of a key also counts the edge for
// the CSR of its source. Rows of the rebuilt CSR are unordered; every consumer (k_pick_neighbor, k_merge_pairs) is
// order independent, so the result is deterministic.
DEVFN u32 hash_edge(u64 key)
{
	key ^= key >> 33;
	key *= 0xff51afd7ed558ccdull;
	key ^= key >> 33;
	key *= 0xc4ceb9fe1a85ec53ull;
	key ^= key >> 33;
	return u32(key);
}

KERNEL k_contract_insert(const u32* __restrict__ src, const u32* __restrict__ dst, const u32* __restrict__ w, const u32* __restrict__ parent, const u32* __restrict__ edge_count, u64 K, u64* table_key, u32* table_w, u32 mask,
    u32* row_count, u32* new_edge_count)
{
	size_t e = GTID;
	if (e >= *edge_count)
		return;
	u64 s = parent[src[e]], d = parent[dst[e]];
	if (s == d)
		return;
	u64 key = s * K + d;
	u32 slot = hash_edge(key) & mask;
	for (;;)
	{
		u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(&table_key[slot]), ~0ull, (unsigned long long)key);
		if (old == ~0ull)
		{
			atomicAdd(&row_count[s], 1u);
			atomicAdd(new_edge_count, 1u);
			break;
		}
		if (old == key)
			break;
		slot = (slot + 1) & mask;
	}
	atomicAdd(&table_w[slot], w[e]);
}

KERNEL k_contract_fill(const u64* __restrict__ table_key, const u32* __restrict__ table_w, u32 table_size, u64 K, const u32* __restrict__ row_offset, u32* row_cursor, u32* src_out, u32* dst_out, u32* w_out)
{
	size_t slot = GTID;
	if (slot >= table_size)
		return;
	u64 key = table_key[slot];
	if (key == ~0ull)
		return;
	u32 s = u32(key / K), d = u32(key % K);
	u32 pos = row_offset[s] + atomicAdd(&row_cursor[s], 1u);
	src_out[pos] = s;
	dst_out[pos] = d;
	w_out[pos] = table_w[slot];
}

KERNEL k_relabel_edges(const u32* __restrict__ src, const u32* __restrict__ dst, const u32* __restrict__ parent, u32 E, u64 K, u64* edge_key, u32* keep)
{
	size_t e = GTID;
	if (e >= E)
		return;
	u64 s = parent[src[e]], d = parent[dst[e]];
	edge_key[e] = s * K + d;
	keep[e] = s != d ? 1u : 0u;
}

KERNEL k_compact_edges(const u64* __restrict__ edge_key, const u32* __restrict__ w, const u32* __restrict__ keep_scanned, u32 total, u32 E, u64* key_out, u32* w_out)
{
	size_t e = GTID;
	if (e >= E)
		return;
	u32 pos = keep_scanned[e];
	u32 next = e + 1 < E ? keep_scanned[e + 1] : total;
	if (next == pos)
		return;
	key_out[pos] = edge_key[e];
	w_out[pos] = w[e];
}

// spatial merge of the leftover small groups along the Morton order of their centres (replaces mergeSpatial's kd-tree
// leaves, partition.cpp:368-480): each still-open group looks at its nearest open neighbours in that order
KERNEL k_leftover_pick(const GroupInfo* __restrict__ info, const u32* __restrict__ order, u32 n, u32 target, u32 max_size, u32* best)
{
	size_t p = GTID;
	if (p >= n)
		return;
	u32 g = order[p];
	best[g] = NONE;
	const GroupInfo& me = info[g];
	if (me.size == 0 || me.size >= target)
		return;
	float best_score = -1.f;
	u32 best_h = NONE;
	for (int d = -4; d <= 4; ++d)
	{
		if (d == 0)
			continue;
		long q = long(p) + d;
		if (q < 0 || q >= long(n))
			continue;
		u32 h = order[q];
		const GroupInfo& other = info[h];
		if (other.size == 0 || me.size + other.size > max_size)
			continue;
		float score = g < h ? merge_score(me, other, 1, true) : merge_score(other, me, 1, true);
		if (score > best_score || (score == best_score && h < best_h))
		{
			best_score = score;
			best_h = h;
		}
	}
	best[g] = best_h;
}

KERNEL k_leftover_claim(const u32* __restrict__ best, u32 K, u32* claim)
{
	size_t g = GTID;
	if (g >= K)
		return;
	if (best[g] != NONE)
		atomicMin(&claim[best[g]], u32(g));
}

// src g merges into dst = best[g] when g won the claim on dst and dst is not itself moving (mutual picks: higher id moves)
KERNEL k_leftover_merge(GroupInfo* info, const u32* __restrict__ best, const u32* __restrict__ claim, u32 K, u32* parent, u32* merge_count)
{
	size_t gg = GTID;
	if (gg >= K)
		return;
	u32 g = u32(gg);
	u32 h = best[g];
	if (h == NONE || claim[h] != g)
		return;
	if (!(best[h] == NONE || (best[h] == g && g > h)))
		return;
	GroupInfo a = info[h], b = info[g];
	a.size += b.size;
	a.vertices += b.vertices;
	merge_bounds(a, b);
	b.size = 0;
	b.vertices = 0;
	b.radius = 0;
	info[h] = a;
	info[g] = b;
	parent[g] = h;
	atomicAdd(merge_count, 1u);
}

KERNEL k_stage_centres(const GroupInfo* __restrict__ info, const u32* __restrict__ root_rank, u32 K, float* centres5, u32* ident, u32* root_ids)
{
	size_t g = GTID;
	if (g >= K)
		return;
	if (info[g].size == 0)
		return;
	u32 r = root_rank[g];
	centres5[r * 5 + 0] = info[g].cx;
	centres5[r * 5 + 1] = info[g].cy;
	centres5[r * 5 + 2] = info[g].cz;
	centres5[r * 5 + 3] = info[g].radius;
	centres5[r * 5 + 4] = 0.f;
	ident[r] = r;
	root_ids[r] = u32(g);
}

KERNEL k_gather_ids(const u32* __restrict__ src, const u32* __restrict__ index, u32* dst, u32 n)
{
	size_t i = GTID;
	if (i >= n)
		return;
	dst[i] = src[index[i]];
}

// ------------------------------------------------------------------------------------------- group numbering + order
KERNEL k_root_flags(const GroupInfo* __restrict__ info, u32* flags, u32 K)
{
	size_t g = GTID;
	if (g >= K)
		return;
	flags[g] = info[g].size ? 1u : 0u;
}

KERNEL k_cluster_part(const u32* __restrict__ label, const u32* __restrict__ root_rank, u32* part, u32* part_last_cluster, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 p = root_rank[label[c]];
	part[c] = p;
	atomicMax(&part_last_cluster[p], u32(c)); // "last cluster center" representative point (clusterlod.h:401-405)
}

DEVFN u32 order_key_f(float f)
{
	u32 u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

DEVFN float order_key_inv(u32 k)
{
	u32 u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
	return __uint_as_float(u);
}

KERNEL k_part_points_minmax(const u32* __restrict__ part_last_cluster, const float* __restrict__ cluster_bounds5, u32 G, u32* minmax)
{
	size_t p = GTID;
	if (p >= G)
		return;
	const float* c = cluster_bounds5 + size_t(part_last_cluster[p]) * 5;
	for (int k = 0; k < 3; ++k)
	{
		atomicMin(&minmax[k], order_key_f(c[k]));
		atomicMax(&minmax[3 + k], order_key_f(c[k]));
	}
}

DEVFN u64 part1by2(u64 x)
{
	x &= 0x000fffffull;
	x = (x ^ (x << 32)) & 0x000f00000000ffffull;
	x = (x ^ (x << 16)) & 0x000f0000ff0000ffull;
	x = (x ^ (x << 8)) & 0x000f00f00f00f00full;
	x = (x ^ (x << 4)) & 0x00c30c30c30c30c3ull;
	x = (x ^ (x << 2)) & 0x0249249249249249ull;
	return x;
}

// computeOrder(morton = true), spatialorder.cpp:25-72
KERNEL k_part_morton(const u32* __restrict__ part_last_cluster, const float* __restrict__ cluster_bounds5, const u32* __restrict__ minmax, u32 G, u64* keys, u32* vals)
{
	size_t p = GTID;
	if (p >= G)
		return;
	float minv[3], maxv[3];
	for (int k = 0; k < 3; ++k)
	{
		minv[k] = order_key_inv(minmax[k]);
		maxv[k] = order_key_inv(minmax[3 + k]);
	}
	float extent = 0.f;
	extent = (maxv[0] - minv[0]) < extent ? extent : (maxv[0] - minv[0]);
	extent = (maxv[1] - minv[1]) < extent ? extent : (maxv[1] - minv[1]);
	extent = (maxv[2] - minv[2]) < extent ? extent : (maxv[2] - minv[2]);
	float scale = extent == 0 ? 0.f : 65535.f / extent;
	const float* v = cluster_bounds5 + size_t(part_last_cluster[p]) * 5;
	int x = int((v[0] - minv[0]) * scale + 0.5f);
	int y = int((v[1] - minv[1]) * scale + 0.5f);
	int z = int((v[2] - minv[2]) * scale + 0.5f);
	keys[p] = part1by2(u64(x)) | (part1by2(u64(y)) << 1) | (part1by2(u64(z)) << 2);
	vals[p] = u32(p);
}

KERNEL k_invert_order(const u32* __restrict__ sorted_part, u32* part_remap, u32 G)
{
	size_t i = GTID;
	if (i >= G)
		return;
	part_remap[sorted_part[i]] = u32(i);
}

KERNEL k_apply_part_remap(u32* part, const u32* __restrict__ part_remap, u32* cluster_ids, u32* counts, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 p = part_remap ? part_remap[part[c]] : part[c];
	part[c] = p;
	cluster_ids[c] = u32(c);
	atomicAdd(&counts[p], 1u);
}

// refined-id cap (clusterlod.h:432-507): bucket a group's clusters by refined id in first-seen order; if there are more than
// `cap` buckets, emit the buckets back to back, starting a new group every `cap` buckets
KERNEL k_refined_cap(const u32* __restrict__ group_offset, u32* group_clusters, const int* __restrict__ cluster_refined, u32 G, u32 cap, u32* scratch_clusters, u32* split_marks, u32* extra_groups)
{
	size_t g = GTID;
	if (g >= G)
		return;
	u32 begin = group_offset[g], end = group_offset[g + 1];
	u32 n = end - begin;
	// first-seen refined keys; groups hold at most a few hundred clusters
	int keys[64];
	u32 nkeys = 0;
	bool overflow = false;
	for (u32 j = 0; j < n && !overflow; ++j)
	{
		int r = cluster_refined[group_clusters[begin + j]];
		bool seen = false;
		for (u32 k = 0; k < nkeys; ++k)
			if (keys[k] == r)
			{
				seen = true;
				break;
			}
		if (!seen)
		{
			if (nkeys == 64)
				overflow = true;
			else
				keys[nkeys++] = r;
		}
	}
	if (!overflow && nkeys <= cap)
		return;

	// slow path: bucket order by first occurrence, clusters keep their relative order inside a bucket
	u32 out = begin;
	u32 buckets_in_current = 0;
	u32 extra = 0;
	for (u32 j = 0; j < n; ++j)
		scratch_clusters[begin + j] = group_clusters[begin + j];
	for (u32 j = 0; j < n; ++j)
	{
		u32 cj = scratch_clusters[begin + j];
		if (cj == NONE)
			continue;
		int r = cluster_refined[cj];
		if (buckets_in_current >= cap)
		{
			split_marks[out] = 1; // a new group starts at this slot
			extra++;
			buckets_in_current = 0;
		}
		for (u32 k = j; k < n; ++k)
		{
			u32 ck = scratch_clusters[begin + k];
			if (ck != NONE && cluster_refined[ck] == r)
			{
				group_clusters[out++] = ck;
				scratch_clusters[begin + k] = NONE;
			}
		}
		buckets_in_current++;
	}
	if (extra)
		atomicAdd(extra_groups, extra);
}

KERNEL k_group_start_flags(const u32* __restrict__ group_offset, u32 G, u32* split_marks)
{
	size_t g = GTID;
	if (g >= G)
		return;
	if (group_offset[g] < group_offset[g + 1])
		split_marks[group_offset[g]] = 1;
}

KERNEL k_emit_group_offsets(const u32* __restrict__ split_marks, const u32* __restrict__ marks_scanned, u32 K, u32 total, u32* new_offsets)
{
	size_t j = GTID;
	if (j > K)
		return;
	if (j == K)
	{
		new_offsets[total] = K;
		return;
	}
	if (split_marks[j])
		new_offsets[marks_scanned[j]] = u32(j);
}

// ---------------------------------------------------------------------------------------------------------------------
static bool refined_fits_single_group(const std::vector<int>& refined, u32 cap)
{
	if (cap == 0)
		return true;
	std::vector<int> seen;
	for (int r : refined)
	{
		if (std::find(seen.begin(), seen.end(), r) == seen.end())
		{
			seen.push_back(r);
			if (seen.size() > cap)
				return false;
		}
	}
	return true;
}

GroupSet partition_clusters(const u32* tri, const u32* cluster_tri_offset, u32 K, const int* cluster_refined, const float* cluster_bounds5, const u32* remap, const float* positions, size_t vertex_count, const Config& config, Workspace& ws)
{
	GroupSet out;
	out.cluster_count = K;
	if (K == 0)
		return out;
	Arena& temp = ws.temp;
	out.group_clusters = ws.persist.alloc<u32>(K);
	out.group_cluster_offset = ws.persist.alloc<u32>(size_t(K) + 1);

	ArenaScope scope(temp);
	u32* scalars = temp.alloc<u32>(16);
	dev_memset(scalars, 0, 16 * sizeof(u32));

	u32 cap = config.partition_max_refined_groups;
	u32 target = config.partition_size;
	u32 max_size = target + target / 3;

	// clusterlod.h:352-385: small pending sets become a single group outright
	if (K <= target)
	{
		std::vector<int> refined = dev_download(cluster_refined, K);
		if (refined_fits_single_group(refined, cap))
		{
			iota(out.group_clusters, K);
			u32 offs[2] = {0, K};
			dev_h2d(out.group_cluster_offset, offs, sizeof(offs));
			out.group_count = 1;
			out.group_cluster_offset_host.assign(offs, offs + 2);
			return out;
		}
	}

	u32 stride = config.max_vertices;
	u32* cv = temp.alloc<u32>(size_t(K) * stride);
	u32* cv_count = temp.alloc<u32>(size_t(K) + 1);
	u32* cv_offset = temp.alloc<u32>(size_t(K) + 1);
	float* center_radius = temp.alloc<float>(size_t(K) * 4);
	LAUNCH(k_cluster_unique, K, tri, cluster_tri_offset, remap, positions, K, stride, cv, cv_count, center_radius);
	exclusive_scan_u32(cv_count, cv_offset, K, scalars, temp);
	u32 P = dev_read(scalars);

	GroupInfo* info = temp.alloc<GroupInfo>(K);
	u32* label = temp.alloc<u32>(K);
	u32* parent = temp.alloc<u32>(K);
	u32* best = temp.alloc<u32>(K);
	float* best_score = temp.alloc<float>(K);
	u64* claim = temp.alloc<u64>(K);
	LAUNCH(k_init_group_info, K, center_radius, cv_count, info, label, K);

	// ---- cluster adjacency: sort (vertex, cluster), expand runs into directed cluster pairs, sort + run-length
	u32 E = 0;
	u32* e_src = nullptr;
	u32* e_dst = nullptr;
	u32* e_w = nullptr;
	u32* e_off = temp.alloc<u32>(size_t(K) + 1);
	u64* e_key = nullptr;
	u64* e_key_tmp = nullptr;
	u32* e_val = nullptr;
	u32* e_val_tmp = nullptr;
	u32* e_flag = nullptr;
	int key_bits = bits_for(u64(K) * u64(K));
	{
		u32* pair_vertex = temp.alloc<u32>(P);
		u32* pair_cluster = temp.alloc<u32>(P);
		u32* pair_tmp_a = temp.alloc<u32>(P);
		u32* pair_tmp_b = temp.alloc<u32>(P);
		LAUNCH(k_emit_vertex_pairs, K, cv, cv_count, cv_offset, K, stride, pair_vertex, pair_cluster);
		radix_sort_pairs<u32>(pair_vertex, pair_tmp_a, pair_cluster, pair_tmp_b, P, 0, bits_for(vertex_count), temp);
		u32* counts = pair_tmp_a;
		LAUNCH(k_shared_counts, P, pair_vertex, P, counts);
		exclusive_scan_u32(counts, counts, P, scalars, temp);
		u32 E0 = dev_read(scalars);
		size_t ecap = std::max<u32>(E0, 1);
		e_key = temp.alloc<u64>(ecap);
		e_key_tmp = temp.alloc<u64>(ecap);
		e_val = temp.alloc<u32>(ecap);
		e_val_tmp = temp.alloc<u32>(ecap);
		e_flag = temp.alloc<u32>(ecap + 1);
		e_src = temp.alloc<u32>(ecap);
		e_dst = temp.alloc<u32>(ecap);
		e_w = temp.alloc<u32>(ecap);
		LAUNCH(k_emit_cluster_edges, P, pair_vertex, pair_cluster, counts, P, u64(K), e_key);
		radix_sort_pairs<u64>(e_key, e_key_tmp, nullptr, nullptr, E0, 0, key_bits, temp);
		LAUNCH(k_edge_run_heads, E0, e_key, nullptr, E0, e_flag);
		exclusive_scan_u32(e_flag, e_flag, E0, scalars, temp);
		E = dev_read(scalars);
		LAUNCH(k_edge_combine, E0, e_key, nullptr, e_flag, E, E0, u64(K), e_src, e_dst, e_w);
	}
	auto rebuild_offsets = [&]() {
		dev_memset(e_off, 0, (size_t(K) + 1) * sizeof(u32));
		LAUNCH(k_edge_src_count, E, e_src, E, e_off);
		exclusive_scan_u32(e_off, e_off, size_t(K) + 1, nullptr, temp);
	};
	rebuild_offsets();

	// ---- merge rounds
	// scalars: [1] merges this round, [3] smallest open size, [8] current edge count, [9] next edge count
	u32 rounds = 0;
	if (E > 0)
	{
		size_t table_cap = 1;
		while (table_cap < size_t(E) * 2)
			table_cap <<= 1;
		u64* table_key = temp.alloc<u64>(table_cap);
		u32* table_w = temp.alloc<u32>(table_cap);
		u32* e_src_alt = temp.alloc<u32>(E);
		u32* e_dst_alt = temp.alloc<u32>(E);
		u32* e_w_alt = temp.alloc<u32>(E);
		u32* row_cursor = temp.alloc<u32>(K);
		dev_h2d(scalars + 8, &E, sizeof(u32));
		u32 E_cur = E;
		for (;;)
		{
			rounds++;
			dev_memset(scalars + 3, 0xff, sizeof(u32));
			LAUNCH(k_pick_neighbor, K, info, e_off, e_dst, e_w, K, target, max_size, config.partition_spatial ? 1 : 0, best, best_score, scalars + 3);
			iota(parent, K);
			dev_memset(scalars + 1, 0, sizeof(u32));
			dev_memset(claim, 0, size_t(K) * sizeof(u64));
			LAUNCH(k_propose, K, info, best, best_score, K, scalars + 3, claim);
			LAUNCH(k_merge_pairs, K, info, best, e_off, e_dst, e_w, K, scalars + 3, claim, parent, scalars + 1);
			u32 merged = dev_read(scalars + 1);
			if (merged == 0 || rounds > 4096)
				break;
			LAUNCH(k_relabel_clusters, K, label, parent, K);
			// contract the group graph
			size_t cap = 1;
			while (cap < size_t(E_cur) * 2)
				cap <<= 1;
			dev_memset(table_key, 0xff, cap * sizeof(u64));
			dev_memset(table_w, 0, cap * sizeof(u32));
			dev_memset(e_off, 0, (size_t(K) + 1) * sizeof(u32));
			dev_memset(row_cursor, 0, size_t(K) * sizeof(u32));
			dev_memset(scalars + 9, 0, sizeof(u32));
			LAUNCH(k_contract_insert, E_cur, e_src, e_dst, e_w, parent, scalars + 8, u64(K), table_key, table_w, u32(cap - 1), e_off, scalars + 9);
			exclusive_scan_u32(e_off, e_off, size_t(K) + 1, nullptr, temp);
			LAUNCH(k_contract_fill, cap, table_key, table_w, u32(cap), u64(K), e_off, row_cursor, e_src_alt, e_dst_alt, e_w_alt);
			dev_d2d(scalars + 8, scalars + 9, sizeof(u32));
			std::swap(e_src, e_src_alt);
			std::swap(e_dst, e_dst_alt);
			std::swap(e_w, e_w_alt);
			// every merge removes at least the two directed edges between the merged pair
			E_cur = E_cur > 2 * merged ? E_cur - 2 * merged : 0;
			if (E_cur == 0)
				break;
		}
	}

	// ---- leftovers: spatially merge open groups (only when positions are used, partition.cpp:599-611)
	if (config.partition_spatial)
	{
		u32* flags = temp.alloc<u32>(size_t(K) + 1);
		u64* mkeys = temp.alloc<u64>(K);
		u64* mkeys_tmp = temp.alloc<u64>(K);
		u32* mvals = temp.alloc<u32>(K);
		u32* mvals_tmp = temp.alloc<u32>(K);
		for (int iter = 0; iter < 8; ++iter)
		{
			ArenaScope iter_scope(temp);
			// Morton order of all live group centres
			LAUNCH(k_root_flags, K, info, flags, K);
			exclusive_scan_u32(flags, flags, K, scalars, temp);
			u32 live = dev_read(scalars);
			if (live <= 1)
				break;
			u32* minmax = scalars + 4;
			u32 init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
			dev_h2d(minmax, init, sizeof(init));
			float* centres5 = temp.alloc<float>(size_t(live) * 5);
			u32* ident = temp.alloc<u32>(live);
			u32* root_ids = temp.alloc<u32>(live);
			u32* order = temp.alloc<u32>(live);
			LAUNCH(k_stage_centres, K, info, flags, K, centres5, ident, root_ids);
			LAUNCH(k_part_points_minmax, live, ident, centres5, live, minmax);
			LAUNCH(k_part_morton, live, ident, centres5, minmax, live, mkeys, mvals);
			radix_sort_pairs<u64>(mkeys, mkeys_tmp, mvals, mvals_tmp, live, 0, 50, temp);
			LAUNCH(k_gather_ids, live, root_ids, mvals, order, live);
			fill(best, NONE, K);
			LAUNCH(k_leftover_pick, live, info, order, live, target, max_size, best);
			iota(parent, K);
			dev_memset(scalars + 1, 0, sizeof(u32));
			u32* lclaim = temp.alloc<u32>(K);
			dev_memset(lclaim, 0xff, size_t(K) * 4);
			LAUNCH(k_leftover_claim, K, best, K, lclaim);
			LAUNCH(k_leftover_merge, K, info, best, lclaim, K, parent, scalars + 1);
			u32 merged = dev_read(scalars + 1);
			if (merged == 0)
				break;
			LAUNCH(k_relabel_clusters, K, label, parent, K);
		}
	}

	// ---- number the groups, order them spatially, list their clusters
	u32* root_rank = temp.alloc<u32>(size_t(K) + 1);
	LAUNCH(k_root_flags, K, info, root_rank, K);
	exclusive_scan_u32(root_rank, root_rank, K, scalars, temp);
	u32 G = dev_read(scalars);
	u32* part = temp.alloc<u32>(K);
	u32* part_last = temp.alloc<u32>(G);
	dev_memset(part_last, 0, size_t(G) * 4);
	LAUNCH(k_cluster_part, K, label, root_rank, part, part_last, K);

	u32* part_remap = nullptr;
	if (config.partition_sort)
	{
		u32* minmax = scalars + 4;
		u32 init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
		dev_h2d(minmax, init, sizeof(init));
		u64* keys = temp.alloc<u64>(G);
		u64* keys_tmp = temp.alloc<u64>(G);
		u32* vals = temp.alloc<u32>(G);
		u32* vals_tmp = temp.alloc<u32>(G);
		part_remap = temp.alloc<u32>(G);
		LAUNCH(k_part_points_minmax, G, part_last, cluster_bounds5, G, minmax);
		LAUNCH(k_part_morton, G, part_last, cluster_bounds5, minmax, G, keys, vals);
		radix_sort_pairs<u64>(keys, keys_tmp, vals, vals_tmp, G, 0, 50, temp);
		LAUNCH(k_invert_order, G, vals, part_remap, G);
	}

	u32* group_offset = temp.alloc<u32>(size_t(G) + 1);
	u32* part_tmp = temp.alloc<u32>(K);
	u32* ids_tmp = temp.alloc<u32>(K);
	dev_memset(group_offset, 0, (size_t(G) + 1) * 4);
	LAUNCH(k_apply_part_remap, K, part, part_remap, out.group_clusters, group_offset, K);
	exclusive_scan_u32(group_offset, group_offset, size_t(G) + 1, nullptr, temp);
	radix_sort_pairs<u32>(part, part_tmp, out.group_clusters, ids_tmp, K, 0, bits_for(G > 1 ? G - 1 : 1), temp);

	// ---- refined-id cap
	u32* marks = temp.alloc<u32>(size_t(K) + 1);
	dev_memset(marks, 0, (size_t(K) + 1) * 4);
	if (cap > 0)
	{
		u32* scratch_clusters = temp.alloc<u32>(K);
		LAUNCH(k_refined_cap, G, group_offset, out.group_clusters, cluster_refined, G, cap, scratch_clusters, marks, scalars + 2);
	}
	LAUNCH(k_group_start_flags, G, group_offset, G, marks);
	u32* marks_scanned = temp.alloc<u32>(size_t(K) + 1);
	exclusive_scan_u32(marks, marks_scanned, K, scalars, temp);
	u32 G_final = dev_read(scalars);
	LAUNCH(k_emit_group_offsets, size_t(K) + 1, marks, marks_scanned, K, G_final, out.group_cluster_offset);

	out.group_count = G_final;
	out.group_cluster_offset_host = dev_download(out.group_cluster_offset, size_t(G_final) + 1);
	out.merge_rounds = rounds;
	return out;
}

} // namespace clodb
