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
// pass (first level) and then, round by round, by hashing inside one persistent kernel. Group sizes follow the same target/max rules, so the result satisfies the same invariants (<= max size, groups
// closed once they reach the target) but is not the same partition bit for bit. Everything downstream of the merge
// (Morton ordering of groups, cluster order inside groups, refined-id cap) follows the reference exactly.
#include "clodb.h"

#include <cfloat>
#include <algorithm>
#ifndef CLODB_EMU
#include <cooperative_groups.h>
#endif

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

#ifndef CLODB_EMU
// Warp-cooperative k_cluster_unique (same results): one warp per cluster. The remapped corner ids go through a shared-memory
// hash table that keeps the smallest corner position per vertex, so the first-occurrence order of the serial loop is
// recovered with a ballot; the centre is summed in that order (three lanes, one coordinate each) to keep its bits.
static const int CU_WARPS = 4;
static __global__ void __launch_bounds__(CU_WARPS * 32) k_cluster_unique_warp(const u32* __restrict__ tri, const u32* __restrict__ cluster_tri_offset, const u32* __restrict__ remap, const float* __restrict__ positions, u32 K, u32 stride,
    u32* cv, u32* cv_count, float* center_radius)
{
	__shared__ u32 s_key[CU_WARPS][512];
	__shared__ u32 s_pos[CU_WARPS][512];
	__shared__ float s_xyz[CU_WARPS][3][128];
	__shared__ float s_centre[CU_WARPS][3];
	const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 c = blockIdx.x * CU_WARPS + w;
	if (c >= K)
		return;
	u32* key = s_key[w];
	u32* pos = s_pos[w];
	for (u32 i = lane; i < 512; i += 32)
		key[i] = NONE, pos[i] = NONE;
	__syncwarp();
	const u32 begin = cluster_tri_offset[c] * 3, n = cluster_tri_offset[c + 1] * 3 - begin;
	const u32 cap = stride < 128 ? stride : 128;
	for (u32 j = lane; j < n; j += 32)
	{
		u32 v = remap[tri[begin + j]];
		u32 h = (v * 0x9E3779B1u) >> 23;
		for (;;)
		{
			u32 old = atomicCAS(&key[h], NONE, v);
			if (old == NONE || old == v)
				break;
			h = (h + 1) & 511;
		}
		atomicMin(&pos[h], j);
	}
	__syncwarp();
	u32 count = 0;
	for (u32 base = 0; base < n; base += 32)
	{
		u32 j = base + lane;
		bool first = false;
		u32 v = 0;
		if (j < n)
		{
			v = remap[tri[begin + j]];
			u32 h = (v * 0x9E3779B1u) >> 23;
			while (key[h] != v)
				h = (h + 1) & 511;
			first = pos[h] == j;
		}
		unsigned mask = __ballot_sync(0xffffffffu, first);
		u32 rank = count + __popc(mask & ((1u << lane) - 1));
		if (first && rank < cap)
		{
			cv[size_t(c) * stride + rank] = v;
			s_xyz[w][0][rank] = positions[size_t(v) * 3 + 0];
			s_xyz[w][1][rank] = positions[size_t(v) * 3 + 1];
			s_xyz[w][2][rank] = positions[size_t(v) * 3 + 2];
		}
		count += __popc(mask);
	}
	count = count < cap ? count : cap;
	__syncwarp();
	if (lane < 3)
	{
		float sum = 0;
		for (u32 j = 0; j < count; ++j)
			sum += s_xyz[w][lane][j];
		if (count)
			sum /= float(count);
		s_centre[w][lane] = sum;
	}
	__syncwarp();
	float cx = s_centre[w][0], cy = s_centre[w][1], cz = s_centre[w][2];
	float r2 = 0;
	for (u32 j = lane; j < count; j += 32)
	{
		float px = s_xyz[w][0][j], py = s_xyz[w][1][j], pz = s_xyz[w][2][j];
		float d2 = (px - cx) * (px - cx) + (py - cy) * (py - cy) + (pz - cz) * (pz - cz);
		r2 = r2 < d2 ? d2 : r2;
	}
	for (int d = 16; d >= 1; d >>= 1)
	{
		float o = __shfl_xor_sync(0xffffffffu, r2, d);
		r2 = r2 < o ? o : r2;
	}
	if (lane == 0)
	{
		cv_count[c] = count;
		center_radius[c * 4 + 0] = cx;
		center_radius[c * 4 + 1] = cy;
		center_radius[c * 4 + 2] = cz;
		center_radius[c * 4 + 3] = sqrtf(r2);
	}
}
#endif

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

KERNEL k_relabel_clusters(u32* label, const u32* __restrict__ parent, u32 K)
{
	size_t c = GTID;
	if (c >= K)
		return;
	label[c] = parent[label[c]];
}

// ---- merge rounds -----------------------------------------------------------------------------------------------------
// Smallest-first agglomeration, bulk-synchronous: groups within 2x of the smallest open size are "movers" this round.
// A mover merges with its pick when the pick is mutual (both movers; the lower id survives), or when the pick is a
// larger, non-moving group that selected it as its best proposer (one proposer per group per round).
//
// The group graph is a flat, unordered list of directed edges (src, dst, shared-vertex weight), symmetric by construction.
// Every step of a round is a reduction that does not depend on the order of that list: the pick is an atomicMax over
// (score, ~neighbour id) keys, which is the order (score desc, min id asc, max id asc) of the edges of one group; the
// contraction inserts the relabelled edges into an open-addressed table that sums the integer weights and then compacts
// the table into the next list. So all rounds of a level run inside ONE persistent cooperative kernel, separated by grid
// barriers, with no host round trip and no CSR rebuild.
struct MergeArgs
{
	GroupInfo* info;
	u32* label;
	u32 *src[2], *dst[2], *w[2]; // edge lists, ping-pong
	u64* table_key;
	u32* table_w;
	u64 *best_key, *claim;
	u32 *best_w, *parent;
	u32* state; // [0] edges in list 0 on entry, [1] edge counter of the next list, [2] smallest open size, [3] merges of the round, [4] rounds run,
	            // [5] which list holds the final contracted graph, [6] its edge count
	u32 K, target, max_size, max_rounds;
	int use_bounds;
};

static const u64 EDGE_EMPTY = ~0ull;

DEVFN u32 hash_edge(u64 key)
{
	key ^= key >> 33;
	key *= 0xff51afd7ed558ccdull;
	key ^= key >> 33;
	key *= 0xc4ceb9fe1a85ec53ull;
	key ^= key >> 33;
	return u32(key);
}

DEVFN bool pick_key_of_edge(const MergeArgs& a, int cur, u32 e, u32* g_out, u64* key_out)
{
	u32 g = a.src[cur][e], h = a.dst[cur][e];
	const GroupInfo me = a.info[g];
	if (me.size == 0 || me.size >= a.target)
		return false;
	const GroupInfo other = a.info[h];
	// a neighbour that has reached the target size still absorbs small groups while the sum fits the maximum: the reference only
	// finalises a group when it is popped off the heap, and large groups are popped last (partition.cpp:556-566)
	if (other.size == 0)
		return false;
	if (me.size + other.size > a.max_size)
		return false;
	float score = g < h ? merge_score(me, other, a.w[cur][e], a.use_bounds != 0) : merge_score(other, me, a.w[cur][e], a.use_bounds != 0);
	if (!(score > 0.f))
		return false;
	*g_out = g;
	*key_out = (u64(__float_as_uint(score)) << 32) | u64(~h);
	return true;
}

// B: best admissible neighbour of every open group
DEVFN void mr_pick(const MergeArgs& a, int cur, u32 e)
{
	u32 g;
	u64 key;
	if (pick_key_of_edge(a, cur, e, &g, &key))
		atomicMax(reinterpret_cast<unsigned long long*>(&a.best_key[g]), (unsigned long long)key);
}

// B2 (edges): the winning edge leaves its weight with the group
DEVFN void mr_pick_weight(const MergeArgs& a, int cur, u32 e)
{
	u32 g;
	u64 key;
	if (pick_key_of_edge(a, cur, e, &g, &key) && a.best_key[g] == key)
		a.best_w[g] = a.w[cur][e];
}

// C: movers propose to non-movers
DEVFN void mr_propose(const MergeArgs& a, u32 g)
{
	u64 key = a.best_key[g];
	if (key == 0)
		return;
	u32 h = ~u32(key);
	u32 limit = a.state[2] * 2;
	if (a.info[g].size > limit || a.info[h].size <= limit)
		return;
	atomicMax(reinterpret_cast<unsigned long long*>(&a.claim[h]), (unsigned long long)((key & 0xffffffff00000000ull) | u64(~g)));
}

// D: merge
DEVFN bool mr_merge(const MergeArgs& a, u32 g)
{
	u64 key = a.best_key[g];
	if (key == 0)
		return false;
	u32 h = ~u32(key);
	u32 limit = a.state[2] * 2;
	if (a.info[g].size > limit)
		return false;
	u32 dst, src;
	if (a.info[h].size <= limit)
	{
		// mover-mover: mutual picks only, handled once by the lower id
		u64 hk = a.best_key[h];
		if (hk == 0 || ~u32(hk) != g || h < g)
			return false;
		dst = g;
		src = h;
	}
	else
	{
		if (u32(~u32(a.claim[h])) != g)
			return false;
		dst = h;
		src = g;
	}
	u32 shared = a.best_w[g]; // weight of (g, h) == weight of (h, g)
	GroupInfo x = a.info[dst], y = a.info[src];
	x.size += y.size;
	x.vertices += y.vertices;
	x.vertices = x.vertices > shared ? x.vertices - shared : 1;
	merge_bounds(x, y);
	y.size = 0;
	y.vertices = 0;
	y.radius = 0;
	a.info[dst] = x;
	a.info[src] = y;
	a.parent[src] = dst;
	return true;
}

// E: relabel every edge to the merged groups, drop self loops, combine parallel edges in the table
DEVFN void mr_contract_insert(const MergeArgs& a, int cur, u32 e, u32 mask)
{
	u64 s = a.parent[a.src[cur][e]], d = a.parent[a.dst[cur][e]];
	if (s == d)
		return;
	u64 key = s * u64(a.K) + d;
	u32 slot = hash_edge(key) & mask;
	for (;;)
	{
		u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(&a.table_key[slot]), (unsigned long long)EDGE_EMPTY, (unsigned long long)key);
		if (old == EDGE_EMPTY || old == key)
			break;
		slot = (slot + 1) & mask;
	}
	atomicAdd(&a.table_w[slot], a.w[cur][e]);
}

HOSTDEVFN u32 table_mask_for(u32 edges)
{
	u32 cap = 2;
	while (cap < edges * 2)
		cap <<= 1;
	return cap - 1;
}

#ifdef CLODB_EMU
KERNEL k_mr_pick(MergeArgs a, int cur, u32 n)
{
	if (GTID < n)
		mr_pick(a, cur, u32(GTID));
}
KERNEL k_mr_pick_weight(MergeArgs a, int cur, u32 n)
{
	if (GTID < n)
		mr_pick_weight(a, cur, u32(GTID));
}
KERNEL k_mr_min_size(MergeArgs a)
{
	u32 g = u32(GTID);
	if (g < a.K && a.best_key[g] != 0)
		atomicMin(&a.state[2], a.info[g].size);
}
KERNEL k_mr_propose(MergeArgs a)
{
	if (GTID < a.K)
		mr_propose(a, u32(GTID));
}
KERNEL k_mr_merge(MergeArgs a)
{
	if (GTID < a.K && mr_merge(a, u32(GTID)))
		a.state[3]++;
}
KERNEL k_mr_insert(MergeArgs a, int cur, u32 n, u32 mask)
{
	if (GTID < n)
		mr_contract_insert(a, cur, u32(GTID), mask);
}
KERNEL k_mr_compact(MergeArgs a, int nxt, u32 cap)
{
	u32 slot = u32(GTID);
	if (slot >= cap || a.table_key[slot] == EDGE_EMPTY)
		return;
	u32 pos = a.state[1]++;
	a.src[nxt][pos] = u32(a.table_key[slot] / a.K);
	a.dst[nxt][pos] = u32(a.table_key[slot] % a.K);
	a.w[nxt][pos] = a.table_w[slot];
	a.table_key[slot] = EDGE_EMPTY;
	a.table_w[slot] = 0;
}
KERNEL k_mr_reset(MergeArgs a)
{
	u32 g = u32(GTID);
	if (g >= a.K)
		return;
	a.label[g] = a.parent[a.label[g]];
}
KERNEL k_mr_reset2(MergeArgs a)
{
	u32 g = u32(GTID);
	if (g >= a.K)
		return;
	a.best_key[g] = 0;
	a.claim[g] = 0;
	a.parent[g] = g;
}

static void merge_rounds(MergeArgs a, u32 E)
{
	int cur = 0;
	u32 rounds = 0;
	while (E > 0 && rounds < a.max_rounds)
	{
		rounds++;
		a.state[2] = 0xffffffffu, a.state[3] = 0, a.state[1] = 0;
		LAUNCH(k_mr_pick, E, a, cur, E);
		LAUNCH(k_mr_pick_weight, E, a, cur, E);
		LAUNCH(k_mr_min_size, a.K, a);
		LAUNCH(k_mr_propose, a.K, a);
		LAUNCH(k_mr_merge, a.K, a);
		if (a.state[3] == 0)
			break;
		u32 mask = table_mask_for(E);
		LAUNCH(k_mr_reset, a.K, a);
		LAUNCH(k_mr_insert, E, a, cur, E, mask);
		LAUNCH(k_mr_compact, size_t(mask) + 1, a, cur ^ 1, mask + 1);
		LAUNCH(k_mr_reset2, a.K, a);
		E = a.state[1];
		cur ^= 1;
	}
	a.state[4] = rounds;
	a.state[5] = u32(cur);
	a.state[6] = E;
}
#else
static const int MERGE_THREADS = 512;

// append-compaction of a grid-stride loop: one atomicAdd per warp
DEVFN u32 warp_append(bool take, u32* counter)
{
	unsigned mask = __ballot_sync(0xffffffffu, take);
	if (!mask)
		return 0;
	u32 lane = threadIdx.x & 31;
	int leader = __ffs(mask) - 1;
	u32 off = 0;
	if (int(lane) == leader)
		off = atomicAdd(counter, u32(__popc(mask)));
	off = __shfl_sync(0xffffffffu, off, leader);
	return off + __popc(mask & ((1u << lane) - 1));
}

static __global__ void __launch_bounds__(MERGE_THREADS, 2) k_merge_rounds(MergeArgs a)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const u32 gsize = gridDim.x * blockDim.x;
	const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x;
	u32 E = *reinterpret_cast<volatile u32*>(a.state);
	int cur = 0;
	u32 rounds = 0;
	while (E > 0 && rounds < a.max_rounds)
	{
		rounds++;
		for (u32 e = gtid; e < E; e += gsize)
			mr_pick(a, cur, e);
		grid.sync();
		if (gtid == 0)
			a.state[1] = 0;
		for (u32 e = gtid; e < E; e += gsize)
			mr_pick_weight(a, cur, e);
		for (u32 base = gtid - (gtid & 31); base < a.K; base += gsize)
		{
			u32 g = base + (gtid & 31);
			u32 size = g < a.K && a.best_key[g] != 0 ? a.info[g].size : 0xffffffffu;
			size = __reduce_min_sync(0xffffffffu, size);
			if ((gtid & 31) == 0 && size != 0xffffffffu)
				atomicMin(&a.state[2], size);
		}
		grid.sync();
		for (u32 g = gtid; g < a.K; g += gsize)
			mr_propose(a, g);
		grid.sync();
		for (u32 base = gtid - (gtid & 31); base < a.K; base += gsize)
		{
			u32 g = base + (gtid & 31);
			bool merged = g < a.K && mr_merge(a, g);
			unsigned m = __ballot_sync(0xffffffffu, merged);
			if ((gtid & 31) == 0 && m)
				atomicAdd(&a.state[3], u32(__popc(m)));
		}
		grid.sync();
		if (*reinterpret_cast<volatile u32*>(a.state + 3) == 0)
			break;
		const u32 mask = table_mask_for(E);
		for (u32 g = gtid; g < a.K; g += gsize)
			a.label[g] = a.parent[a.label[g]];
		for (u32 e = gtid; e < E; e += gsize)
			mr_contract_insert(a, cur, e, mask);
		grid.sync();
		const int nxt = cur ^ 1;
		for (u32 base = gtid - (gtid & 31); base <= mask; base += gsize)
		{
			u32 slot = base + (gtid & 31);
			// tables of fewer than 32 slots (a handful of edges on a small level) end inside the warp's first chunk
			u64 key = slot <= mask ? a.table_key[slot] : EDGE_EMPTY;
			bool take = key != EDGE_EMPTY;
			u32 pos = warp_append(take, a.state + 1);
			if (take)
			{
				a.src[nxt][pos] = u32(key / a.K);
				a.dst[nxt][pos] = u32(key % a.K);
				a.w[nxt][pos] = a.table_w[slot];
				a.table_key[slot] = EDGE_EMPTY;
				a.table_w[slot] = 0;
			}
		}
		for (u32 g = gtid; g < a.K; g += gsize)
		{
			a.best_key[g] = 0;
			a.claim[g] = 0;
			a.parent[g] = g;
		}
		if (gtid == 0)
		{
			a.state[2] = 0xffffffffu;
			a.state[3] = 0;
		}
		grid.sync();
		E = *reinterpret_cast<volatile u32*>(a.state + 1);
		cur = nxt;
	}
	if (gtid == 0)
	{
		a.state[4] = rounds;
		a.state[5] = u32(cur);
		a.state[6] = E;
	}
}

static void merge_rounds(MergeArgs a, u32 E)
{
	static u32 max_blocks = 0;
	if (!max_blocks)
	{
		int per_sm = 0, device = 0, sms = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_rounds, MERGE_THREADS, 0));
		CUDA_CHECK(cudaGetDevice(&device));
		CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
		max_blocks = u32(std::max(1, per_sm) * sms);
	}
	size_t widest = std::max<size_t>(size_t(table_mask_for(E)) + 1, a.K);
	u32 blocks = u32(std::min<size_t>(max_blocks, (widest + MERGE_THREADS - 1) / MERGE_THREADS));
	LAUNCH_COOP(k_merge_rounds, blocks, MERGE_THREADS, a);
}
#endif

// spatial merge of the leftover small groups along the Morton order of their centres (replaces mergeSpatial's kd-tree
// leaves, partition.cpp:368-480): each still-open group looks at its nearest open neighbours in that order
// groups that still have a neighbour in the contracted graph
KERNEL k_mark_connected(const u32* __restrict__ edge_src, u32 E, u32* isolated)
{
	size_t e = GTID;
	if (e < E)
		isolated[edge_src[e]] = 0;
}

KERNEL k_leftover_pick(const GroupInfo* __restrict__ info, const u32* __restrict__ order, const u32* __restrict__ isolated, u32 n, u32 target, u32 max_size, u32* best)
{
	size_t p = GTID;
	if (p >= n)
		return;
	u32 g = order[p];
	best[g] = NONE;
	const GroupInfo& me = info[g];
	if (me.size == 0 || me.size >= target || !isolated[g])
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
		if (other.size == 0 || me.size + other.size > max_size || !isolated[h])
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
// slow path of the refined-id cap (clusterlod.h:432-507): bucket order by first occurrence, clusters keep their relative
// order inside a bucket, a new group starts after every `cap` buckets
DEVFN void refined_cap_split(u32 begin, u32 n, u32* group_clusters, const int* __restrict__ cluster_refined, u32 cap, u32* scratch_clusters, u32* split_marks, u32* extra_groups)
{
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
		atomicAdd(extra_groups, 1u); // one per split partition, as the reference's instrumentation counter counts
}

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
	refined_cap_split(begin, n, group_clusters, cluster_refined, cap, scratch_clusters, split_marks, extra_groups);
}

#ifndef CLODB_EMU
// One warp per group: the distinct refined ids are counted 32 clusters at a time (match_any de-duplicates a chunk, chunk
// leaders are checked against the ids seen so far); only groups above the cap take the serial split.
static __global__ void __launch_bounds__(256) k_refined_cap_warp(const u32* __restrict__ group_offset, u32* group_clusters, const int* __restrict__ cluster_refined, u32 G, u32 cap, u32* scratch_clusters, u32* split_marks, u32* extra_groups)
{
	__shared__ int s_keys[8][64];
	const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 g = blockIdx.x * 8 + w;
	if (g >= G)
		return;
	const u32 begin = group_offset[g], n = group_offset[g + 1] - begin;
	const u32 limit = cap < 63 ? cap : 63;
	u32 nkeys = 0;
	for (u32 base = 0; base < n && nkeys <= limit; base += 32)
	{
		u32 j = base + lane;
		bool valid = j < n;
		int r = valid ? cluster_refined[group_clusters[begin + j]] : 0;
		unsigned peers = __match_any_sync(0xffffffffu, u32(r)) & __ballot_sync(0xffffffffu, valid);
		bool leader = valid && (__ffs(peers) - 1 == int(lane));
		bool seen = false;
		for (u32 k = 0; k < nkeys; ++k)
			seen |= s_keys[w][k] == r;
		bool fresh = leader && !seen;
		unsigned mask = __ballot_sync(0xffffffffu, fresh);
		u32 slot = nkeys + __popc(mask & ((1u << lane) - 1));
		if (fresh && slot < 64)
			s_keys[w][slot] = r;
		nkeys += __popc(mask);
		__syncwarp();
	}
	if (nkeys <= cap)
		return;
	if (lane == 0)
		refined_cap_split(begin, n, group_clusters, cluster_refined, cap, scratch_clusters, split_marks, extra_groups);
}
#endif

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

// Second half of clod::partition (clusterlod.h:396-507), given a partition id per cluster: spatial order of the partitions by the
// centre of their last cluster (meshopt_spatialSortRemap, spatialorder.cpp:218-251), clusters in ascending index inside a
// partition, then the refined-id cap split. Deterministic given `part`; fills out.group_clusters / group_cluster_offset.
static void finish_groups(u32* part, const u32* part_last, u32 G, u32 K, const int* cluster_refined, const float* cluster_bounds5, const Config& config, Workspace& ws, u32* scalars, GroupSet& out)
{
	Arena& temp = ws.temp;
	u32 cap = config.partition_max_refined_groups;
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
#ifdef CLODB_EMU
		LAUNCH(k_refined_cap, G, group_offset, out.group_clusters, cluster_refined, G, cap, scratch_clusters, marks, scalars + 2);
#else
		LAUNCH_GRID(k_refined_cap_warp, (G + 7) / 8, 256, group_offset, out.group_clusters, cluster_refined, G, cap, scratch_clusters, marks, scalars + 2);
#endif
	}
	LAUNCH(k_group_start_flags, G, group_offset, G, marks);
	u32* marks_scanned = temp.alloc<u32>(size_t(K) + 1);
	exclusive_scan_u32(marks, marks_scanned, K, scalars, temp);
	u32 G_final = dev_read(scalars);
	LAUNCH(k_emit_group_offsets, size_t(K) + 1, marks, marks_scanned, K, G_final, out.group_cluster_offset);

	out.group_count = G_final;
	out.refined_splits = cap > 0 && G_final != G ? dev_read(scalars + 2) : 0;
	out.group_cluster_offset_host = dev_download(out.group_cluster_offset, size_t(G_final) + 1);
	out.merge_rounds = 0;
}

KERNEL k_part_last(const u32* __restrict__ part, u32* part_last_cluster, u32 K)
{
	size_t c = GTID;
	if (c < K)
		atomicMax(&part_last_cluster[part[c]], u32(c));
}

GroupSet partition_finish(const u32* part_in, u32 G, u32 K, const int* cluster_refined, const float* cluster_bounds5, const Config& config, Workspace& ws)
{
	GroupSet out;
	out.cluster_count = K;
	if (K == 0 || G == 0)
		return out;
	Arena& temp = ws.temp;
	out.group_clusters = ws.persist.alloc<u32>(K);
	out.group_cluster_offset = ws.persist.alloc<u32>(size_t(K) + 1);
	ArenaScope scope(temp);
	u32* scalars = temp.alloc<u32>(16);
	dev_memset(scalars, 0, 16 * sizeof(u32));
	u32* part = temp.alloc<u32>(K);
	u32* part_last = temp.alloc<u32>(G);
	dev_d2d(part, part_in, size_t(K) * 4);
	dev_memset(part_last, 0, size_t(G) * 4);
	LAUNCH(k_part_last, K, part, part_last, K);
	finish_groups(part, part_last, G, K, cluster_refined, cluster_bounds5, config, ws, scalars, out);
	return out;
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
#ifdef CLODB_EMU
	LAUNCH(k_cluster_unique, K, tri, cluster_tri_offset, remap, positions, K, stride, cv, cv_count, center_radius);
#else
	if (stride <= 128 && config.max_triangles * 3 <= 384)
		LAUNCH_GRID(k_cluster_unique_warp, (K + CU_WARPS - 1) / CU_WARPS, CU_WARPS * 32, tri, cluster_tri_offset, remap, positions, K, stride, cv, cv_count, center_radius);
	else
		LAUNCH(k_cluster_unique, K, tri, cluster_tri_offset, remap, positions, K, stride, cv, cv_count, center_radius);
#endif
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
	// ---- merge rounds: one persistent kernel, no host round trip (state: see MergeArgs)
	u32* final_src[2] = {nullptr, nullptr};
	bool have_graph = false;
	if (E > 0)
	{
		size_t table_cap = 2;
		while (table_cap < size_t(E) * 2)
			table_cap <<= 1;
		MergeArgs ma;
		ma.info = info, ma.label = label;
		ma.src[0] = e_src, ma.dst[0] = e_dst, ma.w[0] = e_w;
		ma.src[1] = temp.alloc<u32>(E), ma.dst[1] = temp.alloc<u32>(E), ma.w[1] = temp.alloc<u32>(E);
		ma.table_key = temp.alloc<u64>(table_cap);
		ma.table_w = temp.alloc<u32>(table_cap);
		ma.best_key = temp.alloc<u64>(K);
		ma.claim = claim;
		ma.best_w = best;
		ma.parent = parent;
		ma.state = scalars + 8;
		ma.K = K, ma.target = target, ma.max_size = max_size, ma.max_rounds = 4096;
		ma.use_bounds = config.partition_spatial ? 1 : 0;
		dev_memset(ma.table_key, 0xff, table_cap * sizeof(u64));
		dev_memset(ma.table_w, 0, table_cap * sizeof(u32));
		dev_memset(ma.best_key, 0, size_t(K) * sizeof(u64));
		dev_memset(claim, 0, size_t(K) * sizeof(u64));
		iota(parent, K);
		u32 st[7] = {E, 0, 0xffffffffu, 0, 0, 0, E};
		dev_h2d(ma.state, st, sizeof(st));
		merge_rounds(ma, E);
		final_src[0] = ma.src[0], final_src[1] = ma.src[1];
		have_graph = true;
	}

	// ---- leftovers: spatially merge open groups (only when positions are used, partition.cpp:599-611) - but only groups that have
	// no neighbour left in the contracted graph (separate components: islands, soups). An open group that is still surrounded by
	// neighbours which cannot take it stays a smaller group: pairing it with a non-adjacent group nearby gives a group of two
	// separate patches, each with its own locked border, and those were the groups with the largest simplification errors
	// (3 of 70 on a 3.4 M-triangle heightfield, all three at the top of the error ranking; the reference's kd-tree leaves produce 1).
	if (config.partition_spatial)
	{
		u32* isolated = temp.alloc<u32>(K);
		fill(isolated, 1u, K);
		if (have_graph)
		{
			std::vector<u32> final_state = dev_download(scalars + 8, 7);
			if (final_state[6])
				LAUNCH(k_mark_connected, final_state[6], final_src[final_state[5] & 1], final_state[6], isolated);
		}
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
			LAUNCH(k_leftover_pick, live, info, order, isolated, live, target, max_size, best);
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

	finish_groups(part, part_last, G, K, cluster_refined, cluster_bounds5, config, ws, scalars, out);
	return out;
}

} // namespace clodb
