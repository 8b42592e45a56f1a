// S3: spatial clusterization (meshlet building) for a batch of independent triangle segments.
//
// Reference semantics: clod::clusterize (clusterlod.h:305-348) -> meshopt_buildMeshletsSpatial
// (ThirdParty/meshoptimizer/src/clusterizer.cpp:1380-1477) -> bvhSplit (:1051-1124), bvhPivot (:985-1033),
// bvhComputeArea/boxMerge (:970-983, :764-805), bvhPackTail (:940-960), then meshopt_optimizeMeshlet (:1682-1770).
//
// B200 formulation. The reference recurses depth-first over one segment; here every tree level of every segment is
// processed breadth-first by flat kernels over the three axis orders:
//   - axis orders come from a stable LSD radix sort on (segment, radixFloat(centroid) >> 2), identical to the
//     reference's 3x10-bit radix (which drops the 2 LSBs and is stable);
//   - prefix/suffix surface areas are segmented min/max scans over the node ranges (min/max are exact and associative,
//     so the parallel scan reproduces the sequential accumulation bit for bit);
//   - the SAH+fill pivot is an argmin over (cost, axis, index) keys, which equals the reference's "first strictly
//     smaller cost wins" scan order;
//   - the stable two-way partition of the other axes is a global exclusive scan of side flags;
//   - nodes of <= max_triangles triangles (leaf tests, vertex-bound splits, tails) are resolved by one thread each.
// Result: the same meshlets, in the same order, with the same optimized triangle/vertex order as the reference.
#include "clodb.h"

#include <cfloat>
#include <algorithm>

namespace clodb
{

static const u32 NODE_DONE = 0xffffffffu;
static const int kMeshletMaxTreeDepth = 50; // clusterizer.cpp:40

struct Box
{
	float min[4];
	float max[4];
};

struct SplitParams
{
	u32 max_vertices, min_triangles, max_triangles;
	float fill_weight;
};

DEVFN u32 radix_float(u32 v)
{
	u32 mask = u32(int(v) >> 31) | 0x80000000u;
	return v ^ mask;
}

// clusterizer.cpp:876-902 (bvhPrepare); the centroid key is what the reference's radix passes sort by
KERNEL k_prepare(const u32* __restrict__ tri, const float* __restrict__ positions, u32 T, Box* boxes, u32* key0, u32* key1, u32* key2)
{
	size_t i = GTID;
	if (i >= T)
		return;
	u32 a = tri[i * 3 + 0], b = tri[i * 3 + 1], c = tri[i * 3 + 2];
	const float* va = positions + size_t(a) * 3;
	const float* vb = positions + size_t(b) * 3;
	const float* vc = positions + size_t(c) * 3;
	Box box;
	u32 keys[3];
	for (int k = 0; k < 3; ++k)
	{
		float mn = va[k] < vb[k] ? va[k] : vb[k];
		mn = vc[k] < mn ? vc[k] : mn;
		float mx = va[k] > vb[k] ? va[k] : vb[k];
		mx = vc[k] > mx ? vc[k] : mx;
		box.min[k] = mn;
		box.max[k] = mx;
		float centroid = (mn + mx) / 2.f;
		keys[k] = radix_float(__float_as_uint(centroid)) >> 2;
	}
	box.min[3] = 0.f;
	box.max[3] = 0.f;
	boxes[i] = box;
	key0[i] = keys[0];
	key1[i] = keys[1];
	key2[i] = keys[2];
}

KERNEL k_widen_keys(const u32* __restrict__ key32, const u32* __restrict__ seg_of_tri, u64* key64, u32 T)
{
	size_t i = GTID;
	if (i >= T)
		return;
	key64[i] = (u64(seg_of_tri[i]) << 30) | key32[i];
}

KERNEL k_segment_of_tri(const u32* __restrict__ seg_offsets, u32 S, u32* seg_of_tri, u32 T)
{
	size_t i = GTID;
	if (i >= T)
		return;
	u32 lo = 0, hi = S; // find s with seg_offsets[s] <= i < seg_offsets[s+1]
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (seg_offsets[mid] <= u32(i))
			lo = mid;
		else
			hi = mid;
	}
	seg_of_tri[i] = lo;
}

KERNEL k_init_nodes(const u32* __restrict__ seg_offsets, u32 S, u32* node_begin, u32* node_count, u32* node_of_pos)
{
	size_t s = GTID;
	if (s >= S)
		return;
	node_begin[s] = seg_offsets[s];
	node_count[s] = seg_offsets[s + 1] - seg_offsets[s];
}

KERNEL k_init_node_of_pos(const u32* __restrict__ seg_of_tri_sorted_pos, u32* node_of_pos, u32 T)
{
	size_t i = GTID;
	if (i >= T)
		return;
	node_of_pos[i] = seg_of_tri_sorted_pos[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// small per-thread open-addressing set of vertex ids (<= 3 * 128 insertions)
struct VertexSet
{
	u32 slots[512];

	DEVFN void clear(VertexSet& s)
	{
		for (int i = 0; i < 512; ++i)
			s.slots[i] = 0xffffffffu;
	}
	// returns 1 when v was not present
	DEVFN u32 insert(VertexSet& s, u32 v)
	{
		u32 h = (v * 0x9E3779B1u) >> 23;
		for (;;)
		{
			u32 cur = s.slots[h];
			if (cur == v)
				return 0;
			if (cur == 0xffffffffu)
			{
				s.slots[h] = v;
				return 1;
			}
			h = (h + 1) & 511;
		}
	}
};

DEVFN float box_area_merge(float* mn, float* mx, const Box& other)
{
	for (int k = 0; k < 3; ++k)
	{
		mn[k] = mn[k] < other.min[k] ? mn[k] : other.min[k];
		mx[k] = mx[k] > other.max[k] ? mx[k] : other.max[k];
	}
	float sx = mx[0] - mn[0], sy = mx[1] - mn[1], sz = mx[2] - mn[2];
	// summation order of the SSE2 boxMerge the reference is compiled with on x86-64 (clusterizer.cpp:764-782)
	return (sx * sy + sy * sz) + sz * sx;
}

DEVFN bool bvh_divisible(u32 count, u32 mn, u32 mx)
{
	return mn * 2 <= mx ? count >= mn : count % mn <= (count / mn) * (mx - mn);
}

// SAH + fill cost of splitting after local index i (clusterizer.cpp:985-1033); returns false when not admissible
DEVFN bool pivot_cost(u32 i, u32 count, u32 mn, u32 mx, bool aligned, float larea, float rarea, u32 lfill_v, bool has_vertices, float fill, u32 maxfill, float* out_cost)
{
	u32 lsplit = i + 1, rsplit = count - (i + 1);
	if (!bvh_divisible(lsplit, mn, mx))
		return false;
	if (aligned && !bvh_divisible(rsplit, mn, mx))
		return false;
	float cost = larea * float(int(lsplit)) + rarea * float(int(rsplit));
	u32 lfill = has_vertices ? lfill_v : lsplit;
	u32 rfill = has_vertices ? lfill_v : rsplit;
	float rmaxfill = 1.f / float(int(maxfill));
	int lrest = int(float(int(lfill + maxfill - 1)) * rmaxfill) * int(maxfill) - int(lfill);
	int rrest = int(float(int(rfill + maxfill - 1)) * rmaxfill) * int(maxfill) - int(rfill);
	cost += fill * (float(lrest) * larea + float(rrest) * rarea);
	*out_cost = cost;
	return true;
}

// bvhPackTail (clusterizer.cpp:940-960) on positions [begin, begin+count) of order
DEVFN void pack_tail(u8* boundary, const u32* order, const u32* tri, u32 begin, u32 count, u32 max_vertices, u32 max_triangles, VertexSet& set)
{
	for (u32 i = 0; i < count;)
	{
		u32 chunk = i + max_triangles <= count ? max_triangles : count - i;
		VertexSet::clear(set);
		u32 used = 0;
		for (u32 j = 0; j < chunk; ++j)
		{
			u32 t = order[begin + i + j];
			used += VertexSet::insert(set, tri[size_t(t) * 3 + 0]);
			used += VertexSet::insert(set, tri[size_t(t) * 3 + 1]);
			used += VertexSet::insert(set, tri[size_t(t) * 3 + 2]);
		}
		u32 take = used <= max_vertices ? chunk : max_vertices / 3;
		boundary[begin + i] = 1;
		for (u32 j = 1; j < take; ++j)
			boundary[begin + i + j] = 0;
		i += take;
	}
}

// Nodes that fit the triangle limit: leaf test, vertex-bound split or tail, all by one thread (bvhSplit for
// count <= max_triangles; the reference's recursion at this size touches <= 128 triangles per node).
// Large nodes that hit the depth limit or found no admissible split are packed here as well.
KERNEL k_resolve_nodes(const u32* __restrict__ node_begin, const u32* __restrict__ node_count, u64* node_best, u32* node_split, u32 node_total, int depth,
    const u32* __restrict__ order0, const u32* __restrict__ order1, const u32* __restrict__ order2, const Box* __restrict__ boxes, const u32* __restrict__ tri, u8* boundary, SplitParams sp, int pass,
    const u8* __restrict__ node_pending)
{
	size_t n = GTID;
	if (n >= node_total)
		return;
	// pass 0 on the GPU: the warp-cooperative leaf test has already settled every node that fits the vertex limit
	if (pass == 0 && node_pending && !node_pending[n])
		return;
	u32 begin = node_begin[n], count = node_count[n];
	VertexSet set;

	if (pass == 1)
	{
		// after the scan-based pivot search over large nodes: no admissible split or depth limit => tail
		if (count <= sp.max_triangles)
			return;
		u64 best = node_best[n];
		if (best == ~u64(0) || depth >= kMeshletMaxTreeDepth)
		{
			pack_tail(boundary, order0, tri, begin, count, sp.max_vertices, sp.max_triangles, set);
			node_split[n] = 0;
		}
		else
		{
			node_split[n] = u32(best & 0x3fffffffu) + 1;
		}
		return;
	}

	if (count > sp.max_triangles)
		return;

	const u32* orders[3] = {order0, order1, order2};

	// leaf test (clusterizer.cpp:1053-1054)
	VertexSet::clear(set);
	u32 used = 0;
	for (u32 j = 0; j < count; ++j)
	{
		u32 t = order0[begin + j];
		used += VertexSet::insert(set, tri[size_t(t) * 3 + 0]);
		used += VertexSet::insert(set, tri[size_t(t) * 3 + 1]);
		used += VertexSet::insert(set, tri[size_t(t) * 3 + 2]);
	}
	if (used <= sp.max_vertices)
	{
		boundary[begin] = 1;
		for (u32 j = 1; j < count; ++j)
			boundary[begin + j] = 0;
		node_split[n] = 0;
		node_best[n] = ~u64(0);
		return;
	}

	// vertex bound: split with vertex-fill cost (clusterizer.cpp:1061-1092)
	u32 mint = sp.max_vertices / 3 < sp.min_triangles ? sp.max_vertices / 3 : sp.min_triangles;
	u32 maxfill = sp.max_vertices;
	bool aligned = count >= mint * 2 && bvh_divisible(count, mint, sp.max_triangles);
	u32 end = aligned ? count - mint : count - 1;

	int bestk = -1;
	u32 bestsplit = 0;
	float bestcost = FLT_MAX;

	float lareas[128];
	u32 lverts[128];

	for (int k = 0; k < 3; ++k)
	{
		const u32* order = orders[k];
		float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		VertexSet::clear(set);
		u32 vcount = 0;
		for (u32 j = 0; j < count; ++j)
		{
			u32 t = order[begin + j];
			lareas[j] = box_area_merge(mn, mx, boxes[t]);
			vcount += VertexSet::insert(set, tri[size_t(t) * 3 + 0]);
			vcount += VertexSet::insert(set, tri[size_t(t) * 3 + 1]);
			vcount += VertexSet::insert(set, tri[size_t(t) * 3 + 2]);
			lverts[j] = vcount;
		}
		// suffix areas are consumed from the right; walk candidates descending while keeping the reference's
		// ascending "first minimum wins" rule by tracking (cost, index)
		float rmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
		float axiscost = FLT_MAX;
		u32 axissplit = 0;
		// rarea for candidate i covers [i+1, count-1]
		u32 covered = count; // suffix currently covers [covered, count-1]
		float rarea = 0.f;
		for (u32 i = end; i-- > mint - 1 + 0;)
		{
			while (covered > i + 1)
			{
				covered--;
				rarea = box_area_merge(rmn, rmx, boxes[order[begin + covered]]);
			}
			float cost;
			if (!pivot_cost(i, count, mint, sp.max_triangles, aligned, lareas[i], rarea, lverts[i], true, sp.fill_weight, maxfill, &cost))
				continue;
			// ascending scan keeps the first strict minimum; descending equivalent: take when cost <= best
			if (cost < FLT_MAX && cost <= axiscost)
			{
				axiscost = cost;
				axissplit = i + 1;
			}
			if (i == 0)
				break;
		}
		if (axissplit && axiscost < bestcost)
		{
			bestk = k;
			bestcost = axiscost;
			bestsplit = axissplit;
		}
	}

	if (bestk < 0 || depth >= kMeshletMaxTreeDepth)
	{
		pack_tail(boundary, order0, tri, begin, count, sp.max_vertices, sp.max_triangles, set);
		node_split[n] = 0;
		node_best[n] = ~u64(0);
		return;
	}

	node_split[n] = bestsplit;
	node_best[n] = (u64(bestk) << 30) | u64(bestsplit - 1);
}

#ifndef CLODB_EMU
// Leaf test of bvhSplit (clusterizer.cpp:1053-1054) for nodes within the triangle limit, one warp per node: the node's
// corners are inserted into a shared-memory vertex set; nodes that also fit the vertex limit become meshlets here, the
// (rare) vertex-bound ones are flagged for the serial k_resolve_nodes path.
static const int LT_WARPS = 8;
static __global__ void __launch_bounds__(LT_WARPS * 32) k_leaf_test_warp(const u32* __restrict__ node_begin, const u32* __restrict__ node_count, u64* node_best, u32* node_split, u32 node_total, const u32* __restrict__ order0,
    const u32* __restrict__ tri, u8* boundary, SplitParams sp, u8* node_pending)
{
	__shared__ u32 s_keys[LT_WARPS][512];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 n = blockIdx.x * LT_WARPS + warp;
	if (n >= node_total)
		return;
	const u32 begin = node_begin[n], count = node_count[n];
	if (count > sp.max_triangles)
	{
		if (lane == 0)
			node_pending[n] = 0;
		return;
	}
	u32* keys = s_keys[warp];
	for (int i = lane; i < 512; i += 32)
		keys[i] = 0xffffffffu;
	__syncwarp();
	u32 fresh = 0;
	for (u32 j = lane; j < count * 3; j += 32)
	{
		u32 t = order0[begin + j / 3];
		u32 v = tri[size_t(t) * 3 + j % 3];
		u32 h = (v * 0x9E3779B1u) >> 23;
		for (;;)
		{
			u32 old = atomicCAS(&keys[h], 0xffffffffu, v);
			if (old == 0xffffffffu)
			{
				fresh++;
				break;
			}
			if (old == v)
				break;
			h = (h + 1) & 511;
		}
	}
	for (int d = 16; d >= 1; d >>= 1)
		fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
	if (fresh <= sp.max_vertices)
	{
		for (u32 j = lane; j < count; j += 32)
			boundary[begin + j] = j == 0 ? 1 : 0;
		if (lane == 0)
		{
			node_split[n] = 0;
			node_best[n] = ~u64(0);
			node_pending[n] = 0;
		}
	}
	else if (lane == 0)
		node_pending[n] = 1;
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// segmented prefix / suffix surface areas over active node ranges (bvhComputeArea, clusterizer.cpp:970-983)

DEVFN u32 node_of(const u32* node_of_pos, size_t p)
{
	return node_of_pos[p];
}

#ifdef CLODB_EMU
static void seg_area_scan(const Box* boxes, const u32* order, const u32* node_of_pos, const u32* node_begin, const u32* node_count, float* out_area, u32 T, bool backward, Arena&)
{
	g_launches += 3;
	float mn[3], mx[3];
	for (u32 j = 0; j < T; ++j)
	{
		u32 p = backward ? T - 1 - j : j;
		u32 n = node_of_pos[p];
		bool head = n == NODE_DONE || (backward ? p == node_begin[n] + node_count[n] - 1 : p == node_begin[n]);
		if (head || j == 0)
			for (int k = 0; k < 3; ++k)
				mn[k] = FLT_MAX, mx[k] = -FLT_MAX;
		out_area[p] = box_area_merge(mn, mx, boxes[order[p]]);
	}
}

static void seg_area_scan_all(const Box* boxes, u32* const* order, const u32* node_of_pos, const u32* node_begin, const u32* node_count, float* areas, u32 T, Arena& temp)
{
	for (int k = 0; k < 3; ++k)
		for (int b = 0; b < 2; ++b)
			seg_area_scan(boxes, order[k], node_of_pos, node_begin, node_count, areas + size_t(k * 2 + b) * (size_t(T) + 1), T, b != 0, temp);
}
#else
struct ScanElem
{
	float mn[3], mx[3];
	u32 flag;
};

DEVFN ScanElem scan_identity()
{
	ScanElem e;
	e.mn[0] = e.mn[1] = e.mn[2] = FLT_MAX;
	e.mx[0] = e.mx[1] = e.mx[2] = -FLT_MAX;
	e.flag = 0;
	return e;
}

// a then b
DEVFN ScanElem scan_combine(const ScanElem& a, const ScanElem& b)
{
	if (b.flag)
		return b;
	ScanElem r;
	for (int k = 0; k < 3; ++k)
	{
		r.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k];
		r.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k];
	}
	r.flag = a.flag;
	return r;
}

DEVFN ScanElem scan_shfl_up(const ScanElem& e, int d)
{
	ScanElem r;
	for (int k = 0; k < 3; ++k)
	{
		r.mn[k] = __shfl_up_sync(0xffffffffu, e.mn[k], d);
		r.mx[k] = __shfl_up_sync(0xffffffffu, e.mx[k], d);
	}
	r.flag = __shfl_up_sync(0xffffffffu, e.flag, d);
	return r;
}

DEVFN ScanElem scan_shfl_down(const ScanElem& e, int d)
{
	ScanElem r;
	for (int k = 0; k < 3; ++k)
	{
		r.mn[k] = __shfl_down_sync(0xffffffffu, e.mn[k], d);
		r.mx[k] = __shfl_down_sync(0xffffffffu, e.mx[k], d);
	}
	r.flag = 0;
	return r;
}

DEVFN ScanElem scan_shfl_idx(const ScanElem& e, int src)
{
	ScanElem r;
	for (int k = 0; k < 3; ++k)
	{
		r.mn[k] = __shfl_sync(0xffffffffu, e.mn[k], src);
		r.mx[k] = __shfl_sync(0xffffffffu, e.mx[k], src);
	}
	r.flag = 0;
	return r;
}

#ifndef CLODB_SA_THREADS
#define CLODB_SA_THREADS 128
#endif
#ifndef CLODB_SA_ITEMS
#define CLODB_SA_ITEMS 8
#endif
#ifndef CLODB_SA_MINBLOCKS
#define CLODB_SA_MINBLOCKS 1
#endif
static const int SA_THREADS = CLODB_SA_THREADS;
static const int SA_ITEMS = CLODB_SA_ITEMS;
static const int SA_TILE = SA_THREADS * SA_ITEMS;

// inclusive block scan of per-thread aggregates; returns the exclusive prefix for this thread; total via smem
DEVFN ScanElem block_scan_exclusive(const ScanElem& agg, ScanElem* smem /* 8 */, ScanElem* block_total)
{
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	ScanElem inc = agg;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		ScanElem t = scan_shfl_up(inc, d);
		if (lane >= d)
			inc = scan_combine(t, inc);
	}
	if (lane == 31)
		smem[warp] = inc;
	__syncthreads();
	ScanElem warp_prefix = scan_identity();
	for (int w = 0; w < warp; ++w)
		warp_prefix = scan_combine(warp_prefix, smem[w]);
	if (block_total)
	{
		ScanElem total = scan_identity();
		for (int w = 0; w < SA_THREADS / 32; ++w)
			total = scan_combine(total, smem[w]);
		*block_total = total;
	}
	ScanElem ex = scan_shfl_up(inc, 1);
	if (lane == 0)
		ex = scan_identity();
	__syncthreads();
	return scan_combine(warp_prefix, ex);
}

struct OpSegBox
{
	DEVFN ScanElem identity()
	{
		return scan_identity();
	}
	DEVFN ScanElem apply(const ScanElem& a, const ScanElem& b)
	{
		return scan_combine(a, b);
	}
	DEVFN bool prefix_independent(const ScanElem& a)
	{
		return a.flag != 0;
	}
};

// ---- look-back over packed box descriptors --------------------------------------------------------------------------
// A tile's aggregate (or inclusive prefix) is published as two self-validating 16-byte words {tag, min xyz} {tag, max xyz}:
// each word is written and read with one aligned 16-byte access, a reader accepts a descriptor when both tags agree, so
// there is no fence and a look-back window costs one L2 round trip (prims.cuh explains the bound). The segment flag does
// not travel: an aggregate with a head inside is published as an inclusive prefix, and the flag of the left-most operand
// of a combination is never inspected.
DEVFN void sa_desc_store(char* desc, u32 tile, u32 tag, const ScanElem& v)
{
	char* p = desc + size_t(tile) * 32;
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(tag), "r"(__float_as_uint(v.mn[0])), "r"(__float_as_uint(v.mn[1])), "r"(__float_as_uint(v.mn[2])) : "memory");
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + 16), "r"(tag), "r"(__float_as_uint(v.mx[0])), "r"(__float_as_uint(v.mx[1])), "r"(__float_as_uint(v.mx[2])) : "memory");
}

// returns the tag when both words carry the same one, 0 otherwise
DEVFN u32 sa_desc_load(const char* desc, u32 tile, ScanElem& v)
{
	const char* p = desc + size_t(tile) * 32;
	u32 t0, t1, a0, a1, a2, b0, b1, b2;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(a0), "=r"(a1), "=r"(a2) : "l"(p) : "memory");
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t1), "=r"(b0), "=r"(b1), "=r"(b2) : "l"(p + 16) : "memory");
	v.mn[0] = __uint_as_float(a0), v.mn[1] = __uint_as_float(a1), v.mn[2] = __uint_as_float(a2);
	v.mx[0] = __uint_as_float(b0), v.mx[1] = __uint_as_float(b1), v.mx[2] = __uint_as_float(b2);
	v.flag = 0;
	return t0 == t1 ? t0 : 0u;
}

// Called by all 32 lanes of warp 0 with the tile's aggregate; returns the tile's exclusive prefix (valid in every lane)
DEVFN ScanElem sa_lookback(u32 tile, const ScanElem& tile_aggregate, char* desc, u32 epoch)
{
	const int lane = threadIdx.x & 31;
	const u32 tag_agg = (epoch << 2) | 1u, tag_inc = (epoch << 2) | 2u;
	if (tile == 0)
	{
		if (lane == 0)
			sa_desc_store(desc, 0, tag_inc, tile_aggregate);
		return scan_identity();
	}
	const bool independent = tile_aggregate.flag != 0;
	if (lane == 0)
		sa_desc_store(desc, tile, independent ? tag_inc : tag_agg, tile_aggregate);
	ScanElem prefix = scan_identity();
	int p = int(tile) - 1;
	for (;;)
	{
		int idx = p - lane;
		ScanElem v = scan_identity();
		u32 t = tag_inc;
		for (;;)
		{
			if (idx >= 0)
				t = sa_desc_load(desc, u32(idx), v);
			if (__all_sync(0xffffffffu, t == tag_agg || t == tag_inc))
				break;
		}
		unsigned inc_mask = __ballot_sync(0xffffffffu, t == tag_inc);
		int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32; // nearest predecessor whose inclusive prefix is known (lanes past the start count)
		if (idx < 0 || lane > first_inc)
			v = scan_identity();
		// ordered reduction: lane L is farther back than lane L-1, so it is the left operand
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			ScanElem o = scan_shfl_down(v, d);
			if (lane + d < 32)
				v = scan_combine(o, v);
		}
		v = scan_shfl_idx(v, 0);
		prefix = scan_combine(v, prefix);
		if (inc_mask)
			break;
		p -= 32;
	}
	if (lane == 0 && !independent)
		sa_desc_store(desc, tile, tag_inc, scan_combine(prefix, tile_aggregate));
	return prefix;
}

// Segmented inclusive min/max scan of the triangle boxes along one axis order -> surface area at every position, in one
// chained pass (prims.cuh): per element it reads order u32 + node id u32 + the gathered 32-byte box and writes one f32.
struct SweepArgs
{
	const u32* order[3];
	float* area; // 6 arrays of T + 1 floats: [axis * 2 + backward]
};

// One thread owns the SA_ITEMS consecutive positions [q, q + SA_ITEMS) of a sweep, q a multiple of SA_ITEMS (the backward sweep counts from the
// padded end, so its chunks are aligned too). Everything a chunk needs is requested before anything is consumed: two
// 16-byte loads of the node ids, two of the order entries, one neighbouring node id, then all box gathers (two 16-byte
// loads = one sector each) are in flight together. Segment heads come from comparing neighbouring node ids (a node is a
// contiguous run of positions with a unique id), so the node tables are not touched.
template <bool BACK>
DEVFN void sa_load_chunk(const Box* __restrict__ boxes, const u32* __restrict__ order, const u32* __restrict__ node_of_pos, u32 T, long q, ScanElem (&e)[SA_ITEMS])
{
	static_assert(SA_ITEMS == 8 || SA_ITEMS == 4, "chunk layout");
	const int N = SA_ITEMS;
	u32 n[N], o[N];
	u32 edge = NODE_DONE; // node id just outside the chunk on the side the sweep comes from
	bool has_edge = false;
	if (q >= 0 && q + N <= long(T))
	{
		const uint4* np = reinterpret_cast<const uint4*>(node_of_pos + q);
		const uint4* op = reinterpret_cast<const uint4*>(order + q);
#pragma unroll
		for (int v = 0; v < N / 4; ++v)
		{
			uint4 nv = __ldg(np + v), ov = __ldg(op + v);
			n[v * 4 + 0] = nv.x, n[v * 4 + 1] = nv.y, n[v * 4 + 2] = nv.z, n[v * 4 + 3] = nv.w;
			o[v * 4 + 0] = ov.x, o[v * 4 + 1] = ov.y, o[v * 4 + 2] = ov.z, o[v * 4 + 3] = ov.w;
		}
	}
	else
	{
#pragma unroll
		for (int i = 0; i < N; ++i)
		{
			long p = q + i;
			bool in = p >= 0 && p < long(T);
			n[i] = in ? __ldg(node_of_pos + p) : NODE_DONE;
			o[i] = in ? __ldg(order + p) : 0u;
		}
	}
	long pe = BACK ? q + N : q - 1;
	if (pe >= 0 && pe < long(T))
	{
		edge = __ldg(node_of_pos + pe);
		has_edge = true;
	}
	float4 lo[N], hi[N];
#pragma unroll
	for (int i = 0; i < N; ++i)
		if (n[i] != NODE_DONE)
		{
			const float4* bp = reinterpret_cast<const float4*>(boxes + o[i]);
			lo[i] = __ldg(bp), hi[i] = __ldg(bp + 1);
		}
#pragma unroll
	for (int k = 0; k < N; ++k)
	{
		const int i = BACK ? N - 1 - k : k; // position inside the chunk of the k-th element in scan order
		long p = q + i;
		ScanElem x = scan_identity();
		if (p >= 0 && p < long(T))
		{
			if (n[i] == NODE_DONE)
				x.flag = 1; // finished nodes are never read by the pivot search: no box gather, and the scan restarts here
			else
			{
				x.mn[0] = lo[i].x, x.mn[1] = lo[i].y, x.mn[2] = lo[i].z;
				x.mx[0] = hi[i].x, x.mx[1] = hi[i].y, x.mx[2] = hi[i].z;
				// the element the sweep visited just before this one
				const int ip = BACK ? i + 1 : i - 1;
				bool head;
				if (ip < 0 || ip > N - 1)
					head = !has_edge || edge != n[i];
				else
					head = (BACK && p + 1 >= long(T)) || n[ip] != n[i];
				x.flag = head ? 1u : 0u;
			}
		}
		e[k] = x;
	}
}

// All six sweeps of a tree level (3 axis orders x forward/backward) in one launch: blockIdx.x = tile * 6 + sweep, so the six
// independent tile chains advance concurrently and hide each other's look-back latency.
static __global__ void __launch_bounds__(SA_THREADS, CLODB_SA_MINBLOCKS) k_sa_chained(const Box* __restrict__ boxes, SweepArgs sw, const u32* __restrict__ node_of_pos,
    u32 T, u32 tiles, char* chain_desc, u32 epoch)
{
	__shared__ ScanElem smem[SA_THREADS / 32];
	__shared__ ScanElem s_prefix;
	const u32 sweep = blockIdx.x % 6, tile = blockIdx.x / 6;
	const bool backward = (sweep & 1) != 0;
	const u32* __restrict__ order = sw.order[sweep >> 1];
	float* out_area = sw.area + size_t(sweep) * (size_t(T) + 1);
	const u32 base = tile * SA_TILE + threadIdx.x * SA_ITEMS;
	const u32 Tpad = (T + u32(SA_ITEMS) - 1) / SA_ITEMS * SA_ITEMS;
	const long q = backward ? long(Tpad) - SA_ITEMS - long(base) : long(base);
	ScanElem e[SA_ITEMS];
	if (backward)
		sa_load_chunk<true>(boxes, order, node_of_pos, T, q, e);
	else
		sa_load_chunk<false>(boxes, order, node_of_pos, T, q, e);
	ScanElem agg = scan_identity();
#pragma unroll
	for (int k = 0; k < SA_ITEMS; ++k)
		agg = scan_combine(agg, e[k]);
	ScanElem total;
	ScanElem ex = block_scan_exclusive(agg, smem, &total);
	if (threadIdx.x < 32)
	{
		size_t region = size_t(sweep) * tiles;
		ScanElem prefix = sa_lookback(tile, total, chain_desc + region * SCAN_CHAIN_VALUE_BYTES, epoch);
		if (threadIdx.x == 0)
			s_prefix = prefix;
	}
	__syncthreads();
	ScanElem prefix = scan_combine(s_prefix, ex);
#pragma unroll
	for (int k = 0; k < SA_ITEMS; ++k)
	{
		prefix = scan_combine(prefix, e[k]);
		long p = q + (backward ? SA_ITEMS - 1 - k : k);
		if (p >= 0 && p < long(T))
		{
			float sx = prefix.mx[0] - prefix.mn[0], sy = prefix.mx[1] - prefix.mn[1], sz = prefix.mx[2] - prefix.mn[2];
			out_area[p] = (sx * sy + sy * sz) + sz * sx;
		}
	}
}

// areas: 6 arrays of T + 1 floats, [axis * 2 + backward]
static void seg_area_scan_all(const Box* boxes, u32* const* order, const u32* node_of_pos, const u32* node_begin, const u32* node_count, float* areas, u32 T, Arena&)
{
	u32 tiles = (((T + 7u) & ~7u) + SA_TILE - 1) / SA_TILE;
	scan_chain_reserve(size_t(tiles) * 6);
	u32 epoch = scan_chain_next_epoch();
	SweepArgs sw;
	for (int k = 0; k < 3; ++k)
		sw.order[k] = order[k];
	sw.area = areas;
	LAUNCH_GRID(k_sa_chained, size_t(tiles) * 6, SA_THREADS, boxes, sw, node_of_pos, T, tiles, g_scan_chain.box_desc, epoch);
}
#endif

// (cost, axis, index) argmin per large node (bvhPivot over count > max_triangles, vertices == NULL)
DEVFN u64 pivot_key_at(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_count, const float* __restrict__ larea, const float* __restrict__ rarea, int axis, size_t p,
    const SplitParams& sp, u32* node_out)
{
	const u64 none = ~u64(0);
	u32 n = node_of_pos[p];
	if (n == NODE_DONE)
		return none;
	u32 count = node_count[n];
	if (count <= sp.max_triangles)
		return none;
	u32 i = u32(p) - node_begin[n];
	u32 mn = sp.min_triangles;
	bool aligned = count >= mn * 2 && bvh_divisible(count, mn, sp.max_triangles);
	u32 end = aligned ? count - mn : count - 1;
	if (i < mn - 1 || i >= end)
		return none;
	float cost;
	if (!pivot_cost(i, count, mn, sp.max_triangles, aligned, larea[p], rarea[p + 1], 0, false, sp.fill_weight, sp.max_triangles, &cost))
		return none;
	if (!(cost < FLT_MAX) || cost < 0.f)
		return none;
	*node_out = n;
	return (u64(__float_as_uint(cost)) << 32) | (u64(axis) << 30) | u64(i);
}

#ifdef CLODB_EMU
KERNEL k_pivot_large(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_count, const float* __restrict__ areas, u64* node_best, u32 T, SplitParams sp)
{
	size_t gi = GTID;
	if (gi >= size_t(T) * 3)
		return;
	int axis = int(gi / T);
	size_t p = gi - size_t(axis) * T;
	u32 n = 0;
	u64 key = pivot_key_at(node_of_pos, node_begin, node_count, areas + size_t(axis * 2) * (size_t(T) + 1), areas + size_t(axis * 2 + 1) * (size_t(T) + 1), axis, p, sp, &n);
	if (key != ~u64(0) && key < node_best[n])
		node_best[n] = key;
}
#else
// Four consecutive positions per thread. Large nodes are long runs of positions, so nearly every warp sits inside one node:
// its 128 keys are reduced with two redux operations and cost one atomicMin. Warps that straddle nodes fall back to a
// match_any reduction per element.
static const int PV_ITEMS = 4;
static __global__ void __launch_bounds__(256) k_pivot_large(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_count, const float* __restrict__ areas, u64* node_best, u32 T,
    SplitParams sp)
{
	const u64 none = ~u64(0);
	const size_t tid = GTID;
	const size_t Tq = (size_t(T) + PV_ITEMS - 1) / PV_ITEMS;
	const u32 lane = threadIdx.x & 31;
	u64 key[PV_ITEMS];
	u32 node[PV_ITEMS];
	u64 best = none;
	u32 nref = NODE_DONE;
	bool same = true;
	if (tid < Tq * 3)
	{
		int axis = int(tid / Tq);
		size_t p0 = (tid - size_t(axis) * Tq) * PV_ITEMS;
		const float* larea = areas + size_t(axis * 2) * (size_t(T) + 1);
		const float* rarea = areas + size_t(axis * 2 + 1) * (size_t(T) + 1);
#pragma unroll
		for (int j = 0; j < PV_ITEMS; ++j)
		{
			key[j] = none;
			node[j] = NODE_DONE;
			if (p0 + j < T)
				key[j] = pivot_key_at(node_of_pos, node_begin, node_count, larea, rarea, axis, p0 + j, sp, &node[j]);
			if (key[j] != none)
			{
				if (nref == NODE_DONE)
					nref = node[j];
				same = same && node[j] == nref;
				best = key[j] < best ? key[j] : best;
			}
		}
	}
	else
	{
#pragma unroll
		for (int j = 0; j < PV_ITEMS; ++j)
			key[j] = none, node[j] = NODE_DONE;
	}
	unsigned have = __ballot_sync(0xffffffffu, nref != NODE_DONE);
	if (!have)
		return;
	u32 warp_ref = __shfl_sync(0xffffffffu, nref, __ffs(have) - 1);
	if (__all_sync(0xffffffffu, nref == NODE_DONE || (same && nref == warp_ref)))
	{
		u32 hi = u32(best >> 32), lo = u32(best);
		u32 mhi = __reduce_min_sync(0xffffffffu, hi);
		u32 mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
		if (lane == 0)
			atomicMin(reinterpret_cast<unsigned long long*>(&node_best[warp_ref]), (unsigned long long)((u64(mhi) << 32) | mlo));
		return;
	}
#pragma unroll
	for (int j = 0; j < PV_ITEMS; ++j)
	{
		unsigned active = __ballot_sync(0xffffffffu, key[j] != none);
		if (key[j] == none)
			continue;
		unsigned peers = __match_any_sync(active, node[j]);
		u64 m = key[j];
		for (int d = 16; d >= 1; d >>= 1)
		{
			u64 other = __shfl_xor_sync(active, m, d);
			unsigned src = lane ^ d;
			if (((peers >> src) & 1u) && other < m)
				m = other;
		}
		// after a butterfly restricted to peers the minimum may not have reached every peer; every lane whose key equals its
		// own reduction result publishes (same address, at most a few per warp)
		if (m == key[j])
			atomicMin(reinterpret_cast<unsigned long long*>(&node_best[node[j]]), (unsigned long long)key[j]);
	}
}
#endif

KERNEL k_reset_best(u64* node_best, u32 n_nodes)
{
	size_t n = GTID;
	if (n >= n_nodes)
		return;
	node_best[n] = ~u64(0);
}

// side flag per triangle for split nodes, read along the node's best axis (clusterizer.cpp:1098-1104)
KERNEL k_mark_sides(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_split, const u64* __restrict__ node_best,
    const u32* __restrict__ order0, const u32* __restrict__ order1, const u32* __restrict__ order2, u8* side, u32 T)
{
	size_t p = GTID;
	if (p >= T)
		return;
	u32 n = node_of_pos[p];
	if (n == NODE_DONE)
		return;
	u32 split = node_split[n];
	if (split == 0)
		return;
	int axis = int((node_best[n] >> 30) & 3u);
	const u32* order = axis == 0 ? order0 : (axis == 1 ? order1 : order2);
	side[order[p]] = (u32(p) - node_begin[n]) >= split ? 1 : 0;
}

// zero-flags of all three axes laid out [axis][pos] for one global exclusive scan
KERNEL k_side_flags(const u32* __restrict__ order0, const u32* __restrict__ order1, const u32* __restrict__ order2, const u8* __restrict__ side, u32* flags, u32 T)
{
	size_t i = GTID;
	if (i >= size_t(T) * 3)
		return;
	u32 axis = u32(i / T);
	u32 p = u32(i - size_t(axis) * T);
	const u32* order = axis == 0 ? order0 : (axis == 1 ? order1 : order2);
	flags[i] = side[order[p]] ? 0u : 1u;
}

// stable two-way partition of every split node, all three axes (bvhPartition, clusterizer.cpp:1035-1049)
KERNEL k_partition(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_split,
    const u32* __restrict__ order0, const u32* __restrict__ order1, const u32* __restrict__ order2, u32* out0, u32* out1, u32* out2,
    const u8* __restrict__ side, const u32* __restrict__ zeros_before, u32 T)
{
	size_t i = GTID;
	if (i >= size_t(T) * 3)
		return;
	u32 axis = u32(i / T);
	u32 p = u32(i - size_t(axis) * T);
	const u32* order = axis == 0 ? order0 : (axis == 1 ? order1 : order2);
	u32* out = axis == 0 ? out0 : (axis == 1 ? out1 : out2);
	u32 t = order[p];
	u32 n = node_of_pos[p];
	u32 split = n == NODE_DONE ? 0 : node_split[n];
	if (split == 0)
	{
		out[p] = t;
		return;
	}
	u32 begin = node_begin[n];
	u32 zeros = zeros_before[size_t(axis) * T + p] - zeros_before[size_t(axis) * T + begin];
	u32 local = p - begin;
	u32 dst = side[t] ? begin + split + (local - zeros) : begin + zeros;
	out[dst] = t;
}

#ifndef CLODB_EMU
// k_side_flags + the 3T-element scan + k_partition in one chained pass per axis: the number of "left" triangles before a
// position inside its node is a segmented exclusive sum (restarting at node heads, found by comparing neighbouring node
// ids), so the destination is known as soon as the tile's look-back returns and the order entry is scattered directly.
// Per element: order u32 + node id u32 + side byte in, order u32 out (13 B instead of 38 B over three kernels).
struct OpSegCount
{
	static const u32 HEAD = 0x80000000u;
	DEVFN u32 identity()
	{
		return 0;
	}
	DEVFN u32 apply(u32 a, u32 b) // a precedes b
	{
		return (b & HEAD) ? b : ((a & HEAD) | ((a + b) & ~HEAD));
	}
	DEVFN bool prefix_independent(u32 a)
	{
		return (a & HEAD) != 0;
	}
};

#ifndef CLODB_PT_THREADS
#define CLODB_PT_THREADS 256
#endif
static const int PT_THREADS = CLODB_PT_THREADS;
static const int PT_ITEMS = 8;
static const int PT_TILE = PT_THREADS * PT_ITEMS;

struct PartitionArgs
{
	const u32* order[3];
	u32* out[3];
};

static __global__ void __launch_bounds__(PT_THREADS) k_partition_chained(PartitionArgs pa, const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_split,
    const u8* __restrict__ side, u32 T, u32 tiles, char* desc, u32 epoch)
{
	__shared__ u32 smem[34];
	__shared__ u32 s_prefix;
	const u32 axis = blockIdx.x % 3, tile = blockIdx.x / 3;
	const u32* __restrict__ order = pa.order[axis];
	u32* __restrict__ out = pa.out[axis];
	const size_t q = size_t(tile) * PT_TILE + size_t(threadIdx.x) * PT_ITEMS;
	u32 n[PT_ITEMS], t[PT_ITEMS], e[PT_ITEMS];
	if (q + PT_ITEMS <= T)
	{
		const uint4* np = reinterpret_cast<const uint4*>(node_of_pos + q);
		const uint4* op = reinterpret_cast<const uint4*>(order + q);
#pragma unroll
		for (int v = 0; v < PT_ITEMS / 4; ++v)
		{
			uint4 nv = __ldg(np + v), ov = __ldg(op + v);
			n[v * 4 + 0] = nv.x, n[v * 4 + 1] = nv.y, n[v * 4 + 2] = nv.z, n[v * 4 + 3] = nv.w;
			t[v * 4 + 0] = ov.x, t[v * 4 + 1] = ov.y, t[v * 4 + 2] = ov.z, t[v * 4 + 3] = ov.w;
		}
	}
	else
	{
#pragma unroll
		for (int i = 0; i < PT_ITEMS; ++i)
		{
			bool in = q + i < T;
			n[i] = in ? __ldg(node_of_pos + q + i) : NODE_DONE;
			t[i] = in ? __ldg(order + q + i) : 0u;
		}
	}
	u32 prev = (q > 0 && q < T) ? __ldg(node_of_pos + q - 1) : NODE_DONE;
	const bool first = q == 0;
	u32 agg = 0;
#pragma unroll
	for (int i = 0; i < PT_ITEMS; ++i)
	{
		bool in = q + i < T;
		u32 z = (in && !side[t[i]]) ? 1u : 0u;
		bool head = in && ((i == 0 ? (first || prev != n[0]) : n[i - 1] != n[i]));
		e[i] = z | (head ? OpSegCount::HEAD : 0u);
		agg = OpSegCount::apply(agg, e[i]);
	}
	u32 total;
	u32 ex = block_exclusive_scan<u32, OpSegCount>(agg, &total, smem);
	if (threadIdx.x < 32)
	{
		u32 prefix = scan_chain_lookback_packed<u32, OpSegCount>(tile, total, desc + (size_t(axis) * tiles) * 16, epoch);
		if (threadIdx.x == 0)
			s_prefix = prefix;
	}
	__syncthreads();
	u32 run = OpSegCount::apply(s_prefix, ex);
#pragma unroll
	for (int i = 0; i < PT_ITEMS; ++i)
	{
		size_t p = q + i;
		u32 zeros = (e[i] & OpSegCount::HEAD) ? 0u : (run & ~OpSegCount::HEAD); // left triangles of the node before p
		run = OpSegCount::apply(run, e[i]);
		if (p >= T)
			continue;
		u32 split = n[i] == NODE_DONE ? 0u : node_split[n[i]];
		if (split == 0)
		{
			out[p] = t[i];
			continue;
		}
		u32 begin = node_begin[n[i]];
		u32 local = u32(p) - begin;
		u32 dst = (e[i] & 1u) ? begin + zeros : begin + split + (local - zeros);
		out[dst] = t[i];
	}
}
#endif

// children numbering input; also tells (any_large) whether some child will still be above the meshlet size, i.e. whether the next
// tree level needs the SAH sweeps: known as soon as the splits are, so it travels to the host with the split count in one read
KERNEL k_split_flags(const u32* __restrict__ node_split, const u32* __restrict__ node_count, u32* flags, u32 n_nodes, u32 max_triangles, u32* any_large)
{
	size_t n = GTID;
	if (n >= n_nodes)
		return;
	u32 split = node_split[n];
	flags[n] = split ? 1u : 0u;
	if (split && (split > max_triangles || node_count[n] - split > max_triangles))
		atomicOr(any_large, 1u);
}

KERNEL k_make_children(const u32* __restrict__ node_begin, const u32* __restrict__ node_count, const u32* __restrict__ node_split, const u32* __restrict__ child_rank,
    u32* new_begin, u32* new_count, u32 n_nodes, u32 max_triangles, u32* any_large)
{
	size_t n = GTID;
	if (n >= n_nodes)
		return;
	u32 split = node_split[n];
	if (!split)
		return;
	u32 c = child_rank[n] * 2;
	new_begin[c] = node_begin[n];
	new_count[c] = split;
	new_begin[c + 1] = node_begin[n] + split;
	new_count[c + 1] = node_count[n] - split;
	if (split > max_triangles || node_count[n] - split > max_triangles)
		atomicOr(any_large, 1u);
}

KERNEL k_update_node_of_pos(const u32* __restrict__ node_of_pos, const u32* __restrict__ node_begin, const u32* __restrict__ node_split, const u32* __restrict__ child_rank, u32* new_node_of_pos, u32 T)
{
	size_t p = GTID;
	if (p >= T)
		return;
	u32 n = node_of_pos[p];
	u32 out = NODE_DONE;
	if (n != NODE_DONE)
	{
		u32 split = node_split[n];
		if (split)
			out = child_rank[n] * 2 + ((u32(p) - node_begin[n]) >= split ? 1 : 0);
	}
	new_node_of_pos[p] = out;
}

// ---------------------------------------------------------------------------------------------------------------------
// meshlet assembly + meshopt_optimizeMeshlet, one thread per cluster

KERNEL k_boundary_flags(const u8* __restrict__ boundary, u32* flags, u32 T)
{
	size_t i = GTID;
	if (i >= T)
		return;
	flags[i] = boundary[i] ? 1u : 0u;
}

KERNEL k_cluster_starts(const u8* __restrict__ boundary, const u32* __restrict__ cluster_rank, u32* cluster_tri_offset, u32 T, u32 K)
{
	size_t i = GTID;
	if (i > T)
		return;
	if (i == T)
	{
		cluster_tri_offset[K] = T;
		return;
	}
	if (boundary[i])
		cluster_tri_offset[cluster_rank[i]] = u32(i);
}

// Segments whose split produced more meshlets than meshopt_buildMeshletsBound allows: the reference then ignores
// boundary marks while it is over budget and lets appendMeshlet cut on the vertex/triangle limits
// (clusterizer.cpp:1441-1466, 362-411). Rare (triangle soups); replayed sequentially by one thread per segment.
KERNEL k_fix_overflow_segments(const u32* __restrict__ seg_offsets, u32 S, const u32* __restrict__ rank, u32 T, u32 K, const u32* __restrict__ order0, const u32* __restrict__ tri, u8* boundary, SplitParams sp, u32* fixed_flag)
{
	size_t s = GTID;
	if (s >= S)
		return;
	u32 begin = seg_offsets[s], end = seg_offsets[s + 1];
	u32 meshlet_count = (end == T ? K : rank[end]) - rank[begin];
	u32 index_count = (end - begin) * 3;
	// clod::clusterize passes min_triangles as the bound's max_triangles (clusterlod.h:307)
	u32 limit_vertices = (index_count + (sp.max_vertices - 2) - 1) / (sp.max_vertices - 2);
	u32 limit_triangles = ((end - begin) + sp.min_triangles - 1) / sp.min_triangles;
	u32 bound = limit_vertices > limit_triangles ? limit_vertices : limit_triangles;
	if (meshlet_count <= bound)
		return;
	*fixed_flag = 1;

	VertexSet set;
	VertexSet::clear(set);
	u32 vertex_count = 0, triangle_count = 0;
	u32 meshlet_offset = 0, meshlet_pending = meshlet_count;
	for (u32 i = begin; i < end; ++i)
	{
		u32 b = boundary[i];
		bool split = i > begin && b == 1;
		if (split && meshlet_offset + meshlet_pending >= bound)
			split = false;
		u32 t = order0[i];
		u32 v0 = tri[size_t(t) * 3 + 0], v1 = tri[size_t(t) * 3 + 1], v2 = tri[size_t(t) * 3 + 2];
		// probe without inserting: count of corners not yet in the meshlet (counted per corner, as the reference does)
		u32 extra = 0;
		for (int k = 0; k < 3; ++k)
		{
			u32 v = k == 0 ? v0 : (k == 1 ? v1 : v2);
			u32 h = (v * 0x9E3779B1u) >> 23;
			bool found = false;
			for (;;)
			{
				u32 cur = set.slots[h];
				if (cur == v)
				{
					found = true;
					break;
				}
				if (cur == 0xffffffffu)
					break;
				h = (h + 1) & 511;
			}
			extra += found ? 0 : 1;
		}
		bool flush = vertex_count + extra > sp.max_vertices || triangle_count >= sp.max_triangles || split;
		if (flush)
		{
			VertexSet::clear(set);
			vertex_count = 0;
			triangle_count = 0;
			meshlet_offset++;
		}
		boundary[i] = (flush || i == begin) ? 1 : 0;
		vertex_count += VertexSet::insert(set, v0);
		vertex_count += VertexSet::insert(set, v1);
		vertex_count += VertexSet::insert(set, v2);
		triangle_count++;
		meshlet_pending -= b;
	}
}


KERNEL k_build_clusters(const u32* __restrict__ cluster_tri_offset, u32 K, const u32* __restrict__ order0, const u32* __restrict__ tri, const u32* __restrict__ seg_of_tri,
    u32* tri_out, u32* cluster_vertex_count, u32* cluster_segment, int optimize)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 begin = cluster_tri_offset[c], count = cluster_tri_offset[c + 1] - begin;

	// meshlet-local tables as appendMeshlet builds them (clusterizer.cpp:362-411): first-occurrence vertex order
	u32 vertices[128];
	u8 indices[128 * 3];
	u32 vcount = 0;
	{
		u32 keys[512];
		u8 vals[512];
		for (int i = 0; i < 512; ++i)
			keys[i] = 0xffffffffu;
		for (u32 j = 0; j < count; ++j)
		{
			u32 t = order0[begin + j];
			for (int k = 0; k < 3; ++k)
			{
				u32 v = tri[size_t(t) * 3 + k];
				u32 h = (v * 0x9E3779B1u) >> 23;
				for (;;)
				{
					if (keys[h] == v)
						break;
					if (keys[h] == 0xffffffffu)
					{
						keys[h] = v;
						vals[h] = u8(vcount);
						vertices[vcount++] = v;
						break;
					}
					h = (h + 1) & 511;
				}
				indices[j * 3 + k] = vals[h];
			}
		}
	}

	if (optimize)
	{
		// meshopt_optimizeMeshlet (clusterizer.cpp:1682-1770)
		u8 cache[128];
		for (u32 i = 0; i < vcount; ++i)
			cache[i] = 0;
		u8 cache_last = 128;
		const u8 cache_cutoff = 3;
		for (u32 i = 0; i < count; ++i)
		{
			int next = -1;
			int next_match = -1;
			for (u32 j = i; j < count; ++j)
			{
				u8 a = indices[j * 3 + 0], b = indices[j * 3 + 1], cc = indices[j * 3 + 2];
				int aok = u8(cache_last - cache[a]) < cache_cutoff;
				int bok = u8(cache_last - cache[b]) < cache_cutoff;
				int cok = u8(cache_last - cache[cc]) < cache_cutoff;
				if (aok + bok + cok > next_match)
				{
					next = int(j);
					next_match = aok + bok + cok;
					if (next_match >= 2)
						break;
				}
			}
			u8 a = indices[next * 3 + 0], b = indices[next * 3 + 1], cc = indices[next * 3 + 2];
			for (int j = next; j > int(i); --j)
			{
				indices[j * 3 + 0] = indices[(j - 1) * 3 + 0];
				indices[j * 3 + 1] = indices[(j - 1) * 3 + 1];
				indices[j * 3 + 2] = indices[(j - 1) * 3 + 2];
			}
			indices[i * 3 + 0] = a;
			indices[i * 3 + 1] = b;
			indices[i * 3 + 2] = cc;
			cache_last++;
			cache[a] = cache_last;
			cache[b] = cache_last;
			cache[cc] = cache_last;
		}
		// the vertex reorder only permutes meshlet-local ids; global ids per corner are unaffected
	}

	for (u32 j = 0; j < count * 3; ++j)
		tri_out[size_t(begin) * 3 + j] = vertices[indices[j]];
	cluster_vertex_count[c] = vcount;
	cluster_segment[c] = seg_of_tri[order0[begin]];
}

#ifndef CLODB_EMU
// Warp-cooperative version of k_build_clusters (same results): one warp per cluster, all tables in shared memory.
//   * local vertex ids in first-occurrence order: every corner inserts its vertex into a shared hash table that keeps the
//     lowest corner index per vertex; a vertex's id is the number of "first" corners before its own first corner;
//   * meshopt_optimizeMeshlet: the serial "next triangle" search becomes ballots over the still-unplaced triangles (each
//     lane owns triangles lane, lane + 32, ...), so "first triangle in remaining order with >= 2 cached vertices, else
//     first with 1, else first" is a find-first-set; the reference's memmove only preserves that remaining order.
static const int BC_WARPS = 4;
static __global__ void __launch_bounds__(BC_WARPS * 32) k_build_clusters_warp(const u32* __restrict__ cluster_tri_offset, u32 K, const u32* __restrict__ order0, const u32* __restrict__ tri, const u32* __restrict__ seg_of_tri,
    u32* tri_out, u32* cluster_vertex_count, u32* cluster_segment, int optimize)
{
	__shared__ u32 s_keys[BC_WARPS][512];
	__shared__ u32 s_min[BC_WARPS][512]; // lowest corner index per table slot, later the vertex's local id
	__shared__ u32 s_verts[BC_WARPS][128];
	__shared__ u8 s_idx[BC_WARPS][384];
	__shared__ u8 s_out[BC_WARPS][384];
	__shared__ u8 s_cache[BC_WARPS][128];

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 c = blockIdx.x * BC_WARPS + warp;
	if (c >= K)
		return;
	u32* keys = s_keys[warp];
	u32* vmin = s_min[warp];
	u32* verts = s_verts[warp];
	u8* idx = s_idx[warp];
	u8* out = s_out[warp];
	u8* cache = s_cache[warp];

	const u32 begin = cluster_tri_offset[c], count = cluster_tri_offset[c + 1] - begin;
	const u32 corners = count * 3;
	for (int i = lane; i < 512; i += 32)
	{
		keys[i] = 0xffffffffu;
		vmin[i] = 0xffffffffu;
	}
	__syncwarp();

	// ---- corners -> table slots (12 corners per lane at most)
	u32 cv[12];
	u32 cslot[12];
#pragma unroll
	for (int r = 0; r < 12; ++r)
	{
		u32 j = lane + 32 * r;
		cv[r] = 0;
		cslot[r] = 0;
		if (j < corners)
		{
			u32 t = order0[begin + j / 3];
			u32 v = tri[size_t(t) * 3 + j % 3];
			u32 h = (v * 0x9E3779B1u) >> 23;
			for (;;)
			{
				u32 old = atomicCAS(&keys[h], 0xffffffffu, v);
				if (old == 0xffffffffu || old == v)
					break;
				h = (h + 1) & 511;
			}
			atomicMin(&vmin[h], j);
			cv[r] = v;
			cslot[r] = h;
		}
	}
	__syncwarp();
	// ---- rank the first occurrences in corner order
	u32 vcount = 0;
	bool first[12];
	u32 rank[12];
#pragma unroll
	for (int r = 0; r < 12; ++r)
	{
		u32 j = lane + 32 * r;
		first[r] = j < corners && vmin[cslot[r]] == j;
		unsigned mask = __ballot_sync(0xffffffffu, first[r]);
		rank[r] = vcount + __popc(mask & ((1u << lane) - 1));
		vcount += __popc(mask);
	}
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 12; ++r)
		if (first[r])
		{
			vmin[cslot[r]] = rank[r]; // slot now holds the local id
			verts[rank[r]] = cv[r];
		}
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 12; ++r)
	{
		u32 j = lane + 32 * r;
		if (j < corners)
			idx[j] = u8(vmin[cslot[r]]);
	}
	for (int i = lane; i < 128; i += 32)
		cache[i] = 0;
	__syncwarp();

	const u8* final_idx = idx;
	if (optimize)
	{
		// meshopt_optimizeMeshlet (clusterizer.cpp:1682-1770)
		unsigned alive = 0; // bit s: triangle lane + 32 * s not placed yet
#pragma unroll
		for (int s4 = 0; s4 < 4; ++s4)
			if (u32(lane + 32 * s4) < count)
				alive |= 1u << s4;
		u8 cache_last = 128;
		for (u32 i = 0; i < count; ++i)
		{
			int match[4];
#pragma unroll
			for (int s4 = 0; s4 < 4; ++s4)
			{
				match[s4] = -1;
				if (alive & (1u << s4))
				{
					u32 t = lane + 32 * s4;
					int aok = u8(cache_last - cache[idx[t * 3 + 0]]) < 3;
					int bok = u8(cache_last - cache[idx[t * 3 + 1]]) < 3;
					int cok = u8(cache_last - cache[idx[t * 3 + 2]]) < 3;
					match[s4] = aok + bok + cok;
				}
			}
			int next = -1;
#pragma unroll
			for (int want = 2; want >= 0 && next < 0; --want)
			{
#pragma unroll
				for (int s4 = 0; s4 < 4; ++s4)
				{
					unsigned m = __ballot_sync(0xffffffffu, want == 2 ? match[s4] >= 2 : match[s4] == want);
					if (m && next < 0)
						next = s4 * 32 + (__ffs(m) - 1);
				}
			}
			u8 a = idx[next * 3 + 0], b = idx[next * 3 + 1], cc = idx[next * 3 + 2];
			if ((next & 31) == lane)
				alive &= ~(1u << (next >> 5));
			cache_last++;
			__syncwarp();
			if (lane == 0)
			{
				out[i * 3 + 0] = a;
				out[i * 3 + 1] = b;
				out[i * 3 + 2] = cc;
				cache[a] = cache_last;
				cache[b] = cache_last;
				cache[cc] = cache_last;
			}
			__syncwarp();
		}
		final_idx = out;
	}

	for (u32 j = lane; j < corners; j += 32)
		tri_out[size_t(begin) * 3 + j] = verts[final_idx[j]];
	if (lane == 0)
	{
		cluster_vertex_count[c] = vcount;
		cluster_segment[c] = seg_of_tri[order0[begin]];
	}
}
#endif

KERNEL k_gather_u32(const u32* __restrict__ src, const u32* __restrict__ index, u32* dst, u32 n)
{
	size_t i = GTID;
	if (i >= n)
		return;
	dst[i] = src[index[i]];
}

// ---------------------------------------------------------------------------------------------------------------------
ClusterSet clusterize(const u32* tri, u32 T, const u32* seg_offsets_host, u32 S, const float* positions, const Config& config, Workspace& ws)
{
	ClusterSet result;
	result.triangle_count = T;
	if (T == 0 || S == 0)
		return result;

	Arena& temp = ws.temp;
	ArenaScope scope(temp);

	SplitParams sp;
	sp.max_vertices = config.max_vertices;
	sp.min_triangles = config.min_triangles;
	sp.max_triangles = config.max_triangles;
	sp.fill_weight = config.cluster_fill_weight;
	if (sp.max_triangles > 128 || sp.max_vertices > 128)
		throw Error("clodb200: clusterize supports max_triangles/max_vertices <= 128");

	for (u32 s = 0; s < S; ++s)
		if (seg_offsets_host[s + 1] <= seg_offsets_host[s])
			throw Error("clodb200: clusterize segments must be non-empty");

	u32* seg_offsets = temp.alloc<u32>(S + 1);
	dev_h2d(seg_offsets, seg_offsets_host, (S + 1) * sizeof(u32));

	Box* boxes = temp.alloc<Box>(T);
	u32* order[3];
	u32* order_alt[3];
	for (int k = 0; k < 3; ++k)
	{
		order[k] = temp.alloc<u32>(T);
		order_alt[k] = temp.alloc<u32>(T);
	}
	u32* seg_of_tri = temp.alloc<u32>(T);
	LAUNCH(k_segment_of_tri, T, seg_offsets, S, seg_of_tri, T);

	// ---- axis orders
	{
		ArenaScope sort_scope(temp);
		u32* key32[3];
		for (int k = 0; k < 3; ++k)
			key32[k] = temp.alloc<u32>(T);
		LAUNCH(k_prepare, T, tri, positions, T, boxes, key32[0], key32[1], key32[2]);
		int seg_bits = S > 1 ? bits_for(S - 1) : 0;
		if (seg_bits + 30 <= 32)
		{
			u32* key_tmp = temp.alloc<u32>(T);
			for (int k = 0; k < 3; ++k)
			{
				iota(order[k], T);
				radix_sort_pairs<u32>(key32[k], key_tmp, order[k], order_alt[k], T, 0, 30, temp);
			}
			if (seg_bits)
			{
				// stable pass on the segment id restores segment-major order (segments were contiguous on input)
				u32* seg_key = temp.alloc<u32>(T);
				u32* seg_tmp = temp.alloc<u32>(T);
				for (int k = 0; k < 3; ++k)
				{
					LAUNCH(k_gather_u32, T, seg_of_tri, order[k], seg_key, T);
					radix_sort_pairs<u32>(seg_key, seg_tmp, order[k], order_alt[k], T, 0, seg_bits, temp);
				}
			}
		}
		else
		{
			u64* key64 = temp.alloc<u64>(T);
			u64* key64_tmp = temp.alloc<u64>(T);
			for (int k = 0; k < 3; ++k)
			{
				LAUNCH(k_widen_keys, T, key32[k], seg_of_tri, key64, T);
				iota(order[k], T);
				radix_sort_pairs<u64>(key64, key64_tmp, order[k], order_alt[k], T, 0, 30 + seg_bits, temp);
			}
		}
	}

	// ---- breadth-first splitting
	u32 node_cap = T / 32 + S + 2;
	u32* node_begin = temp.alloc<u32>(node_cap);
	u32* node_count = temp.alloc<u32>(node_cap);
	u32* node_begin_alt = temp.alloc<u32>(node_cap);
	u32* node_count_alt = temp.alloc<u32>(node_cap);
	u64* node_best = temp.alloc<u64>(node_cap);
	u32* node_split = temp.alloc<u32>(node_cap);
	u32* node_flags = temp.alloc<u32>(node_cap);
	u32* node_total_dev = temp.alloc<u32>(4);
	u8* node_pending = temp.alloc<u8>(node_cap);
	u32* node_of_pos = temp.alloc<u32>(T);
	u32* node_of_pos_alt = temp.alloc<u32>(T);
	u8* boundary = temp.alloc<u8>(T);
	u8* side = temp.alloc<u8>(T);
	float* areas = temp.alloc<float>((size_t(T) + 1) * 6); // prefix / suffix surface areas of the three axis orders
	u32* flags3 = temp.alloc<u32>(size_t(T) * 3);

	LAUNCH(k_init_nodes, S, seg_offsets, S, node_begin, node_count, node_of_pos);
	// positions are segment-major after the sort, so position p belongs to the segment of any triangle stored there
	LAUNCH(k_gather_u32, T, seg_of_tri, order[0], node_of_pos, T);
	dev_memset(boundary, 0, T);

	u32 n_nodes = S;
	u32 max_count = 0;
	for (u32 s = 0; s < S; ++s)
		max_count = std::max(max_count, seg_offsets_host[s + 1] - seg_offsets_host[s]);
	bool any_large = max_count > sp.max_triangles;

	for (int depth = 0; n_nodes > 0; ++depth)
	{
		if (n_nodes > node_cap)
			throw Error("clodb200: clusterize node capacity exceeded");

		LAUNCH(k_reset_best, n_nodes, node_best, n_nodes);
		dev_memset(node_split, 0, size_t(n_nodes) * sizeof(u32));

		if (any_large)
		{
			seg_area_scan_all(boxes, order, node_of_pos, node_begin, node_count, areas, T, temp);
#ifdef CLODB_EMU
			LAUNCH(k_pivot_large, size_t(T) * 3, node_of_pos, node_begin, node_count, areas, node_best, T, sp);
#else
			LAUNCH(k_pivot_large, ((size_t(T) + PV_ITEMS - 1) / PV_ITEMS) * 3, node_of_pos, node_begin, node_count, areas, node_best, T, sp);
#endif
			LAUNCH(k_resolve_nodes, n_nodes, node_begin, node_count, node_best, node_split, n_nodes, depth, order[0], order[1], order[2], boxes, tri, boundary, sp, 1, nullptr);
		}
#ifdef CLODB_EMU
		LAUNCH(k_resolve_nodes, n_nodes, node_begin, node_count, node_best, node_split, n_nodes, depth, order[0], order[1], order[2], boxes, tri, boundary, sp, 0, nullptr);
#else
		LAUNCH_GRID(k_leaf_test_warp, (n_nodes + LT_WARPS - 1) / LT_WARPS, LT_WARPS * 32, node_begin, node_count, node_best, node_split, n_nodes, order[0], tri, boundary, sp, node_pending);
		LAUNCH(k_resolve_nodes, n_nodes, node_begin, node_count, node_best, node_split, n_nodes, depth, order[0], order[1], order[2], boxes, tri, boundary, sp, 0, node_pending);
#endif

		// children numbering
		dev_memset(node_total_dev + 1, 0, sizeof(u32));
		LAUNCH(k_split_flags, n_nodes, node_split, node_count, node_flags, n_nodes, sp.max_triangles, node_total_dev + 1);
		exclusive_scan_u32(node_flags, node_flags, n_nodes, node_total_dev, temp);
		// one read-back per tree level: {number of nodes that split, any child still above the meshlet size}
		std::vector<u32> level_words = dev_download(node_total_dev, 2);
		u32 n_split = level_words[0];
		if (n_split == 0)
			break;

		LAUNCH(k_mark_sides, T, node_of_pos, node_begin, node_split, node_best, order[0], order[1], order[2], side, T);
#ifdef CLODB_EMU
		LAUNCH(k_side_flags, size_t(T) * 3, order[0], order[1], order[2], side, flags3, T);
		exclusive_scan_u32(flags3, flags3, size_t(T) * 3, nullptr, temp);
		LAUNCH(k_partition, size_t(T) * 3, node_of_pos, node_begin, node_split, order[0], order[1], order[2], order_alt[0], order_alt[1], order_alt[2], side, flags3, T);
#else
		{
			u32 tiles = (T + PT_TILE - 1) / PT_TILE;
			u32 epoch = scan_chain_next_epoch();
			scan_chain_reserve(size_t(tiles) * 3);
			PartitionArgs pa;
			for (int k = 0; k < 3; ++k)
				pa.order[k] = order[k], pa.out[k] = order_alt[k];
			LAUNCH_GRID(k_partition_chained, size_t(tiles) * 3, PT_THREADS, pa, node_of_pos, node_begin, node_split, side, T, tiles, g_scan_chain.desc, epoch);
		}
#endif
		LAUNCH(k_make_children, n_nodes, node_begin, node_count, node_split, node_flags, node_begin_alt, node_count_alt, n_nodes, sp.max_triangles, node_total_dev + 2);
		LAUNCH(k_update_node_of_pos, T, node_of_pos, node_begin, node_split, node_flags, node_of_pos_alt, T);

		for (int k = 0; k < 3; ++k)
			std::swap(order[k], order_alt[k]);
		std::swap(node_begin, node_begin_alt);
		std::swap(node_count, node_count_alt);
		std::swap(node_of_pos, node_of_pos_alt);
		n_nodes = n_split * 2;
		any_large = level_words[1] != 0;
		if (depth > kMeshletMaxTreeDepth + 2)
			throw Error("clodb200: clusterize exceeded the tree depth limit");
	}

	// ---- meshlets
	u32* rank = flags3; // reuse
	LAUNCH(k_boundary_flags, T, boundary, rank, T);
	exclusive_scan_u32(rank, rank, T, node_total_dev, temp);
	u32 K = dev_read(node_total_dev);
	{
		dev_memset(node_total_dev + 1, 0, sizeof(u32));
		LAUNCH(k_fix_overflow_segments, S, seg_offsets, S, rank, T, K, order[0], tri, boundary, sp, node_total_dev + 1);
		if (dev_read(node_total_dev + 1))
		{
			LAUNCH(k_boundary_flags, T, boundary, rank, T);
			exclusive_scan_u32(rank, rank, T, node_total_dev, temp);
			K = dev_read(node_total_dev);
		}
	}

	result.cluster_count = K;
	result.tri = ws.persist.alloc<u32>(size_t(T) * 3);
	result.cluster_tri_offset = ws.persist.alloc<u32>(size_t(K) + 1);
	result.cluster_vertex_count = ws.persist.alloc<u32>(K);
	result.cluster_segment = ws.persist.alloc<u32>(K);

	LAUNCH(k_cluster_starts, size_t(T) + 1, boundary, rank, result.cluster_tri_offset, T, K);
#ifdef CLODB_EMU
	LAUNCH(k_build_clusters, K, result.cluster_tri_offset, K, order[0], tri, seg_of_tri, result.tri, result.cluster_vertex_count, result.cluster_segment, config.optimize_clusters ? 1 : 0);
#else
	LAUNCH_GRID(k_build_clusters_warp, (K + BC_WARPS - 1) / BC_WARPS, BC_WARPS * 32, result.cluster_tri_offset, K, order[0], tri, seg_of_tri, result.tri, result.cluster_vertex_count, result.cluster_segment, config.optimize_clusters ? 1 : 0);
#endif
	return result;
}

} // namespace clodb
