// S7: bounding spheres.
//
// Reference semantics: meshopt_computeClusterBounds (ThirdParty/meshoptimizer/src/clusterizer.cpp:1479-1632; only the
// sphere is consumed by clod::boundsCompute, clusterlod.h:270-281), computeBoundingSphere (:176-284) and
// meshopt_computeSphereBounds (:1655-1680) as used by clod::boundsMerge (clusterlod.h:283-303).
// computeBoundingSphere's grow loop is order dependent, so each cluster / group is replayed in index order by one thread:
// same points, same order, same float operations => same bits (the 7-axis extremal search is a min/max and is order
// independent apart from first-index tie breaking, which is reproduced).
#include "clodb.h"

#include <cfloat>

namespace clodb
{

struct SphereAccum
{
	u32 pmin[7], pmax[7];
	float tmin[7], tmax[7];
};

DEVFN void axis_dot(int axis, const float* p, float* out)
{
	const float k = 0.57735026f;
	switch (axis)
	{
	case 0: *out = 1.f * p[0] + 0.f * p[1] + 0.f * p[2]; break;
	case 1: *out = 0.f * p[0] + 1.f * p[1] + 0.f * p[2]; break;
	case 2: *out = 0.f * p[0] + 0.f * p[1] + 1.f * p[2]; break;
	case 3: *out = k * p[0] + k * p[1] + k * p[2]; break;
	case 4: *out = -k * p[0] + k * p[1] + k * p[2]; break;
	case 5: *out = k * p[0] + -k * p[1] + k * p[2]; break;
	default: *out = k * p[0] + k * p[1] + -k * p[2]; break;
	}
}

// Point accessor abstraction: get(i, p[3], &r)
struct ClusterPoints
{
	const u32* tri; // cluster-major index triples
	const float* positions;
	u32 first_corner; // index of the first corner (3 * first triangle)
};

DEVFN void cluster_point(const ClusterPoints& c, const u32* corner_map, u32 i, float* p)
{
	u32 v = c.tri[size_t(c.first_corner) + corner_map[i]];
	p[0] = c.positions[size_t(v) * 3 + 0];
	p[1] = c.positions[size_t(v) * 3 + 1];
	p[2] = c.positions[size_t(v) * 3 + 2];
}

// computeBoundingSphere over `count` points fetched through get_point(ctx, i, p, &r)
template <typename Ctx, void (*GetPoint)(const Ctx&, u32, float*, float*)>
DEVFN void bounding_sphere(const Ctx& ctx, u32 count, float* result)
{
	u32 pmin[7], pmax[7];
	float tmin[7], tmax[7];
	for (int axis = 0; axis < 7; ++axis)
	{
		pmin[axis] = pmax[axis] = 0;
		tmin[axis] = FLT_MAX;
		tmax[axis] = -FLT_MAX;
	}
	for (u32 i = 0; i < count; ++i)
	{
		float p[3], r;
		GetPoint(ctx, i, p, &r);
		for (int axis = 0; axis < 7; ++axis)
		{
			float tp;
			axis_dot(axis, p, &tp);
			float tpmin = tp - r, tpmax = tp + r;
			pmin[axis] = (tpmin < tmin[axis]) ? i : pmin[axis];
			pmax[axis] = (tpmax > tmax[axis]) ? i : pmax[axis];
			tmin[axis] = (tpmin < tmin[axis]) ? tpmin : tmin[axis];
			tmax[axis] = (tpmax > tmax[axis]) ? tpmax : tmax[axis];
		}
	}

	u32 paxis = 0;
	float paxisdr = 0;
	for (int axis = 0; axis < 7; ++axis)
	{
		float p1[3], p2[3], r1, r2;
		GetPoint(ctx, pmin[axis], p1, &r1);
		GetPoint(ctx, pmax[axis], p2, &r2);
		float d2 = (p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]) + (p2[2] - p1[2]) * (p2[2] - p1[2]);
		float dr = sqrtf(d2) + r1 + r2;
		if (dr > paxisdr)
		{
			paxisdr = dr;
			paxis = u32(axis);
		}
	}

	float p1[3], p2[3], r1, r2;
	GetPoint(ctx, pmin[paxis], p1, &r1);
	GetPoint(ctx, pmax[paxis], p2, &r2);
	float paxisd = sqrtf((p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]) + (p2[2] - p1[2]) * (p2[2] - p1[2]));
	float paxisk = paxisd > 0 ? (paxisd + r2 - r1) / (2 * paxisd) : 0.f;

	float center[3] = {p1[0] + (p2[0] - p1[0]) * paxisk, p1[1] + (p2[1] - p1[1]) * paxisk, p1[2] + (p2[2] - p1[2]) * paxisk};
	float radius = paxisdr / 2;

	for (u32 i = 0; i < count; ++i)
	{
		float p[3], r;
		GetPoint(ctx, i, p, &r);
		float d2 = (p[0] - center[0]) * (p[0] - center[0]) + (p[1] - center[1]) * (p[1] - center[1]) + (p[2] - center[2]) * (p[2] - center[2]);
		float d = sqrtf(d2);
		if (d + r > radius)
		{
			float k = d > 0 ? (d + r - radius) / (2 * d) : 0.f;
			center[0] += k * (p[0] - center[0]);
			center[1] += k * (p[1] - center[1]);
			center[2] += k * (p[2] - center[2]);
			radius = (radius + d + r) / 2;
		}
	}
	result[0] = center[0];
	result[1] = center[1];
	result[2] = center[2];
	result[3] = radius;
}

struct ClusterCtx
{
	const u32* tri;
	const float* positions;
	size_t first_corner;
	const unsigned short* corner_of_point; // compacted list of corners of non-degenerate triangles
};

DEVFN void cluster_get_point(const ClusterCtx& c, u32 i, float* p, float* r)
{
	u32 v = c.tri[c.first_corner + c.corner_of_point[i]];
	p[0] = c.positions[size_t(v) * 3 + 0];
	p[1] = c.positions[size_t(v) * 3 + 1];
	p[2] = c.positions[size_t(v) * 3 + 2];
	*r = 0.f;
}

KERNEL k_cluster_bounds(const u32* __restrict__ tri, const u32* __restrict__ cluster_tri_offset, u32 K, const float* __restrict__ positions, float* bounds4)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u32 begin = cluster_tri_offset[c], count = cluster_tri_offset[c + 1] - begin;

	// degenerate (zero-area) triangles are dropped before the sphere fit (clusterizer.cpp:1497-1525)
	unsigned short corners[128 * 3];
	u32 points = 0;
	for (u32 j = 0; j < count && j < 128; ++j)
	{
		u32 a = tri[(size_t(begin) + j) * 3 + 0], b = tri[(size_t(begin) + j) * 3 + 1], cc = tri[(size_t(begin) + j) * 3 + 2];
		const float* p0 = positions + size_t(a) * 3;
		const float* p1 = positions + size_t(b) * 3;
		const float* p2 = positions + size_t(cc) * 3;
		float p10[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
		float p20[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
		float nx = p10[1] * p20[2] - p10[2] * p20[1];
		float ny = p10[2] * p20[0] - p10[0] * p20[2];
		float nz = p10[0] * p20[1] - p10[1] * p20[0];
		float area = sqrtf(nx * nx + ny * ny + nz * nz);
		if (area == 0.f)
			continue;
		corners[points++] = (unsigned short)(j * 3 + 0);
		corners[points++] = (unsigned short)(j * 3 + 1);
		corners[points++] = (unsigned short)(j * 3 + 2);
	}

	float* out = bounds4 + c * 4;
	if (points == 0)
	{
		out[0] = out[1] = out[2] = out[3] = 0.f;
		return;
	}
	ClusterCtx ctx = {tri, positions, size_t(begin) * 3, corners};
	bounding_sphere<ClusterCtx, cluster_get_point>(ctx, points, out);
}

struct GroupCtx
{
	const float* cluster_bounds5;
	const u32* members;
};

DEVFN void group_get_point(const GroupCtx& g, u32 i, float* p, float* r)
{
	const float* b = g.cluster_bounds5 + size_t(g.members[i]) * 5;
	p[0] = b[0];
	p[1] = b[1];
	p[2] = b[2];
	*r = b[3];
}

#ifndef CLODB_EMU
// ---- warp-cooperative sphere fit ---------------------------------------------------------------------------------------
// Points (x, y, z, r) are staged in shared memory in their sequential order. The 7-axis extremal search is a min/max with
// "first index wins" ties, so it is done in parallel (each lane scans a strided subset in ascending order, ties across lanes
// go to the lowest index). The grow loop is order dependent and is replayed by lane 0 over the staged points: same points,
// same order, same float operations as computeBoundingSphere (clusterizer.cpp:176-284) => same bits.
DEVFN void warp_sphere_fit(const float4* pts, u32 count, float* result)
{
	const int lane = threadIdx.x & 31;
	const float k = 0.57735026f;
	u32 pmin[7], pmax[7];
	float tmin[7], tmax[7];
	for (int axis = 0; axis < 7; ++axis)
	{
		pmin[axis] = pmax[axis] = 0xffffffffu;
		tmin[axis] = FLT_MAX;
		tmax[axis] = -FLT_MAX;
	}
	for (u32 i = lane; i < count; i += 32)
	{
		float4 q = pts[i];
		float p[3] = {q.x, q.y, q.z};
		for (int axis = 0; axis < 7; ++axis)
		{
			float tp;
			axis_dot(axis, p, &tp);
			float tpmin = tp - q.w, tpmax = tp + q.w;
			if (tpmin < tmin[axis])
			{
				tmin[axis] = tpmin;
				pmin[axis] = i;
			}
			if (tpmax > tmax[axis])
			{
				tmax[axis] = tpmax;
				pmax[axis] = i;
			}
		}
	}
	(void)k;
	for (int axis = 0; axis < 7; ++axis)
	{
		for (int d = 16; d >= 1; d >>= 1)
		{
			float ov = __shfl_xor_sync(0xffffffffu, tmin[axis], d);
			u32 oi = __shfl_xor_sync(0xffffffffu, pmin[axis], d);
			if (ov < tmin[axis] || (ov == tmin[axis] && oi < pmin[axis]))
			{
				tmin[axis] = ov;
				pmin[axis] = oi;
			}
			ov = __shfl_xor_sync(0xffffffffu, tmax[axis], d);
			oi = __shfl_xor_sync(0xffffffffu, pmax[axis], d);
			if (ov > tmax[axis] || (ov == tmax[axis] && oi < pmax[axis]))
			{
				tmax[axis] = ov;
				pmax[axis] = oi;
			}
		}
		// the sequential search starts from index 0 and only moves on a strict improvement over +-FLT_MAX
		if (pmin[axis] == 0xffffffffu || !(tmin[axis] < FLT_MAX))
			pmin[axis] = 0;
		if (pmax[axis] == 0xffffffffu || !(tmax[axis] > -FLT_MAX))
			pmax[axis] = 0;
	}
	if (lane != 0)
		return;

	u32 paxis = 0;
	float paxisdr = 0;
	for (int axis = 0; axis < 7; ++axis)
	{
		float4 a = pts[pmin[axis]], b = pts[pmax[axis]];
		float d2 = (b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y) + (b.z - a.z) * (b.z - a.z);
		float dr = sqrtf(d2) + a.w + b.w;
		if (dr > paxisdr)
		{
			paxisdr = dr;
			paxis = u32(axis);
		}
	}
	float4 a = pts[pmin[paxis]], b = pts[pmax[paxis]];
	float paxisd = sqrtf((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y) + (b.z - a.z) * (b.z - a.z));
	float paxisk = paxisd > 0 ? (paxisd + b.w - a.w) / (2 * paxisd) : 0.f;
	float center[3] = {a.x + (b.x - a.x) * paxisk, a.y + (b.y - a.y) * paxisk, a.z + (b.z - a.z) * paxisk};
	float radius = paxisdr / 2;
	for (u32 i = 0; i < count; ++i)
	{
		float4 q = pts[i];
		float d2 = (q.x - center[0]) * (q.x - center[0]) + (q.y - center[1]) * (q.y - center[1]) + (q.z - center[2]) * (q.z - center[2]);
		float d = sqrtf(d2);
		if (d + q.w > radius)
		{
			float kk = d > 0 ? (d + q.w - radius) / (2 * d) : 0.f;
			center[0] += kk * (q.x - center[0]);
			center[1] += kk * (q.y - center[1]);
			center[2] += kk * (q.z - center[2]);
			radius = (radius + d + q.w) / 2;
		}
	}
	result[0] = center[0];
	result[1] = center[1];
	result[2] = center[2];
	result[3] = radius;
}

static const int CB_WARPS = 4;
static __global__ void __launch_bounds__(CB_WARPS * 32) k_cluster_bounds_warp(const u32* __restrict__ tri, const u32* __restrict__ cluster_tri_offset, u32 K, const float* __restrict__ positions, float* bounds4)
{
	__shared__ float4 s_pts[CB_WARPS][384];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 c = blockIdx.x * CB_WARPS + warp;
	if (c >= K)
		return;
	float4* pts = s_pts[warp];
	u32 begin = cluster_tri_offset[c], count = cluster_tri_offset[c + 1] - begin;
	if (count > 128)
		count = 128;
	// stage the corners of non-degenerate triangles in order (clusterizer.cpp:1497-1525)
	u32 points = 0;
	for (u32 base = 0; base < count; base += 32)
	{
		u32 j = base + lane;
		bool keep = false;
		float p[9];
		if (j < count)
		{
			for (int k = 0; k < 3; ++k)
			{
				u32 v = tri[(size_t(begin) + j) * 3 + k];
				p[k * 3 + 0] = positions[size_t(v) * 3 + 0];
				p[k * 3 + 1] = positions[size_t(v) * 3 + 1];
				p[k * 3 + 2] = positions[size_t(v) * 3 + 2];
			}
			float p10[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]};
			float p20[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
			float nx = p10[1] * p20[2] - p10[2] * p20[1];
			float ny = p10[2] * p20[0] - p10[0] * p20[2];
			float nz = p10[0] * p20[1] - p10[1] * p20[0];
			float area = sqrtf(nx * nx + ny * ny + nz * nz);
			keep = !(area == 0.f);
		}
		unsigned mask = __ballot_sync(0xffffffffu, keep);
		if (keep)
		{
			u32 at = points + 3 * __popc(mask & ((1u << lane) - 1));
			pts[at + 0] = make_float4(p[0], p[1], p[2], 0.f);
			pts[at + 1] = make_float4(p[3], p[4], p[5], 0.f);
			pts[at + 2] = make_float4(p[6], p[7], p[8], 0.f);
		}
		points += 3 * __popc(mask);
	}
	__syncwarp();
	float* out = bounds4 + size_t(c) * 4;
	if (points == 0)
	{
		if (lane == 0)
			out[0] = out[1] = out[2] = out[3] = 0.f;
		return;
	}
	warp_sphere_fit(pts, points, out);
}

static const int GB_MAX_MEMBERS = 1024;
static __global__ void __launch_bounds__(32) k_group_bounds_merge_warp(const float* __restrict__ cluster_bounds5, const u32* __restrict__ group_cluster_offset, const u32* __restrict__ group_clusters, u32 G, float* out5)
{
	__shared__ float4 s_pts[GB_MAX_MEMBERS];
	const int lane = threadIdx.x;
	const u32 g = blockIdx.x;
	u32 begin = group_cluster_offset[g], count = group_cluster_offset[g + 1] - begin;
	float* out = out5 + size_t(g) * 5;
	if (count == 0)
	{
		if (lane == 0)
			out[0] = out[1] = out[2] = out[3] = out[4] = 0.f;
		return;
	}
	if (count > GB_MAX_MEMBERS)
	{
		// larger than any group the builder configuration produces: serial fit straight from global memory
		if (lane == 0)
		{
			GroupCtx ctx = {cluster_bounds5, group_clusters + begin};
			bounding_sphere<GroupCtx, group_get_point>(ctx, count, out);
			float e = 0.f;
			for (u32 j = 0; j < count; ++j)
			{
				float v = cluster_bounds5[size_t(group_clusters[begin + j]) * 5 + 4];
				e = e < v ? v : e;
			}
			out[4] = e;
		}
		return;
	}
	float error = 0.f;
	for (u32 j = lane; j < count; j += 32)
	{
		const float* b = cluster_bounds5 + size_t(group_clusters[begin + j]) * 5;
		s_pts[j] = make_float4(b[0], b[1], b[2], b[3]);
		error = error < b[4] ? b[4] : error; // max is order independent
	}
	for (int d = 16; d >= 1; d >>= 1)
	{
		float o = __shfl_xor_sync(0xffffffffu, error, d);
		error = error < o ? o : error;
	}
	__syncwarp();
	warp_sphere_fit(s_pts, count, out);
	if (lane == 0)
		out[4] = error;
}
#endif

KERNEL k_group_bounds_merge(const float* __restrict__ cluster_bounds5, const u32* __restrict__ group_cluster_offset, const u32* __restrict__ group_clusters, u32 G, float* out5)
{
	size_t g = GTID;
	if (g >= G)
		return;
	u32 begin = group_cluster_offset[g], count = group_cluster_offset[g + 1] - begin;
	float* out = out5 + g * 5;
	if (count == 0)
	{
		out[0] = out[1] = out[2] = out[3] = out[4] = 0.f;
		return;
	}
	GroupCtx ctx = {cluster_bounds5, group_clusters + begin};
	bounding_sphere<GroupCtx, group_get_point>(ctx, count, out);
	float error = 0.f;
	for (u32 j = 0; j < count; ++j)
	{
		float e = cluster_bounds5[size_t(group_clusters[begin + j]) * 5 + 4];
		error = error < e ? e : error; // std::max(result.error, e), clusterlod.h:300
	}
	out[4] = error;
}

void cluster_bounds(const u32* tri, const u32* cluster_tri_offset, u32 cluster_count, const float* positions, float* bounds4)
{
#ifdef CLODB_EMU
	LAUNCH(k_cluster_bounds, cluster_count, tri, cluster_tri_offset, cluster_count, positions, bounds4);
#else
	LAUNCH_GRID(k_cluster_bounds_warp, (cluster_count + CB_WARPS - 1) / CB_WARPS, CB_WARPS * 32, tri, cluster_tri_offset, cluster_count, positions, bounds4);
#endif
}

void group_bounds_merge(const float* cluster_bounds5, const u32* group_cluster_offset, const u32* group_clusters, u32 group_count, float* out5)
{
#ifdef CLODB_EMU
	LAUNCH(k_group_bounds_merge, group_count, cluster_bounds5, group_cluster_offset, group_clusters, group_count, out5);
#else
	// groups hold at most partition_size + max cluster count of a partition (512 on the builder path); larger ones take
	// the serial kernel
	LAUNCH_GRID(k_group_bounds_merge_warp, group_count, 32, cluster_bounds5, group_cluster_offset, group_clusters, group_count, out5);
#endif
}

} // namespace clodb
