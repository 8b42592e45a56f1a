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

void cluster_bounds(const u32* tri, const u32* cluster_tri_offset, u32 cluster_count, const float* positions, float* bounds4)
{
	LAUNCH(k_cluster_bounds, cluster_count, tri, cluster_tri_offset, cluster_count, positions, bounds4);
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

void group_bounds_merge(const float* cluster_bounds5, const u32* group_cluster_offset, const u32* group_clusters, u32 group_count, float* out5)
{
	LAUNCH(k_group_bounds_merge, group_count, cluster_bounds5, group_cluster_offset, group_clusters, group_count, out5);
}

} // namespace clodb
