// S6: boundary-locked, attribute-aware quadric edge-collapse simplification of ALL groups of a DAG level at once.
//
// Reference semantics: clod::simplify (clusterlod.h:601-659) -> meshopt_simplifyWithAttributes == meshopt_simplifyEdge
// (ThirdParty/meshoptimizer/src/simplifier.cpp:2340-2623) with options Sparse | ErrorAbsolute | Permissive, one call per
// group. Callees followed: buildSparseRemap :242-294, updateEdgeAdjacency :52-109, buildPositionRemap :201-240,
// classifyVertices :364-542, rescalePositions :550-606, rescaleAttributes :608-624, quadric math :716-1056,
// fill*Quadrics :1057-1175, hasTriangleFlips :1196-1234, getComplexTarget :1285-1298, pickEdgeCollapses :1322-1392,
// rankEdgeCollapses :1394-1470, sortEdgeCollapses :1472-1517, performEdgeCollapses :1519-1656, updateQuadrics :1658-1698,
// remapIndexBuffer :1807-1839, remapEdgeLoops :1841-1860.
//
// B200 formulation: the groups of a level are independent meshes. They are concatenated into flat arrays ("sparse"
// vertices numbered in first-occurrence order per group, exactly the ids the reference's Sparse mode would assign, plus a
// per-group base) and every step of the reference's pass structure becomes a grid-wide kernel over all groups:
//   * all per-vertex float accumulations (quadrics) are gathers over adjacency lists sorted by corner index, i.e. the same
//     summation order as the reference's triangle-order scatter => same bits;
//   * the 12-bit counting sort is a stable radix sort on (group, key);
//   * the sequential greedy collapse scan is resolved as a lexicographically-first dependency wavefront: a candidate is
//     decided once every lower-ranked candidate that could change its lock state or its flip test is decided, which gives
//     the same accepted set as the serial scan; the serial scan's early-out conditions are prefix sums over the decided
//     sequence and are applied as a per-group cut afterwards.
#include "clodb.h"

#ifndef CLODB_EMU
#include <cooperative_groups.h>
#endif
#include <cfloat>
#include <cmath>
#include <algorithm>

namespace clodb
{

static const u32 NONE = 0xffffffffu;

// upper bound on dependency-wavefront rounds per pass; candidates still undecided after that are deferred to the next pass
static u32 config_max_rounds()
{
	static u32 value = 0;
	if (!value)
	{
		const char* e = getenv("CLODB200_MAX_ROUNDS");
		value = e ? u32(atoi(e)) : 4096u;
		if (!value)
			value = 1;
	}
	return value;
}

enum VertexKind
{
	Kind_Manifold,
	Kind_Border,
	Kind_Seam,
	Kind_Complex,
	Kind_Locked,
	Kind_Count
};

CONSTANT unsigned char kCanCollapse[Kind_Count][Kind_Count] = {
    {1, 1, 1, 1, 1},
    {0, 1, 0, 0, 1},
    {0, 0, 1, 0, 1},
    {0, 0, 0, 1, 1},
    {0, 0, 0, 0, 0},
};

CONSTANT unsigned char kHasOpposite[Kind_Count][Kind_Count] = {
    {1, 1, 1, 1, 1},
    {1, 0, 1, 0, 0},
    {1, 1, 1, 0, 1},
    {1, 0, 0, 0, 0},
    {1, 0, 1, 0, 0},
};

struct Vector3
{
	float x, y, z;
};

struct Quadric
{
	float a00, a11, a22;
	float a10, a20, a21;
	float b0, b1, b2, c;
	float w;
};

struct QuadricGrad
{
	float gx, gy, gz, gw;
};

enum CandidateStatus
{
	Status_Undecided = 0,
	Status_Performed = 1,
	Status_Flip = 2,
	Status_Locked = 3,
};

// ------------------------------------------------------------------------------------------------- quadric math
DEVFN float normalize3(Vector3& v)
{
	float length = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	if (length > 0)
	{
		v.x /= length;
		v.y /= length;
		v.z /= length;
	}
	return length;
}

DEVFN void quadric_zero(Quadric& Q)
{
	Q.a00 = Q.a11 = Q.a22 = Q.a10 = Q.a20 = Q.a21 = Q.b0 = Q.b1 = Q.b2 = Q.c = Q.w = 0.f;
}

DEVFN void quadric_add(Quadric& Q, const Quadric& R)
{
	Q.a00 += R.a00;
	Q.a11 += R.a11;
	Q.a22 += R.a22;
	Q.a10 += R.a10;
	Q.a20 += R.a20;
	Q.a21 += R.a21;
	Q.b0 += R.b0;
	Q.b1 += R.b1;
	Q.b2 += R.b2;
	Q.c += R.c;
	Q.w += R.w;
}

DEVFN float quadric_eval(const Quadric& Q, const Vector3& v)
{
	float rx = Q.b0;
	float ry = Q.b1;
	float rz = Q.b2;
	rx += Q.a10 * v.y;
	ry += Q.a21 * v.z;
	rz += Q.a20 * v.x;
	rx *= 2;
	ry *= 2;
	rz *= 2;
	rx += Q.a00 * v.x;
	ry += Q.a11 * v.y;
	rz += Q.a22 * v.z;
	float r = Q.c;
	r += rx * v.x;
	r += ry * v.y;
	r += rz * v.z;
	return r;
}

DEVFN float quadric_error(const Quadric& Q, const Vector3& v)
{
	float r = quadric_eval(Q, v);
	float s = Q.w == 0.f ? 0.f : 1.f / Q.w;
	return fabsf(r) * s;
}

DEVFN float quadric_error_attr(const Quadric& Q, const QuadricGrad* G, u32 attribute_count, const Vector3& v, const float* va)
{
	float r = quadric_eval(Q, v);
	for (u32 k = 0; k < attribute_count; ++k)
	{
		float a = va[k];
		float g = v.x * G[k].gx + v.y * G[k].gy + v.z * G[k].gz + G[k].gw;
		r += a * (a * Q.w - 2 * g);
	}
	return fabsf(r);
}

DEVFN void quadric_from_plane(Quadric& Q, float a, float b, float c, float d, float w)
{
	float aw = a * w;
	float bw = b * w;
	float cw = c * w;
	float dw = d * w;
	Q.a00 = a * aw;
	Q.a11 = b * bw;
	Q.a22 = c * cw;
	Q.a10 = a * bw;
	Q.a20 = a * cw;
	Q.a21 = b * cw;
	Q.b0 = a * dw;
	Q.b1 = b * dw;
	Q.b2 = c * dw;
	Q.c = d * dw;
	Q.w = w;
}

DEVFN void quadric_from_point(Quadric& Q, float x, float y, float z, float w)
{
	Q.a00 = Q.a11 = Q.a22 = w;
	Q.a10 = Q.a20 = Q.a21 = 0;
	Q.b0 = -x * w;
	Q.b1 = -y * w;
	Q.b2 = -z * w;
	Q.c = (x * x + y * y + z * z) * w;
	Q.w = w;
}

DEVFN void quadric_from_triangle(Quadric& Q, const Vector3& p0, const Vector3& p1, const Vector3& p2, float weight)
{
	Vector3 p10 = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z};
	Vector3 p20 = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
	Vector3 normal = {p10.y * p20.z - p10.z * p20.y, p10.z * p20.x - p10.x * p20.z, p10.x * p20.y - p10.y * p20.x};
	float area = normalize3(normal);
	float distance = normal.x * p0.x + normal.y * p0.y + normal.z * p0.z;
	quadric_from_plane(Q, normal.x, normal.y, normal.z, -distance, sqrtf(area) * weight);
}

DEVFN void quadric_from_triangle_edge(Quadric& Q, const Vector3& p0, const Vector3& p1, const Vector3& p2, float weight)
{
	Vector3 p10 = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z};
	float lengthsq = p10.x * p10.x + p10.y * p10.y + p10.z * p10.z;
	float length = sqrtf(lengthsq);
	Vector3 p20 = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
	float p20p = p20.x * p10.x + p20.y * p10.y + p20.z * p10.z;
	Vector3 perp = {p20.x * lengthsq - p10.x * p20p, p20.y * lengthsq - p10.y * p20p, p20.z * lengthsq - p10.z * p20p};
	normalize3(perp);
	float distance = perp.x * p0.x + perp.y * p0.y + perp.z * p0.z;
	quadric_from_plane(Q, perp.x, perp.y, perp.z, -distance, length * weight);
}

// quadricFromAttributes (simplifier.cpp:905-985); G must hold attribute_count entries
DEVFN void quadric_from_attributes(Quadric& Q, QuadricGrad* G, const Vector3& p0, const Vector3& p1, const Vector3& p2, const float* va0, const float* va1, const float* va2, u32 attribute_count)
{
	Vector3 p10 = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z};
	Vector3 p20 = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
	Vector3 normal = {p10.y * p20.z - p10.z * p20.y, p10.z * p20.x - p10.x * p20.z, p10.x * p20.y - p10.y * p20.x};
	float area = sqrtf(normal.x * normal.x + normal.y * normal.y + normal.z * normal.z) * 0.5f;
	float w = area;
	const Vector3& v0 = p10;
	const Vector3& v1 = p20;
	float d00 = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z;
	float d01 = v0.x * v1.x + v0.y * v1.y + v0.z * v1.z;
	float d11 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
	float denom = d00 * d11 - d01 * d01;
	float denomr = denom == 0 ? 0.f : 1.f / denom;
	float gx1 = (d11 * v0.x - d01 * v1.x) * denomr;
	float gx2 = (d00 * v1.x - d01 * v0.x) * denomr;
	float gy1 = (d11 * v0.y - d01 * v1.y) * denomr;
	float gy2 = (d00 * v1.y - d01 * v0.y) * denomr;
	float gz1 = (d11 * v0.z - d01 * v1.z) * denomr;
	float gz2 = (d00 * v1.z - d01 * v0.z) * denomr;

	quadric_zero(Q);
	Q.w = w;

	for (u32 k = 0; k < attribute_count; ++k)
	{
		float a0 = va0[k], a1 = va1[k], a2 = va2[k];
		float gx = gx1 * (a1 - a0) + gx2 * (a2 - a0);
		float gy = gy1 * (a1 - a0) + gy2 * (a2 - a0);
		float gz = gz1 * (a1 - a0) + gz2 * (a2 - a0);
		float gw = a0 - p0.x * gx - p0.y * gy - p0.z * gz;
		Q.a00 += w * (gx * gx);
		Q.a11 += w * (gy * gy);
		Q.a22 += w * (gz * gz);
		Q.a10 += w * (gy * gx);
		Q.a20 += w * (gz * gx);
		Q.a21 += w * (gz * gy);
		Q.b0 += w * (gx * gw);
		Q.b1 += w * (gy * gw);
		Q.b2 += w * (gz * gw);
		Q.c += w * (gw * gw);
		G[k].gx = w * gx;
		G[k].gy = w * gy;
		G[k].gz = w * gz;
		G[k].gw = w * gw;
	}
}

DEVFN bool has_triangle_flip(const Vector3& a, const Vector3& b, const Vector3& c, const Vector3& d)
{
	Vector3 eb = {b.x - a.x, b.y - a.y, b.z - a.z};
	Vector3 ec = {c.x - a.x, c.y - a.y, c.z - a.z};
	Vector3 ed = {d.x - a.x, d.y - a.y, d.z - a.z};
	Vector3 nbc = {eb.y * ec.z - eb.z * ec.y, eb.z * ec.x - eb.x * ec.z, eb.x * ec.y - eb.y * ec.x};
	Vector3 nbd = {eb.y * ed.z - eb.z * ed.y, eb.z * ed.x - eb.x * ed.z, eb.x * ed.y - eb.y * ed.x};
	float ndp = nbc.x * nbd.x + nbc.y * nbd.y + nbc.z * nbd.z;
	float abc = nbc.x * nbc.x + nbc.y * nbc.y + nbc.z * nbc.z;
	float abd = nbd.x * nbd.x + nbd.y * nbd.y + nbd.z * nbd.z;
	return ndp <= 0.25f * sqrtf(abc * abd);
}

DEVFN u32 get_complex_target(u32 v, u32 target, const u32* remap, const u32* loop, const u32* loopback)
{
	u32 r = remap[target];
	if (loop[v] != NONE && remap[loop[v]] == r)
		return loop[v];
	else if (loopback[v] != NONE && remap[loopback[v]] == r)
		return loopback[v];
	else
		return target;
}

// -------------------------------------------------------------------------------------------------- hash table (u64 keys)
DEVFN u32 hash_u64(u64 k)
{
	k ^= k >> 33;
	k *= 0xff51afd7ed558ccdULL;
	k ^= k >> 33;
	k *= 0xc4ceb9fe1a85ec53ULL;
	k ^= k >> 33;
	return u32(k);
}

// returns the slot holding `key`, claiming an empty slot if needed (keys never equal ~0)
DEVFN u32 table_insert(u64* keys, u32 mask, u64 key)
{
	u32 h = hash_u64(key) & mask;
	for (;;)
	{
		unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[h]), ~0ull, (unsigned long long)key);
		if (prev == ~0ull || prev == key)
			return h;
		h = (h + 1) & mask;
	}
}

// ---------------------------------------------------------------------------------------- sparse remap (per group)
KERNEL k_tri_group(const u32* __restrict__ group_tri_offset, u32 G, u32* tri_group, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 lo = 0, hi = G;
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (group_tri_offset[mid] <= u32(t))
			lo = mid;
		else
			hi = mid;
	}
	tri_group[t] = lo;
}

KERNEL k_corner_insert(const u32* __restrict__ gtri, const u32* __restrict__ tri_group, u64* table_keys, u32* table_first, u32 mask, u32* corner_slot, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u64 key = (u64(tri_group[c / 3]) << 32) | gtri[c];
	u32 slot = table_insert(table_keys, mask, key);
	atomicMin(&table_first[slot], u32(c));
	corner_slot[c] = slot;
}

KERNEL k_corner_first_flag(const u32* __restrict__ corner_slot, const u32* __restrict__ table_first, u32* flag, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	flag[c] = table_first[corner_slot[c]] == u32(c) ? 1u : 0u;
}

KERNEL k_assign_sparse(const u32* __restrict__ gtri, const u32* __restrict__ tri_group, const u32* __restrict__ corner_slot, const u32* __restrict__ table_first, const u32* __restrict__ sparse_id,
    u32* idx, u32* sv_global, u32* sv_group, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 fc = table_first[corner_slot[c]];
	u32 s = sparse_id[fc];
	idx[c] = s;
	if (fc == u32(c))
	{
		sv_global[s] = gtri[c];
		sv_group[s] = tri_group[c / 3];
	}
}

// ---------------------------------------------------------------------------------------------------- adjacency
// CSR over sparse vertices of the corners that reference them (optionally through remap); lists sorted by corner index
KERNEL k_adj_count(const u32* __restrict__ idx, const u32* __restrict__ remap, u32* counts, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 v = idx[c];
	if (remap)
		v = remap[v];
	atomicAdd(&counts[v], 1u);
}

KERNEL k_adj_fill(const u32* __restrict__ idx, const u32* __restrict__ remap, const u32* __restrict__ offsets, u32* cursor, u32* adj_corner, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 v = idx[c];
	if (remap)
		v = remap[v];
	u32 slot = atomicAdd(&cursor[v], 1u);
	adj_corner[offsets[v] + slot] = u32(c);
}

KERNEL k_adj_sort(const u32* __restrict__ offsets, u32* adj_corner, u32 vertex_count)
{
	size_t v = GTID;
	if (v >= vertex_count)
		return;
	u32 begin = offsets[v], end = offsets[v + 1];
	for (u32 i = begin + 1; i < end; ++i)
	{
		u32 x = adj_corner[i];
		u32 j = i;
		while (j > begin && adj_corner[j - 1] > x)
		{
			adj_corner[j] = adj_corner[j - 1];
			--j;
		}
		adj_corner[j] = x;
	}
}

DEVFN u32 corner_next(u32 c)
{
	return (c % 3 == 2) ? c - 2 : c + 1;
}

DEVFN u32 corner_prev(u32 c)
{
	return (c % 3 == 0) ? c + 2 : c - 1;
}

// -------------------------------------------------------------------------------------------- position remap + wedges
KERNEL k_posremap_insert(const u32* __restrict__ sv_global, const u32* __restrict__ sv_group, const u32* __restrict__ global_remap, u64* table_keys, u32* table_min, u32 mask, u32* vertex_slot, u32 vertex_count)
{
	size_t s = GTID;
	if (s >= vertex_count)
		return;
	u64 key = (u64(sv_group[s]) << 32) | global_remap[sv_global[s]];
	u32 slot = table_insert(table_keys, mask, key);
	atomicMin(&table_min[slot], u32(s));
	vertex_slot[s] = slot;
}

KERNEL k_posremap_resolve(const u32* __restrict__ vertex_slot, const u32* __restrict__ table_min, u32* remap, u32* sort_key, u32* sort_val, u32 vertex_count)
{
	size_t s = GTID;
	if (s >= vertex_count)
		return;
	u32 r = table_min[vertex_slot[s]];
	remap[s] = r;
	sort_key[s] = r;
	sort_val[s] = u32(s);
}

// buildPositionRemap's wedge loop (simplifier.cpp:228-239): r -> largest member -> ... -> smallest member -> r
KERNEL k_wedge_from_sorted(const u32* __restrict__ sorted_key, const u32* __restrict__ sorted_val, u32* wedge, u32 vertex_count)
{
	size_t p = GTID;
	if (p >= vertex_count)
		return;
	u32 r = sorted_key[p], i = sorted_val[p];
	if (i == r)
	{
		size_t q = p;
		while (q + 1 < vertex_count && sorted_key[q + 1] == r)
			++q;
		wedge[r] = sorted_val[q];
	}
	else
	{
		wedge[i] = sorted_val[p - 1];
	}
}

// ------------------------------------------------------------------------------------------------- classification
DEVFN bool has_edge(const u32* adj_off, const u32* adj_corner, const u32* idx, u32 a, u32 b)
{
	for (u32 e = adj_off[a]; e < adj_off[a + 1]; ++e)
		if (idx[corner_next(adj_corner[e])] == b)
			return true;
	return false;
}

DEVFN bool has_edge_wedge(const u32* adj_off, const u32* adj_corner, const u32* idx, u32 a, u32 b, const u32* remap, const u32* wedge)
{
	u32 v = a;
	do
	{
		for (u32 e = adj_off[v]; e < adj_off[v + 1]; ++e)
			if (remap[idx[corner_next(adj_corner[e])]] == remap[b])
				return true;
		v = wedge[v];
	} while (v != a);
	return false;
}

KERNEL k_open_edges(const u32* __restrict__ idx, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner, u32* in_count, u32* in_vertex, u32* out_count, u32* out_vertex, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 vertex = idx[c];
	u32 target = idx[corner_next(u32(c))];
	if (target == vertex)
	{
		atomicAdd(&in_count[vertex], 2u);
		atomicAdd(&out_count[vertex], 2u);
	}
	else if (!has_edge(adj_off, adj_corner, idx, target, vertex))
	{
		atomicAdd(&in_count[target], 1u);
		in_vertex[target] = vertex; // only consumed when the count is exactly 1 (single writer)
		atomicAdd(&out_count[vertex], 1u);
		out_vertex[vertex] = target;
	}
}

KERNEL k_open_finalize(const u32* __restrict__ in_count, const u32* __restrict__ in_vertex, const u32* __restrict__ out_count, const u32* __restrict__ out_vertex, u32* loop, u32* loopback, u32 vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	loopback[i] = in_count[i] == 0 ? NONE : (in_count[i] == 1 ? in_vertex[i] : u32(i));
	loop[i] = out_count[i] == 0 ? NONE : (out_count[i] == 1 ? out_vertex[i] : u32(i));
}

KERNEL k_classify_primary(const u32* __restrict__ remap, const u32* __restrict__ wedge, const u32* __restrict__ loop, const u32* __restrict__ loopback, u8* kind, u32 vertex_count)
{
	size_t ii = GTID;
	if (ii >= vertex_count)
		return;
	u32 i = u32(ii);
	if (remap[i] != i)
		return;
	const u32* openinc = loopback;
	const u32* openout = loop;
	u8 result;
	if (wedge[i] == i)
	{
		u32 openi = openinc[i], openo = openout[i];
		if (openi == NONE && openo == NONE)
			result = Kind_Manifold;
		else if (openi != NONE && openo != NONE && remap[openi] == remap[openo] && openi != i)
			result = Kind_Seam;
		else if (openi != i && openo != i)
			result = Kind_Border;
		else
			result = Kind_Locked;
	}
	else if (wedge[wedge[i]] == i)
	{
		u32 w = wedge[i];
		u32 openiv = openinc[i], openov = openout[i];
		u32 openiw = openinc[w], openow = openout[w];
		if (openiv != NONE && openiv != i && openov != NONE && openov != i && openiw != NONE && openiw != w && openow != NONE && openow != w)
		{
			if (remap[openiv] == remap[openow] && remap[openov] == remap[openiw] && remap[openiv] != remap[openov])
				result = Kind_Seam;
			else
				result = Kind_Locked;
		}
		else
			result = Kind_Locked;
	}
	else
		result = Kind_Locked;
	kind[i] = result;
}

KERNEL k_classify_copy(const u32* __restrict__ remap, u8* kind, u32 vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	if (remap[i] != u32(i))
		kind[i] = kind[remap[i]];
}

// permissive mode: seam/locked vertices without protected wedges or border edges become complex (simplifier.cpp:489-522)
KERNEL k_classify_permissive(const u32* __restrict__ remap, const u32* __restrict__ wedge, const u32* __restrict__ sv_global, const u8* __restrict__ vertex_lock,
    const u32* __restrict__ idx, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner, u8* kind, u32 vertex_count)
{
	size_t ii = GTID;
	if (ii >= vertex_count)
		return;
	u32 i = u32(ii);
	if (remap[i] != i)
		return;
	if (kind[i] != Kind_Seam && kind[i] != Kind_Locked)
		return;
	bool protect = false;
	u32 v = i;
	do
	{
		protect |= vertex_lock && (vertex_lock[sv_global[v]] & 2) != 0;
		v = wedge[v];
	} while (v != i);
	do
	{
		for (u32 e = adj_off[v]; e < adj_off[v + 1]; ++e)
			protect |= !has_edge_wedge(adj_off, adj_corner, idx, idx[corner_next(adj_corner[e])], v, remap, wedge);
		v = wedge[v];
	} while (v != i);
	if (!protect)
		kind[i] = Kind_Complex;
}

KERNEL k_classify_lock_primary(const u32* __restrict__ remap, const u32* __restrict__ sv_global, const u8* __restrict__ vertex_lock, u8* kind, u32 vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	if (vertex_lock[sv_global[i]] & 1)
		kind[remap[i]] = Kind_Locked;
}

KERNEL k_classify_lock_spread(const u32* __restrict__ remap, u8* kind, u32 vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	if (kind[remap[i]] == Kind_Locked)
		kind[i] = Kind_Locked;
}

// ------------------------------------------------------------------------------------------------------ rescale
DEVFN u32 float_order_key(float f)
{
	u32 u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

DEVFN float float_from_order_key(u32 k)
{
	u32 u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
	return __uint_as_float(u);
}

KERNEL k_group_minmax(const u32* __restrict__ sv_global, const u32* __restrict__ sv_group, const float* __restrict__ positions, u32* gmin, u32* gmax, u32 vertex_count)
{
	size_t s = GTID;
	if (s >= vertex_count)
		return;
	const float* p = positions + size_t(sv_global[s]) * 3;
	u32 g = sv_group[s];
#ifndef CLODB_EMU
	// sparse vertices are numbered group by group, so a warp almost always sits in one group: reduce in registers and
	// issue one atomic per warp and component instead of 32 contending ones
	unsigned active = __activemask();
	u32 g0 = __shfl_sync(active, g, __ffs(active) - 1);
	if (active == 0xffffffffu && __all_sync(active, g == g0))
	{
		for (int k = 0; k < 3; ++k)
		{
			u32 key = float_order_key(p[k]);
			u32 lo = key, hi = key;
			for (int d = 16; d >= 1; d >>= 1)
			{
				u32 olo = __shfl_xor_sync(0xffffffffu, lo, d), ohi = __shfl_xor_sync(0xffffffffu, hi, d);
				lo = olo < lo ? olo : lo;
				hi = ohi > hi ? ohi : hi;
			}
			if ((threadIdx.x & 31) == 0)
			{
				atomicMin(&gmin[g * 3 + k], lo);
				atomicMax(&gmax[g * 3 + k], hi);
			}
		}
		return;
	}
#endif
	for (int k = 0; k < 3; ++k)
	{
		// the reference keeps the first value unless strictly smaller/larger; with -0/+0 ties the order key orders -0 < +0,
		// which can only differ from it by the sign of a zero extent contribution
		atomicMin(&gmin[g * 3 + k], float_order_key(p[k]));
		atomicMax(&gmax[g * 3 + k], float_order_key(p[k]));
	}
}

KERNEL k_group_extent(const u32* __restrict__ gmin, const u32* __restrict__ gmax, float* group_min, float* group_extent, u32 G)
{
	size_t g = GTID;
	if (g >= G)
		return;
	float extent = 0.f;
	for (int k = 0; k < 3; ++k)
	{
		float mn = float_from_order_key(gmin[g * 3 + k]), mx = float_from_order_key(gmax[g * 3 + k]);
		group_min[g * 3 + k] = mn;
		extent = (mx - mn) < extent ? extent : (mx - mn);
	}
	group_extent[g] = extent;
}

KERNEL k_rescale(const u32* __restrict__ sv_global, const u32* __restrict__ sv_group, const float* __restrict__ positions, const float* __restrict__ group_min, const float* __restrict__ group_extent,
    const float* __restrict__ attributes, u32 attribute_stride, u32 attribute_count, const u32* __restrict__ attribute_remap, const float* __restrict__ attribute_weights,
    Vector3* vpos, float* vattr, u32 vertex_count)
{
	size_t s = GTID;
	if (s >= vertex_count)
		return;
	u32 gv = sv_global[s];
	u32 g = sv_group[s];
	float extent = group_extent[g];
	float scale = extent == 0 ? 0.f : 1.f / extent;
	const float* p = positions + size_t(gv) * 3;
	Vector3 r;
	r.x = (p[0] - group_min[g * 3 + 0]) * scale;
	r.y = (p[1] - group_min[g * 3 + 1]) * scale;
	r.z = (p[2] - group_min[g * 3 + 2]) * scale;
	vpos[s] = r;
	for (u32 k = 0; k < attribute_count; ++k)
	{
		u32 rk = attribute_remap[k];
		float a = attributes[size_t(gv) * attribute_stride + rk];
		vattr[s * attribute_count + k] = a * attribute_weights[rk];
	}
}

// ------------------------------------------------------------------------------------------------------ quadrics
// vertex_quadrics of canonical vertex r: face quadrics in corner order, then the point quadric, then edge quadrics in
// (triangle, edge) order - the accumulation order of fillFaceQuadrics/fillVertexQuadrics/fillEdgeQuadrics.
KERNEL k_fill_vertex_quadrics(const u32* __restrict__ idx, const u32* __restrict__ remap, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner, const Vector3* __restrict__ vpos,
    const u8* __restrict__ kind, const u32* __restrict__ loop, const u32* __restrict__ loopback, Quadric* vertex_quadrics, u32 vertex_count)
{
	size_t rr = GTID;
	if (rr >= vertex_count)
		return;
	u32 r = u32(rr);
	Quadric acc;
	quadric_zero(acc);
	if (remap[r] != r)
	{
		vertex_quadrics[r] = acc;
		return;
	}
	u32 begin = adj_off[r], end = adj_off[r + 1];
	for (u32 e = begin; e < end; ++e)
	{
		u32 t = adj_corner[e] / 3;
		u32 i0 = idx[t * 3 + 0], i1 = idx[t * 3 + 1], i2 = idx[t * 3 + 2];
		Quadric Q;
		quadric_from_triangle(Q, vpos[i0], vpos[i1], vpos[i2], 1.f);
		quadric_add(acc, Q);
	}
	{
		const Vector3& p = vpos[r];
		float w = acc.w * 1e-7f;
		Quadric Q;
		quadric_from_point(Q, p.x, p.y, p.z, w);
		quadric_add(acc, Q);
	}
	for (u32 e = begin; e < end;)
	{
		u32 t = adj_corner[e] / 3;
		// corners of triangle t that map to r are adjacent in the sorted list
		u32 e_end = e;
		while (e_end < end && adj_corner[e_end] / 3 == t)
			++e_end;
		u32 tri[3] = {idx[t * 3 + 0], idx[t * 3 + 1], idx[t * 3 + 2]};
		for (int ed = 0; ed < 3; ++ed)
		{
			int nx = ed == 2 ? 0 : ed + 1;
			int nn = nx == 2 ? 0 : nx + 1;
			u32 i0 = tri[ed], i1 = tri[nx];
			u8 k0 = kind[i0], k1 = kind[i1];
			if (k0 != Kind_Border && k0 != Kind_Seam && k1 != Kind_Border && k1 != Kind_Seam)
				continue;
			if ((k0 == Kind_Border || k0 == Kind_Seam) && loop[i0] != i1)
				continue;
			if ((k1 == Kind_Border || k1 == Kind_Seam) && loopback[i1] != i0)
				continue;
			bool hit0 = remap[i0] == r, hit1 = remap[i1] == r;
			if (!hit0 && !hit1)
				continue;
			u32 i2 = tri[nn];
			const float kEdgeWeightSeam = 0.5f;
			const float kEdgeWeightBorder = 10.f;
			float edgeWeight = (k0 == Kind_Border || k1 == Kind_Border) ? kEdgeWeightBorder : kEdgeWeightSeam;
			Quadric Q;
			quadric_from_triangle_edge(Q, vpos[i0], vpos[i1], vpos[i2], edgeWeight);
			Quadric QT;
			quadric_from_triangle(QT, vpos[i0], vpos[i1], vpos[i2], edgeWeight);
			QT.w = 0;
			quadric_add(Q, QT);
			if (hit0)
				quadric_add(acc, Q);
			if (hit1)
				quadric_add(acc, Q);
		}
		e = e_end;
	}
	vertex_quadrics[r] = acc;
}

// attribute quadrics/gradients per wedge vertex, accumulated in corner order (fillAttributeQuadrics)
KERNEL k_fill_attribute_quadrics(const u32* __restrict__ idx, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner, const Vector3* __restrict__ vpos, const float* __restrict__ vattr, u32 attribute_count,
    Quadric* attribute_quadrics, QuadricGrad* attribute_gradients, u32 vertex_count)
{
	size_t v = GTID;
	if (v >= vertex_count)
		return;
	Quadric acc;
	quadric_zero(acc);
	QuadricGrad gacc[32];
	for (u32 k = 0; k < attribute_count; ++k)
		gacc[k].gx = gacc[k].gy = gacc[k].gz = gacc[k].gw = 0.f;
	for (u32 e = adj_off[v]; e < adj_off[v + 1]; ++e)
	{
		u32 t = adj_corner[e] / 3;
		u32 i0 = idx[t * 3 + 0], i1 = idx[t * 3 + 1], i2 = idx[t * 3 + 2];
		Quadric QA;
		QuadricGrad G[32];
		quadric_from_attributes(QA, G, vpos[i0], vpos[i1], vpos[i2], &vattr[size_t(i0) * attribute_count], &vattr[size_t(i1) * attribute_count], &vattr[size_t(i2) * attribute_count], attribute_count);
		quadric_add(acc, QA);
		for (u32 k = 0; k < attribute_count; ++k)
		{
			gacc[k].gx += G[k].gx;
			gacc[k].gy += G[k].gy;
			gacc[k].gz += G[k].gz;
			gacc[k].gw += G[k].gw;
		}
	}
	attribute_quadrics[v] = acc;
	for (u32 k = 0; k < attribute_count; ++k)
		attribute_gradients[v * attribute_count + k] = gacc[k];
}

// ------------------------------------------------------------------------------------------------ pass: pick / rank
struct GroupState
{
	u32 tri_begin;     // first triangle of the group in the current index buffer
	u32 tri_count;     // current triangle count
	u32 target_tris;   // target triangle count
	u32 active;        // still simplifying
	u32 cand_begin;    // first candidate of the group this pass
	u32 cand_count;
	u32 cut;           // global sorted position where the serial scan would have stopped
	float result_error;
	u32 win_begin;     // sorted positions [win_begin, win_end) are the candidates examined by the current wavefront run
	u32 win_end;
};

// one slot per triangle edge; flag + direction info (pickEdgeCollapses)
KERNEL k_pick_flags(const u32* __restrict__ idx, const u32* __restrict__ tri_group, const GroupState* __restrict__ groups, const u32* __restrict__ remap, const u8* __restrict__ kind,
    const u32* __restrict__ loop, const u32* __restrict__ loopback, u32* flags, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 out = 0;
	u32 t = u32(c / 3);
	if (groups[tri_group[t]].active)
	{
		u32 i0 = idx[c];
		u32 i1 = idx[corner_next(u32(c))];
		if (remap[i0] != remap[i1])
		{
			u8 k0 = kind[i0], k1 = kind[i1];
			bool ok = (kCanCollapse[k0][k1] | kCanCollapse[k1][k0]) != 0;
			if (ok && kHasOpposite[k0][k1] && remap[i1] > remap[i0])
				ok = false;
			if (ok && (k0 == Kind_Border || k0 == Kind_Seam) && k1 != Kind_Manifold && loop[i0] != i1)
				ok = false;
			if (ok && (k1 == Kind_Border || k1 == Kind_Seam) && k0 != Kind_Manifold && loopback[i1] != i0)
				ok = false;
			out = ok ? 1u : 0u;
		}
	}
	flags[c] = out;
}

KERNEL k_pick_emit(const u32* __restrict__ idx, const u32* __restrict__ flags_scanned, const u32* __restrict__ tri_group, const u8* __restrict__ kind, u32 total, u32* cand_v0, u32* cand_v1, u8* cand_bidi, u32* cand_group, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 pos = flags_scanned[c];
	u32 next = c + 1 < corners ? flags_scanned[c + 1] : total;
	if (next == pos)
		return;
	u32 i0 = idx[c];
	u32 i1 = idx[corner_next(u32(c))];
	u8 k0 = kind[i0], k1 = kind[i1];
	if (kCanCollapse[k0][k1] & kCanCollapse[k1][k0])
	{
		cand_v0[pos] = i0;
		cand_v1[pos] = i1;
		cand_bidi[pos] = 1;
	}
	else
	{
		cand_v0[pos] = kCanCollapse[k0][k1] ? i0 : i1;
		cand_v1[pos] = kCanCollapse[k0][k1] ? i1 : i0;
		cand_bidi[pos] = 0;
	}
	cand_group[pos] = tri_group[c / 3];
}

KERNEL k_group_cand_ranges(GroupState* groups, const u32* __restrict__ flags_scanned, u32 total, u32 G, size_t corners)
{
	size_t g = GTID;
	if (g >= G)
		return;
	GroupState& gs = groups[g];
	size_t c0 = size_t(gs.tri_begin) * 3, c1 = size_t(gs.tri_begin + gs.tri_count) * 3;
	u32 b = c0 < corners ? flags_scanned[c0] : total;
	u32 e = c1 < corners ? flags_scanned[c1] : total;
	gs.cand_begin = b;
	gs.cand_count = e - b;
	gs.cut = e;
}

KERNEL k_rank(u32* cand_v0, u32* cand_v1, const u8* __restrict__ cand_bidi, const u32* __restrict__ cand_group, float* cand_error, u32* sort_key, u32* sort_val, u32 cand_total,
    const Vector3* __restrict__ vpos, const float* __restrict__ vattr, const Quadric* __restrict__ vertex_quadrics, const Quadric* __restrict__ attribute_quadrics, const QuadricGrad* __restrict__ attribute_gradients,
    u32 attribute_count, const u32* __restrict__ remap, const u32* __restrict__ wedge, const u8* __restrict__ kind, const u32* __restrict__ loop, const u32* __restrict__ loopback)
{
	size_t ci = GTID;
	if (ci >= cand_total)
		return;
	u32 i0 = cand_v0[ci];
	u32 i1 = cand_v1[ci];
	bool bidi = cand_bidi[ci] != 0;

	float ei = quadric_error(vertex_quadrics[remap[i0]], vpos[i1]);
	float ej = bidi ? quadric_error(vertex_quadrics[remap[i1]], vpos[i0]) : FLT_MAX;

	if (attribute_count)
	{
		ei += quadric_error_attr(attribute_quadrics[i0], &attribute_gradients[size_t(i0) * attribute_count], attribute_count, vpos[i1], &vattr[size_t(i1) * attribute_count]);
		ej += bidi ? quadric_error_attr(attribute_quadrics[i1], &attribute_gradients[size_t(i1) * attribute_count], attribute_count, vpos[i0], &vattr[size_t(i0) * attribute_count]) : 0;

		if (kind[i0] == Kind_Seam)
		{
			u32 s0 = wedge[i0];
			u32 s1 = loop[i0] == i1 ? loopback[s0] : loop[s0];
			s1 = (s1 != NONE) ? s1 : wedge[i1];
			ei += quadric_error_attr(attribute_quadrics[s0], &attribute_gradients[size_t(s0) * attribute_count], attribute_count, vpos[s1], &vattr[size_t(s1) * attribute_count]);
			ej += bidi ? quadric_error_attr(attribute_quadrics[s1], &attribute_gradients[size_t(s1) * attribute_count], attribute_count, vpos[s0], &vattr[size_t(s0) * attribute_count]) : 0;
		}
		else
		{
			if (kind[i0] == Kind_Complex)
				for (u32 v = wedge[i0]; v != i0; v = wedge[v])
				{
					u32 t = get_complex_target(v, i1, remap, loop, loopback);
					ei += quadric_error_attr(attribute_quadrics[v], &attribute_gradients[size_t(v) * attribute_count], attribute_count, vpos[t], &vattr[size_t(t) * attribute_count]);
				}
			if (kind[i1] == Kind_Complex && bidi)
				for (u32 v = wedge[i1]; v != i1; v = wedge[v])
				{
					u32 t = get_complex_target(v, i0, remap, loop, loopback);
					ej += quadric_error_attr(attribute_quadrics[v], &attribute_gradients[size_t(v) * attribute_count], attribute_count, vpos[t], &vattr[size_t(t) * attribute_count]);
				}
		}
	}

	bool rev = bidi & (ej < ei);
	cand_v0[ci] = rev ? i1 : i0;
	cand_v1[ci] = rev ? i0 : i1;
	float error = ej < ei ? ej : ei;
	cand_error[ci] = error;

	// sortEdgeCollapses key: top 12 bits of exponent+mantissa, clamped (simplifier.cpp:1478-1493)
	const u32 sort_bits = 12;
	const u32 sort_bins = 2048 + 512;
	u32 key = (__float_as_uint(error) << 1) >> (32 - sort_bits);
	key = key < sort_bins ? key : sort_bins - 1;
	sort_key[ci] = (cand_group[ci] << sort_bits) | key;
	sort_val[ci] = u32(ci);
}

// ---- the examined prefixes as one dense index space ---------------------------------------------------------------------
// Everything after the wavefront (the cut scans, the roll-back, the quadric updates) only concerns the sorted positions the
// wavefront has examined: per group the prefix [cand_begin, win_end) of its candidates, typically a third of the list. Those
// prefixes are laid end to end: window position j of group g <-> sorted position cand_begin[g] + (j - woff[g]). The kernels below
// run over j in [0, W) and find their group by a binary search over the G + 1 offsets (cache resident).
KERNEL k_window_lengths(const GroupState* __restrict__ groups, u32 G, u32* lengths)
{
	size_t g = GTID;
	if (g > G)
		return;
	u32 len = 0;
	if (g < G)
	{
		const GroupState& gs = groups[g];
		if (gs.active && gs.cand_count)
			len = gs.win_end - gs.cand_begin;
	}
	lengths[g] = len; // lengths[G] = 0: after the exclusive scan it holds W
}

DEVFN u32 window_group(const u32* __restrict__ woff, u32 G, u32 j)
{
	u32 lo = 0, hi = G;
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (woff[mid] <= j)
			lo = mid;
		else
			hi = mid;
	}
	return lo;
}

// ------------------------------------------------------------------------------------- pass: greedy collapse wavefront
KERNEL k_collapse_init(u32* collapse_remap, u8* collapse_locked, u32 vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	collapse_remap[i] = u32(i);
	collapse_locked[i] = 0;
}

// publish, per canonical vertex, the lowest sorted position among still-undecided candidates touching it
DEVFN void wave_publish_item(u32 k, const u32* __restrict__ sorted_cand, const u32* __restrict__ cand_v0, const u32* __restrict__ cand_v1, const u32* __restrict__ remap, u64* vmin_any, u64* vmin_src, u32 round_tag)
{
	u32 c = sorted_cand[k];
	u32 r0 = remap[cand_v0[c]], r1 = remap[cand_v1[c]];
	u64 value = (u64(~round_tag) << 32) | u64(k);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_any[r0]), (unsigned long long)value);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_any[r1]), (unsigned long long)value);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_src[r0]), (unsigned long long)value);
}

DEVFN u32 wave_min(const u64* vmin, u32 v, u32 round_tag)
{
	u64 x = vmin[v];
	return u32(x >> 32) == ~round_tag ? u32(x) : NONE;
}

// One dependency-wavefront step for the undecided candidate at sorted position k; returns true while it stays undecided.
DEVFN bool wave_decide_item(u32 k, const u32* __restrict__ sorted_cand, u8* status, const u32* __restrict__ cand_v0, const u32* __restrict__ cand_v1, const u32* __restrict__ remap, const u32* __restrict__ wedge, const u8* __restrict__ kind,
    const u32* __restrict__ loop, const u32* __restrict__ loopback, const Vector3* __restrict__ vpos, const u32* __restrict__ idx, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner,
    const u64* __restrict__ vmin_any, const u64* __restrict__ vmin_src, u32 round_tag, u32* collapse_remap, u8* collapse_locked)
{
	u32 c = sorted_cand[k];
	u32 i0 = cand_v0[c], i1 = cand_v1[c];
	u32 r0 = remap[i0], r1 = remap[i1];

	// a lock set by any decided collapse is final: locks are only ever set by lower-ranked candidates (see header)
	if (collapse_locked[r0] | collapse_locked[r1])
	{
		status[k] = Status_Locked;
		return false;
	}
	if (wave_min(vmin_any, r0, round_tag) != k || wave_min(vmin_any, r1, round_tag) != k)
		return true;
	// the flip test reads collapse_remap of r0's neighbours: wait for lower-ranked collapses that could move them
	for (u32 e = adj_off[r0]; e < adj_off[r0 + 1]; ++e)
	{
		u32 corner = adj_corner[e];
		u32 a = remap[idx[corner_next(corner)]], b = remap[idx[corner_prev(corner)]];
		if (wave_min(vmin_src, a, round_tag) < k || wave_min(vmin_src, b, round_tag) < k)
			return true;
	}

	// hasTriangleFlips(adjacency, vertex_positions, collapse_remap, r0, r1)
	{
		const Vector3& v0 = vpos[r0];
		const Vector3& v1 = vpos[r1];
		for (u32 e = adj_off[r0]; e < adj_off[r0 + 1]; ++e)
		{
			u32 corner = adj_corner[e];
			u32 a = collapse_remap[remap[idx[corner_next(corner)]]];
			u32 b = collapse_remap[remap[idx[corner_prev(corner)]]];
			if (a == r1 || b == r1 || a == b)
				continue;
			if (has_triangle_flip(vpos[a], vpos[b], v0, v1))
			{
				status[k] = Status_Flip;
				return false;
			}
		}
	}

	u8 kd = kind[i0];
	if (kd == Kind_Complex)
	{
		u32 v = i0;
		do
		{
			collapse_remap[v] = get_complex_target(v, i1, remap, loop, loopback);
			v = wedge[v];
		} while (v != i0);
	}
	else if (kd == Kind_Seam)
	{
		u32 s0 = wedge[i0];
		u32 s1 = loop[i0] == i1 ? loopback[s0] : loop[s0];
		s1 = (s1 != NONE) ? s1 : wedge[i1];
		collapse_remap[i0] = i1;
		collapse_remap[s0] = s1;
	}
	else
	{
		collapse_remap[i0] = i1;
	}
	collapse_locked[r0] = 1;
	collapse_locked[r1] = 1;
	status[k] = Status_Performed;
	return false;
}

#ifdef CLODB_EMU
// development emulation: one publish + one decide sweep over the window list per round
KERNEL k_wave_publish(const u32* __restrict__ list, u32 n, const u32* __restrict__ sorted_cand, const u8* __restrict__ status, const u32* __restrict__ cand_v0, const u32* __restrict__ cand_v1, const u32* __restrict__ remap,
    u64* vmin_any, u64* vmin_src, u32 round_tag, u32* undecided_count)
{
	size_t i = GTID;
	if (i >= n || status[list[i]] != Status_Undecided)
		return;
	wave_publish_item(list[i], sorted_cand, cand_v0, cand_v1, remap, vmin_any, vmin_src, round_tag);
	atomicAdd(undecided_count, 1u);
}

KERNEL k_wave_decide(const u32* __restrict__ list, u32 n, const u32* __restrict__ sorted_cand, u8* status, const u32* __restrict__ cand_v0, const u32* __restrict__ cand_v1, const u32* __restrict__ remap, const u32* __restrict__ wedge, const u8* __restrict__ kind,
    const u32* __restrict__ loop, const u32* __restrict__ loopback, const Vector3* __restrict__ vpos, const u32* __restrict__ idx, const u32* __restrict__ adj_off, const u32* __restrict__ adj_corner,
    const u64* __restrict__ vmin_any, const u64* __restrict__ vmin_src, u32 round_tag, u32* collapse_remap, u8* collapse_locked)
{
	size_t i = GTID;
	if (i >= n || status[list[i]] != Status_Undecided)
		return;
	wave_decide_item(list[i], sorted_cand, status, cand_v0, cand_v1, remap, wedge, kind, loop, loopback, vpos, idx, adj_off, adj_corner, vmin_any, vmin_src, round_tag, collapse_remap, collapse_locked);
}
#else
// ---- persistent wavefront kernel ------------------------------------------------------------------------------------------
// All wavefront rounds of a pass in one cooperative launch. What a round costs is a chain of dependent gathers per undecided
// candidate and the barrier between rounds, so the kernel is organised around both:
//   * the work list holds 16-byte records {sorted position, r0, r1, candidate}: the position classes of both endpoints are
//     resolved once, when the candidate is listed, and a round starts with one coalesced record load followed by four
//     INDEPENDENT gathers (two lock bytes, two vertex minima);
//   * a candidate that stays undecided publishes itself for the NEXT round while it is appended to the next list (vertex
//     minima are double buffered by round parity), so a round needs one grid barrier instead of two;
//   * only candidates that are minimal at both endpoints walk the adjacency of r0; they are compacted inside the warp and
//     handled by 8 cooperating lanes each (the two walks have ~6-12 entries);
//   * once the list is short (tail_threshold) the other CTAs leave and CTA 0 finishes the remaining rounds alone, separated by
//     __syncthreads() instead of grid barriers: the tail is a long sequence of almost empty rounds (dependency chains).
// Mutable arrays are read with ld.global.cg: in the single-CTA tail the values were written by other warps of the same SM
// between two __syncthreads(), and L1 must not serve an older copy.
struct WaveEntry
{
	u32 k, r0, r1, c;
};

// state[0..2] list counters (round r appends through state[r % 3]); [3] running round tag; [4] rounds of this launch;
// [5] undecided left; [6] rounds of all launches; [7] max rounds in a launch; [8..9] (u64) work items (candidate x round)
struct WaveArgs
{
	u8* status;
	const u32 *cand_v0, *cand_v1, *remap, *wedge;
	const u8* kind;
	const u32 *loop, *loopback;
	const Vector3* vpos;
	const u32 *idx, *adj_off, *adj_corner;
	u64* vmin_any[2];
	u64* vmin_src[2];
	u32* collapse_remap;
	u8* collapse_locked;
	WaveEntry* list[2];
	u32* state;
	u32 max_rounds, tail_threshold;
	u32* round_log; // optional (debug): per round {active count, globaltimer ns low word at round end}
};

DEVFN u32 wave_min_cg(const u64* vmin, u32 v, u32 round_tag)
{
	u64 x = __ldcg(reinterpret_cast<const unsigned long long*>(vmin + v));
	return u32(x >> 32) == ~round_tag ? u32(x) : NONE;
}

DEVFN void wave_publish_entry(const WaveEntry& e, u64* vmin_any, u64* vmin_src, u32 round_tag)
{
	unsigned long long value = (u64(~round_tag) << 32) | u64(e.k);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_any[e.r0]), value);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_any[e.r1]), value);
	atomicMin(reinterpret_cast<unsigned long long*>(&vmin_src[e.r0]), value);
}

// lists the undecided candidates of the current windows as records and publishes them for the first round; one thread per
// position of the examined prefixes (window index space, see k_window_lengths)
static __global__ void __launch_bounds__(256) k_wave_list(const u32* __restrict__ woff, u32 G, u32 W, const u32* __restrict__ sorted_cand, const GroupState* __restrict__ groups, const u8* __restrict__ status,
    const u32* __restrict__ cand_v0, const u32* __restrict__ cand_v1, const u32* __restrict__ remap, WaveArgs a)
{
	size_t jj = GTID;
	bool take = false;
	WaveEntry e = {0, 0, 0, 0};
	if (jj < W)
	{
		const u32 j = u32(jj);
		const u32 g = window_group(woff, G, j);
		const GroupState& gs = groups[g];
		const u32 k = gs.cand_begin + (j - woff[g]);
		take = k >= gs.win_begin && status[k] == Status_Undecided; // k < win_end by construction
		if (take)
		{
			u32 c = sorted_cand[k];
			e.k = k;
			e.c = c;
			e.r0 = remap[cand_v0[c]];
			e.r1 = remap[cand_v1[c]];
		}
	}
	unsigned mask = __ballot_sync(0xffffffffu, take);
	if (mask)
	{
		u32 lane = threadIdx.x & 31;
		u32 off = 0;
		if (lane == u32(__ffs(mask) - 1))
			off = atomicAdd(a.state, u32(__popc(mask)));
		off = __shfl_sync(0xffffffffu, off, __ffs(mask) - 1);
		if (take)
		{
			a.list[0][off + __popc(mask & ((1u << lane) - 1))] = e;
			u32 tag = a.state[3] + 1;
			wave_publish_entry(e, a.vmin_any[tag & 1], a.vmin_src[tag & 1], tag);
		}
	}
}

// The part of the decide step that walks the adjacency of r0, executed by the 8 lanes of gmask for a candidate that is minimal
// at both endpoints. Returns 1 while the candidate has to wait for a lower-ranked collapse next to it, else 0 (decided).
DEVFN int wave_walk8(const WaveArgs& a, u32 k, u32 r0, u32 r1, u32 c, unsigned gmask, u32 lane8, const u64* __restrict__ src_cur, u32 tag)
{
	const u32 begin = a.adj_off[r0], end = a.adj_off[r0 + 1];
	bool wait = false;
	for (u32 e = begin + lane8; e < end; e += 8)
	{
		u32 corner = a.adj_corner[e];
		u32 va = a.remap[a.idx[corner_next(corner)]], vb = a.remap[a.idx[corner_prev(corner)]];
		wait |= wave_min_cg(src_cur, va, tag) < k || wave_min_cg(src_cur, vb, tag) < k;
	}
	if (__ballot_sync(gmask, wait))
		return 1;
	bool flip = false;
	{
		const Vector3 v0 = a.vpos[r0];
		const Vector3 v1 = a.vpos[r1];
		for (u32 e = begin + lane8; e < end; e += 8)
		{
			u32 corner = a.adj_corner[e];
			u32 va = __ldcg(&a.collapse_remap[a.remap[a.idx[corner_next(corner)]]]);
			u32 vb = __ldcg(&a.collapse_remap[a.remap[a.idx[corner_prev(corner)]]]);
			if (va == r1 || vb == r1 || va == vb)
				continue;
			flip |= has_triangle_flip(a.vpos[va], a.vpos[vb], v0, v1);
		}
	}
	if (__ballot_sync(gmask, flip))
	{
		if (lane8 == 0)
			a.status[k] = Status_Flip;
		return 0;
	}
	if (lane8 == 0)
	{
		u32 i0 = a.cand_v0[c], i1 = a.cand_v1[c];
		u8 kd = a.kind[i0];
		if (kd == Kind_Complex)
		{
			u32 v = i0;
			do
			{
				a.collapse_remap[v] = get_complex_target(v, i1, a.remap, a.loop, a.loopback);
				v = a.wedge[v];
			} while (v != i0);
		}
		else if (kd == Kind_Seam)
		{
			u32 s0 = a.wedge[i0];
			u32 s1 = a.loop[i0] == i1 ? a.loopback[s0] : a.loop[s0];
			s1 = (s1 != NONE) ? s1 : a.wedge[i1];
			a.collapse_remap[i0] = i1;
			a.collapse_remap[s0] = s1;
		}
		else
		{
			a.collapse_remap[i0] = i1;
		}
		a.collapse_locked[r0] = 1;
		a.collapse_locked[r1] = 1;
		a.status[k] = Status_Performed;
	}
	return 0;
}

// one round over cur[0..n): decide, and append + publish (for round tag + 1) what stays undecided. Executed by `nthreads`
// threads (the whole grid, or one CTA in the tail); tid is the thread's rank among them.
DEVFN void wave_round(const WaveArgs& a, const WaveEntry* cur, u32 n, WaveEntry* next, u32* counter, u32 tag, u32 tid, u32 nthreads)
{
	const u32 lane = threadIdx.x & 31;
	const u32 sub = lane >> 3, lane8 = lane & 7;
	const unsigned gmask = 0xffu << (sub * 8);
	const u64* any_cur = a.vmin_any[tag & 1];
	const u64* src_cur = a.vmin_src[tag & 1];
	u64* any_nxt = a.vmin_any[(tag + 1) & 1];
	u64* src_nxt = a.vmin_src[(tag + 1) & 1];
	for (u32 base = tid - lane; base < n; base += nthreads)
	{
		const u32 i = base + lane;
		const bool valid = i < n;
		WaveEntry e = {0, 0, 0, 0};
		bool undecided = false, ready = false;
		if (valid)
		{
			uint4 raw = __ldcg(reinterpret_cast<const uint4*>(cur + i));
			e.k = raw.x, e.r0 = raw.y, e.r1 = raw.z, e.c = raw.w;
			// four independent gathers. A lock set by any decided collapse is final: locks are only ever set by lower-ranked
			// candidates (see the file header)
			u32 l0 = __ldcg(&a.collapse_locked[e.r0]), l1 = __ldcg(&a.collapse_locked[e.r1]);
			u32 m0 = wave_min_cg(any_cur, e.r0, tag), m1 = wave_min_cg(any_cur, e.r1, tag);
			if (l0 | l1)
				a.status[e.k] = Status_Locked;
			else if (m0 != e.k || m1 != e.k)
				undecided = true;
			else
				ready = true;
		}
		unsigned rmask = __ballot_sync(0xffffffffu, ready);
		while (rmask)
		{
			// the sub-th ready lane of this batch of (up to) four is handled by the 8 lanes of group `sub`
			const u32 src = __fns(rmask, 0, sub + 1);
			const bool active = src != 0xffffffffu;
			const u32 from = active ? src : 0;
			u32 k = __shfl_sync(0xffffffffu, e.k, from), r0 = __shfl_sync(0xffffffffu, e.r0, from), r1 = __shfl_sync(0xffffffffu, e.r1, from), c = __shfl_sync(0xffffffffu, e.c, from);
			int res = 0;
			if (active)
				res = wave_walk8(a, k, r0, r1, c, gmask, lane8, src_cur, tag);
			__syncwarp();
			for (u32 s = 0; s < 4; ++s)
			{
				u32 owner = __fns(rmask, 0, s + 1);
				int r = __shfl_sync(0xffffffffu, res, s * 8);
				if (owner == lane)
					undecided = r != 0;
			}
			// drop the (up to) four handled bits
			for (int s = 0; s < 4 && rmask; ++s)
				rmask &= rmask - 1;
		}
		unsigned umask = __ballot_sync(0xffffffffu, undecided);
		if (umask)
		{
			int leader = __ffs(umask) - 1;
			u32 off = 0;
			if (int(lane) == leader)
				off = atomicAdd(counter, u32(__popc(umask)));
			off = __shfl_sync(0xffffffffu, off, leader);
			if (undecided)
			{
				__stcg(reinterpret_cast<uint4*>(next + off + __popc(umask & ((1u << lane) - 1))), make_uint4(e.k, e.r0, e.r1, e.c));
				wave_publish_entry(e, any_nxt, src_nxt, tag + 1);
			}
		}
	}
}

// The same round with one candidate per 8-lane group (all lanes of the group hold the same record): used when the list is short
// compared with the grid, where a round costs one chain of dependent loads whatever its size. Nothing is serialised inside a
// warp, and the adjacency walk starts right after the four gathers.
DEVFN void wave_round_groups(const WaveArgs& a, const WaveEntry* cur, u32 n, WaveEntry* next, u32* counter, u32 tag, u32 tid, u32 nthreads)
{
	const u32 lane = threadIdx.x & 31;
	const u32 sub = lane >> 3, lane8 = lane & 7;
	const unsigned gmask = 0xffu << (sub * 8);
	const u64* any_cur = a.vmin_any[tag & 1];
	const u64* src_cur = a.vmin_src[tag & 1];
	u64* any_nxt = a.vmin_any[(tag + 1) & 1];
	u64* src_nxt = a.vmin_src[(tag + 1) & 1];
	for (u32 base = (tid >> 5) * 4; base < n; base += (nthreads >> 5) * 4)
	{
		const u32 i = base + sub;
		const bool valid = i < n;
		WaveEntry e = {0, 0, 0, 0};
		bool undecided = false;
		if (valid)
		{
			uint4 raw = __ldcg(reinterpret_cast<const uint4*>(cur + i));
			e.k = raw.x, e.r0 = raw.y, e.r1 = raw.z, e.c = raw.w;
			u32 l0 = __ldcg(&a.collapse_locked[e.r0]), l1 = __ldcg(&a.collapse_locked[e.r1]);
			u32 m0 = wave_min_cg(any_cur, e.r0, tag), m1 = wave_min_cg(any_cur, e.r1, tag);
			if (l0 | l1)
			{
				if (lane8 == 0)
					a.status[e.k] = Status_Locked;
			}
			else if (m0 != e.k || m1 != e.k)
				undecided = true;
			else
				undecided = wave_walk8(a, e.k, e.r0, e.r1, e.c, gmask, lane8, src_cur, tag) != 0;
		}
		__syncwarp();
		const bool keep = undecided && lane8 == 0;
		unsigned umask = __ballot_sync(0xffffffffu, keep);
		if (umask)
		{
			int leader = __ffs(umask) - 1;
			u32 off = 0;
			if (int(lane) == leader)
				off = atomicAdd(counter, u32(__popc(umask)));
			off = __shfl_sync(0xffffffffu, off, leader);
			if (keep)
			{
				__stcg(reinterpret_cast<uint4*>(next + off + __popc(umask & ((1u << lane) - 1))), make_uint4(e.k, e.r0, e.r1, e.c));
				wave_publish_entry(e, any_nxt, src_nxt, tag + 1);
			}
		}
	}
}

static const int WAVE_THREADS = 1024; // one CTA per SM: the grid barrier costs one atomic per CTA
static __global__ void __launch_bounds__(WAVE_THREADS, 1) k_wave_rounds(WaveArgs a)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const u32 gsize = gridDim.x * blockDim.x;
	const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x;
	volatile u32* vstate = a.state;
	u32 n = vstate[0];
	u32 tag = vstate[3];
	u32 round = 0;
	unsigned long long items = 0;
	bool tail = false;
	while (n != 0 && round < a.max_rounds)
	{
		if (gridDim.x > 1 && n <= a.tail_threshold)
		{
			tail = true;
			break;
		}
		++round;
		++tag;
		items += n;
		if (gtid == 0)
			vstate[(round + 1) % 3] = 0; // last read after the barrier of round - 2
		if (n <= gsize / 4) // at most two candidates per 8-lane group
			wave_round_groups(a, a.list[(round - 1) & 1], n, a.list[round & 1], a.state + round % 3, tag, gtid, gsize);
		else
			wave_round(a, a.list[(round - 1) & 1], n, a.list[round & 1], a.state + round % 3, tag, gtid, gsize);
		grid.sync();
		if (a.round_log && gtid == 0 && round < 256)
		{
			unsigned long long t;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
			a.round_log[round * 2] = n;
			a.round_log[round * 2 + 1] = u32(t);
		}
		n = vstate[round % 3];
	}
	if (tail)
	{
		// every CTA took the same decision (n was read after a grid barrier): the others are done, CTA 0 finishes alone
		if (blockIdx.x != 0)
			return;
		while (n != 0 && round < a.max_rounds)
		{
			++round;
			++tag;
			items += n;
			if (threadIdx.x == 0)
				vstate[(round + 1) % 3] = 0;
			wave_round_groups(a, a.list[(round - 1) & 1], n, a.list[round & 1], a.state + round % 3, tag, threadIdx.x, blockDim.x);
			__threadfence();
			__syncthreads();
			if (a.round_log && threadIdx.x == 0 && round < 256)
			{
				unsigned long long t;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
				a.round_log[round * 2] = n | 0x80000000u;
				a.round_log[round * 2 + 1] = u32(t);
			}
			n = vstate[round % 3];
			__syncthreads(); // nobody resets the counter of round + 2 (== round - 1) before everybody has read this one
		}
	}
	if (gtid == 0)
	{
		a.state[3] = tag;
		a.state[4] = round;
		a.state[5] = n;
		a.state[6] += round;
		a.state[7] = a.state[7] > round ? a.state[7] : round;
		*reinterpret_cast<unsigned long long*>(a.state + 8) = items;
	}
}
#endif



// ---- candidate windows -------------------------------------------------------------------------------------------------
// The serial scan of performEdgeCollapses stops long before the end of the sorted list (typically after 20-35 % of it),
// and a candidate's fate depends only on lower-ranked candidates. So the wavefront only runs over a per-group prefix
// window of the sorted list; if a group's stop position is not inside its window, the window is extended and the
// wavefront continues with the new candidates (the decided ones stay decided). The result is identical to running the
// wavefront over the whole list.
KERNEL k_window_init(GroupState* groups, u32 G)
{
	size_t g = GTID;
	if (g >= G)
		return;
	GroupState& gs = groups[g];
	u32 goal = gs.tri_count > gs.target_tris ? gs.tri_count - gs.target_tris : 0;
	u32 w = goal + goal / 8 + 64;
	gs.win_begin = gs.cand_begin;
	gs.win_end = gs.cand_begin + (w < gs.cand_count ? w : gs.cand_count);
}

// Only UNDECIDED candidates are listed: when one group extends its window, the windows of the other groups are unchanged
// and everything inside them has been decided by the previous run; running a decided candidate through the decide step
// again would find its own locks and relabel a performed collapse as locked.
KERNEL k_wave_window(const u32* __restrict__ sorted_cand, const u32* __restrict__ cand_group, const GroupState* __restrict__ groups, const u8* __restrict__ status, u32* list, u32* counter, u32 cand_total)
{
	size_t kk = GTID;
	bool take = false;
	if (kk < cand_total && status[kk] == Status_Undecided)
	{
		const GroupState& gs = groups[cand_group[sorted_cand[kk]]];
		take = u32(kk) >= gs.win_begin && u32(kk) < gs.win_end;
	}
#ifdef CLODB_EMU
	if (take)
		list[atomicAdd(counter, 1u)] = u32(kk);
#else
	unsigned mask = __ballot_sync(0xffffffffu, take);
	if (mask)
	{
		u32 lane = threadIdx.x & 31;
		u32 off = 0;
		if (lane == u32(__ffs(mask) - 1))
			off = atomicAdd(counter, u32(__popc(mask)));
		off = __shfl_sync(0xffffffffu, off, __ffs(mask) - 1);
		if (take)
			list[off + __popc(mask & ((1u << lane) - 1))] = u32(kk);
	}
#endif
}

// after k_cut_find: groups whose stop position is not inside their window get the next window
KERNEL k_window_check(GroupState* groups, u32 G, u32* any_extend)
{
	size_t g = GTID;
	if (g >= G)
		return;
	GroupState& gs = groups[g];
	if (!gs.active || gs.cand_count == 0)
		return;
	u32 end = gs.cand_begin + gs.cand_count;
	if (gs.cut >= gs.win_end && gs.win_end < end)
	{
		u32 goal = gs.tri_count > gs.target_tris ? gs.tri_count - gs.target_tris : 0;
		u32 grow = goal / 2 + 64;
		gs.win_begin = gs.win_end;
		gs.win_end = end - gs.win_end < grow ? end : gs.win_end + grow;
		gs.cut = end;
		atomicOr(any_extend, 1u);
	}
}

// per window position: triangle weight / flip flag / tagged error for the prefix scans behind the serial early-outs
KERNEL k_cut_inputs(const u32* __restrict__ woff, const GroupState* __restrict__ groups, u32 G, u32 W, const u32* __restrict__ sorted_cand, const u8* __restrict__ status, const u32* __restrict__ cand_v0,
    const float* __restrict__ cand_error, const u8* __restrict__ kind, u32* tri_weight, u32* flip_flag, u64* tagged_error)
{
	size_t jj = GTID;
	if (jj >= W)
		return;
	u32 j = u32(jj);
	u32 g = window_group(woff, G, j);
	u32 k = groups[g].cand_begin + (j - woff[g]);
	u32 c = sorted_cand[k];
	u8 st = status[k];
	tri_weight[j] = st == Status_Performed ? (kind[cand_v0[c]] == Kind_Border ? 1u : 2u) : 0u;
	flip_flag[j] = st == Status_Flip ? 1u : 0u;
	tagged_error[j] = (u64(g + 1) << 32) | (st == Status_Performed ? u64(__float_as_uint(cand_error[c])) : 0ull);
}

// evaluates the break conditions of performEdgeCollapses (simplifier.cpp:1533-1557) at every examined sorted position
KERNEL k_cut_find(const u32* __restrict__ woff, u32 G, u32 W, const u32* __restrict__ sorted_cand, const u8* __restrict__ status, const float* __restrict__ cand_error, GroupState* groups,
    const u32* __restrict__ tri_prefix, const u32* __restrict__ flip_prefix, const u64* __restrict__ error_prefix)
{
	size_t jj = GTID;
	if (jj >= W)
		return;
	u32 j = u32(jj);
	u32 g = window_group(woff, G, j);
	GroupState& gs = groups[g];
	u32 wbase = woff[g];
	u32 base = gs.cand_begin;
	u32 k = base + (j - wbase);
	u32 c = sorted_cand[k];
	u32 goal = gs.tri_count - gs.target_tris; // triangle_collapse_goal
	u32 tris_before = tri_prefix[j] - tri_prefix[wbase];
	u32 flips_before = flip_prefix[j] - flip_prefix[wbase];
	float error = cand_error[c];

	bool stop = false;
	if (tris_before >= goal)
		stop = true;
	else
	{
		u32 edge_goal = goal / 2 + flips_before;
		float error_goal = edge_goal < gs.cand_count ? 1.5f * cand_error[sorted_cand[base + edge_goal]] : FLT_MAX;
		u64 pe = error_prefix[j];
		float result_error = gs.result_error;
		if (u32(pe >> 32) == g + 1)
		{
			float m = __uint_as_float(u32(pe));
			result_error = result_error < m ? m : result_error;
		}
		if (error > error_goal && error > result_error && tris_before > goal / 6)
			stop = true;
	}
	// undecided candidates (round cap reached) also end the pass for their group: nothing after them is trustworthy
	if (status[k] == Status_Undecided)
		stop = true;
	// every position past the first stop also stops: test the current minimum first so only a few atomics are issued
	if (stop && k < *reinterpret_cast<volatile u32*>(&gs.cut))
		atomicMin(&gs.cut, k);
}

// roll back collapses past the cut; count the accepted ones per group, fold their errors into the group result and list them
// (accepted[0..*accepted_count), any order: a vertex takes part in at most one performed collapse per pass, so the quadric
// updates of different accepted collapses touch different vertices). Counters are updated once per (warp, group).
KERNEL k_cut_apply(const u32* __restrict__ woff, u32 G, u32 W, const u32* __restrict__ sorted_cand, u8* status, const u32* __restrict__ cand_v0, const float* __restrict__ cand_error,
    const u32* __restrict__ wedge, const u8* __restrict__ kind, GroupState* groups, u32* group_collapses, u32* group_error_bits, u32* collapse_remap, u32* accepted, u32* accepted_count)
{
	size_t jj = GTID;
	bool accept = false;
	u32 g = 0, k = 0, error_bits = 0;
	if (jj < W)
	{
		u32 j = u32(jj);
		g = window_group(woff, G, j);
		k = groups[g].cand_begin + (j - woff[g]);
		if (status[k] == Status_Performed)
		{
			u32 c = sorted_cand[k];
			if (k >= groups[g].cut)
			{
				u32 i0 = cand_v0[c];
				u8 kd = kind[i0];
				if (kd == Kind_Complex)
				{
					u32 v = i0;
					do
					{
						collapse_remap[v] = v;
						v = wedge[v];
					} while (v != i0);
				}
				else if (kd == Kind_Seam)
				{
					collapse_remap[i0] = i0;
					collapse_remap[wedge[i0]] = wedge[i0];
				}
				else
					collapse_remap[i0] = i0;
				status[k] = Status_Locked;
			}
			else
			{
				accept = true;
				error_bits = __float_as_uint(cand_error[c]);
			}
		}
	}
#ifdef CLODB_EMU
	if (accept)
	{
		accepted[atomicAdd(accepted_count, 1u)] = k;
		atomicAdd(&group_collapses[g], 1u);
		atomicMax(&group_error_bits[g], error_bits);
	}
#else
	const u32 lane = threadIdx.x & 31;
	unsigned remaining = __ballot_sync(0xffffffffu, accept);
	if (remaining)
	{
		int leader = __ffs(remaining) - 1;
		u32 off = 0;
		if (int(lane) == leader)
			off = atomicAdd(accepted_count, u32(__popc(remaining)));
		off = __shfl_sync(0xffffffffu, off, leader);
		if (accept)
			accepted[off + __popc(remaining & ((1u << lane) - 1))] = k;
	}
	while (remaining)
	{
		int leader = __ffs(remaining) - 1;
		u32 g0 = __shfl_sync(0xffffffffu, g, leader);
		unsigned same = __ballot_sync(0xffffffffu, accept && g == g0);
		if (accept && g == g0)
		{
			u32 mx = __reduce_max_sync(same, error_bits);
			if (int(lane) == leader)
			{
				atomicAdd(&group_collapses[g0], u32(__popc(same)));
				atomicMax(&group_error_bits[g0], mx);
			}
		}
		remaining &= ~same;
	}
#endif
}

// updateQuadrics (simplifier.cpp:1658-1698), one thread per accepted collapse; wedges are merged in ascending vertex order
KERNEL k_update_quadrics(const u32* __restrict__ accepted, const u32* __restrict__ accepted_count, const u32* __restrict__ sorted_cand, const u32* __restrict__ cand_v0,
    const u32* __restrict__ remap, const u32* __restrict__ wedge, const u32* __restrict__ collapse_remap, Quadric* vertex_quadrics, Quadric* attribute_quadrics, QuadricGrad* attribute_gradients, u32 attribute_count)
{
	// one thread per ACCEPTED collapse (the list k_cut_apply compacted): every thread that survives the bound check has work, so the
	// chains of dependent scattered loads of many collapses are in flight together. The grid covers the examined prefixes (an
	// upper bound known on the host); the list length stays on the device.
	size_t jj = GTID;
	if (jj >= *accepted_count)
		return;
	u32 k = accepted[jj];
	u32 i0 = cand_v0[sorted_cand[k]];
	u32 r0 = remap[i0];
	u32 r1 = remap[collapse_remap[r0]];
	quadric_add(vertex_quadrics[r1], vertex_quadrics[r0]);
	if (!attribute_count)
		return;
	// visit the wedge ring of r0 in ascending vertex id: r0 is the smallest, the ring links descend from the largest
	u32 ring[64];
	u32 n = 0;
	u32 v = r0;
	do
	{
		if (n < 64)
			ring[n++] = v;
		v = wedge[v];
	} while (v != r0);
	// ring = r0, largest, ..., smallest-but-r0 ; ascending order = r0 then the tail reversed
	for (u32 j = 0; j < n; ++j)
	{
		u32 w = j == 0 ? ring[0] : ring[n - j];
		u32 t = collapse_remap[w];
		if (t == w)
			continue;
		quadric_add(attribute_quadrics[t], attribute_quadrics[w]);
		for (u32 a = 0; a < attribute_count; ++a)
		{
			QuadricGrad& G = attribute_gradients[size_t(t) * attribute_count + a];
			const QuadricGrad& R = attribute_gradients[size_t(w) * attribute_count + a];
			G.gx += R.gx;
			G.gy += R.gy;
			G.gz += R.gz;
			G.gw += R.gw;
		}
	}
}

KERNEL k_remap_loops(const u32* __restrict__ loop_in, u32* loop_out, const u32* __restrict__ collapse_remap, u32 vertex_count)
{
	size_t ii = GTID;
	if (ii >= vertex_count)
		return;
	u32 i = u32(ii);
	u32 l = loop_in[i];
	u32 out = l;
	if (l != NONE)
	{
		u32 r = collapse_remap[l];
		if (i == r)
			out = (loop_in[l] != NONE) ? collapse_remap[loop_in[l]] : NONE;
		else
			out = r;
	}
	loop_out[i] = out;
}

// remapIndexBuffer: apply collapses, drop triangles that became degenerate by position
KERNEL k_remap_triangles(u32* idx, const u32* __restrict__ collapse_remap, const u32* __restrict__ remap, const u32* __restrict__ tri_group, const u32* __restrict__ group_collapses, u32* keep, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	// groups that performed no collapse this pass leave the loop before remapIndexBuffer (simplifier.cpp:2483-2502)
	if (group_collapses[tri_group[t]] == 0)
	{
		keep[t] = 1;
		return;
	}
	u32 v0 = collapse_remap[idx[t * 3 + 0]];
	u32 v1 = collapse_remap[idx[t * 3 + 1]];
	u32 v2 = collapse_remap[idx[t * 3 + 2]];
	idx[t * 3 + 0] = v0;
	idx[t * 3 + 1] = v1;
	idx[t * 3 + 2] = v2;
	u32 r0 = remap[v0], r1 = remap[v1], r2 = remap[v2];
	keep[t] = (r0 != r1 && r0 != r2 && r1 != r2) ? 1u : 0u;
}

KERNEL k_compact_triangles(const u32* __restrict__ idx, const u32* __restrict__ tri_group, const u32* __restrict__ keep_scanned, u32 total, u32* idx_out, u32* tri_group_out, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 pos = keep_scanned[t];
	u32 next = t + 1 < T ? keep_scanned[t + 1] : total;
	if (next == pos)
		return;
	idx_out[pos * 3 + 0] = idx[t * 3 + 0];
	idx_out[pos * 3 + 1] = idx[t * 3 + 1];
	idx_out[pos * 3 + 2] = idx[t * 3 + 2];
	tri_group_out[pos] = tri_group[t];
}

KERNEL k_group_after_pass(GroupState* groups, const u32* __restrict__ keep_scanned, u32 total, const u32* __restrict__ group_collapses, const u32* __restrict__ group_error_bits, u32 T, u32 G, u32* any_active)
{
	size_t g = GTID;
	if (g >= G)
		return;
	GroupState& gs = groups[g];
	u32 b = gs.tri_begin < T ? keep_scanned[gs.tri_begin] : total;
	u32 e = gs.tri_begin + gs.tri_count < T ? keep_scanned[gs.tri_begin + gs.tri_count] : total;
	bool was_active = gs.active != 0;
	gs.tri_begin = b;
	gs.tri_count = e - b;
	if (!was_active)
		return;
	float m = __uint_as_float(group_error_bits[g]);
	gs.result_error = gs.result_error < m ? m : gs.result_error;
	// loop conditions of meshopt_simplifyEdge (simplifier.cpp:2474-2520)
	bool more = gs.cand_count != 0 && group_collapses[g] != 0 && gs.tri_count > gs.target_tris;
	gs.active = more ? 1u : 0u;
	if (more)
		atomicOr(any_active, 1u);
}

KERNEL k_init_groups(GroupState* groups, const u32* __restrict__ group_tri_offset, u32 G, float ratio, u32* any_active)
{
	size_t g = GTID;
	if (g >= G)
		return;
	GroupState gs;
	gs.tri_begin = group_tri_offset[g];
	gs.tri_count = group_tri_offset[g + 1] - group_tri_offset[g];
	// clusterlod.h:713-715: size_t((merged.size() / 3) * simplify_ratio) * 3, at least one triangle
	u32 target = u32(float(gs.tri_count) * ratio);
	if (gs.tri_count && target < 1)
		target = 1;
	gs.target_tris = target;
	gs.active = gs.tri_count > target ? 1u : 0u;
	gs.cand_begin = gs.cand_count = 0;
	gs.cut = 0;
	gs.result_error = 0.f;
	groups[g] = gs;
	if (gs.active)
		atomicOr(any_active, 1u);
}

KERNEL k_finalize_output(const u32* __restrict__ idx, const u32* __restrict__ sv_global, u32* out, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	out[c] = sv_global[idx[c]];
}

KERNEL k_group_results(const GroupState* __restrict__ groups, const float* __restrict__ group_extent, u32* out_tri_offset, float* out_error, u32 G)
{
	size_t g = GTID;
	if (g >= G)
		return;
	out_tri_offset[g] = groups[g].tri_begin;
	if (g + 1 == G)
		out_tri_offset[G] = groups[g].tri_begin + groups[g].tri_count;
	// ErrorAbsolute: sqrtf(result_error) * vertex_scale (simplifier.cpp:2607-2609)
	out_error[g] = sqrtf(groups[g].result_error) * group_extent[g];
}

// =====================================================================================================================
// Sloppy fallback: clod::simplifyFallback (clusterlod.h:567-599) -> meshopt_simplifySloppy (simplifier.cpp:2644-2774) on the
// de-indexed group (one "vertex" per corner), + meshopt_simplifyScale (:2934-2944). Callees: rescalePositions :550-606,
// computeVertexIds :2083-2101, countTriangles :2103-2117, fillVertexCells :2119-2143, fillCellQuadrics :2165-2193,
// fillCellRemap :2233-2248, filterTriangles :2278-2321, interpolate :2323-2329.
// All groups of the level whose edge-collapse result misses the target go through it together: their corners are laid out
// back to back, every step is one data-parallel kernel over all of them, and the per-group grid-size search advances in
// lock step with its state kept on the device (one host read per search round, not per group). The reference's sequential
// semantics are recovered by
//   * lowest-corner-wins hash tables (cell numbering, duplicate-triangle filter; one table region per group) + order
//     preserving compaction,
//   * per-cell quadric sums taken over the cell's corners in ascending corner order (the serial accumulation order),
//   * (error, corner) lexicographic argmin per cell.
struct SlGroup
{
	u32 tri_begin, tri_count; // in the concatenated triangle space of the fallback groups
	u32 src_tri_begin;        // first triangle of the group in the level's merged index list
	u32 target_tris;
	u32 mm[6]; // order keys of min xyz / max xyz
	float minv[3];
	float extent, scale;
	int min_grid, max_grid, next_grid, cur_grid;
	u32 min_triangles, max_triangles;
	u32 count; // non-degenerate triangles counted in the running search round
	u32 phase; // 0 = initial count (grid 1) pending, 1 = searching, 2 = done
	int pass;
	u32 cell_table_base, cell_table_mask, tri_table_base, tri_table_mask;
	u32 max_error_bits;
	u32 kept;
};

// float -> int as x86-64 converts it (cvttss2si): out-of-range and NaN give INT_MIN. The reference's search runs on the host
// CPU; CUDA's conversion saturates instead, which would steer the clamp the other way.
DEVFN int x86_float_to_int(float f)
{
	if (!(f > -2147483648.f && f < 2147483648.f))
		return int(0x80000000u);
	return int(f);
}

KERNEL k_sl_gather_corners(const u32* __restrict__ gtri, const u32* __restrict__ tri_group, const SlGroup* __restrict__ groups, u32 n, u32* corner_vertex)
{
	size_t i = GTID;
	if (i >= n)
		return;
	u32 t = u32(i / 3);
	const SlGroup& g = groups[tri_group[t]];
	corner_vertex[i] = gtri[(size_t(g.src_tri_begin) + (t - g.tri_begin)) * 3 + (i - size_t(t) * 3)];
}

KERNEL k_sl_minmax(const u32* __restrict__ corner_vertex, const u32* __restrict__ tri_group, const float* __restrict__ positions, u32 n, SlGroup* groups)
{
	size_t i = GTID;
	if (i >= n)
		return;
	const float* p = positions + size_t(corner_vertex[i]) * 3;
	u32 g = tri_group[i / 3];
#ifndef CLODB_EMU
	unsigned active = __activemask();
	u32 g0 = __shfl_sync(active, g, __ffs(active) - 1);
	if (active == 0xffffffffu && __all_sync(active, g == g0))
	{
		for (int k = 0; k < 3; ++k)
		{
			u32 key = float_order_key(p[k]);
			u32 lo = key, hi = key;
			for (int d = 16; d >= 1; d >>= 1)
			{
				u32 olo = __shfl_xor_sync(0xffffffffu, lo, d), ohi = __shfl_xor_sync(0xffffffffu, hi, d);
				lo = olo < lo ? olo : lo;
				hi = ohi > hi ? ohi : hi;
			}
			if ((threadIdx.x & 31) == 0)
			{
				atomicMin(&groups[g].mm[k], lo);
				atomicMax(&groups[g].mm[3 + k], hi);
			}
		}
		return;
	}
#endif
	for (int k = 0; k < 3; ++k)
	{
		atomicMin(&groups[g].mm[k], float_order_key(p[k]));
		atomicMax(&groups[g].mm[3 + k], float_order_key(p[k]));
	}
}

// rescalePositions extent/scale + the start of the grid-size search (simplifier.cpp:2690-2703)
KERNEL k_sl_setup(SlGroup* groups, u32 S)
{
	size_t s = GTID;
	if (s >= S)
		return;
	SlGroup& g = groups[s];
	float extent = 0.f;
	for (int k = 0; k < 3; ++k)
	{
		float mn = float_from_order_key(g.mm[k]), mx = float_from_order_key(g.mm[3 + k]);
		g.minv[k] = mn;
		extent = (mx - mn) < extent ? extent : (mx - mn);
	}
	g.extent = extent;
	g.scale = extent == 0 ? 0.f : 1.f / extent;
	size_t target_index_count = size_t(g.target_tris) * 3;
	size_t target_cell_count = target_index_count / 6;
	g.min_grid = 1;
	g.max_grid = 1025;
	g.max_triangles = g.tri_count;
	g.min_triangles = 0;
	g.next_grid = x86_float_to_int(sqrtf(float(target_cell_count)) + 0.5f);
	g.cur_grid = 1;
	g.count = 0;
	g.phase = 0;
	g.pass = 0;
	g.max_error_bits = 0;
	g.kept = 0;
}

KERNEL k_sl_rescale(const u32* __restrict__ corner_vertex, const u32* __restrict__ tri_group, const SlGroup* __restrict__ groups, const float* __restrict__ positions, u32 n, Vector3* vpos)
{
	size_t i = GTID;
	if (i >= n)
		return;
	const SlGroup& g = groups[tri_group[i / 3]];
	const float* p = positions + size_t(corner_vertex[i]) * 3;
	Vector3 r;
	r.x = (p[0] - g.minv[0]) * g.scale;
	r.y = (p[1] - g.minv[1]) * g.scale;
	r.z = (p[2] - g.minv[2]) * g.scale;
	vpos[i] = r;
}

// computeVertexIds for one corner; locked corners keep an id of their own (corner index inside the group)
DEVFN u32 sl_vertex_id(const Vector3& v, bool locked, u32 local_corner, int grid_size)
{
	float cell_scale = float(grid_size - 1);
	int xi = int(v.x * cell_scale + 0.5f);
	int yi = int(v.y * cell_scale + 0.5f);
	int zi = int(v.z * cell_scale + 0.5f);
	if (locked)
		return (1u << 30) | local_corner;
	return (u32(xi) << 20) | (u32(yi) << 10) | u32(zi);
}

// one search round: countTriangles with every searching group's current grid size
KERNEL k_sl_count_triangles(const Vector3* __restrict__ vpos, const u32* __restrict__ corner_vertex, const u8* __restrict__ locks, const u32* __restrict__ tri_group, SlGroup* groups, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 gi = tri_group[t];
	const SlGroup& g = groups[gi];
	bool hit = false;
	if (g.phase != 2)
	{
		u32 local = u32(t - g.tri_begin) * 3;
		u32 id[3];
		for (int k = 0; k < 3; ++k)
			id[k] = sl_vertex_id(vpos[t * 3 + k], locks && (locks[corner_vertex[t * 3 + k]] & 1) != 0, local + k, g.cur_grid);
		hit = (id[0] != id[1]) & (id[0] != id[2]) & (id[1] != id[2]);
	}
#ifndef CLODB_EMU
	// triangles of a group are contiguous: one atomic per warp and group
	unsigned active = __activemask();
	unsigned peers = __match_any_sync(active, gi);
	unsigned votes = __ballot_sync(active, hit) & peers;
	if (votes && (threadIdx.x & 31) == __ffs(peers) - 1)
		atomicAdd(&groups[gi].count, u32(__popc(votes)));
#else
	if (hit)
		atomicAdd(&groups[gi].count, 1u);
#endif
}

// three point interpolation of the grid-size search (simplifier.cpp:2323-2329)
DEVFN float sloppy_interpolate(float y, float x0, float y0, float x1, float y1, float x2, float y2)
{
	float num = (y1 - y) * (x1 - x2) * (x1 - x0) * (y2 - y0);
	float den = (y2 - y) * (x1 - x2) * (y0 - y1) + (y0 - y) * (x1 - x0) * (y1 - y2);
	return x1 + (den == 0.f ? 0.f : num / den);
}

// guided search for the grid size (simplifier.cpp:2697-2741; target_error = FLT_MAX => min_grid = 1), one step per round
KERNEL k_sl_search_step(SlGroup* groups, u32 S, u32* any_searching)
{
	size_t s = GTID;
	if (s >= S)
		return;
	SlGroup& g = groups[s];
	if (g.phase == 2)
		return;
	const int kInterpolationPasses = 5;
	size_t target_tris = (size_t(g.target_tris) * 3) / 3;
	u32 triangles = g.count;
	g.count = 0;
	if (g.phase == 0)
	{
		g.min_triangles = triangles;
		g.phase = 1;
	}
	else
	{
		int grid_size = g.cur_grid;
		float tip = sloppy_interpolate(float(target_tris), float(g.min_grid), float(g.min_triangles), float(grid_size), float(triangles), float(g.max_grid), float(g.max_triangles));
		if (triangles <= target_tris)
		{
			g.min_grid = grid_size;
			g.min_triangles = triangles;
		}
		else
		{
			g.max_grid = grid_size;
			g.max_triangles = triangles;
		}
		g.next_grid = (g.pass < kInterpolationPasses) ? x86_float_to_int(tip + 0.5f) : (g.min_grid + g.max_grid) / 2;
		g.pass++;
	}
	if (g.pass >= 10 + kInterpolationPasses || g.min_triangles >= target_tris || g.max_grid - g.min_grid <= 1)
	{
		g.phase = 2;
		g.cur_grid = g.min_grid;
		return;
	}
	int grid_size = g.next_grid;
	g.cur_grid = (grid_size <= g.min_grid) ? g.min_grid + 1 : (grid_size >= g.max_grid ? g.max_grid - 1 : grid_size);
	*any_searching = 1u;
}

// hash table regions: per group, a power of two >= 2 * corners (cells) and >= 2 * triangles (duplicate filter)
KERNEL k_sl_ids(const Vector3* __restrict__ vpos, const u32* __restrict__ corner_vertex, const u8* __restrict__ locks, const u32* __restrict__ tri_group, const SlGroup* __restrict__ groups, u32 n, u32* ids)
{
	size_t i = GTID;
	if (i >= n)
		return;
	const SlGroup& g = groups[tri_group[i / 3]];
	ids[i] = sl_vertex_id(vpos[i], locks && (locks[corner_vertex[i]] & 1) != 0, u32(i) - g.tri_begin * 3, g.min_grid);
}

DEVFN u32 sl_hash(u32 h)
{
	h ^= h >> 13;
	h *= 0x5bd1e995u;
	h ^= h >> 15;
	return h;
}

// cell table: lowest corner per distinct vertex id of a group (groups whose search ended at zero triangles are skipped)
KERNEL k_sl_cell_insert(const u32* __restrict__ ids, const u32* __restrict__ tri_group, const SlGroup* __restrict__ groups, u32 n, u32* table_id, u32* table_first, u32* slot_of)
{
	size_t i = GTID;
	if (i >= n)
		return;
	const SlGroup& g = groups[tri_group[i / 3]];
	slot_of[i] = 0xffffffffu;
	if (g.min_triangles == 0)
		return;
	u32 id = ids[i];
	u32 h = sl_hash(id) & g.cell_table_mask;
	for (;;)
	{
		u32 old = atomicCAS(&table_id[g.cell_table_base + h], 0xffffffffu, id);
		if (old == 0xffffffffu || old == id)
			break;
		h = (h + 1) & g.cell_table_mask;
	}
	atomicMin(&table_first[g.cell_table_base + h], u32(i));
	slot_of[i] = g.cell_table_base + h;
}

KERNEL k_sl_first_flags(const u32* __restrict__ slot_of, const u32* __restrict__ table_first, u32 n, u32* flags)
{
	size_t i = GTID;
	if (i >= n)
		return;
	flags[i] = (slot_of[i] != 0xffffffffu && table_first[slot_of[i]] == u32(i)) ? 1u : 0u;
}

KERNEL k_sl_assign_cells(const u32* __restrict__ slot_of, const u32* __restrict__ table_first, const u32* __restrict__ first_rank, u32 n, u32* cells)
{
	size_t i = GTID;
	if (i >= n)
		return;
	cells[i] = slot_of[i] != 0xffffffffu ? first_rank[table_first[slot_of[i]]] : 0xffffffffu;
}

// (cell, corner) entries of fillCellQuadrics: one per corner, or one (weight 3) for a triangle inside a single cell
KERNEL k_sl_quadric_entries(const u32* __restrict__ cells, u32 T, u32* entry_cell, u32* entry_corner)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 c0 = cells[t * 3 + 0], c1 = cells[t * 3 + 1], c2 = cells[t * 3 + 2];
	bool single = (c0 == c1) & (c0 == c2);
	entry_cell[t * 3 + 0] = c0;
	entry_cell[t * 3 + 1] = single ? 0xffffffffu : c1;
	entry_cell[t * 3 + 2] = single ? 0xffffffffu : c2;
	entry_corner[t * 3 + 0] = u32(t * 3 + 0);
	entry_corner[t * 3 + 1] = u32(t * 3 + 1);
	entry_corner[t * 3 + 2] = u32(t * 3 + 2);
}

KERNEL k_sl_cell_heads(const u32* __restrict__ sorted_cell, u32 n, u32* cell_begin, u32 cell_count)
{
	size_t i = GTID;
	if (i > n)
		return;
	u32 cur = i < n ? sorted_cell[i] : 0xffffffffu;
	u32 prev = i > 0 ? sorted_cell[i - 1] : 0xffffffffu;
	if (i == 0 || cur != prev)
	{
		if (cur != 0xffffffffu)
			cell_begin[cur] = u32(i);
		if (i > 0 && prev != 0xffffffffu && cur == 0xffffffffu)
			cell_begin[cell_count] = u32(i); // end of the valid entries
	}
	if (i == n && prev != 0xffffffffu)
		cell_begin[cell_count] = u32(n);
}

KERNEL k_sl_cell_quadrics(const u32* __restrict__ sorted_cell, const u32* __restrict__ sorted_corner, const u32* __restrict__ cell_begin, const u32* __restrict__ cells, const Vector3* __restrict__ vpos, u32 cell_count,
    u32 n, Quadric* cell_quadrics)
{
	size_t c = GTID;
	if (c >= cell_count)
		return;
	Quadric Q;
	quadric_zero(Q);
	for (u32 e = cell_begin[c]; e < n && sorted_cell[e] == u32(c); ++e)
	{
		u32 corner = sorted_corner[e];
		u32 t = corner / 3;
		u32 c0 = cells[t * 3 + 0], c1 = cells[t * 3 + 1], c2 = cells[t * 3 + 2];
		bool single = (c0 == c1) & (c0 == c2);
		Quadric R;
		quadric_from_triangle(R, vpos[t * 3 + 0], vpos[t * 3 + 1], vpos[t * 3 + 2], single ? 3.f : 1.f);
		quadric_add(Q, R);
	}
	cell_quadrics[c] = Q;
}

KERNEL k_sl_cell_remap(const u32* __restrict__ cells, const Quadric* __restrict__ cell_quadrics, const Vector3* __restrict__ vpos, u32 n, u64* cell_best)
{
	size_t i = GTID;
	if (i >= n)
		return;
	u32 cell = cells[i];
	if (cell == 0xffffffffu)
		return;
	float error = quadric_error(cell_quadrics[cell], vpos[i]);
	u64 key = (u64(__float_as_uint(error)) << 32) | u64(u32(i));
	atomicMin(reinterpret_cast<unsigned long long*>(&cell_best[cell]), (unsigned long long)key);
}

KERNEL k_sl_max_error(const u64* __restrict__ cell_best, const u32* __restrict__ tri_group, u32 cell_count, SlGroup* groups)
{
	size_t c = GTID;
	if (c >= cell_count)
		return;
	u64 best = cell_best[c];
	atomicMax(&groups[tri_group[u32(best) / 3]].max_error_bits, u32(best >> 32));
}

// rotated (lowest corner id first) output triple per non-degenerate triangle; duplicates keep their first occurrence
KERNEL k_sl_triangles(const u32* __restrict__ cells, const u64* __restrict__ cell_best, const u32* __restrict__ tri_group, const SlGroup* __restrict__ groups, u32 T, u32* triple, u64* table_key, u32* table_first, u32* slot_of)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 c0 = cells[t * 3 + 0], c1 = cells[t * 3 + 1], c2 = cells[t * 3 + 2];
	slot_of[t] = 0xffffffffu;
	if (c0 == 0xffffffffu || !(c0 != c1 && c0 != c2 && c1 != c2))
		return;
	u32 a = u32(cell_best[c0]), b = u32(cell_best[c1]), c = u32(cell_best[c2]);
	if (b < a && b < c)
	{
		u32 tmp = a;
		a = b, b = c, c = tmp;
	}
	else if (c < a && c < b)
	{
		u32 tmp = c;
		c = b, b = a, a = tmp;
	}
	triple[t * 3 + 0] = a;
	triple[t * 3 + 1] = b;
	triple[t * 3 + 2] = c;
	// corner ids inside a group are < 2^21 (checked by the caller), so the group-local triple packs exactly into one 64-bit key
	const SlGroup& g = groups[tri_group[t]];
	u32 base = g.tri_begin * 3;
	u32 la = a - base, lb = b - base, lc = c - base;
	u32 h = ((la * 73856093u) ^ (lb * 19349663u) ^ (lc * 83492791u)) & g.tri_table_mask;
	u64 key = (u64(la) << 42) | (u64(lb) << 21) | u64(lc);
	for (;;)
	{
		u64 old = atomicCAS(reinterpret_cast<unsigned long long*>(&table_key[g.tri_table_base + h]), ~0ull, (unsigned long long)key);
		if (old == ~0ull || old == key)
			break;
		h = (h + 1) & g.tri_table_mask;
	}
	atomicMin(&table_first[g.tri_table_base + h], u32(t));
	slot_of[t] = g.tri_table_base + h;
}

KERNEL k_sl_keep_flags(const u32* __restrict__ slot_of, const u32* __restrict__ table_first, u32 T, u32* keep)
{
	size_t t = GTID;
	if (t >= T)
		return;
	keep[t] = (slot_of[t] != 0xffffffffu && table_first[slot_of[t]] == u32(t)) ? 1u : 0u;
}

// kept triangles of all fallback groups, compacted in order (so group by group): global vertex ids
KERNEL k_sl_emit(const u32* __restrict__ triple, const u32* __restrict__ keep_scanned, u32 total, const u32* __restrict__ corner_vertex, u32 T, u32* out_tri)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 pos = keep_scanned[t];
	u32 next = t + 1 < T ? keep_scanned[t + 1] : total;
	if (next == pos)
		return;
	out_tri[size_t(pos) * 3 + 0] = corner_vertex[triple[t * 3 + 0]];
	out_tri[size_t(pos) * 3 + 1] = corner_vertex[triple[t * 3 + 1]];
	out_tri[size_t(pos) * 3 + 2] = corner_vertex[triple[t * 3 + 2]];
}

KERNEL k_sl_group_kept(const u32* __restrict__ keep_scanned, u32 total, u32 T, SlGroup* groups, u32 S)
{
	size_t s = GTID;
	if (s >= S)
		return;
	SlGroup& g = groups[s];
	u32 end = g.tri_begin + g.tri_count;
	g.kept = (end < T ? keep_scanned[end] : total) - keep_scanned[g.tri_begin];
}

// level output after the fallback: group g takes its triangles either from the edge-collapse result or from the compacted
// fallback output (src_is_fallback), new_offset gives the reassembled layout
KERNEL k_sl_assemble(const u32* __restrict__ new_offset, const u32* __restrict__ src_begin, const u8* __restrict__ src_is_fallback, u32 G, const u32* __restrict__ collapsed_tri, const u32* __restrict__ fallback_tri, u32* out_tri, u32 total)
{
	size_t j = GTID;
	if (j >= total)
		return;
	u32 lo = 0, hi = G;
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (new_offset[mid] <= u32(j))
			lo = mid;
		else
			hi = mid;
	}
	const u32* src = (src_is_fallback[lo] ? fallback_tri : collapsed_tri) + (size_t(src_begin[lo]) + (j - new_offset[lo])) * 3;
	out_tri[j * 3 + 0] = src[0];
	out_tri[j * 3 + 1] = src[1];
	out_tri[j * 3 + 2] = src[2];
}

static float host_uint_as_float(u32 u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

struct SloppyResult
{
	u32* tri = nullptr; // kept triangles of the fallback groups back to back, in group order (temp arena)
	std::vector<u32> kept;
	std::vector<float> error; // absolute error per fallback group (before the sloppy error factor)
};

// sel: the groups of the level that take the fallback (ascending); group_tri_offset_host: the level's merged index list layout.
static SloppyResult sloppy_groups(const u32* gtri, const u32* group_tri_offset_host, const std::vector<u32>& sel, const std::vector<u32>& target_tris, const DeviceMesh& mesh, const u8* locks, Arena& temp)
{
	SloppyResult res;
	u32 S = u32(sel.size());
	std::vector<SlGroup> host(S);
	u64 T64 = 0, cell_table_total = 0, tri_table_total = 0;
	for (u32 s = 0; s < S; ++s)
	{
		SlGroup& g = host[s];
		memset(&g, 0, sizeof(g));
		u32 Tg = group_tri_offset_host[sel[s] + 1] - group_tri_offset_host[sel[s]];
		if (size_t(Tg) * 3 >= (1u << 21))
			throw Error("clodb200: sloppy fallback supports groups of up to 699050 triangles");
		g.tri_begin = u32(T64);
		g.tri_count = Tg;
		g.src_tri_begin = group_tri_offset_host[sel[s]];
		g.target_tris = target_tris[s];
		for (int k = 0; k < 3; ++k)
			g.mm[k] = 0xffffffffu;
		size_t cells = 1, tris = 1;
		while (cells < size_t(Tg) * 6)
			cells <<= 1;
		while (tris < size_t(Tg) * 2)
			tris <<= 1;
		g.cell_table_base = u32(cell_table_total);
		g.cell_table_mask = u32(cells - 1);
		g.tri_table_base = u32(tri_table_total);
		g.tri_table_mask = u32(tris - 1);
		cell_table_total += cells;
		tri_table_total += tris;
		T64 += Tg;
	}
	if (T64 * 3 >= (u64(1) << 32) || cell_table_total >= (u64(1) << 32))
		throw Error("clodb200: sloppy fallback batch exceeds 2^32 corners");
	u32 T = u32(T64), n = T * 3;

	SlGroup* groups = temp.alloc<SlGroup>(S);
	dev_h2d(groups, host.data(), size_t(S) * sizeof(SlGroup));
	u32* group_tri_begin = temp.alloc<u32>(size_t(S) + 1);
	{
		std::vector<u32> begins(size_t(S) + 1);
		for (u32 s = 0; s < S; ++s)
			begins[s] = host[s].tri_begin;
		begins[S] = T;
		dev_h2d(group_tri_begin, begins.data(), begins.size() * sizeof(u32));
	}
	u32* scalars = temp.alloc<u32>(8);
	u32* tri_group = temp.alloc<u32>(T);
	u32* corner_vertex = temp.alloc<u32>(n);
	Vector3* vpos = temp.alloc<Vector3>(n);
	res.tri = temp.alloc<u32>(size_t(n) + 3);
	LAUNCH(k_tri_group, T, group_tri_begin, S, tri_group, T);
	LAUNCH(k_sl_gather_corners, n, gtri, tri_group, groups, n, corner_vertex);

	// rescalePositions over the de-indexed subsets
	LAUNCH(k_sl_minmax, n, corner_vertex, tri_group, mesh.positions, n, groups);
	LAUNCH(k_sl_setup, S, groups, S);
	LAUNCH(k_sl_rescale, n, corner_vertex, tri_group, groups, mesh.positions, n, vpos);

	// grid-size search: initial count at grid 1 + at most 15 guided rounds, all groups in lock step
	for (int round = 0; round < 17; ++round)
	{
		LAUNCH(k_sl_count_triangles, T, vpos, corner_vertex, locks, tri_group, groups, T);
		dev_memset(scalars, 0, sizeof(u32));
		LAUNCH(k_sl_search_step, S, groups, S, scalars);
		if (dev_read(scalars) == 0)
			break;
	}

	ArenaScope scope(temp);
	// cells in first-occurrence order
	u32* ids = temp.alloc<u32>(n);
	u32* table_id = temp.alloc<u32>(cell_table_total);
	u32* table_first = temp.alloc<u32>(cell_table_total);
	u32* slot_of = temp.alloc<u32>(n);
	u32* flags = temp.alloc<u32>(size_t(n) + 1);
	u32* cells = temp.alloc<u32>(n);
	LAUNCH(k_sl_ids, n, vpos, corner_vertex, locks, tri_group, groups, n, ids);
	dev_memset(table_id, 0xff, cell_table_total * 4);
	dev_memset(table_first, 0xff, cell_table_total * 4);
	LAUNCH(k_sl_cell_insert, n, ids, tri_group, groups, n, table_id, table_first, slot_of);
	LAUNCH(k_sl_first_flags, n, slot_of, table_first, n, flags);
	exclusive_scan_u32(flags, flags, n, scalars + 6, temp);
	u32 cell_count = dev_read(scalars + 6);
	LAUNCH(k_sl_assign_cells, n, slot_of, table_first, flags, n, cells);

	u32 kept_total = 0;
	if (cell_count)
	{
		// per-cell quadrics, summed in ascending corner order
		u32* entry_cell = ids; // ids, slot_of are dead from here on
		u32* entry_corner = slot_of;
		u32* entry_cell_tmp = temp.alloc<u32>(n);
		u32* entry_corner_tmp = temp.alloc<u32>(n);
		u32* cell_begin = temp.alloc<u32>(size_t(cell_count) + 1);
		Quadric* cell_quadrics = temp.alloc<Quadric>(cell_count);
		u64* cell_best = temp.alloc<u64>(cell_count);
		LAUNCH(k_sl_quadric_entries, T, cells, T, entry_cell, entry_corner);
		// keys are cell numbers or 0xffffffff (entries that do not count): all 32 bits take part
		radix_sort_pairs<u32>(entry_cell, entry_cell_tmp, entry_corner, entry_corner_tmp, n, 0, 32, temp);
		dev_memset(cell_begin, 0, (size_t(cell_count) + 1) * 4);
		LAUNCH(k_sl_cell_heads, size_t(n) + 1, entry_cell, n, cell_begin, cell_count);
		LAUNCH(k_sl_cell_quadrics, cell_count, entry_cell, entry_corner, cell_begin, cells, vpos, cell_count, n, cell_quadrics);

		// best vertex per cell, error
		dev_memset(cell_best, 0xff, size_t(cell_count) * 8);
		LAUNCH(k_sl_cell_remap, n, cells, cell_quadrics, vpos, n, cell_best);
		LAUNCH(k_sl_max_error, cell_count, cell_best, tri_group, cell_count, groups);

		// triangles: drop degenerate and duplicate ones, keep order
		u32* triple = entry_cell_tmp;
		u64* tt_key = temp.alloc<u64>(tri_table_total);
		u32* tt_first = temp.alloc<u32>(tri_table_total);
		u32* tslot = temp.alloc<u32>(T);
		u32* keep = temp.alloc<u32>(size_t(T) + 1);
		dev_memset(tt_key, 0xff, tri_table_total * 8);
		dev_memset(tt_first, 0xff, tri_table_total * 4);
		LAUNCH(k_sl_triangles, T, cells, cell_best, tri_group, groups, T, triple, tt_key, tt_first, tslot);
		LAUNCH(k_sl_keep_flags, T, tslot, tt_first, T, keep);
		exclusive_scan_u32(keep, keep, T, scalars + 6, temp);
		kept_total = dev_read(scalars + 6);
		LAUNCH(k_sl_emit, T, triple, keep, kept_total, corner_vertex, T, res.tri);
		LAUNCH(k_sl_group_kept, S, keep, kept_total, T, groups, S);
	}

	host = dev_download(groups, S);
	res.kept.resize(S);
	res.error.resize(S);
	for (u32 s = 0; s < S; ++s)
	{
		res.kept[s] = host[s].kept;
		// a search that ends at zero triangles returns the empty mesh with error 1 (simplifier.cpp:2743-2750)
		res.error[s] = host[s].min_triangles == 0 ? 1.f * host[s].extent : sqrtf(host_uint_as_float(host[s].max_error_bits)) * host[s].extent;
	}
	return res;
}

// ---------------------------------------------------------------------------------------------------------------------
static void build_adjacency(const u32* idx, size_t corners, const u32* remap, u32 vertex_count, u32* adj_off, u32* adj_corner, bool sorted, Arena& temp)
{
	ArenaScope scope(temp);
	u32* cursor = temp.alloc<u32>(vertex_count);
	dev_memset(adj_off, 0, (size_t(vertex_count) + 1) * sizeof(u32));
	dev_memset(cursor, 0, size_t(vertex_count) * sizeof(u32));
	LAUNCH(k_adj_count, corners, idx, remap, adj_off, corners);
	exclusive_scan_u32(adj_off, adj_off, size_t(vertex_count) + 1, nullptr, temp);
	LAUNCH(k_adj_fill, corners, idx, remap, adj_off, cursor, adj_corner, corners);
	if (sorted)
		LAUNCH(k_adj_sort, vertex_count, adj_off, adj_corner, vertex_count);
}

thread_local SimplifyStats g_simplify_stats;

SimplifyOutput simplify_groups(const u32* gtri, const u32* group_tri_offset_host, u32 G, const DeviceMesh& mesh, const u32* global_remap, const u8* locks, const Config& config, Workspace& ws)
{
	SimplifyOutput out;
	out.group_count = G;
	if (G == 0)
		return out;
	u32 T = group_tri_offset_host[G];
	size_t corners = size_t(T) * 3;

	out.tri = ws.persist.alloc<u32>(corners);
	out.group_tri_offset = ws.persist.alloc<u32>(size_t(G) + 1);
	out.group_error = ws.persist.alloc<float>(G);

	Arena& temp = ws.temp;
	ArenaScope scope(temp);

	u32* group_tri_offset = temp.alloc<u32>(size_t(G) + 1);
	dev_h2d(group_tri_offset, group_tri_offset_host, (size_t(G) + 1) * sizeof(u32));
	u32* scalars = temp.alloc<u32>(8); // [0] totals, [1] any_active, [2] undecided

	GroupState* groups = temp.alloc<GroupState>(G);
	const size_t main_mark = temp.mark(); // allocations of the edge-collapse passes start here (released before the fallback)
	dev_memset(scalars, 0, 8 * sizeof(u32));
	LAUNCH(k_init_groups, G, groups, group_tri_offset, G, config.simplify_ratio, scalars + 1);

	u32* tri_group = temp.alloc<u32>(T);
	u32* tri_group_alt = temp.alloc<u32>(T);
	LAUNCH(k_tri_group, T, group_tri_offset, G, tri_group, T);

	// ---- sparse ids in first-occurrence order per group (buildSparseRemap)
	u32* idx = temp.alloc<u32>(corners);
	u32* idx_alt = temp.alloc<u32>(corners);
	u32 vertex_count = 0;
	u32* sv_global;
	u32* sv_group;
	{
		size_t table_size = 1;
		while (table_size < corners + corners / 4)
			table_size <<= 1;
		u32* sparse_id = temp.alloc<u32>(corners);
		sv_global = temp.alloc<u32>(corners); // trimmed below by re-allocation order: kept simple, sized by corners
		ArenaScope table_scope(temp);
		u64* table_keys = temp.alloc<u64>(table_size);
		u32* table_first = temp.alloc<u32>(table_size);
		u32* corner_slot = temp.alloc<u32>(corners);
		dev_memset(table_keys, 0xff, table_size * sizeof(u64));
		dev_memset(table_first, 0xff, table_size * sizeof(u32));
		LAUNCH(k_corner_insert, corners, gtri, tri_group, table_keys, table_first, u32(table_size - 1), corner_slot, corners);
		LAUNCH(k_corner_first_flag, corners, corner_slot, table_first, sparse_id, corners);
		exclusive_scan_u32(sparse_id, sparse_id, corners, scalars, temp);
		vertex_count = dev_read(scalars);
		// sv_group aliases the tail of the sparse_id buffer after use? keep separate for clarity
		sv_group = idx_alt; // scratch until the first compaction; copied out below
		LAUNCH(k_assign_sparse, corners, gtri, tri_group, corner_slot, table_first, sparse_id, idx, sv_global, sv_group, corners);
	}
	u32* sv_group_final = temp.alloc<u32>(vertex_count);
	dev_d2d(sv_group_final, sv_group, size_t(vertex_count) * sizeof(u32));
	sv_group = sv_group_final;

	u32 A = 0;
	u32 attribute_remap_host[32];
	for (u32 i = 0; i < mesh.attribute_count && i < 32; ++i)
		if (mesh.attribute_weights[i] > 0)
			attribute_remap_host[A++] = i;
	if (!mesh.attributes)
		A = 0;

	// ---- per-vertex state
	u32* remap = temp.alloc<u32>(vertex_count);
	u32* wedge = temp.alloc<u32>(vertex_count);
	u8* kind = temp.alloc<u8>(vertex_count);
	u32* loop = temp.alloc<u32>(vertex_count);
	u32* loopback = temp.alloc<u32>(vertex_count);
	u32* loop_alt = temp.alloc<u32>(vertex_count);
	u32* loopback_alt = temp.alloc<u32>(vertex_count);
	Vector3* vpos = temp.alloc<Vector3>(vertex_count);
	float* vattr = temp.alloc<float>(size_t(vertex_count) * (A ? A : 1));
	Quadric* vertex_quadrics = temp.alloc<Quadric>(vertex_count);
	Quadric* attribute_quadrics = A ? temp.alloc<Quadric>(vertex_count) : nullptr;
	QuadricGrad* attribute_gradients = A ? temp.alloc<QuadricGrad>(size_t(vertex_count) * A) : nullptr;
	u32* adj_off = temp.alloc<u32>(size_t(vertex_count) + 1);
	u32* adj_corner = temp.alloc<u32>(corners);
	u32* collapse_remap = temp.alloc<u32>(vertex_count);
	u8* collapse_locked = temp.alloc<u8>(vertex_count);
	// per-vertex minima of the wavefront; the CUDA kernel double-buffers them by round parity
	u64* vmin_any = temp.alloc<u64>(size_t(vertex_count) * 2);
	u64* vmin_src = temp.alloc<u64>(size_t(vertex_count) * 2);
	float* group_min = temp.alloc<float>(size_t(G) * 3);
	float* group_extent = temp.alloc<float>(G);
	u32* group_collapses = temp.alloc<u32>(G);
	u32* group_error_bits = temp.alloc<u32>(G);

	// ---- adjacency (wedge topology), position remap, wedges, classification
	build_adjacency(idx, corners, nullptr, vertex_count, adj_off, adj_corner, true, temp);
	{
		ArenaScope s2(temp);
		size_t table_size = 1;
		while (table_size < size_t(vertex_count) * 2)
			table_size <<= 1;
		u64* table_keys = temp.alloc<u64>(table_size);
		u32* table_min = temp.alloc<u32>(table_size);
		u32* vertex_slot = temp.alloc<u32>(vertex_count);
		u32* sort_key = temp.alloc<u32>(vertex_count);
		u32* sort_val = temp.alloc<u32>(vertex_count);
		u32* sort_key_tmp = temp.alloc<u32>(vertex_count);
		u32* sort_val_tmp = temp.alloc<u32>(vertex_count);
		dev_memset(table_keys, 0xff, table_size * sizeof(u64));
		dev_memset(table_min, 0xff, table_size * sizeof(u32));
		LAUNCH(k_posremap_insert, vertex_count, sv_global, sv_group, global_remap, table_keys, table_min, u32(table_size - 1), vertex_slot, vertex_count);
		LAUNCH(k_posremap_resolve, vertex_count, vertex_slot, table_min, remap, sort_key, sort_val, vertex_count);
		radix_sort_pairs<u32>(sort_key, sort_key_tmp, sort_val, sort_val_tmp, vertex_count, 0, bits_for(vertex_count), temp);
		LAUNCH(k_wedge_from_sorted, vertex_count, sort_key, sort_val, wedge, vertex_count);
	}
	{
		ArenaScope s2(temp);
		u32* in_count = temp.alloc<u32>(vertex_count);
		u32* in_vertex = temp.alloc<u32>(vertex_count);
		u32* out_count = temp.alloc<u32>(vertex_count);
		u32* out_vertex = temp.alloc<u32>(vertex_count);
		dev_memset(in_count, 0, size_t(vertex_count) * 4);
		dev_memset(out_count, 0, size_t(vertex_count) * 4);
		LAUNCH(k_open_edges, corners, idx, adj_off, adj_corner, in_count, in_vertex, out_count, out_vertex, corners);
		LAUNCH(k_open_finalize, vertex_count, in_count, in_vertex, out_count, out_vertex, loop, loopback, vertex_count);
	}
	LAUNCH(k_classify_primary, vertex_count, remap, wedge, loop, loopback, kind, vertex_count);
	LAUNCH(k_classify_copy, vertex_count, remap, kind, vertex_count);
	if (config.simplify_permissive)
	{
		LAUNCH(k_classify_permissive, vertex_count, remap, wedge, sv_global, locks, idx, adj_off, adj_corner, kind, vertex_count);
		LAUNCH(k_classify_copy, vertex_count, remap, kind, vertex_count);
	}
	if (locks)
	{
		LAUNCH(k_classify_lock_primary, vertex_count, remap, sv_global, locks, kind, vertex_count);
		LAUNCH(k_classify_lock_spread, vertex_count, remap, kind, vertex_count);
	}

	// ---- rescale + quadrics
	{
		ArenaScope s2(temp);
		u32* gmin = temp.alloc<u32>(size_t(G) * 3);
		u32* gmax = temp.alloc<u32>(size_t(G) * 3);
		u32* attribute_remap = temp.alloc<u32>(32);
		float* attribute_weights = temp.alloc<float>(32);
		dev_memset(gmin, 0xff, size_t(G) * 12);
		dev_memset(gmax, 0, size_t(G) * 12);
		dev_h2d(attribute_remap, attribute_remap_host, sizeof(attribute_remap_host));
		dev_h2d(attribute_weights, mesh.attribute_weights, sizeof(float) * 32);
		LAUNCH(k_group_minmax, vertex_count, sv_global, sv_group, mesh.positions, gmin, gmax, vertex_count);
		LAUNCH(k_group_extent, G, gmin, gmax, group_min, group_extent, G);
		LAUNCH(k_rescale, vertex_count, sv_global, sv_group, mesh.positions, group_min, group_extent, mesh.attributes, mesh.attribute_stride, A, attribute_remap, attribute_weights, vpos, vattr, vertex_count);
	}
	if (A)
		LAUNCH(k_fill_attribute_quadrics, vertex_count, idx, adj_off, adj_corner, vpos, vattr, A, attribute_quadrics, attribute_gradients, vertex_count);
	build_adjacency(idx, corners, remap, vertex_count, adj_off, adj_corner, true, temp);
	LAUNCH(k_fill_vertex_quadrics, vertex_count, idx, remap, adj_off, adj_corner, vpos, kind, loop, loopback, vertex_quadrics, vertex_count);

	// ---- passes
	u32* cand_v0 = temp.alloc<u32>(corners);
	u32* cand_v1 = temp.alloc<u32>(corners);
	u8* cand_bidi = temp.alloc<u8>(corners);
	u32* cand_group = temp.alloc<u32>(corners);
	float* cand_error = temp.alloc<float>(corners);
	u32* sort_key = temp.alloc<u32>(corners);
	u32* sort_val = temp.alloc<u32>(corners);
	u32* sort_key_tmp = temp.alloc<u32>(corners);
	u32* sort_val_tmp = temp.alloc<u32>(corners);
	u8* status = temp.alloc<u8>(corners);
	u32* flags = temp.alloc<u32>(corners + 1);
	u32* tri_weight = temp.alloc<u32>(corners);
	u32* flip_flag = temp.alloc<u32>(corners);
	u64* tagged_error = temp.alloc<u64>(corners);

	dev_memset(vmin_any, 0xff, size_t(vertex_count) * 16);
	dev_memset(vmin_src, 0xff, size_t(vertex_count) * 16);
	u32* wave_state = temp.alloc<u32>(12);
	dev_memset(wave_state, 0, 12 * sizeof(u32));
#ifndef CLODB_EMU
	static u32 wave_max_blocks = 0;
	if (!wave_max_blocks)
	{
		int per_sm = 0, device = 0, sms = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave_rounds, WAVE_THREADS, 0));
		CUDA_CHECK(cudaGetDevice(&device));
		CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
		wave_max_blocks = u32(std::max(1, per_sm) * sms);
	}
#endif

	u32 cur_T = T;
	u32 round_tag = 0;
	int group_bits = bits_for(G > 1 ? G - 1 : 1);
	bool any_active = dev_read(scalars + 1) != 0;
	g_simplify_stats = SimplifyStats();

	while (any_active && cur_T > 0)
	{
		size_t cur_corners = size_t(cur_T) * 3;
		g_simplify_stats.passes++;

		// note: throughout the passes adjacency reflects welded topology of the result-in-progress
		if (g_simplify_stats.passes > 1)
			build_adjacency(idx, cur_corners, remap, vertex_count, adj_off, adj_corner, false, temp);

		LAUNCH(k_pick_flags, cur_corners, idx, tri_group, groups, remap, kind, loop, loopback, flags, cur_corners);
		exclusive_scan_u32(flags, flags, cur_corners, scalars, temp);
		u32 cand_total = dev_read(scalars);
		LAUNCH(k_group_cand_ranges, G, groups, flags, cand_total, G, cur_corners);
		if (cand_total == 0)
			break;
		LAUNCH(k_pick_emit, cur_corners, idx, flags, tri_group, kind, cand_total, cand_v0, cand_v1, cand_bidi, cand_group, cur_corners);
		LAUNCH(k_rank, cand_total, cand_v0, cand_v1, cand_bidi, cand_group, cand_error, sort_key, sort_val, cand_total, vpos, vattr, vertex_quadrics, attribute_quadrics, attribute_gradients, A, remap, wedge, kind, loop, loopback);
		radix_sort_pairs<u32>(sort_key, sort_key_tmp, sort_val, sort_val_tmp, cand_total, 0, 12 + group_bits, temp);

		LAUNCH(k_collapse_init, vertex_count, collapse_remap, collapse_locked, vertex_count);
		dev_memset(status, 0, cand_total);

		u32 rounds = 0;
		LAUNCH(k_window_init, G, groups, G);
		ArenaScope pass_scope(temp);
		u32* woff = temp.alloc<u32>(size_t(G) + 2); // window offsets of the examined prefixes (see k_window_lengths)
		u32 W = 0;
#ifdef CLODB_EMU
		u32* wave_list[2] = {tri_weight, flip_flag}; // the cut-scan inputs are not live during the rounds: reuse them as work lists
#else
		WaveEntry* wave_list[2] = {temp.alloc<WaveEntry>(cand_total), temp.alloc<WaveEntry>(cand_total)};
#endif
		for (;;)
		{
			// ---- the examined prefixes of all groups as one dense index space (k_window_lengths)
			LAUNCH(k_window_lengths, size_t(G) + 1, groups, G, woff);
			exclusive_scan_u32(woff, woff, size_t(G) + 1, scalars, temp);
			W = dev_read(scalars);
			// ---- wavefront over the current windows
			dev_memset(wave_state, 0, 3 * sizeof(u32));
#ifdef CLODB_EMU
			LAUNCH(k_wave_window, (size_t(cand_total) + 31) / 32 * 32, sort_val, cand_group, groups, status, wave_list[0], wave_state, cand_total);
			u32 listed = dev_read(wave_state);
			for (u32 r = 0; listed != 0; )
			{
				round_tag++;
				rounds++;
				r++;
				dev_memset(scalars + 2, 0, sizeof(u32));
				LAUNCH(k_wave_publish, listed, wave_list[0], listed, sort_val, status, cand_v0, cand_v1, remap, vmin_any, vmin_src, round_tag, scalars + 2);
				LAUNCH(k_wave_decide, listed, wave_list[0], listed, sort_val, status, cand_v0, cand_v1, remap, wedge, kind, loop, loopback, vpos, idx, adj_off, adj_corner, vmin_any, vmin_src, round_tag, collapse_remap, collapse_locked);
				if (dev_read(scalars + 2) == 0 || r >= config_max_rounds())
					break;
			}
#else
			{
				WaveArgs wa;
				wa.status = status, wa.cand_v0 = cand_v0, wa.cand_v1 = cand_v1, wa.remap = remap, wa.wedge = wedge, wa.kind = kind, wa.loop = loop, wa.loopback = loopback;
				wa.vpos = vpos, wa.idx = idx, wa.adj_off = adj_off, wa.adj_corner = adj_corner, wa.collapse_remap = collapse_remap, wa.collapse_locked = collapse_locked;
				wa.vmin_any[0] = vmin_any, wa.vmin_any[1] = vmin_any + vertex_count;
				wa.vmin_src[0] = vmin_src, wa.vmin_src[1] = vmin_src + vertex_count;
				wa.list[0] = wave_list[0];
				wa.list[1] = wave_list[1];
				wa.state = wave_state;
				wa.max_rounds = config_max_rounds();
				static const u32 tail_threshold = getenv("CLODB200_WAVE_TAIL") ? u32(atoi(getenv("CLODB200_WAVE_TAIL"))) : 256u;
				wa.tail_threshold = tail_threshold;
				static const bool log_rounds = getenv("CLODB200_DEBUG_ROUNDS") != nullptr;
				wa.round_log = nullptr;
				if (log_rounds)
				{
					wa.round_log = temp.alloc<u32>(512);
					dev_memset(wa.round_log, 0, 512 * 4);
				}
				static const u32 blocks_per_sm_cap = getenv("CLODB200_WAVE_BLOCKS") ? u32(atoi(getenv("CLODB200_WAVE_BLOCKS"))) : 0u;
				LAUNCH_GRID(k_wave_list, (W + 255) / 256, 256, woff, G, W, sort_val, groups, status, cand_v0, cand_v1, remap, wa);
				u32 blocks = std::min<u32>(blocks_per_sm_cap ? std::min(wave_max_blocks, blocks_per_sm_cap) : wave_max_blocks, std::max<u32>(1u, (W + WAVE_THREADS - 1) / WAVE_THREADS));
				LAUNCH_COOP(k_wave_rounds, blocks, WAVE_THREADS, wa);
				if (g_profile)
				{
					std::vector<u32> ws_now = dev_download(wave_state, 12);
					profile_set_last_work((size_t(ws_now[9]) << 32) | ws_now[8]);
				}
				if (log_rounds)
				{
					std::vector<u32> lg = dev_download(wa.round_log, 512);
					fprintf(stderr, "wave T %u cands %u blocks %u:", cur_T, cand_total, blocks);
					for (u32 r = 1; r < 256 && lg[r * 2]; ++r)
						fprintf(stderr, " %u%s/%.0fus", lg[r * 2] & 0x7fffffffu, (lg[r * 2] >> 31) ? "t" : "", r > 1 ? (lg[r * 2 + 1] - lg[r * 2 - 1]) / 1e3 : 0.0);
					fprintf(stderr, "\n");
				}
			}
#endif

			// ---- where would the serial scan have stopped? (over the examined prefixes only)
			LAUNCH(k_cut_inputs, W, woff, groups, G, W, sort_val, status, cand_v0, cand_error, kind, tri_weight, flip_flag, tagged_error);
			exclusive_scan_u32(tri_weight, tri_weight, W, nullptr, temp);
			exclusive_scan_u32(flip_flag, flip_flag, W, nullptr, temp);
			exclusive_scan<u64, OpMaxU64>(tagged_error, tagged_error, W, nullptr, temp);
			LAUNCH(k_cut_find, W, woff, G, W, sort_val, status, cand_error, groups, tri_weight, flip_flag, tagged_error);
			dev_memset(scalars + 3, 0, sizeof(u32));
			LAUNCH(k_window_check, G, groups, G, scalars + 3);
			if (dev_read(scalars + 3) == 0)
				break;
			g_simplify_stats.window_extensions++;
		}
#ifdef CLODB_EMU
		g_simplify_stats.rounds += rounds;
		g_simplify_stats.max_rounds = std::max(g_simplify_stats.max_rounds, rounds);
#endif
		if (getenv("CLODB200_DEBUG_CUT"))
		{
			std::vector<GroupState> gh = dev_download(groups, G);
			double scanned = 0, total = 0, window = 0;
			for (u32 g = 0; g < G; ++g)
				if (gh[g].active && gh[g].cand_count)
				{
					u32 cut = std::min(gh[g].cut, gh[g].cand_begin + gh[g].cand_count);
					scanned += cut - gh[g].cand_begin;
					window += gh[g].win_end - gh[g].cand_begin;
					total += gh[g].cand_count;
				}
			fprintf(stderr, "pass %u: T %u cands %u scanned %.0f (%.1f%%) window %.0f rounds %u ext %u\n", g_simplify_stats.passes, cur_T, cand_total, scanned, total ? 100.0 * scanned / total : 0.0, window, rounds, g_simplify_stats.window_extensions);
		}
		dev_memset(group_collapses, 0, size_t(G) * 4);
		dev_memset(group_error_bits, 0, size_t(G) * 4);
		u32* accepted = flip_flag; // the cut scans are done with it
		dev_memset(scalars + 4, 0, sizeof(u32));
		LAUNCH(k_cut_apply, (size_t(W) + 31) / 32 * 32, woff, G, W, sort_val, status, cand_v0, cand_error, wedge, kind, groups, group_collapses, group_error_bits, collapse_remap, accepted, scalars + 4);

		LAUNCH(k_update_quadrics, W, accepted, scalars + 4, sort_val, cand_v0, remap, wedge, collapse_remap, vertex_quadrics, attribute_quadrics, attribute_gradients, A);
		LAUNCH(k_remap_loops, vertex_count, loop, loop_alt, collapse_remap, vertex_count);
		LAUNCH(k_remap_loops, vertex_count, loopback, loopback_alt, collapse_remap, vertex_count);
		std::swap(loop, loop_alt);
		std::swap(loopback, loopback_alt);

		u32* keep = flags;
		LAUNCH(k_remap_triangles, cur_T, idx, collapse_remap, remap, tri_group, group_collapses, keep, cur_T);
		exclusive_scan_u32(keep, keep, cur_T, scalars, temp);
		u32 new_T = dev_read(scalars);
		LAUNCH(k_compact_triangles, cur_T, idx, tri_group, keep, new_T, idx_alt, tri_group_alt, cur_T);
		dev_memset(scalars + 1, 0, sizeof(u32));
		LAUNCH(k_group_after_pass, G, groups, keep, new_T, group_collapses, group_error_bits, cur_T, G, scalars + 1);
		std::swap(idx, idx_alt);
		std::swap(tri_group, tri_group_alt);
		cur_T = new_T;
		any_active = dev_read(scalars + 1) != 0;
	}

#ifndef CLODB_EMU
	{
		std::vector<u32> st = dev_download(wave_state, 8);
		g_simplify_stats.rounds = st[6];
		g_simplify_stats.max_rounds = st[7];
	}
#endif
	LAUNCH(k_finalize_output, size_t(cur_T) * 3, idx, sv_global, out.tri, size_t(cur_T) * 3);
	LAUNCH(k_group_results, G, groups, group_extent, out.group_tri_offset, out.group_error, G);
	out.triangle_count = cur_T;

	// ---- sloppy fallback for groups that are still above their target (clusterlod.h:623-628)
	if (config.simplify_fallback_sloppy)
	{
		std::vector<GroupState> gh = dev_download(groups, G);
		std::vector<u32> sel, sel_target;
		for (u32 g = 0; g < G; ++g)
			if (gh[g].tri_count > gh[g].target_tris)
			{
				sel.push_back(g);
				sel_target.push_back(gh[g].target_tris);
			}
		if (!sel.empty())
		{
			// everything the edge-collapse passes allocated is dead now (its results live in the persist arena)
			temp.release(main_mark);
			std::vector<float> new_error = dev_download(out.group_error, G);
			SloppyResult sl = sloppy_groups(gtri, group_tri_offset_host, sel, sel_target, mesh, locks, temp);
			g_simplify_stats.sloppy_groups += u32(sel.size());
			// reassemble the level output group by group (the fallback may return more or fewer triangles)
			std::vector<u32> new_offset(size_t(G) + 1, 0), src_begin(G);
			std::vector<u8> src_is_fallback(G, 0);
			u32 fallback_begin = 0;
			for (u32 g = 0, s = 0; g < G; ++g)
			{
				u32 count = gh[g].tri_count;
				src_begin[g] = gh[g].tri_begin;
				if (s < sel.size() && sel[s] == g)
				{
					count = sl.kept[s];
					src_begin[g] = fallback_begin;
					src_is_fallback[g] = 1;
					fallback_begin += count;
					new_error[g] = sl.error[s] * config.simplify_error_factor_sloppy;
					++s;
				}
				new_offset[g + 1] = new_offset[g] + count;
			}
			if (size_t(new_offset[G]) > size_t(T))
				throw Error("clodb200: sloppy fallback produced more triangles than its input");
			u32* d_offset = temp.alloc<u32>(size_t(G) + 1);
			u32* d_src_begin = temp.alloc<u32>(G);
			u8* d_is_fallback = temp.alloc<u8>(G);
			u32* assembled = temp.alloc<u32>(size_t(new_offset[G]) * 3 + 3);
			dev_h2d(d_offset, new_offset.data(), (size_t(G) + 1) * 4);
			dev_h2d(d_src_begin, src_begin.data(), size_t(G) * 4);
			dev_h2d(d_is_fallback, src_is_fallback.data(), G);
			LAUNCH(k_sl_assemble, new_offset[G], d_offset, d_src_begin, d_is_fallback, G, out.tri, sl.tri, assembled, new_offset[G]);
			dev_d2d(out.tri, assembled, size_t(new_offset[G]) * 12);
			dev_h2d(out.group_tri_offset, new_offset.data(), (size_t(G) + 1) * 4);
			dev_h2d(out.group_error, new_error.data(), size_t(G) * 4);
			out.triangle_count = new_offset[G];
		}
	}
	return out;
}

} // namespace clodb
