// MikkTSpace tangent stream for the simplification attributes (a1 of SURVEY.md §8a).
//
// Reference: GenerateMikkTangents, BasicRenderer/src/Mesh/ClusterLODUtilities.cpp:655-737, which runs
// genTangSpaceDefault (BasicRenderer/src/Utilities/mikktspace.cpp:224-412) over the indexed triangle list and then sums
// the per-corner tangents per vertex in (face, corner) order (:630-653), normalises, and falls back to a normal-derived
// tangent (:559-579). The reference is a serial flood fill over the mesh; what it computes is, per corner, a function of
// the fan of triangles around the corner's welded vertex, so it is restated here as data-parallel passes:
//
//   weld        class of a vertex = all vertices with == position, normal and uv (mikktspace.cpp:451-576, 578-694; the
//               hash grid and the recursive median split are only accelerators of that equivalence)
//   tri info    orientation / degenerate-uv flags and normalised first-order derivatives per triangle (:946-1012)
//   neighbours  edges sorted by (min, max, triangle); first-fit pairing of opposite half edges per run (:1502-1595)
//   groups      Build4RuleGroups' groups are the connected components of the corner graph (a node per corner, links
//               through the two edges at the vertex, cut where the orientation flag changes) (:1071-1193). Every
//               component is a path or a cycle, so each corner walks its own fan instead of a flood fill
//   tspaces     per corner: members of its sub-group (angular threshold, :1232-1281) in ascending triangle order, angle
//               weighted sum of the projected derivatives (EvalTspace, :1369-1441)
//   degenerates copy from the first good corner with the same welded vertex (:1820-1861)
//   vertices    ordered per-vertex accumulation and finalisation (ClusterLODUtilities.cpp:630-653, 706-734)
//
// Float arithmetic keeps the reference's expression order (-fmad=false); the corner angle uses the C library's acosf
// algorithm (mk_acosf below). Known deviations, both outside what the reference's own tests or validation pin: (1) a welded class
// is represented by its lowest vertex index, the reference by whichever member its median split visits first, which can
// only change the sign of a zero; (2) triangles with a degenerate uv mapping ("group with anything", :963, 1007) take the
// orientation of the first group that reaches them in the reference's serial flood order (:1160-1172); here they take the
// orientation offered by the neighbouring group with the lowest seed corner, iterated to a fixed point, which is the
// same whenever the offers do not conflict across a chain of such triangles. Quads never occur on this path
// (MikkGetNumVerticesOfFace returns 3, ClusterLODUtilities.cpp:600-603).
#include "clodb.h"

#include <cmath>

namespace clodb
{

static const u32 MK_EMPTY = 0xffffffffu;
static const u8 MK_DEGENERATE = 1; // MARK_DEGENERATE
static const u8 MK_ANY = 4;        // GROUP_WITH_ANY
static const u8 MK_ORIENT = 8;     // ORIENT_PRESERVING
static const u32 MK_FAN = 24;      // fan members kept in registers/local memory before the storage-free slow path
static const float MK_FLT_MIN = 1.17549435e-38f;

#ifdef CLODB_EMU
#define MK_MEMFN inline
#else
#define MK_MEMFN __device__ __forceinline__
#endif

DEVFN const float* mk_vertex(const u8* vertices, u32 stride, u32 v)
{
	return reinterpret_cast<const float*>(vertices + size_t(v) * stride);
}

DEVFN bool mk_not_zero(float x)
{
	return fabsf(x) > MK_FLT_MIN;
}

DEVFN bool mk_vnot_zero(float x, float y, float z)
{
	return mk_not_zero(x) || mk_not_zero(y) || mk_not_zero(z);
}

DEVFN float mk_length(float x, float y, float z)
{
	return sqrtf(x * x + y * y + z * z);
}

// The reference's `(float)acos(fCos)` (mikktspace.cpp:1421) sits in a C++ translation unit, so overload resolution picks
// acos(float), i.e. the C library's acosf. This is the fdlibm algorithm (e_acosf.c, Sun Microsystems' rational
// approximation) in float arithmetic; checked bit for bit against glibc 2.39's acosf over every float in [-1, 1]
// (tests/test_mikk.py re-checks a sample wherever the tests run). Needs IEEE mul/add/div/sqrt without contraction.
DEVFN float mk_acosf(float x)
{
	const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
	const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f;
	const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
	const u32 hx = __float_as_uint(x), ix = hx & 0x7fffffffu;
	const bool negative = (hx >> 31) != 0;
	if (ix == 0x3f800000u)
		return negative ? pi + 2.0f * pio2_lo : 0.0f;
	if (ix > 0x3f800000u)
		return (x - x) / (x - x);
	if (ix < 0x3f000000u)
	{
		if (ix <= 0x32800000u)
			return pio2_hi + pio2_lo;
		float z = x * x;
		float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		float r = p / q;
		return pio2_hi - (x - (pio2_lo - x * r));
	}
	if (negative)
	{
		float z = (1.0f + x) * 0.5f;
		float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		float s = sqrtf(z);
		float r = p / q;
		float w = r * s - pio2_lo;
		return pi - 2.0f * (s + w);
	}
	float z = (1.0f - x) * 0.5f;
	float s = sqrtf(z);
	float df = __uint_as_float(__float_as_uint(s) & 0xfffff000u);
	float c = (z - df * df) / (s + df);
	float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
	float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
	float r = p / q;
	float w = r * s + c;
	return 2.0f * (df + w);
}

// ---------------------------------------------------------------------------------------------------------- weld
DEVFN u32 mk_hash8(const float* p)
{
	u32 h = 0x9E3779B1u;
	for (int k = 0; k < 8; ++k)
	{
		u32 a = __float_as_uint(p[k] + 0.0f); // -0 and +0 compare equal, so they must hash alike
		h = (h ^ (h >> 15)) * 0x85EBCA77u + a * 0xC2B2AE3Du;
	}
	h ^= h >> 16;
	h *= 0x7FEB352Du;
	h ^= h >> 15;
	return h;
}

KERNEL k_mk_weld_insert(const u8* __restrict__ vertices, u32 stride, size_t vertex_count, u32* table_rep, u32* table_min, u32 mask, u32* slot_out)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	const float* p = mk_vertex(vertices, stride, u32(i));
	float key[8];
	bool nan = false;
	for (int k = 0; k < 8; ++k)
	{
		key[k] = p[k];
		nan |= key[k] != key[k];
	}
	if (nan)
	{
		slot_out[i] = MK_EMPTY;
		return;
	}
	u32 h = mk_hash8(key) & mask;
	for (;;)
	{
		u32 rep = atomicCAS(&table_rep[h], MK_EMPTY, u32(i));
		if (rep == MK_EMPTY || rep == u32(i))
			break;
		const float* q = mk_vertex(vertices, stride, rep);
		bool same = true;
		for (int k = 0; k < 8; ++k)
			same &= q[k] == key[k];
		if (same)
			break;
		h = (h + 1) & mask;
	}
	atomicMin(&table_min[h], u32(i));
	slot_out[i] = h;
}

KERNEL k_mk_weld_resolve(u32* weld, const u32* __restrict__ table_min, size_t vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	u32 slot = weld[i];
	weld[i] = slot == MK_EMPTY ? u32(i) : table_min[slot];
}

// ------------------------------------------------------------------------------------------------------ tri info
// One thread per triangle: welded corner ids, degenerate mark (:285-298), first-order derivatives and flags
// (:968-1012), the undirected edge keys for the neighbour sort (:1508-1517) and the lowest good corner of every welded
// vertex (the search of DegenEpilogue, :1835-1842).
KERNEL k_mk_tri_info(const u8* __restrict__ vertices, u32 stride, const u32* __restrict__ indices, const u32* __restrict__ weld, u32 triangle_count, int vertex_bits,
    u32* wtri, u8* tflag, float* tos, float* tot, u64* edge_key, u32* edge_val, u32* first_good, u32* counters)
{
	size_t t = GTID;
	if (t >= triangle_count)
		return;
	u32 w[3];
	for (int k = 0; k < 3; ++k)
	{
		w[k] = weld[indices[t * 3 + k]];
		wtri[t * 3 + k] = w[k];
	}
	const float* v1 = mk_vertex(vertices, stride, w[0]);
	const float* v2 = mk_vertex(vertices, stride, w[1]);
	const float* v3 = mk_vertex(vertices, stride, w[2]);
	bool e01 = v1[0] == v2[0] && v1[1] == v2[1] && v1[2] == v2[2];
	bool e02 = v1[0] == v3[0] && v1[1] == v3[1] && v1[2] == v3[2];
	bool e12 = v2[0] == v3[0] && v2[1] == v3[1] && v2[2] == v3[2];
	u8 flag = 0;
	float os[3] = {0.f, 0.f, 0.f}, ot[3] = {0.f, 0.f, 0.f};
	if (e01 || e02 || e12)
	{
		flag = MK_DEGENERATE;
		for (int k = 0; k < 3; ++k)
		{
			edge_key[t * 3 + k] = ~u64(0);
			edge_val[t * 3 + k] = u32(t * 3 + k);
		}
	}
	else
	{
		flag = MK_ANY; // assumed bad
		const float t21x = v2[6] - v1[6];
		const float t21y = v2[7] - v1[7];
		const float t31x = v3[6] - v1[6];
		const float t31y = v3[7] - v1[7];
		const float d1[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
		const float d2[3] = {v3[0] - v1[0], v3[1] - v1[1], v3[2] - v1[2]};
		const float area2 = t21x * t31y - t21y * t31x;
		float vos[3], vot[3];
		for (int k = 0; k < 3; ++k)
		{
			vos[k] = t31y * d1[k] - t21y * d2[k];    // eq 18
			vot[k] = (-t31x) * d1[k] + t21x * d2[k]; // eq 19
		}
		if (area2 > 0)
			flag |= MK_ORIENT;
		if (mk_not_zero(area2))
		{
			const float abs_area = fabsf(area2);
			const float len_os = mk_length(vos[0], vos[1], vos[2]);
			const float len_ot = mk_length(vot[0], vot[1], vot[2]);
			const float fs = (flag & MK_ORIENT) == 0 ? -1.0f : 1.0f;
			if (mk_not_zero(len_os))
			{
				float s = fs / len_os;
				os[0] = s * vos[0], os[1] = s * vos[1], os[2] = s * vos[2];
			}
			if (mk_not_zero(len_ot))
			{
				float s = fs / len_ot;
				ot[0] = s * vot[0], ot[1] = s * vot[1], ot[2] = s * vot[2];
			}
			const float mag_s = len_os / abs_area;
			const float mag_t = len_ot / abs_area;
			if (mk_not_zero(mag_s) && mk_not_zero(mag_t))
				flag &= u8(~MK_ANY);
		}
		if (flag & MK_ANY)
			atomicAdd(&counters[0], 1u);
		for (int k = 0; k < 3; ++k)
		{
			u32 a = w[k], b = w[k == 2 ? 0 : k + 1];
			u32 lo = a < b ? a : b, hi = a < b ? b : a;
			edge_key[t * 3 + k] = (u64(lo) << vertex_bits) | u64(hi);
			edge_val[t * 3 + k] = u32(t * 3 + k);
			atomicMin(&first_good[w[k]], u32(t * 3 + k));
		}
	}
	tflag[t] = flag;
	for (int k = 0; k < 3; ++k)
	{
		tos[t * 3 + k] = os[k];
		tot[t * 3 + k] = ot[k];
	}
}

// ---------------------------------------------------------------------------------------------------- neighbours
// GetEdge (:1703-1732): which edge of the triangle has the two given endpoints, and its direction in the triangle.
DEVFN void mk_get_edge(u32& i0_out, u32& i1_out, int& edge_out, const u32* idx, u32 i0_in, u32 i1_in)
{
	if (idx[0] == i0_in || idx[0] == i1_in)
	{
		if (idx[1] == i0_in || idx[1] == i1_in)
		{
			edge_out = 0;
			i0_out = idx[0];
			i1_out = idx[1];
		}
		else
		{
			edge_out = 2;
			i0_out = idx[2];
			i1_out = idx[0];
		}
	}
	else
	{
		edge_out = 1;
		i0_out = idx[1];
		i1_out = idx[2];
	}
}

// The entries of one undirected edge are contiguous and in ascending triangle order after the stable sort; the thread at
// the head of each run replays the reference's first-fit pairing over its run (:1554-1594). Runs write disjoint slots.
KERNEL k_mk_pair_edges(const u64* __restrict__ edge_key, const u32* __restrict__ edge_val, size_t entries, const u32* __restrict__ wtri, int vertex_bits, int* nbr)
{
	size_t e = GTID;
	if (e >= entries)
		return;
	u64 key = edge_key[e];
	if (key == ~u64(0))
		return;
	if (e > 0 && edge_key[e - 1] == key)
		return;
	size_t end = e + 1;
	while (end < entries && edge_key[end] == key)
		++end;
	if (end - e < 2)
		return;
	const u32 i0 = u32(key >> vertex_bits), i1 = u32(key & ((u64(1) << vertex_bits) - 1));
	for (size_t i = e; i < end; ++i)
	{
		u32 f = edge_val[i] / 3;
		u32 a0, a1;
		int edge_a;
		mk_get_edge(a0, a1, edge_a, wtri + size_t(f) * 3, i0, i1);
		if (nbr[size_t(f) * 3 + edge_a] != -1)
			continue;
		for (size_t j = i + 1; j < end; ++j)
		{
			u32 t = edge_val[j] / 3;
			u32 b0, b1;
			int edge_b;
			mk_get_edge(b1, b0, edge_b, wtri + size_t(t) * 3, i0, i1); // flipped
			if (a0 == b0 && a1 == b1 && nbr[size_t(t) * 3 + edge_b] == -1)
			{
				nbr[size_t(f) * 3 + edge_a] = int(t);
				nbr[size_t(t) * 3 + edge_b] = int(f);
				break;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------- corner graph
DEVFN u32 mk_corner_of(const u32* __restrict__ wtri, u32 t, u32 w)
{
	const u32* p = wtri + size_t(t) * 3;
	return p[0] == w ? 0u : (p[1] == w ? 1u : 2u);
}

// Enumerates the corners of the group of corner (t0, j0): the start corner, then the fan in the direction of the edge
// leaving the vertex, then (unless the fan closed) the fan in the direction of the edge arriving at it. A link is followed
// while the neighbour has the group's orientation flag (:1186-1187) and `blocked` does not veto it.
template <typename Visit, typename Blocked>
DEVFN void mk_walk_fan(const u32* __restrict__ wtri, const int* __restrict__ nbr, const u8* __restrict__ tflag, u32 t0, u32 j0, u32 w, bool orient, u32 max_steps, bool include_start, Visit& visit, Blocked& blocked)
{
	if (include_start)
		visit(t0 * 3 + j0);
	bool closed = false;
	u32 t = t0, j = j0;
	for (u32 step = 0; step < max_steps; ++step)
	{
		int n = nbr[size_t(t) * 3 + j];
		if (n < 0 || ((tflag[n] & MK_ORIENT) != 0) != orient || blocked(u32(n)))
			break;
		if (u32(n) == t0)
		{
			closed = true;
			break;
		}
		u32 nj = mk_corner_of(wtri, u32(n), w);
		visit(u32(n) * 3 + nj);
		t = u32(n);
		j = nj;
	}
	if (closed)
		return;
	t = t0, j = j0;
	for (u32 step = 0; step < max_steps; ++step)
	{
		int n = nbr[size_t(t) * 3 + (j == 0 ? 2 : j - 1)];
		if (n < 0 || ((tflag[n] & MK_ORIENT) != 0) != orient || blocked(u32(n)) || u32(n) == t0)
			break;
		u32 nj = mk_corner_of(wtri, u32(n), w);
		visit(u32(n) * 3 + nj);
		t = u32(n);
		j = nj;
	}
}

struct MkNeverBlocked
{
	MK_MEMFN bool operator()(u32) const
	{
		return false;
	}
};

// ---- "group with anything" triangles: orientation taken from the group that claims them (:1160-1172) ------------------
struct MkMinSeed
{
	const u8* tflag;
	u32 seed;
	MK_MEMFN void operator()(u32 corner)
	{
		if ((tflag[corner / 3] & MK_ANY) == 0 && corner < seed)
			seed = corner;
	}
};

struct MkUnresolved
{
	const u8* state;
	MK_MEMFN bool operator()(u32 t) const
	{
		return state[t] == 1;
	}
};

KERNEL k_mk_any_init(const u8* __restrict__ tflag, u8* state, u32 triangle_count)
{
	size_t t = GTID;
	if (t >= triangle_count)
		return;
	u8 f = tflag[t];
	state[t] = ((f & MK_ANY) != 0 && (f & MK_DEGENERATE) == 0) ? 1 : 0;
}

// One relaxation round: every unresolved triangle collects, through each of its six links, the seed (lowest non-wildcard
// corner) of the group on the other side, and adopts the orientation of the lowest seed.
KERNEL k_mk_any_round(const u32* __restrict__ wtri, const int* __restrict__ nbr, u8* tflag, const u8* __restrict__ state_in, u8* state_out, u32 triangle_count, u32* changed)
{
	size_t a = GTID;
	if (a >= triangle_count)
		return;
	u8 s = state_in[a];
	state_out[a] = s;
	if (s != 1)
		return;
	u32 best_seed = MK_EMPTY;
	bool best_orient = false;
	MkUnresolved blocked = {state_in};
	for (u32 i = 0; i < 3; ++i)
	{
		u32 w = wtri[a * 3 + i];
		for (int dir = 0; dir < 2; ++dir)
		{
			int n = nbr[a * 3 + (dir == 0 ? i : (i == 0 ? 2 : i - 1))];
			if (n < 0 || state_in[n] == 1)
				continue;
			bool orient = (tflag[n] & MK_ORIENT) != 0;
			u32 nj = mk_corner_of(wtri, u32(n), w);
			// the group on that side: the fan continuing away from `a` (the other direction leads back through `a`)
			MkMinSeed seed = {tflag, MK_EMPTY};
			seed(u32(n) * 3 + nj);
			u32 t = u32(n), j = nj;
			for (u32 step = 0; step < triangle_count; ++step)
			{
				int m = nbr[size_t(t) * 3 + (dir == 0 ? j : (j == 0 ? 2 : j - 1))];
				if (m < 0 || u32(m) == u32(a) || ((tflag[m] & MK_ORIENT) != 0) != orient || blocked(u32(m)))
					break;
				u32 mj = mk_corner_of(wtri, u32(m), w);
				seed(u32(m) * 3 + mj);
				t = u32(m);
				j = mj;
			}
			if (seed.seed < best_seed)
			{
				best_seed = seed.seed;
				best_orient = orient;
			}
		}
	}
	if (best_seed != MK_EMPTY)
	{
		tflag[a] = u8((tflag[a] & ~MK_ORIENT) | (best_orient ? MK_ORIENT : 0));
		state_out[a] = 2;
		atomicAdd(changed, 1u);
	}
}

// ------------------------------------------------------------------------------------------------- corner terms
// Per corner of a good triangle: the triangle's derivatives projected into the tangent plane of the welded vertex and
// normalised (:1248-1252 and again :1398-1401), and the corner angle between the projected edges (:1403-1421).
// cps = {os.x, os.y, os.z, angle}, cpt = {ot.x, ot.y, ot.z}.
KERNEL k_mk_corner_terms(const u8* __restrict__ vertices, u32 stride, const u32* __restrict__ wtri, const u8* __restrict__ tflag, const float* __restrict__ tos, const float* __restrict__ tot,
    size_t corners, float* cps, float* cpt)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	size_t t = c / 3;
	u32 i = u32(c - t * 3);
	if (tflag[t] & MK_DEGENERATE)
		return;
	const u32* idx = wtri + t * 3;
	const float* p1 = mk_vertex(vertices, stride, idx[i]);
	const float n[3] = {p1[3], p1[4], p1[5]};
	float os[3] = {tos[t * 3], tos[t * 3 + 1], tos[t * 3 + 2]};
	float ot[3] = {tot[t * 3], tot[t * 3 + 1], tot[t * 3 + 2]};
	float ds = n[0] * os[0] + n[1] * os[1] + n[2] * os[2];
	float dt = n[0] * ot[0] + n[1] * ot[1] + n[2] * ot[2];
	for (int k = 0; k < 3; ++k)
	{
		os[k] = os[k] - ds * n[k];
		ot[k] = ot[k] - dt * n[k];
	}
	if (mk_vnot_zero(os[0], os[1], os[2]))
	{
		float s = 1 / mk_length(os[0], os[1], os[2]);
		os[0] = s * os[0], os[1] = s * os[1], os[2] = s * os[2];
	}
	if (mk_vnot_zero(ot[0], ot[1], ot[2]))
	{
		float s = 1 / mk_length(ot[0], ot[1], ot[2]);
		ot[0] = s * ot[0], ot[1] = s * ot[1], ot[2] = s * ot[2];
	}
	const float* p2 = mk_vertex(vertices, stride, idx[i < 2 ? i + 1 : 0]);
	const float* p0 = mk_vertex(vertices, stride, idx[i > 0 ? i - 1 : 2]);
	float v1[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]};
	float v2[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
	float d1 = n[0] * v1[0] + n[1] * v1[1] + n[2] * v1[2];
	for (int k = 0; k < 3; ++k)
		v1[k] = v1[k] - d1 * n[k];
	if (mk_vnot_zero(v1[0], v1[1], v1[2]))
	{
		float s = 1 / mk_length(v1[0], v1[1], v1[2]);
		v1[0] = s * v1[0], v1[1] = s * v1[1], v1[2] = s * v1[2];
	}
	float d2 = n[0] * v2[0] + n[1] * v2[1] + n[2] * v2[2];
	for (int k = 0; k < 3; ++k)
		v2[k] = v2[k] - d2 * n[k];
	if (mk_vnot_zero(v2[0], v2[1], v2[2]))
	{
		float s = 1 / mk_length(v2[0], v2[1], v2[2]);
		v2[0] = s * v2[0], v2[1] = s * v2[1], v2[2] = s * v2[2];
	}
	float fcos = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
	fcos = fcos > 1 ? 1 : (fcos < (-1) ? (-1) : fcos);
	float angle = mk_acosf(fcos);
	cps[c * 4 + 0] = os[0];
	cps[c * 4 + 1] = os[1];
	cps[c * 4 + 2] = os[2];
	cps[c * 4 + 3] = angle;
	cpt[c * 3 + 0] = ot[0];
	cpt[c * 3 + 1] = ot[1];
	cpt[c * 3 + 2] = ot[2];
}

// ----------------------------------------------------------------------------------------------------- tspaces
struct MkCollect
{
	u32 members[MK_FAN];
	u32 count;
	MK_MEMFN void operator()(u32 corner)
	{
		if (count < MK_FAN)
			members[count] = corner;
		++count;
	}
};

// slow path visitor: the lowest member above `floor_corner`
struct MkNextAbove
{
	u32 floor_corner; // MK_EMPTY: no floor yet
	u32 best;
	MK_MEMFN void operator()(u32 corner)
	{
		if ((floor_corner == MK_EMPTY || corner > floor_corner) && corner < best)
			best = corner;
	}
};

struct MkTspaceSum
{
	const u8* tflag;
	const float* cps;
	const float* cpt;
	float thres_cos;
	u32 self;
	bool self_any;
	float self_os[3], self_ot[3];
	float res[3];
	bool seeded;

	MK_MEMFN void add(u32 m)
	{
		if (tflag[m / 3] & MK_ANY)
			return; // only valid triangles contribute (:1383); they are members of every sub-group regardless
		seeded = true;
		const float* ps = cps + size_t(m) * 4;
		if (!self_any && m != self)
		{
			const float* pt = cpt + size_t(m) * 3;
			float cos_s = self_os[0] * ps[0] + self_os[1] * ps[1] + self_os[2] * ps[2];
			float cos_t = self_ot[0] * pt[0] + self_ot[1] * pt[1] + self_ot[2] * pt[2];
			if (!(cos_s > thres_cos && cos_t > thres_cos))
				return;
		}
		float angle = ps[3];
		res[0] = res[0] + angle * ps[0];
		res[1] = res[1] + angle * ps[1];
		res[2] = res[2] + angle * ps[2];
	}
};

// One thread per corner: tangent (vOs of the sub-group's tangent space) and orientation sign as handed to
// m_setTSpaceBasic (:398-399). Corners no group reaches keep the initial space (1,0,0) with bOrient = 0 (:340-346).
KERNEL k_mk_corner_tspace(const u32* __restrict__ wtri, const int* __restrict__ nbr, const u8* __restrict__ tflag, const float* __restrict__ cps, const float* __restrict__ cpt,
    size_t corners, float thres_cos, float* ctan)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 t = u32(c / 3), i = u32(c - size_t(t) * 3);
	u8 flag = tflag[t];
	if (flag & MK_DEGENERATE)
		return;
	const bool orient = (flag & MK_ORIENT) != 0;
	const u32 w = wtri[c];
	const u32 max_steps = u32(corners / 3);
	MkNeverBlocked never;

	MkTspaceSum sum;
	sum.tflag = tflag;
	sum.cps = cps;
	sum.cpt = cpt;
	sum.thres_cos = thres_cos;
	sum.self = u32(c);
	sum.self_any = (flag & MK_ANY) != 0;
	for (int k = 0; k < 3; ++k)
	{
		sum.self_os[k] = cps[c * 4 + k];
		sum.self_ot[k] = cpt[c * 3 + k];
		sum.res[k] = 0.0f;
	}
	sum.seeded = false;

	MkCollect fan;
	fan.count = 0;
	mk_walk_fan(wtri, nbr, tflag, t, i, w, orient, max_steps, true, fan, never);
	if (fan.count <= MK_FAN)
	{
		for (u32 a = 1; a < fan.count; ++a)
		{
			u32 v = fan.members[a];
			u32 b = a;
			while (b > 0 && fan.members[b - 1] > v)
			{
				fan.members[b] = fan.members[b - 1];
				--b;
			}
			fan.members[b] = v;
		}
		for (u32 a = 0; a < fan.count; ++a)
			sum.add(fan.members[a]);
	}
	else
	{
		// high-valence vertex: selection by repeated walks, no storage (quadratic like the reference's own loop, :1260)
		u32 floor_corner = MK_EMPTY;
		for (u32 a = 0; a < fan.count; ++a)
		{
			MkNextAbove next = {floor_corner, MK_EMPTY};
			mk_walk_fan(wtri, nbr, tflag, t, i, w, orient, max_steps, true, next, never);
			if (next.best == MK_EMPTY)
				break;
			sum.add(next.best);
			floor_corner = next.best;
		}
	}

	float out[4] = {1.0f, 0.0f, 0.0f, -1.0f};
	if (sum.seeded)
	{
		float r[3] = {sum.res[0], sum.res[1], sum.res[2]};
		if (mk_vnot_zero(r[0], r[1], r[2]))
		{
			float s = 1 / mk_length(r[0], r[1], r[2]);
			r[0] = s * r[0], r[1] = s * r[1], r[2] = s * r[2];
		}
		out[0] = r[0], out[1] = r[1], out[2] = r[2];
		out[3] = orient ? 1.0f : -1.0f;
	}
	for (int k = 0; k < 4; ++k)
		ctan[c * 4 + k] = out[k];
}

// ------------------------------------------------------------------------------------------------- per vertex
KERNEL k_mk_count_corners(const u32* __restrict__ indices, size_t corners, u32* counts)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	atomicAdd(&counts[indices[c]], 1u);
}

KERNEL k_mk_fill_corners(const u32* __restrict__ indices, size_t corners, const u32* __restrict__ offsets, u32* cursor, u32* list)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 v = indices[c];
	list[offsets[v] + atomicAdd(&cursor[v], 1u)] = u32(c);
}

// MikkSetTSpaceBasic is called for faces and corners in ascending order (mikktspace.cpp:366-404), so the per-vertex
// float sums (ClusterLODUtilities.cpp:647-652) run over the vertex's corners in ascending corner id; then :706-734.
KERNEL k_mk_vertex_tangent(const u8* __restrict__ vertices, u32 stride, size_t vertex_count, const u32* __restrict__ offsets, u32* list, const u32* __restrict__ wtri, const u8* __restrict__ tflag,
    const u32* __restrict__ first_good, const float* __restrict__ ctan, float* tangents4)
{
	size_t v = GTID;
	if (v >= vertex_count)
		return;
	u32 begin = offsets[v], end = offsets[v + 1];
	for (u32 a = begin + 1; a < end; ++a)
	{
		u32 x = list[a];
		u32 b = a;
		while (b > begin && list[b - 1] > x)
		{
			list[b] = list[b - 1];
			--b;
		}
		list[b] = x;
	}
	float acc[3] = {0.0f, 0.0f, 0.0f};
	float signs = 0.0f;
	for (u32 a = begin; a < end; ++a)
	{
		u32 c = list[a];
		float val[4] = {1.0f, 0.0f, 0.0f, -1.0f};
		u32 src = c;
		if (tflag[c / 3] & MK_DEGENERATE)
			src = first_good[wtri[c]];
		if (src != MK_EMPTY)
			for (int k = 0; k < 4; ++k)
				val[k] = ctan[size_t(src) * 4 + k];
		acc[0] += val[0];
		acc[1] += val[1];
		acc[2] += val[2];
		signs += val[3];
	}
	float tangent[3] = {acc[0], acc[1], acc[2]};
	const float len_sq = tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2];
	const bool finite = fabsf(tangent[0]) <= 3.402823466e+38f && fabsf(tangent[1]) <= 3.402823466e+38f && fabsf(tangent[2]) <= 3.402823466e+38f;
	if (end == begin || len_sq <= 1e-20f || !finite)
	{
		// BuildFallbackTangentFromNormal (ClusterLODUtilities.cpp:559-579) over NormalizeOrFallback (:528-557)
		const float* p = mk_vertex(vertices, stride, u32(v));
		float n[3] = {p[3], p[4], p[5]};
		const float nl = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
		if (nl <= 1e-20f)
		{
			n[0] = 0.0f, n[1] = 0.0f, n[2] = 1.0f; // the fallback (0,0,1) is already unit length: 1/sqrt(1) scales by 1
		}
		else
		{
			const float inv = 1.0f / sqrtf(nl);
			n[0] = n[0] * inv, n[1] = n[1] * inv, n[2] = n[2] * inv;
		}
		float axis[3] = {0.0f, 0.0f, 1.0f};
		if (!(fabsf(n[2]) < 0.999f))
			axis[1] = 1.0f, axis[2] = 0.0f;
		tangent[0] = axis[1] * n[2] - axis[2] * n[1];
		tangent[1] = axis[2] * n[0] - axis[0] * n[2];
		tangent[2] = axis[0] * n[1] - axis[1] * n[0];
		const float tl = tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2];
		if (tl <= 1e-20f)
		{
			tangent[0] = 1.0f, tangent[1] = 0.0f, tangent[2] = 0.0f;
		}
		else
		{
			const float inv = 1.0f / sqrtf(tl);
			tangent[0] = tangent[0] * inv, tangent[1] = tangent[1] * inv, tangent[2] = tangent[2] * inv;
		}
	}
	else
	{
		const float inv = 1.0f / sqrtf(len_sq);
		tangent[0] *= inv;
		tangent[1] *= inv;
		tangent[2] *= inv;
	}
	tangents4[v * 4 + 0] = tangent[0];
	tangents4[v * 4 + 1] = tangent[1];
	tangents4[v * 4 + 2] = tangent[2];
	tangents4[v * 4 + 3] = signs < 0.0f ? -1.0f : 1.0f;
}

KERNEL k_mk_acosf(const float* __restrict__ in, float* out, size_t n)
{
	size_t i = GTID;
	if (i >= n)
		return;
	out[i] = mk_acosf(in[i]);
}

void mikk_acosf(const float* in, float* out, size_t n)
{
	LAUNCH(k_mk_acosf, n, in, out, n);
}

size_t mikk_temp_bytes(size_t vertex_count, size_t index_count)
{
	size_t table = 1;
	while (table < vertex_count * 2)
		table <<= 1;
	// weld 4V + tables 8*table (released early), then per corner: wtri 4, nbr 4, edge sort 24, terms 28, tangent 16, list 4,
	// per triangle 25, per vertex 12; the sort scratch and the weld tables overlap with later stages
	return vertex_count * 24 + table * 8 + index_count * 96 + (index_count / 3) * 32 + (size_t(64) << 20);
}

bool mikk_tangents(const u8* vertices, u32 vertex_stride, size_t vertex_count, const u32* indices, size_t index_count, float* tangents4, Arena& temp, float* corner_tangents4)
{
	// GenerateMikkTangents' own preconditions (ClusterLODUtilities.cpp:665-675); an out-of-range index was rejected at upload
	if (vertex_stride < 32 || vertex_stride % 4 || index_count == 0 || index_count % 3 != 0 || vertex_count == 0)
		return false;
	if (index_count / 3 >= (size_t(1) << 29))
		throw Error("clodb200: MikkTSpace corner ids need triangle_count < 2^29 (mikktspace.cpp:172-176)");
	ArenaScope scope(temp);
	const size_t corners = index_count;
	const u32 T = u32(index_count / 3);
	const int vertex_bits = bits_for(vertex_count > 1 ? vertex_count - 1 : 1);
	// fThresCos for genTangSpaceDefault's 180 degrees (:227, 241)
	const float thres_cos = std::cos((180.0f * float(3.1415926535897932384626433832795)) / 180.0f); // cos(float): cosf, as in the reference

	u32* weld = temp.alloc<u32>(vertex_count);
	{
		ArenaScope tables(temp);
		size_t table_size = 1;
		while (table_size < vertex_count * 2)
			table_size <<= 1;
		u32* table_rep = temp.alloc<u32>(table_size);
		u32* table_min = temp.alloc<u32>(table_size);
		dev_memset(table_rep, 0xff, table_size * sizeof(u32));
		dev_memset(table_min, 0xff, table_size * sizeof(u32));
		LAUNCH(k_mk_weld_insert, vertex_count, vertices, vertex_stride, vertex_count, table_rep, table_min, u32(table_size - 1), weld);
		LAUNCH(k_mk_weld_resolve, vertex_count, weld, table_min, vertex_count);
	}

	u32* wtri = temp.alloc<u32>(corners);
	int* nbr = temp.alloc<int>(corners);
	u8* tflag = temp.alloc<u8>(T);
	float* tos = temp.alloc<float>(corners);
	float* tot = temp.alloc<float>(corners);
	u32* first_good = temp.alloc<u32>(vertex_count);
	u32* counters = temp.alloc<u32>(4);
	dev_memset(first_good, 0xff, vertex_count * sizeof(u32));
	dev_memset(counters, 0, 4 * sizeof(u32));
	dev_memset(nbr, 0xff, corners * sizeof(int));
	{
		ArenaScope edges(temp);
		u64* edge_key = temp.alloc<u64>(corners);
		u64* edge_key_tmp = temp.alloc<u64>(corners);
		u32* edge_val = temp.alloc<u32>(corners);
		u32* edge_val_tmp = temp.alloc<u32>(corners);
		LAUNCH(k_mk_tri_info, T, vertices, vertex_stride, indices, weld, T, vertex_bits, wtri, tflag, tos, tot, edge_key, edge_val, first_good, counters);
		// stable: equal (lo, hi) keep ascending corner id, i.e. ascending triangle (the reference's third sort key, :1540-1552)
		radix_sort_pairs<u64>(edge_key, edge_key_tmp, edge_val, edge_val_tmp, corners, 0, 2 * vertex_bits, temp);
		LAUNCH(k_mk_pair_edges, corners, edge_key, edge_val, corners, wtri, vertex_bits, nbr);
	}

	if (dev_read(counters) != 0)
	{
		ArenaScope any(temp);
		u8* state_a = temp.alloc<u8>(T);
		u8* state_b = temp.alloc<u8>(T);
		LAUNCH(k_mk_any_init, T, tflag, state_a, T);
		for (int round = 0; round < 256; ++round)
		{
			dev_memset(counters + 1, 0, sizeof(u32));
			LAUNCH(k_mk_any_round, T, wtri, nbr, tflag, state_a, state_b, T, counters + 1);
			u8* s = state_a;
			state_a = state_b;
			state_b = s;
			if (dev_read(counters + 1) == 0)
				break;
		}
	}

	float* cps = temp.alloc<float>(corners * 4);
	float* cpt = temp.alloc<float>(corners * 3);
	float* ctan = temp.alloc<float>(corners * 4);
	LAUNCH(k_mk_corner_terms, corners, vertices, vertex_stride, wtri, tflag, tos, tot, corners, cps, cpt);
	if (corner_tangents4)
		dev_memset(ctan, 0, corners * 16);
	LAUNCH(k_mk_corner_tspace, corners, wtri, nbr, tflag, cps, cpt, corners, thres_cos, ctan);
	if (corner_tangents4)
		dev_d2d(corner_tangents4, ctan, corners * 16);

	u32* offsets = temp.alloc<u32>(vertex_count + 1);
	u32* cursor = temp.alloc<u32>(vertex_count);
	u32* list = temp.alloc<u32>(corners);
	dev_memset(offsets, 0, (vertex_count + 1) * sizeof(u32));
	dev_memset(cursor, 0, vertex_count * sizeof(u32));
	LAUNCH(k_mk_count_corners, corners, indices, corners, offsets);
	exclusive_scan_u32(offsets, offsets, vertex_count + 1, nullptr, temp);
	LAUNCH(k_mk_fill_corners, corners, indices, corners, offsets, cursor, list);
	LAUNCH(k_mk_vertex_tangent, vertex_count, vertices, vertex_stride, vertex_count, offsets, list, wtri, tflag, first_good, ctan, tangents4);
	return true;
}

} // namespace clodb
