// Outer boundary of the cluster-LOD build (see artifacts.h). Reference: BuildClusterLODArtifactsFromGeometry,
// BasicRenderer/src/Mesh/ClusterLODUtilities.cpp:5325-5766 ("CLU.cpp" below) and the functions it runs around clodBuildEx:
//   a16 BuildClusterLODGroupOutput            CLU.cpp:856-1805   buckets, page binning, segments, SoA page blobs
//   a17 BuildClusterLODTraversalHierarchy     CLU.cpp:4606-4963  parent errors, 8-wide traversal nodes
//   a18 FinalizeMeshWidePagePacking           CLU.cpp:2313-2540  greedy re-binning of segments into mesh pages
//
// Division of work. The DAG build keeps every level's cluster tables in HBM (no index list crosses PCIe). The host only
// ever sees O(clusters) metadata: it replays the reference's serial decisions (bucket order, greedy page fills, segment
// runs, node layout) on per-meshlet byte counts, which fixes where every byte of every mesh page goes. One kernel then
// writes all pages straight from the level tables and the interleaved vertex stream: a warp per meshlet rebuilds the
// first-occurrence vertex table (clodLocalIndices semantics, clusterlod.h:972-1023) in shared memory and streams out
// positions, oct-encoded normals, colours, UV bitstreams, triangle bytes and the 64-byte descriptor. The reference's
// intermediate per-group page blobs (a16) are never materialised: a18 copies their meshlet payloads verbatim into the mesh
// pages, so writing the mesh pages directly gives the same bytes.
#include "artifacts.h"

#include <chrono>
#include <memory>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <fstream>

namespace clodb
{

// ---- reference PODs (sizes checked against the reference headers in tests/test_abi.py) --------------------------------
struct ClodGroup // ClusterLODGroup, ClusterLODShaderTypes.h:119-139
{
	float bounds[5];
	u32 firstMeshlet, meshletCount;
	int depth;
	u32 firstGroupVertex, groupVertexCount, firstSegment, segmentCount, terminalSegmentCount, flags, pageMapBase, pageCount;
	int parentGroupId;
	float maxParentError, representationError;
};
struct ClodSegment // ClusterLODGroupSegment, ClusterLODShaderTypes.h:110-116
{
	int refinedGroup;
	u32 firstMeshletInPage, meshletCount, pageIndex;
};
struct ClodChunk // ClusterLODGroupChunk, ClusterLODShaderTypes.h:97-104
{
	u32 groupVertexCount, meshletCount, meshletTrianglesByteCount, compressedPositionQuantExp, compressedFlags;
};
struct ClodNode // ClusterLODNode, ClusterLODTypes.h:36-56
{
	u32 isGroup, indexOrOffset, countMinusOne, ownerGroupId;
	float cullingSphere[4], lodBoundingSphere[4], maxQuadricError, padding[3];
};
struct ClodNodeRange
{
	u32 offset, count;
};
struct ClodDiskLocator // ClusterLODGroupDiskLocator, ClusterLODTypes.h:68-73
{
	u64 blobOffset;
	u32 blobSizeBytes, reserved;
};
static_assert(sizeof(ClodGroup) == 76 && sizeof(ClodSegment) == 16 && sizeof(ClodChunk) == 20 && sizeof(ClodNode) == 64 && sizeof(ClodDiskLocator) == 16, "reference POD layout");

static const u32 kPageSize = 256u * 1024u; // CLOD_PAGE_SIZE, shaders/Common/defines.h
static const u32 kPageHeaderSize = 64, kDescriptorSize = 64, kUvDescriptorSize = 32;
static const u32 kAttrNormal = 1u << 0, kAttrJoints = 1u << 1, kAttrWeights = 1u << 2, kAttrColor = 1u << 3; // CLOD_PAGE_ATTRIBUTE_*, ClusterLODShaderTypes.h:15-18
static const u32 kSkinJointOffset = 24, kSkinInfluenceBytes = 64; // skinning vertex: pos f3, normal f3, then PackedSkinningInfluences (CLU.cpp:50-56, :1100)
static const u32 kMaxSkinInfluences = 8;
static const u32 kMaxLevels = 48;

// ---- device side --------------------------------------------------------------------------------------------------------
struct LevelTable
{
	const u32* tri[kMaxLevels];
};

// one meshlet as the kernels see it: where its triangles live and where its payload goes
struct MeshletJob
{
	u32 level, tri_begin, tri_count, vertex_count;
	u32 group;     // owning group id (descriptor sourceGroupLocalIndex, CLU.cpp:2182)
	u32 tri_word;  // triangleCount:16 | refinedGroup+1:16 (CLU.cpp:1631-1635)
	u32 page;      // mesh page index
	u32 slot;      // meshlet index inside the page
	u32 pos_cursor, attr_cursor, tri_cursor;
	u32 bone_cursor;  // first entry of the meshlet's bone list in the page's bone-index stream (descriptor boneListOffset)
	float bounds[4];
	u32 bone_count;   // distinct joints with a positive weight among the meshlet's vertices (descriptor boneCount)
	u32 bone_scratch; // where the pre-pass left the sorted list (u32 index into the scratch array)
	u32 pad[2];
};

struct PageRecord
{
	u64 base;                  // byte offset of the page in the output buffer
	u32 header[16];            // CLodPageHeader
	u32 uv_stream[kMaxUvSets]; // uvBitstreamOffsets
	u32 pad[2];
};

struct UvJob // per (meshlet, uv set)
{
	u32 bit_cursor;
	float min_u, min_v;
	u32 bits; // bitsU | bitsV << 8
};

struct VertexStreams
{
	const u8* vertices;
	u32 stride;
	u32 normal_offset, color_offset; // byte offsets, 0xffffffff = absent
	u32 uv_set_count;
	const float* uv_values[kMaxUvSets];
	u32 uv_stride[kMaxUvSets];
	// skinned meshes (CLU.cpp:1091-1120): a second vertex stream that carries two uint4 joints and two float4 weights per vertex
	const u8* skinning; // null = no skinning stream
	u32 skinning_stride;
	u32 skinning_count; // vertices beyond this read as zero influences (:1107-1110)
};

// The meshlet's bone list (CLU.cpp:1240-1265): joints that carry a positive weight on any of its vertices, sorted ascending, unique.
// Built by insertion into `list` (room for 8 entries per vertex); returns the count. One thread per meshlet: skinned meshes are
// character-sized and the lists are a few dozen entries.
DEVFN u32 meshlet_bone_list(const VertexStreams& vs, const u32* vertices, u32 count, u32* list)
{
	u32 n = 0;
	for (u32 vi = 0; vi < count; ++vi)
	{
		if (vertices[vi] >= vs.skinning_count)
			continue;
		const u8* sv = vs.skinning + size_t(vertices[vi]) * vs.skinning_stride + kSkinJointOffset;
		const u32* joints = reinterpret_cast<const u32*>(sv);
		const float* weights = reinterpret_cast<const float*>(sv + 32);
		for (u32 k = 0; k < kMaxSkinInfluences; ++k)
		{
			if (!(weights[k] > 0.0f))
				continue;
			u32 j = joints[k];
			u32 lo = 0, hi = n;
			while (lo < hi)
			{
				u32 mid = (lo + hi) / 2;
				if (list[mid] < j)
					lo = mid + 1;
				else
					hi = mid;
			}
			if (lo < n && list[lo] == j)
				continue;
			for (u32 q = n; q > lo; --q)
				list[q] = list[q - 1];
			list[lo] = j;
			++n;
		}
	}
	return n;
}

// joints (two uint4) and weights (two float4) of one meshlet vertex into the page arrays; vertices without a skinning record get zeros
DEVFN void write_skinning(const VertexStreams& vs, u32 vertex, u32* joints_out, u32* weights_out)
{
	if (vertex < vs.skinning_count)
	{
		const u32* src = reinterpret_cast<const u32*>(vs.skinning + size_t(vertex) * vs.skinning_stride + kSkinJointOffset);
		for (int k = 0; k < 8; ++k)
		{
			joints_out[k] = src[k];
			weights_out[k] = src[8 + k];
		}
	}
	else
		for (int k = 0; k < 8; ++k)
			joints_out[k] = weights_out[k] = 0;
}

KERNEL k_split_vertex_streams(const u8* __restrict__ vertices, u32 stride, size_t V, float* positions3, float* attributes, u32 astride, u32 with_normals, const float* __restrict__ tangents4)
{
	size_t i = GTID;
	if (i >= V)
		return;
	const float* src = reinterpret_cast<const float*>(vertices + i * stride);
	if (positions3)
	{
		positions3[i * 3 + 0] = src[0];
		positions3[i * 3 + 1] = src[1];
		positions3[i * 3 + 2] = src[2];
	}
	if (attributes)
	{
		float* dst = attributes + i * astride;
		u32 k = 0;
		if (with_normals)
		{
			dst[0] = src[3];
			dst[1] = src[4];
			dst[2] = src[5];
			k = 3;
		}
		if (tangents4)
			for (int j = 0; j < 4; ++j)
				dst[k + j] = tangents4[i * 4 + j];
	}
}

KERNEL k_write_tangent_columns(const float* __restrict__ tangents4, size_t vertex_count, float* attributes, u32 attribute_stride, u32 tangent_column)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	for (u32 j = 0; j < 4; ++j)
		attributes[i * attribute_stride + tangent_column + j] = tangents4[i * 4 + j];
}

void write_tangent_columns(const float* tangents4, size_t vertex_count, float* attributes, u32 attribute_stride, u32 tangent_column)
{
	LAUNCH(k_write_tangent_columns, vertex_count, tangents4, vertex_count, attributes, attribute_stride, tangent_column);
}

void split_vertex_streams(const u8* vertices, u32 vertex_stride, size_t vertex_count, float* positions3, float* attributes, u32 attribute_stride, bool with_normals, const float* tangents4)
{
	LAUNCH(k_split_vertex_streams, vertex_count, vertices, vertex_stride, vertex_count, positions3, attributes, attribute_stride, with_normals ? 1u : 0u, tangents4);
}

// OctEncodeNormal + PackOctNormalSnorm16 (CLU.cpp:458-493), same expression order
HOSTDEVFN u32 pack_oct_normal(float nx, float ny, float nz)
{
	const float denom = fabsf(nx) + fabsf(ny) + fabsf(nz);
	if (denom > 1e-8f)
	{
		nx /= denom;
		ny /= denom;
		nz /= denom;
	}
	if (nz < 0.0f)
	{
		const float ox = nx;
		nx = (1.0f - fabsf(ny)) * (ox >= 0.0f ? 1.0f : -1.0f);
		ny = (1.0f - fabsf(ox)) * (ny >= 0.0f ? 1.0f : -1.0f);
	}
	float cx = fmaxf(-1.0f, fminf(1.0f, nx)), cy = fmaxf(-1.0f, fminf(1.0f, ny));
	int qx = int(roundf(cx * 32767.0f)), qy = int(roundf(cy * 32767.0f));
	return u32(uint16_t(int16_t(qx))) | (u32(uint16_t(int16_t(qy))) << 16);
}

// PackColorUnorm8 (CLU.cpp:495-506)
HOSTDEVFN u32 pack_color(float r, float g, float b)
{
	float c[3] = {r, g, b};
	u32 q[3];
	for (int k = 0; k < 3; ++k)
	{
		float v = c[k] < 0.0f ? 0.0f : (1.0f < c[k] ? 1.0f : c[k]);
		q[k] = u32(lroundf(v * 255.0f));
	}
	return q[0] | (q[1] << 8) | (q[2] << 16) | (0xFFu << 24);
}

// QuantizeUvOffset (CLU.cpp:508-516)
HOSTDEVFN u32 quantize_uv_offset(float value)
{
	long long scaled = llround(double(value) * 65535.0);
	if (scaled < 0)
		scaled = 0;
	if (scaled > 0xffffffffll)
		scaled = 0xffffffffll;
	return u32(scaled);
}

DEVFN void append_bits(u32* words, u64 bit_cursor, u32 value, u32 bit_count) // AppendBits, CLU.cpp:82-113, on zeroed words
{
	if (bit_count == 0)
		return;
	const u32 bit_offset = u32(bit_cursor & 31);
	const u64 word = bit_cursor >> 5;
	const u64 mask = bit_count >= 32 ? 0xffffffffull : ((1ull << bit_count) - 1ull);
	const u64 clamped = u64(value) & mask;
	atomicOr(&words[word], u32(clamped << bit_offset));
	if (bit_offset + bit_count > 32)
		atomicOr(&words[word + 1], u32(clamped >> (32 - bit_offset)));
}

DEVFN u32 vertex_hash(u32 v)
{
	return (v * 0x9E3779B1u) >> 23; // 9 bits
}

// (group, vertex) set shared by all meshlets: the number of first insertions per group is ClusterLODGroup::groupVertexCount
// (distinct vertices of the group, CLU.cpp:886-901, 984)
// returns true when this call inserted the pair (exactly one caller per distinct pair sees true)
DEVFN bool group_vertex_insert(u64* table, u64 mask, u32 group, u32 v)
{
	u64 key = (u64(group) << 32) | v;
	u64 h = (key * 0x9E3779B97F4A7C15ull) >> 20;
	for (;;)
	{
		h &= mask;
		unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&table[h]), ~0ull, (unsigned long long)key);
		if (old == ~0ull)
			return true;
		if (old == key)
			return false;
		h++;
	}
}

// slot of an inserted (group, vertex) pair (the pre-pass inserted every pair the meshlets reference)
DEVFN u64 group_vertex_slot(const u64* table, u64 mask, u32 group, u32 v)
{
	u64 key = (u64(group) << 32) | v;
	u64 h = (key * 0x9E3779B97F4A7C15ull) >> 20;
	for (;;)
	{
		h &= mask;
		if (table[h] == key)
			return h;
		h++;
	}
}

// RecalculateGroupNormals (CLU.cpp:739-822, preserveImportedNormals = false): per group vertex the UNNORMALISED face normals of the
// group's triangles, summed in the reference's order - meshlets in bucket order, triangles in meshlet order - so the float sums
// carry the same bits. One thread per group walks its meshlets sequentially (groups are independent); the sums live at the
// (group, vertex) slots of the pre-pass table. Off the default path (the reference never clears preserveImportedNormals itself).
struct GroupSpan
{
	u32 first, count; // meshlets [first, first + count) in group-bucket order
};

KERNEL k_group_normal_sums(const GroupSpan* __restrict__ spans, u32 G, const MeshletJob* __restrict__ jobs, LevelTable levels, VertexStreams vs, const u64* __restrict__ table, u64 table_mask, float* sums3)
{
	size_t g = GTID;
	if (g >= G)
		return;
	const GroupSpan span = spans[g];
	for (u32 m = span.first; m < span.first + span.count; ++m)
	{
		const MeshletJob job = jobs[m];
		const u32* idx = levels.tri[job.level] + size_t(job.tri_begin) * 3;
		for (u32 t = 0; t < job.tri_count; ++t)
		{
			const u32 v0 = idx[t * 3 + 0], v1 = idx[t * 3 + 1], v2 = idx[t * 3 + 2];
			const float* p0 = reinterpret_cast<const float*>(vs.vertices + size_t(v0) * vs.stride);
			const float* p1 = reinterpret_cast<const float*>(vs.vertices + size_t(v1) * vs.stride);
			const float* p2 = reinterpret_cast<const float*>(vs.vertices + size_t(v2) * vs.stride);
			const float e10x = p1[0] - p0[0], e10y = p1[1] - p0[1], e10z = p1[2] - p0[2];
			const float e20x = p2[0] - p0[0], e20y = p2[1] - p0[1], e20z = p2[2] - p0[2];
			const float nx = e10y * e20z - e10z * e20y, ny = e10z * e20x - e10x * e20z, nz = e10x * e20y - e10y * e20x;
			const u32 vv[3] = {v0, v1, v2};
			for (int c = 0; c < 3; ++c)
			{
				float* acc = sums3 + group_vertex_slot(table, table_mask, u32(g), vv[c]) * 3;
				acc[0] += nx;
				acc[1] += ny;
				acc[2] += nz;
			}
		}
	}
}

// NormalizeOrFallback (CLU.cpp:528-548) of the group sum with the source normal as fallback
DEVFN void recomputed_normal(const float* sum3, const float* source, float& x, float& y, float& z)
{
	const float len_sq = sum3[0] * sum3[0] + sum3[1] * sum3[1] + sum3[2] * sum3[2];
	if (len_sq <= 1e-20f)
	{
		const float fallback_sq = source[0] * source[0] + source[1] * source[1] + source[2] * source[2];
		if (fallback_sq <= 1e-20f)
		{
			x = 0.0f, y = 0.0f, z = 1.0f;
			return;
		}
		const float inv = 1.0f / sqrtf(fallback_sq);
		x = source[0] * inv, y = source[1] * inv, z = source[2] * inv;
		return;
	}
	const float inv = 1.0f / sqrtf(len_sq);
	x = sum3[0] * inv, y = sum3[1] * inv, z = sum3[2] * inv;
}

// what the page writer needs to replace the source normals by the recomputed ones (all null/zero on the default path)
struct NormalRecompute
{
	const u64* table;
	u64 table_mask;
	const float* sums3;
};

// ---- scalar kernels (one thread per meshlet): the emulation build runs these ----------------------------------------------
struct LocalTable
{
	u32 keys[512];
	u16 vals[512];
	u32 vertices[256];
	u32 count;
};

DEVFN u32 local_table_build(LocalTable& t, const u32* idx, u32 n, u8* local_ids)
{
	for (int i = 0; i < 512; ++i)
		t.keys[i] = 0xffffffffu;
	t.count = 0;
	for (u32 j = 0; j < n; ++j)
	{
		u32 v = idx[j];
		u32 h = vertex_hash(v);
		for (;;)
		{
			if (t.keys[h] == v)
				break;
			if (t.keys[h] == 0xffffffffu)
			{
				t.keys[h] = v;
				t.vals[h] = u16(t.count);
				if (t.count < 256)
					t.vertices[t.count] = v;
				t.count++;
				break;
			}
			h = (h + 1) & 511;
		}
		if (local_ids)
			local_ids[j] = u8(t.vals[h]);
	}
	return t.count;
}

KERNEL k_meshlet_prepass(const MeshletJob* __restrict__ jobs, u32 M, LevelTable levels, VertexStreams vs, u64* table, u64 table_mask, u32* group_vertex_count, float* uv_ranges, u32* errors,
    u32* bone_scratch, u32* bone_counts)
{
	size_t m = GTID;
	if (m >= M)
		return;
	MeshletJob job = jobs[m];
	const u32* idx = levels.tri[job.level] + size_t(job.tri_begin) * 3;
	LocalTable t;
	u32 count = local_table_build(t, idx, job.tri_count * 3, nullptr);
	if (count != job.vertex_count || count > 256)
	{
		atomicOr(errors, 1u);
		return;
	}
	u32 inserted = 0;
	for (u32 vi = 0; vi < count; ++vi)
		inserted += group_vertex_insert(table, table_mask, job.group, t.vertices[vi]) ? 1u : 0u;
	if (inserted)
		atomicAdd(&group_vertex_count[job.group], inserted);
	if (vs.skinning)
		bone_counts[m] = meshlet_bone_list(vs, t.vertices, count, bone_scratch + job.bone_scratch);
	for (u32 s = 0; s < vs.uv_set_count; ++s)
	{
		float mn_u = FLT_MAX, mn_v = FLT_MAX, mx_u = -FLT_MAX, mx_v = -FLT_MAX;
		for (u32 vi = 0; vi < count; ++vi)
		{
			const float* uv = vs.uv_values[s] + size_t(t.vertices[vi]) * vs.uv_stride[s];
			mn_u = fminf(mn_u, uv[0]);
			mn_v = fminf(mn_v, uv[1]);
			mx_u = fmaxf(mx_u, uv[0]);
			mx_v = fmaxf(mx_v, uv[1]);
		}
		float* out = uv_ranges + (m * vs.uv_set_count + s) * 4;
		out[0] = mn_u, out[1] = mn_v, out[2] = mx_u, out[3] = mx_v;
	}
}

DEVFN void write_descriptor(u32* desc, const MeshletJob& job)
{
	// CLodMeshletDescriptor (ClusterLODShaderTypes.h:49-75), filled as CLU.cpp:1618-1641 / :2176-2182
	desc[0] = job.pos_cursor;
	desc[1] = job.attr_cursor;
	desc[2] = job.tri_cursor;
	desc[3] = job.bone_cursor; // boneListOffset
	desc[4] = desc[5] = desc[6] = 0;
	desc[7] = (job.vertex_count & 0xFFu) << 24;
	desc[8] = job.tri_word;
	desc[9] = job.bone_count;
	desc[10] = job.group;
	desc[11] = 0;
	desc[12] = __float_as_uint(job.bounds[0]);
	desc[13] = __float_as_uint(job.bounds[1]);
	desc[14] = __float_as_uint(job.bounds[2]);
	desc[15] = __float_as_uint(job.bounds[3]);
}

KERNEL k_write_pages(const MeshletJob* __restrict__ jobs, u32 M, LevelTable levels, VertexStreams vs, const PageRecord* __restrict__ pages, const UvJob* __restrict__ uv_jobs, u8* out, u32* errors,
    const u32* __restrict__ bone_scratch, NormalRecompute recompute)
{
	size_t m = GTID;
	if (m >= M)
		return;
	MeshletJob job = jobs[m];
	const PageRecord& pr = pages[job.page];
	u8* page = out + pr.base;
	const u32* hdr = pr.header;
	const u32* idx = levels.tri[job.level] + size_t(job.tri_begin) * 3;
	LocalTable t;
	u8* tri_out = page + hdr[12] + job.tri_cursor;
	u32 count = local_table_build(t, idx, job.tri_count * 3, tri_out);
	if (count != job.vertex_count || count > 256)
	{
		atomicOr(errors, 1u);
		return;
	}
	if (job.slot == 0)
	{
		u32* h = reinterpret_cast<u32*>(page);
		for (int k = 0; k < 16; ++k)
			h[k] = hdr[k];
		for (u32 s = 0; s < vs.uv_set_count; ++s)
			reinterpret_cast<u32*>(page + hdr[11])[s] = pr.uv_stream[s];
	}
	write_descriptor(reinterpret_cast<u32*>(page + hdr[4]) + size_t(job.slot) * 16, job);
	float* pos = reinterpret_cast<float*>(page + hdr[6] + job.pos_cursor);
	u32* nrm = hdr[7] ? reinterpret_cast<u32*>(page + hdr[7]) + job.attr_cursor : nullptr;
	u32* col = hdr[8] ? reinterpret_cast<u32*>(page + hdr[8]) + job.attr_cursor : nullptr;
	for (u32 vi = 0; vi < count; ++vi)
	{
		const u8* v = vs.vertices + size_t(t.vertices[vi]) * vs.stride;
		const float* p = reinterpret_cast<const float*>(v);
		pos[vi * 3 + 0] = p[0];
		pos[vi * 3 + 1] = p[1];
		pos[vi * 3 + 2] = p[2];
		if (nrm)
		{
			const float* n = reinterpret_cast<const float*>(v + vs.normal_offset);
			if (recompute.sums3)
			{
				float x, y, z;
				recomputed_normal(recompute.sums3 + group_vertex_slot(recompute.table, recompute.table_mask, job.group, t.vertices[vi]) * 3, n, x, y, z);
				nrm[vi] = pack_oct_normal(x, y, z);
			}
			else
				nrm[vi] = pack_oct_normal(n[0], n[1], n[2]);
		}
		if (col)
		{
			const float* c = reinterpret_cast<const float*>(v + vs.color_offset);
			col[vi] = pack_color(c[0], c[1], c[2]);
		}
		if (hdr[9])
			write_skinning(vs, t.vertices[vi], reinterpret_cast<u32*>(page + hdr[9]) + size_t(job.attr_cursor + vi) * 8, reinterpret_cast<u32*>(page + hdr[10]) + size_t(job.attr_cursor + vi) * 8);
	}
	for (u32 b = 0; b < job.bone_count; ++b)
		reinterpret_cast<u32*>(page + hdr[13])[job.bone_cursor + b] = bone_scratch[job.bone_scratch + b];
	for (u32 s = 0; s < vs.uv_set_count; ++s)
	{
		UvJob uj = uv_jobs[m * vs.uv_set_count + s];
		u32* d = reinterpret_cast<u32*>(page + hdr[5]) + (size_t(job.slot) * vs.uv_set_count + s) * 8;
		d[0] = uj.bit_cursor;
		d[1] = __float_as_uint(uj.min_u);
		d[2] = __float_as_uint(uj.min_v);
		d[3] = d[4] = __float_as_uint(1.0f / 65535.0f); // CLOD_UV_QUANTIZATION_INV_SCALE
		d[5] = uj.bits;
		d[6] = d[7] = 0;
		u32 bits_u = uj.bits & 0xFF, bits_v = (uj.bits >> 8) & 0xFF;
		u32 max_u = bits_u >= 32 ? 0xFFFFFFFFu : ((1u << bits_u) - 1u), max_v = bits_v >= 32 ? 0xFFFFFFFFu : ((1u << bits_v) - 1u);
		u32* words = reinterpret_cast<u32*>(page + pr.uv_stream[s]);
		for (u32 vi = 0; vi < count; ++vi)
		{
			const float* uv = vs.uv_values[s] + size_t(t.vertices[vi]) * vs.uv_stride[s];
			u32 eu = quantize_uv_offset(fmaxf(0.0f, uv[0] - uj.min_u)), ev = quantize_uv_offset(fmaxf(0.0f, uv[1] - uj.min_v));
			eu = eu < max_u ? eu : max_u;
			ev = ev < max_v ? ev : max_v;
			u64 cursor = u64(uj.bit_cursor) + u64(vi) * (bits_u + bits_v);
			append_bits(words, cursor, eu, bits_u);
			append_bits(words, cursor + bits_u, ev, bits_v);
		}
	}
}

#ifndef CLODB_EMU
// ---- warp kernels (sm_100a): one warp per meshlet, 4 meshlets per CTA --------------------------------------------------
static const int MW_WARPS = 4;
struct WarpTable
{
	u32 keys[512];
	u32 first[512]; // lowest corner that references the key
	u32 vertices[256];
	u8 local[512];
};

// builds the first-occurrence vertex table of one meshlet; returns the vertex count. slot_of[k] = hash slot of corner
// k * 32 + lane (kept in registers by the caller), so corner -> local id is t.local[slot]
DEVFN u32 warp_table_build(WarpTable& t, const u32* __restrict__ idx, u32 n, int lane, u16 (&slot_of)[12])
{
	for (int i = lane; i < 512; i += 32)
	{
		t.keys[i] = 0xffffffffu;
		t.first[i] = 0xffffffffu;
	}
	__syncwarp();
#pragma unroll
	for (int k = 0; k < 12; ++k)
	{
		u32 i = k * 32 + lane;
		if (i < n)
		{
			u32 v = __ldg(idx + i);
			u32 h = vertex_hash(v);
			for (;;)
			{
				u32 old = atomicCAS(&t.keys[h], 0xffffffffu, v);
				if (old == 0xffffffffu || old == v)
					break;
				h = (h + 1) & 511;
			}
			atomicMin(&t.first[h], i);
			slot_of[k] = u16(h);
		}
	}
	__syncwarp();
	u32 count = 0;
#pragma unroll
	for (int k = 0; k < 12; ++k)
	{
		u32 i = k * 32 + lane;
		bool is_first = i < n && t.first[slot_of[k]] == i;
		u32 mask = __ballot_sync(0xffffffffu, is_first);
		if (is_first)
		{
			u32 rank = count + __popc(mask & ((1u << lane) - 1u));
			if (rank < 256)
				t.vertices[rank] = t.keys[slot_of[k]];
			t.local[slot_of[k]] = u8(rank);
		}
		count += __popc(mask);
	}
	__syncwarp();
	return count;
}

static __global__ void __launch_bounds__(MW_WARPS * 32) k_meshlet_prepass_warp(const MeshletJob* __restrict__ jobs, u32 M, LevelTable levels, VertexStreams vs, u64* table, u64 table_mask, u32* group_vertex_count, float* uv_ranges, u32* errors,
    u32* bone_scratch, u32* bone_counts)
{
	__shared__ WarpTable s_tables[MW_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 m = blockIdx.x * MW_WARPS + warp;
	if (m >= M)
		return;
	MeshletJob job = jobs[m];
	WarpTable& t = s_tables[warp];
	u16 slot_of[12];
	u32 n = job.tri_count * 3;
	if (n > 384)
	{
		if (lane == 0)
			atomicOr(errors, 2u);
		return;
	}
	u32 count = warp_table_build(t, levels.tri[job.level] + size_t(job.tri_begin) * 3, n, lane, slot_of);
	if (count != job.vertex_count || count > 256)
	{
		if (lane == 0)
			atomicOr(errors, 1u);
		return;
	}
	// one counter update per meshlet: every lane's first insertions are summed across the warp first (the group counter is
	// shared by up to 512 meshlets x 128 vertices, per-vertex atomics on it serialise)
	u32 inserted = 0;
	for (u32 vi = lane; vi < count; vi += 32)
		inserted += group_vertex_insert(table, table_mask, job.group, t.vertices[vi]) ? 1u : 0u;
	for (int o = 16; o > 0; o >>= 1)
		inserted += __shfl_xor_sync(0xffffffffu, inserted, o);
	if (lane == 0 && inserted)
		atomicAdd(&group_vertex_count[job.group], inserted);
	if (vs.skinning && lane == 0)
		bone_counts[m] = meshlet_bone_list(vs, t.vertices, count, bone_scratch + job.bone_scratch);
	for (u32 s = 0; s < vs.uv_set_count; ++s)
	{
		float mn_u = FLT_MAX, mn_v = FLT_MAX, mx_u = -FLT_MAX, mx_v = -FLT_MAX;
		for (u32 vi = lane; vi < count; vi += 32)
		{
			const float* uv = vs.uv_values[s] + size_t(t.vertices[vi]) * vs.uv_stride[s];
			mn_u = fminf(mn_u, uv[0]);
			mn_v = fminf(mn_v, uv[1]);
			mx_u = fmaxf(mx_u, uv[0]);
			mx_v = fmaxf(mx_v, uv[1]);
		}
		for (int o = 16; o > 0; o >>= 1)
		{
			mn_u = fminf(mn_u, __shfl_xor_sync(0xffffffffu, mn_u, o));
			mn_v = fminf(mn_v, __shfl_xor_sync(0xffffffffu, mn_v, o));
			mx_u = fmaxf(mx_u, __shfl_xor_sync(0xffffffffu, mx_u, o));
			mx_v = fmaxf(mx_v, __shfl_xor_sync(0xffffffffu, mx_v, o));
		}
		if (lane == 0)
		{
			float* out = uv_ranges + (size_t(m) * vs.uv_set_count + s) * 4;
			out[0] = mn_u, out[1] = mn_v, out[2] = mx_u, out[3] = mx_v;
		}
	}
}

static __global__ void __launch_bounds__(MW_WARPS * 32) k_write_pages_warp(const MeshletJob* __restrict__ jobs, u32 M, LevelTable levels, VertexStreams vs, const PageRecord* __restrict__ pages, const UvJob* __restrict__ uv_jobs, u8* out, u32* errors,
    const u32* __restrict__ bone_scratch, NormalRecompute recompute)
{
	__shared__ WarpTable s_tables[MW_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 m = blockIdx.x * MW_WARPS + warp;
	if (m >= M)
		return;
	MeshletJob job = jobs[m];
	WarpTable& t = s_tables[warp];
	u16 slot_of[12];
	u32 n = job.tri_count * 3;
	if (n > 384)
	{
		if (lane == 0)
			atomicOr(errors, 2u);
		return;
	}
	u32 count = warp_table_build(t, levels.tri[job.level] + size_t(job.tri_begin) * 3, n, lane, slot_of);
	if (count != job.vertex_count || count > 256)
	{
		if (lane == 0)
			atomicOr(errors, 1u);
		return;
	}
	const PageRecord& pr = pages[job.page];
	u8* page = out + pr.base;
	u32 hdr_desc = pr.header[4], hdr_uvdesc = pr.header[5], hdr_pos = pr.header[6], hdr_nrm = pr.header[7], hdr_col = pr.header[8], hdr_uvdir = pr.header[11], hdr_tri = pr.header[12];
	const u32 hdr_joints = pr.header[9], hdr_weights = pr.header[10], hdr_bones = pr.header[13];

	// triangle bytes: corner k * 32 + lane -> local id; consecutive lanes write consecutive bytes
	u8* tri_out = page + hdr_tri + job.tri_cursor;
#pragma unroll
	for (int k = 0; k < 12; ++k)
	{
		u32 i = k * 32 + lane;
		if (i < n)
			tri_out[i] = t.local[slot_of[k]];
	}
	if (job.slot == 0)
	{
		if (lane < 16)
			reinterpret_cast<u32*>(page)[lane] = pr.header[lane];
		if (u32(lane) < vs.uv_set_count)
			reinterpret_cast<u32*>(page + hdr_uvdir)[lane] = pr.uv_stream[lane];
	}
	if (lane == 0)
		write_descriptor(reinterpret_cast<u32*>(page + hdr_desc) + size_t(job.slot) * 16, job);

	float* pos = reinterpret_cast<float*>(page + hdr_pos + job.pos_cursor);
	u32* nrm = hdr_nrm ? reinterpret_cast<u32*>(page + hdr_nrm) + job.attr_cursor : nullptr;
	u32* col = hdr_col ? reinterpret_cast<u32*>(page + hdr_col) + job.attr_cursor : nullptr;
	for (u32 vi = lane; vi < count; vi += 32)
	{
		const u8* v = vs.vertices + size_t(t.vertices[vi]) * vs.stride;
		const float* p = reinterpret_cast<const float*>(v);
		float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
		pos[vi * 3 + 0] = px;
		pos[vi * 3 + 1] = py;
		pos[vi * 3 + 2] = pz;
		if (nrm)
		{
			const float* nn = reinterpret_cast<const float*>(v + vs.normal_offset);
			if (recompute.sums3)
			{
				const float source[3] = {__ldg(nn), __ldg(nn + 1), __ldg(nn + 2)};
				float x, y, z;
				recomputed_normal(recompute.sums3 + group_vertex_slot(recompute.table, recompute.table_mask, job.group, t.vertices[vi]) * 3, source, x, y, z);
				nrm[vi] = pack_oct_normal(x, y, z);
			}
			else
				nrm[vi] = pack_oct_normal(__ldg(nn), __ldg(nn + 1), __ldg(nn + 2));
		}
		if (col)
		{
			const float* c = reinterpret_cast<const float*>(v + vs.color_offset);
			col[vi] = pack_color(__ldg(c), __ldg(c + 1), __ldg(c + 2));
		}
		if (hdr_joints)
			write_skinning(vs, t.vertices[vi], reinterpret_cast<u32*>(page + hdr_joints) + size_t(job.attr_cursor + vi) * 8, reinterpret_cast<u32*>(page + hdr_weights) + size_t(job.attr_cursor + vi) * 8);
	}
	for (u32 b = lane; b < job.bone_count; b += 32)
		reinterpret_cast<u32*>(page + hdr_bones)[job.bone_cursor + b] = bone_scratch[job.bone_scratch + b];
	for (u32 s = 0; s < vs.uv_set_count; ++s)
	{
		UvJob uj = uv_jobs[size_t(m) * vs.uv_set_count + s];
		if (lane == 0)
		{
			u32* d = reinterpret_cast<u32*>(page + hdr_uvdesc) + (size_t(job.slot) * vs.uv_set_count + s) * 8;
			d[0] = uj.bit_cursor;
			d[1] = __float_as_uint(uj.min_u);
			d[2] = __float_as_uint(uj.min_v);
			d[3] = d[4] = __float_as_uint(1.0f / 65535.0f);
			d[5] = uj.bits;
			d[6] = d[7] = 0;
		}
		u32 bits_u = uj.bits & 0xFF, bits_v = (uj.bits >> 8) & 0xFF;
		u32 max_u = bits_u >= 32 ? 0xFFFFFFFFu : ((1u << bits_u) - 1u), max_v = bits_v >= 32 ? 0xFFFFFFFFu : ((1u << bits_v) - 1u);
		u32* words = reinterpret_cast<u32*>(page + pr.uv_stream[s]);
		for (u32 vi = lane; vi < count; vi += 32)
		{
			const float* uv = vs.uv_values[s] + size_t(t.vertices[vi]) * vs.uv_stride[s];
			u32 eu = quantize_uv_offset(fmaxf(0.0f, __ldg(uv) - uj.min_u)), ev = quantize_uv_offset(fmaxf(0.0f, __ldg(uv + 1) - uj.min_v));
			eu = eu < max_u ? eu : max_u;
			ev = ev < max_v ? ev : max_v;
			u64 cursor = u64(uj.bit_cursor) + u64(vi) * (bits_u + bits_v);
			append_bits(words, cursor, eu, bits_u);
			append_bits(words, cursor + bits_u, ev, bits_v);
		}
	}
}
#endif

// ---- host side: the reference's serial bookkeeping on per-meshlet byte counts -------------------------------------------
namespace
{

size_t align4(size_t v)
{
	return (v + 3u) & ~size_t(3);
}

u32 bits_needed_for_range(u32 range) // BitsNeededForRange, CLU.cpp:58-65
{
	if (range == 0)
		return 1;
	return 32u - u32(__builtin_clz(range));
}

struct PageTotals // PageTotals / TriangleMeshPageBuildTotals, CLU.cpp:1338-1346, 1883-1893
{
	u32 meshlets = 0, position_bytes = 0, vertex_count = 0, triangle_bytes = 0, bone_indices = 0;
	u64 uv_bits[kMaxUvSets] = {};
};

// ComputePageBlobSize, CLU.cpp:368-417
size_t page_blob_size(u32 mask, u32 uv_set_count, const PageTotals& t)
{
	size_t size = kPageHeaderSize;
	size = align4(size) + align4(size_t(t.meshlets) * kDescriptorSize);
	if (uv_set_count > 0)
		size = align4(size) + align4(size_t(t.meshlets) * uv_set_count * kUvDescriptorSize);
	size = align4(size) + align4(size_t(t.position_bytes));
	if (mask & kAttrNormal)
		size = align4(size) + align4(size_t(t.vertex_count) * 4);
	if (mask & kAttrColor)
		size = align4(size) + align4(size_t(t.vertex_count) * 4);
	if (mask & kAttrJoints)
		size = align4(size) + align4(size_t(t.vertex_count) * 32);
	if (mask & kAttrWeights)
		size = align4(size) + align4(size_t(t.vertex_count) * 32);
	if (uv_set_count > 0)
	{
		size = align4(size) + align4(size_t(uv_set_count) * 4);
		for (u32 s = 0; s < uv_set_count; ++s)
			size = align4(size) + align4(size_t((t.uv_bits[s] + 31ull) / 32ull) * 4);
	}
	size = align4(size) + align4(size_t(t.bone_indices) * 4); // bone index stream
	size = align4(size) + align4(size_t(t.triangle_bytes));
	return align4(size);
}

// stream offsets of a page, CLU.cpp:2097-2135 (identical to :1510-1571)
void page_layout(u32 mask, u32 uv_set_count, const PageTotals& t, PageRecord& pr, u32& total_size)
{
	const bool has_normals = (mask & kAttrNormal) != 0, has_colors = (mask & kAttrColor) != 0, has_uv = uv_set_count > 0;
	const u32 descriptor_offset = u32(align4(kPageHeaderSize));
	const size_t descriptor_bytes = size_t(t.meshlets) * kDescriptorSize;
	const u32 uv_descriptor_offset = has_uv ? u32(align4(descriptor_offset + descriptor_bytes)) : 0u;
	const size_t uv_descriptor_bytes = has_uv ? size_t(t.meshlets) * uv_set_count * kUvDescriptorSize : 0u;
	const u32 position_offset = u32(align4(has_uv ? (uv_descriptor_offset + uv_descriptor_bytes) : (descriptor_offset + descriptor_bytes)));
	const size_t position_bytes = t.position_bytes;
	const u32 normal_offset = has_normals ? u32(align4(position_offset + position_bytes)) : 0u;
	const size_t normal_bytes = has_normals ? size_t(t.vertex_count) * 4 : 0u;
	const u32 color_offset = has_colors ? u32(align4(has_normals ? (normal_offset + normal_bytes) : (position_offset + position_bytes))) : 0u;
	const size_t color_bytes = has_colors ? size_t(t.vertex_count) * 4 : 0u;
	const bool has_joints = (mask & kAttrJoints) != 0, has_weights = (mask & kAttrWeights) != 0;
	const size_t attributes_end = has_colors ? (color_offset + color_bytes) : (has_normals ? (normal_offset + normal_bytes) : (position_offset + position_bytes));
	const u32 joint_offset = has_joints ? u32(align4(attributes_end)) : 0u;
	const size_t joint_bytes = has_joints ? size_t(t.vertex_count) * 32 : 0u;
	const u32 weight_offset = has_weights ? u32(align4(has_joints ? (joint_offset + joint_bytes) : attributes_end)) : 0u;
	const size_t weight_bytes = has_weights ? size_t(t.vertex_count) * 32 : 0u;
	const size_t streams_end = has_weights ? (weight_offset + weight_bytes) : (has_joints ? (joint_offset + joint_bytes) : attributes_end);
	const u32 uv_directory_offset = has_uv ? u32(align4(streams_end)) : 0u;
	size_t uv_cursor = has_uv ? align4(size_t(uv_directory_offset) + size_t(uv_set_count) * 4) : align4(streams_end);
	for (u32 s = 0; s < uv_set_count; ++s)
	{
		pr.uv_stream[s] = u32(uv_cursor);
		uv_cursor = align4(uv_cursor + size_t((t.uv_bits[s] + 31ull) / 32ull) * 4);
	}
	const u32 bone_offset = u32(align4(uv_cursor));
	const u32 triangle_offset = u32(align4(bone_offset + size_t(t.bone_indices) * 4));
	total_size = u32(align4(triangle_offset + t.triangle_bytes));
	u32* h = pr.header; // CLodPageHeader, ClusterLODShaderTypes.h:26-45
	memset(h, 0, 64);
	h[0] = t.meshlets;
	h[1] = 1; // CLOD_POSITION_FORMAT_FLOAT3
	h[2] = mask;
	h[3] = uv_set_count;
	h[4] = descriptor_offset;
	h[5] = uv_descriptor_offset;
	h[6] = position_offset;
	h[7] = normal_offset;
	h[8] = color_offset;
	h[9] = joint_offset;
	h[10] = weight_offset;
	h[11] = uv_directory_offset;
	h[12] = triangle_offset;
	h[13] = bone_offset;
}

// computeBoundingSphere with 7 axes = meshopt_computeSphereBounds (clusterizer.cpp:176-284, 1655-1680); the points are
// {center xyz} with stride `stride` floats, radii likewise
void sphere_bounds(float result[4], const float* points, size_t count, size_t stride, const float* radii, size_t radii_stride)
{
	static const float kAxes[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.57735026f, 0.57735026f, 0.57735026f}, {-0.57735026f, 0.57735026f, 0.57735026f}, {0.57735026f, -0.57735026f, 0.57735026f},
	    {0.57735026f, 0.57735026f, -0.57735026f}};
	result[0] = result[1] = result[2] = result[3] = 0.f;
	if (count == 0)
		return;
	size_t pmin[7], pmax[7];
	float tmin[7], tmax[7];
	for (int axis = 0; axis < 7; ++axis)
	{
		pmin[axis] = pmax[axis] = 0;
		tmin[axis] = FLT_MAX;
		tmax[axis] = -FLT_MAX;
	}
	for (size_t i = 0; i < count; ++i)
	{
		const float* p = points + i * stride;
		float r = radii[i * radii_stride];
		for (int axis = 0; axis < 7; ++axis)
		{
			const float* ax = kAxes[axis];
			float tp = ax[0] * p[0] + ax[1] * p[1] + ax[2] * p[2];
			float tpmin = tp - r, tpmax = tp + r;
			pmin[axis] = (tpmin < tmin[axis]) ? i : pmin[axis];
			pmax[axis] = (tpmax > tmax[axis]) ? i : pmax[axis];
			tmin[axis] = (tpmin < tmin[axis]) ? tpmin : tmin[axis];
			tmax[axis] = (tpmax > tmax[axis]) ? tpmax : tmax[axis];
		}
	}
	size_t paxis = 0;
	float paxisdr = 0;
	for (int axis = 0; axis < 7; ++axis)
	{
		const float* p1 = points + pmin[axis] * stride;
		const float* p2 = points + pmax[axis] * stride;
		float r1 = radii[pmin[axis] * radii_stride], r2 = radii[pmax[axis] * radii_stride];
		float d2 = (p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]) + (p2[2] - p1[2]) * (p2[2] - p1[2]);
		float dr = sqrtf(d2) + r1 + r2;
		if (dr > paxisdr)
		{
			paxisdr = dr;
			paxis = size_t(axis);
		}
	}
	const float* p1 = points + pmin[paxis] * stride;
	const float* p2 = points + pmax[paxis] * stride;
	float r1 = radii[pmin[paxis] * radii_stride], r2 = radii[pmax[paxis] * radii_stride];
	float paxisd = sqrtf((p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]) + (p2[2] - p1[2]) * (p2[2] - p1[2]));
	float paxisk = paxisd > 0 ? (paxisd + r2 - r1) / (2 * paxisd) : 0.f;
	float center[3] = {p1[0] + (p2[0] - p1[0]) * paxisk, p1[1] + (p2[1] - p1[1]) * paxisk, p1[2] + (p2[2] - p1[2]) * paxisk};
	float radius = paxisdr / 2;
	for (size_t i = 0; i < count; ++i)
	{
		const float* p = points + i * stride;
		float r = radii[i * radii_stride];
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
	result[0] = center[0], result[1] = center[1], result[2] = center[2], result[3] = radius;
}

// meshopt_spatialClusterPoints (spatialorder.cpp:307-341; computeOrder :25-68, splitPoints :158-214)
void split_points(u32* destination, u32* orderx, u32* ordery, u32* orderz, const u64* keys, size_t count, std::vector<u32>& temp, std::vector<u8>& sides, size_t cluster_size)
{
	if (count <= cluster_size)
	{
		memcpy(destination, orderx, count * sizeof(u32));
		return;
	}
	u32* axes[3] = {orderx, ordery, orderz};
	int bestk = -1;
	u32 bestdim = 0;
	for (int k = 0; k < 3; ++k)
	{
		const u32 mask = (1u << 20) - 1;
		u32 dim = (u32(keys[axes[k][count - 1]] >> (k * 20)) & mask) - (u32(keys[axes[k][0]] >> (k * 20)) & mask);
		if (dim >= bestdim)
		{
			bestk = k;
			bestdim = dim;
		}
	}
	size_t split = ((count / 2) + cluster_size - 1) / cluster_size * cluster_size;
	for (size_t i = 0; i < split; ++i)
		sides[axes[bestk][i]] = 0;
	for (size_t i = split; i < count; ++i)
		sides[axes[bestk][i]] = 1;
	for (int k = 0; k < 3; ++k)
	{
		if (k == bestk)
			continue;
		u32* axis = axes[k];
		memcpy(temp.data(), axis, sizeof(u32) * count);
		size_t l = 0, r = split;
		for (size_t i = 0; i < count; ++i)
		{
			u8 side = sides[temp[i]];
			axis[side ? r : l] = temp[i];
			l += 1;
			l -= side;
			r += side;
		}
	}
	split_points(destination, orderx, ordery, orderz, keys, split, temp, sides, cluster_size);
	split_points(destination + split, orderx + split, ordery + split, orderz + split, keys, count - split, temp, sides, cluster_size);
}

void spatial_cluster_points(u32* destination, const float* points, size_t count, size_t stride, size_t cluster_size)
{
	float minv[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, maxv[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (size_t i = 0; i < count; ++i)
		for (int j = 0; j < 3; ++j)
		{
			float vj = points[i * stride + j];
			minv[j] = minv[j] > vj ? vj : minv[j];
			maxv[j] = maxv[j] < vj ? vj : maxv[j];
		}
	float extent = 0.f;
	for (int j = 0; j < 3; ++j)
		extent = (maxv[j] - minv[j]) < extent ? extent : (maxv[j] - minv[j]);
	float scale = extent == 0 ? 0.f : 65535.f / extent;
	std::vector<u64> keys(count);
	for (size_t i = 0; i < count; ++i)
	{
		const float* v = points + i * stride;
		int x = int((v[0] - minv[0]) * scale + 0.5f), y = int((v[1] - minv[1]) * scale + 0.5f), z = int((v[2] - minv[2]) * scale + 0.5f);
		keys[i] = (u64(x) << 0) | (u64(y) << 20) | (u64(z) << 40);
	}
	std::vector<u32> order(count * 3);
	for (int k = 0; k < 3; ++k)
	{
		u32* o = order.data() + size_t(k) * count;
		for (size_t i = 0; i < count; ++i)
			o[i] = u32(i);
		// two stable 8-bit radix passes over the low 16 bits of the axis key == one stable sort by that 16-bit key
		std::stable_sort(o, o + count, [&](u32 a, u32 b) { return uint16_t(keys[a] >> (k * 20)) < uint16_t(keys[b] >> (k * 20)); });
	}
	std::vector<u32> temp(count);
	std::vector<u8> sides(count);
	split_points(destination, order.data(), order.data() + count, order.data() + 2 * count, keys.data(), count, temp, sides, cluster_size);
}

struct MeshletRec
{
	u32 level, tri_begin, tri_count, vertex_count;
	int refined;
	float bounds[4];
};

struct GroupRec
{
	int depth;
	float simplified[5];
	u32 first, count;
};

struct ArtifactSink : DagSink
{
	std::vector<const u32*> level_tri;
	std::vector<MeshletRec> meshlets; // emission order: group-major, callback cluster order
	std::vector<GroupRec> groups;

	bool wants_indices() const override
	{
		return false;
	}
	void begin_level(const ClusterSet& level, int) override
	{
		level_tri.push_back(level.tri);
	}
	bool emit_level_bulk(const LevelBulk& b, std::vector<int>& group_ids) override
	{
		const u32 level = u32(level_tri.size() - 1);
		const size_t base_group = groups.size(), base_meshlet = meshlets.size();
		if (base_meshlet == 0)
		{
			// the DAG halves per level: all levels together hold about twice the clusters of the first. Growing the vectors level by
			// level would copy tens of MB at every reallocation.
			meshlets.reserve(size_t(b.cluster_count) * 2 + size_t(b.cluster_count) / 4 + 1024);
			groups.reserve(size_t(b.group_count) * 2 + size_t(b.group_count) / 2 + 64);
		}
		groups.resize(base_group + b.group_count);
		meshlets.resize(base_meshlet + b.cluster_count);
		host_parallel_for(b.group_count, 16, [&](size_t g_begin, size_t g_end, size_t) {
			for (size_t g = g_begin; g < g_end; ++g)
			{
				const u32 first = b.group_cluster_offset[g], last = b.group_cluster_offset[g + 1];
				GroupRec& gr = groups[base_group + g];
				gr.depth = b.depth;
				memcpy(gr.simplified, b.group_bounds5 + g * 5, sizeof(gr.simplified));
				gr.first = u32(base_meshlet + first);
				gr.count = last - first;
				for (u32 j = first; j < last; ++j)
				{
					const u32 c = b.group_clusters[j];
					MeshletRec& m = meshlets[base_meshlet + j];
					m.level = level;
					m.tri_begin = b.cluster_tri_offset[c];
					m.tri_count = b.cluster_tri_offset[c + 1] - b.cluster_tri_offset[c];
					m.vertex_count = b.cluster_vertex_count[c];
					m.refined = b.refined[c];
					const bool precise = b.use_precise && b.refined[c] != -1;
					const float* src = precise ? b.precise4 + size_t(c) * 4 : b.bounds5 + size_t(c) * 5;
					m.bounds[0] = src[0];
					m.bounds[1] = src[1];
					m.bounds[2] = src[2];
					m.bounds[3] = src[3];
				}
				group_ids[g] = int(base_group + g); // CaptureOutputContext::nextGroupId, CLU.cpp:5498
			}
		});
		return true;
	}
	int group(const DagGroup& group, const DagCluster* clusters, size_t cluster_count, size_t) override
	{
		GroupRec g;
		g.depth = group.depth;
		memcpy(g.simplified, group.simplified, sizeof(g.simplified));
		g.first = u32(meshlets.size());
		g.count = u32(cluster_count);
		u32 level = u32(level_tri.size() - 1);
		for (size_t i = 0; i < cluster_count; ++i)
		{
			MeshletRec m;
			m.level = level;
			m.tri_begin = level_cluster_tri_offset[cluster_ids[i]];
			m.tri_count = u32(clusters[i].index_count / 3);
			m.vertex_count = u32(clusters[i].vertex_count);
			m.refined = clusters[i].refined;
			memcpy(m.bounds, clusters[i].bounds, sizeof(m.bounds));
			meshlets.push_back(m);
		}
		groups.push_back(g);
		return int(groups.size() - 1); // CaptureOutputContext::nextGroupId, CLU.cpp:5498
	}
};

template <typename T>
void put_blob(Artifacts& a, const char* name, const std::vector<T>& v)
{
	std::vector<u8>& b = a.blobs[name];
	b.resize(v.size() * sizeof(T));
	if (!v.empty())
		memcpy(b.data(), v.data(), b.size());
}

ClodNode internal_node(const std::vector<ClodNode>& nodes, u32 child_offset, u32 child_count) // CLU.cpp:4823-4861 / :4865-4906
{
	ClodNode node;
	memset(&node, 0, sizeof(node));
	node.isGroup = 0;
	node.indexOrOffset = child_offset;
	node.countMinusOne = child_count - 1;
	const ClodNode* children = &nodes[child_offset];
	float max_err = 0.f;
	for (u32 c = 0; c < child_count; ++c)
		max_err = std::max(max_err, children[c].maxQuadricError);
	node.maxQuadricError = max_err;
	const size_t stride = sizeof(ClodNode) / sizeof(float);
	float cull[4], lod[4];
	sphere_bounds(cull, children[0].cullingSphere, child_count, stride, children[0].cullingSphere + 3, stride);
	sphere_bounds(lod, children[0].lodBoundingSphere, child_count, stride, children[0].lodBoundingSphere + 3, stride);
	for (int k = 0; k < 3; ++k)
	{
		node.cullingSphere[k] = cull[k];
		node.lodBoundingSphere[k] = lod[k];
	}
	node.cullingSphere[3] = cull[3] * (1.0f + 1e-5f);
	node.lodBoundingSphere[3] = lod[3] * (1.0f + 1e-5f);
	return node;
}

struct Hierarchy
{
	std::vector<ClodNode> nodes;
	std::vector<ClodNodeRange> ranges;
	std::vector<u32> level_roots;
	u32 max_depth = 0, max_traversal_depth = 0;
};

u32 traversal_depth(const std::vector<ClodNode>& nodes, u32 index) // ComputeCLodTraversalDepth, CLU.cpp:1819-1846
{
	if (index >= nodes.size())
		return 0;
	const ClodNode& node = nodes[index];
	if (node.isGroup != 0)
		return 1;
	u32 best = 0;
	for (u32 c = 0; c <= node.countMinusOne; ++c)
		best = std::max(best, traversal_depth(nodes, node.indexOrOffset + c));
	return 1 + best;
}

// BuildClusterLODTraversalHierarchy (CLU.cpp:4606-4963) + its validation (:4964-5253) for mesh-only builds
void build_hierarchy(std::vector<ClodGroup>& groups, const std::vector<ClodSegment>& segments, const std::vector<float>& segment_bounds, Hierarchy& h)
{
	const u32 width = 8; // TraversalNodeFanout, CLU.cpp:5440
	const u32 G = u32(groups.size());
	h.max_depth = 0;
	for (const ClodGroup& g : groups)
		h.max_depth = std::max(h.max_depth, u32(g.depth));
	const u32 levels = h.max_depth + 1;

	struct Leaf
	{
		u32 segment, owner;
		int refined;
	};
	std::vector<std::vector<Leaf>> leaves(levels);
	for (u32 g = 0; g < G; ++g)
		for (u32 s = 0; s < groups[g].segmentCount; ++s)
			leaves[u32(groups[g].depth)].push_back({groups[g].firstSegment + s, g, segments[groups[g].firstSegment + s].refinedGroup});
	for (u32 d = 0; d < levels; ++d)
		if (leaves[d].empty())
			throw Error("Cluster LOD: missing traversal leaves for an intermediate depth; compact depths or handle gaps.");

	std::vector<float> parent_error(G, 0.0f);
	std::vector<int> parent_id(G, -1);
	for (u32 g = 0; g < G; ++g)
	{
		const float e = groups[g].bounds[4];
		for (u32 s = 0; s < groups[g].segmentCount; ++s)
		{
			const ClodSegment& seg = segments[groups[g].firstSegment + s];
			if (seg.refinedGroup >= 0)
			{
				u32 child = u32(seg.refinedGroup);
				if (child >= G)
					throw Error("Cluster LOD: segment refines into a group that does not exist");
				if (e >= parent_error[child])
				{
					parent_error[child] = e;
					parent_id[child] = int(g);
				}
			}
		}
	}
	for (u32 g = 0; g < G; ++g)
	{
		if (parent_id[g] < 0)
			parent_error[g] = FLT_MAX;
		groups[g].parentGroupId = parent_id[g];
		groups[g].maxParentError = parent_error[g];
	}

	h.ranges.assign(levels, ClodNodeRange{0, 0});
	h.level_roots.resize(levels);
	for (u32 d = 0; d < levels; ++d)
		h.level_roots[d] = 1 + d;
	u32 node_offset = 1 + levels;
	for (u32 d = 0; d < levels; ++d)
	{
		u32 leaf_count = u32(leaves[d].size()), node_count = leaf_count, iter = leaf_count;
		while (iter > 1)
		{
			iter = (iter + width - 1) / width;
			node_count += iter;
		}
		node_count--;
		h.ranges[d].offset = node_offset;
		h.ranges[d].count = node_count;
		node_offset += node_count;
	}
	ClodNode zero;
	memset(&zero, 0, sizeof(zero));
	h.nodes.assign(node_offset, zero);

	for (u32 d = 0; d < levels; ++d)
	{
		const std::vector<Leaf>& lv = leaves[d];
		const u32 leaf_count = u32(lv.size());
		u32 write_offset = h.ranges[d].offset, last_layer = write_offset;
		for (u32 i = 0; i < leaf_count; ++i)
		{
			const Leaf& info = lv[i];
			const ClodGroup& grp = groups[info.owner];
			ClodNode& node = (leaf_count == 1) ? h.nodes[1 + d] : h.nodes[write_offset++];
			node = zero;
			node.isGroup = 2;
			node.indexOrOffset = info.segment;
			node.countMinusOne = info.refined >= 0 ? u32(info.refined + 1) : 0u;
			node.ownerGroupId = info.owner;
			const float* sb = &segment_bounds[size_t(info.segment) * 4];
			const float sx = sb[0], sy = sb[1], sz = sb[2], sr = sb[3];
			const float gx = grp.bounds[0], gy = grp.bounds[1], gz = grp.bounds[2], gr = grp.bounds[3];
			const float dx = gx - sx, dy = gy - sy, dz = gz - sz;
			const float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
			float cx, cy, cz, cr;
			if (dist + gr <= sr)
				cx = sx, cy = sy, cz = sz, cr = sr;
			else if (dist + sr <= gr)
				cx = gx, cy = gy, cz = gz, cr = gr;
			else
			{
				cr = (dist + sr + gr) * 0.5f;
				const float t = (cr - sr) / std::max(dist, 1e-12f);
				cx = sx + dx * t;
				cy = sy + dy * t;
				cz = sz + dz * t;
			}
			cr *= (1.0f + 1e-5f);
			node.cullingSphere[0] = cx, node.cullingSphere[1] = cy, node.cullingSphere[2] = cz, node.cullingSphere[3] = cr;
			node.lodBoundingSphere[0] = gx, node.lodBoundingSphere[1] = gy, node.lodBoundingSphere[2] = gz, node.lodBoundingSphere[3] = gr;
			node.maxQuadricError = parent_error[info.owner];
		}
		if (leaf_count == 1)
			write_offset++;

		u32 iter = leaf_count;
		std::vector<u32> partitioned;
		std::vector<ClodNode> scratch;
		while (iter > 1)
		{
			const u32 last_count = iter;
			ClodNode* last_nodes = &h.nodes[last_layer];
			partitioned.resize(last_count);
			spatial_cluster_points(partitioned.data(), last_nodes->cullingSphere, last_count, sizeof(ClodNode) / sizeof(float), width);
			scratch.assign(last_nodes, last_nodes + last_count);
			for (u32 n = 0; n < last_count; ++n)
				last_nodes[n] = scratch[partitioned[n]];
			iter = (last_count + width - 1) / width;
			u32 new_base = (iter == 1) ? 1 + d : write_offset;
			for (u32 n = 0; n < iter; ++n)
			{
				const u32 child_begin = n * width, child_end = std::min(child_begin + width, last_count);
				h.nodes[new_base + n] = internal_node(h.nodes, last_layer + child_begin, child_end - child_begin);
			}
			last_layer = write_offset;
			write_offset += iter;
		}
		write_offset--;
		if (h.ranges[d].offset + h.ranges[d].count != write_offset)
			throw Error("Cluster LOD: traversal node allocation mismatch (range/count).");
	}

	// top hierarchy over the per-depth roots (CLU.cpp:4913-4960)
	std::vector<u32> layer(h.level_roots);
	while (layer.size() > width)
	{
		std::vector<u32> next;
		for (u32 begin = 0; begin < layer.size(); begin += width)
		{
			const u32 child_count = std::min<u32>(width, u32(layer.size()) - begin), child_offset = layer[begin];
			for (u32 c = 1; c < child_count; ++c)
				if (layer[begin + c] != child_offset + c)
					throw Error("Cluster LOD: expected contiguous node ids while building top hierarchy");
			ClodNode parent = internal_node(h.nodes, child_offset, child_count);
			next.push_back(u32(h.nodes.size()));
			h.nodes.push_back(parent);
		}
		layer.swap(next);
	}
	for (u32 c = 1; c < layer.size(); ++c)
		if (layer[c] != layer[0] + c)
			throw Error("Cluster LOD: expected contiguous root children in top hierarchy");
	h.nodes[0] = internal_node(h.nodes, layer[0], u32(layer.size()));
	h.max_traversal_depth = traversal_depth(h.nodes, 0);

	// validation (CLU.cpp:4978-5253): the reference throws on any of these
	u32 violations = 0;
	for (u32 g = 0; g < G; ++g)
	{
		const ClodGroup& group = groups[g];
		if (size_t(group.firstSegment) + group.segmentCount > segments.size())
		{
			violations++;
			continue;
		}
		for (u32 s = 0; s < group.segmentCount; ++s)
		{
			const ClodSegment& seg = segments[group.firstSegment + s];
			if (seg.refinedGroup < 0)
				continue;
			const float pe = group.bounds[4], ce = groups[u32(seg.refinedGroup)].bounds[4];
			const bool fp = std::isfinite(pe) && pe < FLT_MAX * 0.5f, fc = std::isfinite(ce) && ce < FLT_MAX * 0.5f;
			if (fp && fc && !(pe > ce))
				violations++;
		}
	}
	std::vector<u8> reachable(h.nodes.size(), 0);
	std::vector<u32> stack(1, 0u);
	while (!stack.empty())
	{
		u32 index = stack.back();
		stack.pop_back();
		if (index >= h.nodes.size() || reachable[index])
			continue;
		reachable[index] = 1;
		const ClodNode& node = h.nodes[index];
		if (node.isGroup == 0)
		{
			const u32 child_count = node.countMinusOne + 1;
			if (child_count == 0 || child_count > width || size_t(node.indexOrOffset) + child_count > h.nodes.size())
			{
				violations++;
				continue;
			}
			float max_child = 0.f;
			for (u32 c = 0; c < child_count; ++c)
			{
				max_child = std::max(max_child, h.nodes[node.indexOrOffset + c].maxQuadricError);
				stack.push_back(node.indexOrOffset + c);
			}
			if (node.maxQuadricError + 1.0e-8f < max_child)
				violations++;
			continue;
		}
		if (node.ownerGroupId >= G || node.isGroup != 2 || node.indexOrOffset >= segments.size())
			violations++;
	}
	if (violations)
		throw Error("Cluster LOD: runtime hierarchy validation failed (" + std::to_string(violations) + " violations)");
}

} // namespace

void build_artifacts(const DeviceGeometry& geo, const BuilderSettings& settings, Workspace& ws, Artifacts& out, BuildStats& stats)
{
	out.blobs.clear();
	out.page_bytes = 0;
	memset(out.stats, 0, sizeof(out.stats));

	const u32 flags = geo.vertex_flags, stride = geo.vertex_stride;
	const bool has_normals = (flags & kVertexNormals) != 0 && stride >= 24;
	const bool has_texcoords = (flags & kVertexTexcoords) != 0 && stride >= 32;
	const u32 color_offset = 24 + ((flags & kVertexTexcoords) ? 8u : 0u);
	const bool has_colors = (flags & kVertexColors) != 0 && stride >= color_offset + 12;
	const bool recompute_normals = has_normals && !settings.preserve_imported_normals; // CLU.cpp:5352

	// builder configuration (CLU.cpp:5426-5460)
	Config config;
	config.simplify_error_merge_additive = std::max(0.0f, settings.lod_error_merge_additive);
	config.simplify_error_merge_previous = std::max(0.0f, settings.lod_error_merge_previous);
	config.partition_size = std::max<u32>(384u, std::max<u32>(1u, settings.partition_size_floor));

	// CLODB200_LEVEL_TIMES: host wall time of the outer stages on stderr (diagnostics only)
	const bool trace = getenv("CLODB200_LEVEL_TIMES") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!trace)
			return;
		dev_sync();
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "artifacts: %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	ArtifactSink sink;
	build_dag(config, geo.mesh, geo.indices, geo.index_count, ws, sink, stats);
	lap("DAG build");
	const u32 M = u32(sink.meshlets.size()), G = u32(sink.groups.size());
	if (M == 0 || G == 0)
		throw Error("clodb200: the DAG build produced no groups");
	if (sink.level_tri.size() > kMaxLevels)
		throw Error("clodb200: DAG deeper than " + std::to_string(kMaxLevels) + " levels");

	// ---- a16, first half: bucket order (first-seen `refined`, stable) fixes the meshlet order of every group (CLU.cpp:905-984)
	std::vector<u32> ordered(M); // position in group-bucket order -> emission index
	host_parallel_for(G, 64, [&](size_t g_begin, size_t g_end, size_t) {
		std::vector<int> bucket_refined;
		std::vector<u32> bucket_count, bucket_of;
		for (u32 g = u32(g_begin); g < u32(g_end); ++g)
		{
			const GroupRec& gr = sink.groups[g];
			bucket_refined.clear();
			bucket_count.clear();
			bucket_of.resize(gr.count);
			for (u32 j = 0; j < gr.count; ++j)
			{
				int r = sink.meshlets[gr.first + j].refined;
				u32 b = 0;
				while (b < bucket_refined.size() && bucket_refined[b] != r)
					b++;
				if (b == bucket_refined.size())
				{
					bucket_refined.push_back(r);
					bucket_count.push_back(0);
				}
				bucket_of[j] = b;
				bucket_count[b]++;
			}
			u32 sum = 0;
			for (u32& c : bucket_count)
			{
				u32 v = c;
				c = sum;
				sum += v;
			}
			for (u32 j = 0; j < gr.count; ++j)
				ordered[gr.first + bucket_count[bucket_of[j]]++] = gr.first + j;
		}
	});

	LevelTable levels;
	memset(&levels, 0, sizeof(levels));
	for (size_t l = 0; l < sink.level_tri.size(); ++l)
		levels.tri[l] = sink.level_tri[l];

	VertexStreams vs;
	memset(&vs, 0, sizeof(vs));
	vs.vertices = geo.vertices;
	vs.stride = stride;
	vs.normal_offset = has_normals ? 12u : 0xffffffffu;
	vs.color_offset = has_colors ? color_offset : 0xffffffffu;
	// group UV sets: the importer's sets, else the legacy interleaved UV0 (CLU.cpp:1037-1071)
	vs.uv_set_count = geo.uv_set_count;
	for (u32 s = 0; s < geo.uv_set_count; ++s)
	{
		vs.uv_values[s] = geo.uv_values[s];
		vs.uv_stride[s] = geo.uv_stride[s];
	}
	if (vs.uv_set_count == 0 && has_texcoords)
	{
		vs.uv_set_count = 1;
		vs.uv_values[0] = reinterpret_cast<const float*>(geo.vertices + 24);
		vs.uv_stride[0] = stride / 4;
	}
	const u32 U = vs.uv_set_count;
	// hasSkinningStream, CLU.cpp:1092-1097
	const bool has_skinning = geo.skinning_vertices != nullptr && geo.skinning_stride >= kSkinJointOffset + kSkinInfluenceBytes && geo.skinning_vertex_count > 0;
	if (has_skinning)
	{
		vs.skinning = geo.skinning_vertices;
		vs.skinning_stride = geo.skinning_stride;
		vs.skinning_count = u32(std::min<size_t>(geo.skinning_vertex_count, 0xffffffffu));
	}
	const u32 mask = (has_normals ? kAttrNormal : 0u) | (has_colors ? kAttrColor : 0u) | (has_skinning ? (kAttrJoints | kAttrWeights) : 0u);

	Arena& temp = ws.temp;
	ArenaScope scope(temp);

	// jobs in group-bucket order; placement fields are filled after the packing below
	// not value-initialised: every element is written (memset included) by the parallel fill below
	// ... and in the workspace's pinned staging buffer (idle once the DAG is built: the level downloads have been consumed), because
	// the array is uploaded twice (128 MB on C3)
	dev_d2h_async_wait();
	ws.stage.reserve(size_t(M) * sizeof(MeshletJob) + 256);
	MeshletJob* const jobs = reinterpret_cast<MeshletJob*>(ws.stage.base);
	std::unique_ptr<u32[]> group_of_storage(new u32[M]);
	u32* const group_of = group_of_storage.get();
	size_t vertex_refs = 0, triangle_total = 0;
	host_parallel_for(G, 64, [&](size_t g_begin, size_t g_end, size_t) {
		for (u32 g = u32(g_begin); g < u32(g_end); ++g)
			for (u32 j = 0; j < sink.groups[g].count; ++j)
				group_of[sink.groups[g].first + j] = g;
	});
	std::vector<size_t> part_vertex_refs(host_parallel_chunks(M, 16384), 0), part_triangles(host_parallel_chunks(M, 16384), 0);
	host_parallel_for(M, 16384, [&](size_t m_begin, size_t m_end, size_t chunk) {
	size_t vertex_refs = 0, triangle_total = 0;
	for (u32 m = u32(m_begin); m < u32(m_end); ++m)
	{
		const MeshletRec& r = sink.meshlets[ordered[m]];
		MeshletJob& job = jobs[m];
		memset(&job, 0, sizeof(job));
		job.level = r.level;
		job.tri_begin = r.tri_begin;
		job.tri_count = r.tri_count;
		job.vertex_count = r.vertex_count;
		job.group = group_of[m];
		const u32 encoded = r.refined >= 0 ? u32(r.refined + 1) : 0u;
		job.tri_word = (r.tri_count & 0xFFFFu) | (encoded << 16);
		memcpy(job.bounds, r.bounds, sizeof(job.bounds));
		vertex_refs += r.vertex_count;
		triangle_total += r.tri_count;
	}
	part_vertex_refs[chunk] = vertex_refs;
	part_triangles[chunk] = triangle_total;
	});
	for (size_t c = 0; c < part_vertex_refs.size(); ++c)
	{
		vertex_refs += part_vertex_refs[c];
		triangle_total += part_triangles[c];
	}

	// skinned meshes: room for every meshlet's bone list (at most 8 joints per vertex) behind a prefix of the vertex counts
	u32* d_bone_scratch = nullptr;
	u32* d_bone_counts = nullptr;
	if (has_skinning)
	{
		u64 cursor = 0;
		for (u32 m = 0; m < M; ++m)
		{
			jobs[m].bone_scratch = u32(cursor);
			cursor += u64(jobs[m].vertex_count) * kMaxSkinInfluences;
		}
		if (cursor > 0xffffffffull)
			throw Error("clodb200: skinned mesh too large for the bone-list scratch (more than 2^32 influence slots)");
		d_bone_scratch = temp.alloc<u32>(size_t(cursor) + 1);
		d_bone_counts = temp.alloc<u32>(M);
		dev_memset(d_bone_counts, 0, size_t(M) * 4);
	}
	lap("bucket order + jobs (host)");
	// ---- device pre-pass: distinct vertices per group, UV ranges per meshlet
	MeshletJob* d_jobs = temp.alloc<MeshletJob>(M);
	dev_h2d(d_jobs, jobs, size_t(M) * sizeof(MeshletJob));
	u64 table_size = 1;
	while (table_size < u64(vertex_refs) * 2 + 16)
		table_size <<= 1;
	u64* d_table = temp.alloc<u64>(table_size);
	u32* d_group_vertices = temp.alloc<u32>(G);
	float* d_uv_ranges = temp.alloc<float>(size_t(M) * std::max(U, 1u) * 4);
	u32* d_errors = temp.alloc<u32>(4);
	dev_memset(d_table, 0xff, table_size * 8);
	dev_memset(d_group_vertices, 0, size_t(G) * 4);
	dev_memset(d_errors, 0, 16);
#ifdef CLODB_EMU
	LAUNCH(k_meshlet_prepass, M, d_jobs, M, levels, vs, d_table, table_size - 1, d_group_vertices, d_uv_ranges, d_errors, d_bone_scratch, d_bone_counts);
#else
	LAUNCH_GRID(k_meshlet_prepass_warp, (M + MW_WARPS - 1) / MW_WARPS, MW_WARPS * 32, d_jobs, M, levels, vs, d_table, table_size - 1, d_group_vertices, d_uv_ranges, d_errors, d_bone_scratch, d_bone_counts);
#endif
	if (dev_read(d_errors))
		throw Error("clodb200: meshlet vertex table does not match the cluster's vertex count");
	std::vector<u32> group_vertices = dev_download(d_group_vertices, G);
	std::vector<float> uv_ranges;
	if (U)
		uv_ranges = dev_download(d_uv_ranges, size_t(M) * U * 4);
	if (has_skinning)
	{
		std::vector<u32> bone_counts = dev_download(d_bone_counts, M);
		for (u32 m = 0; m < M; ++m)
			jobs[m].bone_count = bone_counts[m];
	}

	lap("meshlet pre-pass + read-back");
	// per-(meshlet, set) UV compression parameters (CLU.cpp:1268-1306)
	std::vector<UvJob> uv_jobs(size_t(M) * U);
	std::vector<u32> uv_bits_total(size_t(M) * U); // totalUvBits / vertex = bitsU + bitsV
	host_parallel_for(size_t(M) * U, 32768, [&](size_t i_begin, size_t i_end, size_t) {
	for (size_t i = i_begin; i < i_end; ++i)
	{
		const float* r = &uv_ranges[i * 4];
		float min_u = r[0], min_v = r[1], max_u = r[2], max_v = r[3];
		if (jobs[i / U].vertex_count == 0)
			min_u = min_v = max_u = max_v = 0.0f;
		const float range_u = std::max(0.0f, max_u - min_u), range_v = std::max(0.0f, max_v - min_v);
		const u32 bits_u = bits_needed_for_range(quantize_uv_offset(range_u)), bits_v = bits_needed_for_range(quantize_uv_offset(range_v));
		uv_jobs[i].bit_cursor = 0;
		uv_jobs[i].min_u = min_u;
		uv_jobs[i].min_v = min_v;
		uv_jobs[i].bits = (bits_u & 0xFFu) | ((bits_v & 0xFFu) << 8);
		uv_bits_total[i] = bits_u + bits_v;
	}
	});
	auto add_meshlet = [&](PageTotals& t, u32 m) {
		t.meshlets++;
		t.position_bytes += jobs[m].vertex_count * 12;
		t.vertex_count += jobs[m].vertex_count;
		t.triangle_bytes += jobs[m].tri_count * 3;
		t.bone_indices += jobs[m].bone_count;
		for (u32 s = 0; s < U; ++s)
			t.uv_bits[s] += u64(jobs[m].vertex_count) * uv_bits_total[size_t(m) * U + s];
	};

	// ---- a16, second half: greedy group pages, segments, segment spheres, group records (CLU.cpp:1348-1495)
	// Phase 1 (host threads, one group at a time each): everything that only depends on the group's own meshlets. Phase 2 (serial):
	// the running totals (firstMeshlet, firstGroupVertex, firstSegment) and the concatenation in group order.
	struct GroupLayout
	{
		std::vector<ClodSegment> segs; // final order: terminal segments first (stable)
		std::vector<u32> seg_first;
		std::vector<float> seg_bounds;
		u32 page_count = 0, terminal_segments = 0, tri_bytes = 0;
	};
	std::vector<GroupLayout> layouts(G);
	host_parallel_for(G, 32, [&](size_t g_begin, size_t g_end, size_t) {
		std::vector<u32> page_first;
		std::vector<ClodSegment> gsegs;
		std::vector<u32> gseg_first;
		std::vector<u32> order;
		for (u32 g = u32(g_begin); g < u32(g_end); ++g)
		{
			const GroupRec& gr = sink.groups[g];
			GroupLayout& lay = layouts[g];
			page_first.assign(1, gr.first);
			PageTotals cur;
			for (u32 m = gr.first; m < gr.first + gr.count; ++m)
			{
				PageTotals cand = cur;
				add_meshlet(cand, m);
				if (page_blob_size(mask, U, cand) > kPageSize && cur.meshlets > 0)
				{
					page_first.push_back(m);
					cur = PageTotals();
					add_meshlet(cur, m);
					continue;
				}
				cur = cand;
			}
			page_first.push_back(gr.first + gr.count);
			gsegs.clear();
			gseg_first.clear();
			for (u32 pi = 0; pi + 1 < page_first.size(); ++pi)
			{
				u32 b = page_first[pi], e = page_first[pi + 1], run = b;
				while (run < e)
				{
					const int tag = sink.meshlets[ordered[run]].refined;
					u32 run_end = run + 1;
					while (run_end < e && sink.meshlets[ordered[run_end]].refined == tag)
						run_end++;
					gsegs.push_back(ClodSegment{tag, run - b, run_end - run, pi});
					gseg_first.push_back(run);
					run = run_end;
				}
			}
			// terminal segments first, stable (CLU.cpp:1446-1451)
			order.resize(gsegs.size());
			for (u32 i = 0; i < order.size(); ++i)
				order[i] = i;
			std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return (gsegs[a].refinedGroup < 0) > (gsegs[b].refinedGroup < 0); });
			lay.page_count = u32(page_first.size() - 1);
			for (u32 m = gr.first; m < gr.first + gr.count; ++m)
				lay.tri_bytes += jobs[m].tri_count * 3;
			bool leading = true;
			lay.segs.reserve(order.size());
			lay.seg_first.reserve(order.size());
			lay.seg_bounds.reserve(order.size() * 4);
			for (u32 i : order)
			{
				const ClodSegment& seg = gsegs[i];
				if (seg.refinedGroup < 0 && leading)
					lay.terminal_segments++;
				else
					leading = false;
				lay.segs.push_back(seg);
				lay.seg_first.push_back(gseg_first[i]);
				float sphere[4];
				sphere_bounds(sphere, jobs[gseg_first[i]].bounds, seg.meshletCount, sizeof(MeshletJob) / 4, jobs[gseg_first[i]].bounds + 3, sizeof(MeshletJob) / 4);
				lay.seg_bounds.insert(lay.seg_bounds.end(), sphere, sphere + 4);
			}
		}
	});
	std::vector<ClodGroup> groups(G);
	std::vector<ClodChunk> chunks(G);
	std::vector<ClodSegment> segments;
	std::vector<u32> segment_first; // first meshlet (group-bucket order index) of each segment
	std::vector<float> segment_bounds;
	{
		size_t total_segments = 0;
		for (u32 g = 0; g < G; ++g)
			total_segments += layouts[g].segs.size();
		segments.reserve(total_segments);
		segment_first.reserve(total_segments);
		segment_bounds.reserve(total_segments * 4);
		u32 cumulative_meshlets = 0, cumulative_vertices = 0;
		for (u32 g = 0; g < G; ++g)
		{
			const GroupRec& gr = sink.groups[g];
			const GroupLayout& lay = layouts[g];
			ClodGroup& out_group = groups[g];
			memset(&out_group, 0, sizeof(out_group));
			memcpy(out_group.bounds, gr.simplified, sizeof(out_group.bounds));
			out_group.depth = gr.depth;
			out_group.firstMeshlet = cumulative_meshlets;
			out_group.meshletCount = gr.count;
			out_group.firstGroupVertex = cumulative_vertices;
			out_group.groupVertexCount = group_vertices[g];
			out_group.firstSegment = u32(segments.size());
			out_group.segmentCount = u32(lay.segs.size());
			out_group.terminalSegmentCount = lay.terminal_segments;
			out_group.pageCount = lay.page_count;
			out_group.parentGroupId = -1;
			cumulative_meshlets += gr.count;
			cumulative_vertices += group_vertices[g];
			chunks[g] = ClodChunk{group_vertices[g], gr.count, lay.tri_bytes, 1u, has_normals ? 4u : 0u}; // CLOD_COMPRESSED_NORMALS = 1 << 2
			segments.insert(segments.end(), lay.segs.begin(), lay.segs.end());
			segment_first.insert(segment_first.end(), lay.seg_first.begin(), lay.seg_first.end());
			segment_bounds.insert(segment_bounds.end(), lay.seg_bounds.begin(), lay.seg_bounds.end());
		}
	}
	layouts.clear();
	lap("group pages + segments (host)");

	// ---- a17
	Hierarchy hier;
	build_hierarchy(groups, segments, segment_bounds, hier);

	lap("hierarchy (host)");
	// ---- a18: mesh-wide greedy packing of segments, groups visited by (depth, parentGroupId, index) (CLU.cpp:2374-2470)
	// The greedy fill is sequential in the reference and stays so here, but over per-segment totals (summed on host threads
	// first; integer sums, so the order does not matter); the per-meshlet placement inside the finished pages runs on host
	// threads afterwards, one page at a time each.
	const u32 S = u32(segments.size());
	std::vector<PageTotals> segment_totals(S);
	host_parallel_for(S, 2048, [&](size_t s_begin, size_t s_end, size_t) {
		for (u32 si = u32(s_begin); si < u32(s_end); ++si)
		{
			PageTotals t;
			for (u32 k = 0; k < segments[si].meshletCount; ++k)
				add_meshlet(t, segment_first[si] + k);
			segment_totals[si] = t;
		}
	});
	auto add_totals = [&](PageTotals& t, const PageTotals& o) {
		t.meshlets += o.meshlets;
		t.position_bytes += o.position_bytes;
		t.vertex_count += o.vertex_count;
		t.triangle_bytes += o.triangle_bytes;
		t.bone_indices += o.bone_indices;
		for (u32 s = 0; s < U; ++s)
			t.uv_bits[s] += o.uv_bits[s];
	};
	std::vector<PageRecord> pages;
	std::vector<u64> page_offsets(1, 0);
	std::vector<std::vector<u32>> group_pages(G);
	std::vector<u32> page_segments;           // segment indices, page by page
	std::vector<u32> page_segment_offsets(1, 0);
	{
		std::vector<u32> group_order(G);
		for (u32 g = 0; g < G; ++g)
			group_order[g] = g;
		std::stable_sort(group_order.begin(), group_order.end(), [&](u32 a, u32 b) {
			if (groups[a].depth != groups[b].depth)
				return groups[a].depth < groups[b].depth;
			if (groups[a].parentGroupId != groups[b].parentGroupId)
				return groups[a].parentGroupId < groups[b].parentGroupId;
			return a < b;
		});
		PageTotals cur;
		size_t current_begin = 0; // page_segments[current_begin..) is the page being filled
		auto flush = [&]() {
			if (page_segments.size() == current_begin)
				return;
			PageRecord pr;
			memset(&pr, 0, sizeof(pr));
			u32 size = 0;
			page_layout(mask, U, cur, pr, size);
			if (size > kPageSize)
				throw Error("clodb200: a segment does not fit a 256 KiB page");
			pr.base = page_offsets.back();
			const u32 page_index = u32(pages.size());
			for (size_t q = current_begin; q < page_segments.size(); ++q)
				group_pages[jobs[segment_first[page_segments[q]]].group].push_back(page_index);
			pages.push_back(pr);
			page_offsets.push_back(pr.base + size);
			page_segment_offsets.push_back(u32(page_segments.size()));
			current_begin = page_segments.size();
			cur = PageTotals();
		};
		for (u32 g : group_order)
		{
			const ClodGroup& group = groups[g];
			for (u32 si = group.firstSegment; si < group.firstSegment + group.segmentCount; ++si)
			{
				const ClodSegment& seg = segments[si];
				if (seg.meshletCount == 0)
					continue;
				PageTotals cand = cur;
				add_totals(cand, segment_totals[si]);
				if (page_blob_size(mask, U, cand) > kPageSize && page_segments.size() != current_begin)
				{
					flush();
					cand = segment_totals[si];
				}
				page_segments.push_back(si);
				cur = cand;
			}
		}
		flush();
	}
	host_parallel_for(pages.size(), 8, [&](size_t p_begin, size_t p_end, size_t) {
		for (u32 page_index = u32(p_begin); page_index < u32(p_end); ++page_index)
		{
			u32 slot = 0, pos_cursor = 0, attr_cursor = 0, tri_cursor = 0, bone_cursor = 0;
			u64 uv_cursor[kMaxUvSets] = {};
			for (u32 q = page_segment_offsets[page_index]; q < page_segment_offsets[page_index + 1]; ++q)
			{
				const u32 si = page_segments[q];
				ClodSegment& seg = segments[si];
				seg.pageIndex = page_index;
				seg.firstMeshletInPage = slot;
				for (u32 k = 0; k < seg.meshletCount; ++k)
				{
					u32 m = segment_first[si] + k;
					MeshletJob& job = jobs[m];
					job.page = page_index;
					job.slot = slot++;
					job.pos_cursor = pos_cursor;
					job.attr_cursor = attr_cursor;
					job.tri_cursor = tri_cursor;
					job.bone_cursor = bone_cursor;
					bone_cursor += job.bone_count;
					pos_cursor += job.vertex_count * 12;
					attr_cursor += job.vertex_count;
					tri_cursor += job.tri_count * 3;
					for (u32 s = 0; s < U; ++s)
					{
						uv_jobs[size_t(m) * U + s].bit_cursor = u32(uv_cursor[s]);
						uv_cursor[s] += u64(job.vertex_count) * uv_bits_total[size_t(m) * U + s];
					}
				}
			}
		}
	});
	const u32 page_count = u32(pages.size());
	std::vector<u32> page_refs, page_ref_offsets;
	for (u32 g = 0; g < G; ++g)
	{
		page_ref_offsets.push_back(u32(page_refs.size()));
		std::vector<u32>& refs = group_pages[g];
		std::sort(refs.begin(), refs.end());
		refs.erase(std::unique(refs.begin(), refs.end()), refs.end());
		if (refs.empty())
		{
			groups[g].pageMapBase = 0;
			groups[g].pageCount = 0;
			continue;
		}
		groups[g].pageMapBase = refs.front();
		groups[g].pageCount = refs.back() - refs.front() + 1;
		page_refs.insert(page_refs.end(), refs.begin(), refs.end());
	}
	page_ref_offsets.push_back(u32(page_refs.size()));

	lap("mesh page packing (host)");
	// ---- page bytes on the device, one read-back
	const size_t total_bytes = size_t(page_offsets.back());
	u8* d_out = temp.alloc<u8>(total_bytes);
	PageRecord* d_pages = temp.alloc<PageRecord>(page_count);
	UvJob* d_uv_jobs = temp.alloc<UvJob>(size_t(M) * std::max(U, 1u));
	dev_memset(d_out, 0, total_bytes);
	dev_h2d(d_jobs, jobs, size_t(M) * sizeof(MeshletJob));
	dev_h2d(d_pages, pages.data(), size_t(page_count) * sizeof(PageRecord));
	if (U)
		dev_h2d(d_uv_jobs, uv_jobs.data(), uv_jobs.size() * sizeof(UvJob));
	NormalRecompute recompute = {nullptr, 0, nullptr};
	if (recompute_normals)
	{
		std::vector<GroupSpan> spans(G);
		for (u32 g = 0; g < G; ++g)
			spans[g] = GroupSpan{sink.groups[g].first, sink.groups[g].count};
		GroupSpan* d_spans = temp.alloc<GroupSpan>(G);
		float* d_sums = temp.alloc<float>(size_t(table_size) * 3);
		dev_h2d(d_spans, spans.data(), size_t(G) * sizeof(GroupSpan));
		dev_memset(d_sums, 0, size_t(table_size) * 12);
		LAUNCH(k_group_normal_sums, G, d_spans, G, d_jobs, levels, vs, d_table, table_size - 1, d_sums);
		recompute = NormalRecompute{d_table, table_size - 1, d_sums};
	}
#ifdef CLODB_EMU
	LAUNCH(k_write_pages, M, d_jobs, M, levels, vs, d_pages, d_uv_jobs, d_out, d_errors, d_bone_scratch, recompute);
#else
	LAUNCH_GRID(k_write_pages_warp, (M + MW_WARPS - 1) / MW_WARPS, MW_WARPS * 32, d_jobs, M, levels, vs, d_pages, d_uv_jobs, d_out, d_errors, d_bone_scratch, recompute);
#endif
	lap("page writer kernel");
	out.pages.reserve(total_bytes + 16);
	dev_d2h(out.pages.base, d_out, total_bytes);
	lap("page read-back");
	out.page_bytes = total_bytes;
	if (dev_read(d_errors))
		throw Error("clodb200: page writer found an inconsistent meshlet");

	put_blob(out, "groups", groups);
	put_blob(out, "segments", segments);
	put_blob(out, "segmentBounds", segment_bounds);
	put_blob(out, "groupChunks", chunks);
	put_blob(out, "groupPageReferences", page_refs);
	put_blob(out, "groupPageReferenceOffsets", page_ref_offsets);
	put_blob(out, "nodes", hier.nodes);
	put_blob(out, "lodNodeRanges", hier.ranges);
	put_blob(out, "lodLevelRoots", hier.level_roots);
	std::vector<float> object_sphere(hier.nodes[0].cullingSphere, hier.nodes[0].cullingSphere + 4); // CLU.cpp:1807-1817
	put_blob(out, "objectBoundingSphere", object_sphere);
	std::vector<u32> counts = {page_count, page_count, 0u, hier.max_depth, hier.max_traversal_depth};
	put_blob(out, "counts", counts);
	put_blob(out, "meshPageOffsets", page_offsets);

	out.stats[0] = M;
	out.stats[1] = G;
	out.stats[2] = segments.size();
	out.stats[3] = page_count;
	out.stats[4] = total_bytes;
	out.stats[5] = vertex_refs;
	out.stats[6] = triangle_total;
	u64 group_vertex_total = 0;
	for (u32 v : group_vertices)
		group_vertex_total += v;
	out.stats[7] = group_vertex_total;
	out.stats[8] = stats.levels;
	out.stats[9] = stats.simplified_triangles;
	out.stats[10] = stats.d2h_bytes + total_bytes + size_t(G) * 4 + uv_ranges.size() * 4;
	out.stats[11] = hier.nodes.size();
}

// ---- cache files ------------------------------------------------------------------------------------------------------------
namespace
{
template <typename T>
void write_pod(std::vector<u8>& out, const T& v)
{
	const u8* p = reinterpret_cast<const u8*>(&v);
	out.insert(out.end(), p, p + sizeof(T));
}
void write_vector(std::vector<u8>& out, const std::vector<u8>& bytes, size_t element_size)
{
	write_pod(out, u64(bytes.size() / element_size));
	out.insert(out.end(), bytes.begin(), bytes.end());
}
void write_string(std::vector<u8>& out, const std::string& s)
{
	write_pod(out, u64(s.size()));
	out.insert(out.end(), s.begin(), s.end());
}
const std::vector<u8>& blob(const Artifacts& a, const char* name)
{
	static const std::vector<u8> empty;
	auto it = a.blobs.find(name);
	return it == a.blobs.end() ? empty : it->second;
}
std::vector<ClodDiskLocator> page_locators(const Artifacts& a)
{
	const std::vector<u8>& ob = blob(a, "meshPageOffsets");
	const u64* offsets = reinterpret_cast<const u64*>(ob.data());
	size_t pages = ob.size() / 8 ? ob.size() / 8 - 1 : 0;
	const u64 first_blob = 16 + u64(pages) * sizeof(ClodDiskLocator); // ContainerHeader + directory
	std::vector<ClodDiskLocator> loc(pages);
	for (size_t p = 0; p < pages; ++p)
		loc[p] = ClodDiskLocator{first_blob + offsets[p], u32(offsets[p + 1] - offsets[p]), 0u};
	return loc;
}
} // namespace

std::vector<u8> serialize_cache_metadata(const Artifacts& a, const CacheIdentity& id, const std::string& container_file_name)
{
	std::vector<u8> out;
	write_pod(out, u32(47)); // kSchemaVersion, CLodCache.h:15
	write_pod(out, id.build_config_hash);
	write_vector(out, blob(a, "groups"), sizeof(ClodGroup));
	write_vector(out, blob(a, "segments"), sizeof(ClodSegment));
	write_vector(out, blob(a, "segmentBounds"), 16);
	const std::vector<u8>& sphere = blob(a, "objectBoundingSphere");
	out.insert(out.end(), sphere.begin(), sphere.end());
	const std::vector<u8>& chunks = blob(a, "groupChunks");
	write_pod(out, u8(chunks.empty() ? 0 : 1));
	if (!chunks.empty())
		write_vector(out, chunks, sizeof(ClodChunk));
	write_vector(out, std::vector<u8>(), sizeof(ClodDiskLocator)); // groupDiskLocators: empty for container caches
	std::vector<ClodDiskLocator> loc = page_locators(a);
	write_pod(out, u64(loc.size()));
	out.insert(out.end(), reinterpret_cast<const u8*>(loc.data()), reinterpret_cast<const u8*>(loc.data() + loc.size()));
	write_vector(out, blob(a, "groupPageReferences"), 4);
	write_vector(out, blob(a, "groupPageReferenceOffsets"), 4);
	const u32* counts = reinterpret_cast<const u32*>(blob(a, "counts").data());
	write_pod(out, counts[0]);
	write_pod(out, counts[1]);
	write_pod(out, counts[2]);
	write_string(out, id.source_identifier);
	write_string(out, id.prim_path);
	write_string(out, id.subset_name);
	write_pod(out, id.build_config_hash);
	write_string(out, container_file_name);
	write_vector(out, blob(a, "nodes"), sizeof(ClodNode));
	write_vector(out, blob(a, "lodNodeRanges"), sizeof(ClodNodeRange));
	write_vector(out, blob(a, "lodLevelRoots"), 4);
	write_pod(out, counts[3]);
	write_pod(out, counts[4]);
	return out;
}

void write_cache_container(const Artifacts& a, const std::string& path)
{
	std::vector<ClodDiskLocator> loc = page_locators(a);
	std::ofstream file(path, std::ios::binary | std::ios::trunc);
	if (!file.is_open())
		throw Error("clodb200: cannot open " + path);
	u32 header[4] = {0x444F4C43u, 4u, 0u, u32(loc.size())}; // ContainerHeader, CLodCache.cpp:252-259
	file.write(reinterpret_cast<const char*>(header), sizeof(header));
	file.write(reinterpret_cast<const char*>(loc.data()), std::streamsize(loc.size() * sizeof(ClodDiskLocator)));
	file.write(a.pages.base, std::streamsize(a.page_bytes));
	if (!file.good())
		throw Error("clodb200: failed writing " + path);
}

} // namespace clodb
