// S8 (first part): meshlet-local index extraction.
// Reference: clodLocalIndices (BasicRenderer/include/ThirdParty/meshoptimizer/clusterlod.h:972-1023): vertices[] = the
// distinct global indices of a cluster in first-occurrence order, triangles[i] = position of indices[i] in vertices[]
// (the reference's 1024-entry cache is only an accelerator and does not change the result).
#include "clodb.h"

namespace clodb
{

// one thread per cluster; clusters hold <= 128 triangles / <= 256 distinct vertices (u8 local ids)
KERNEL k_local_indices(const u32* __restrict__ indices, const u64* __restrict__ cluster_index_offset, u32 K, u32 vertex_capacity, u32* vertices, u8* triangles, u32* vertex_count)
{
	size_t c = GTID;
	if (c >= K)
		return;
	u64 begin = cluster_index_offset[c], end = cluster_index_offset[c + 1];
	u32 keys[512];
	u8 vals[512];
	for (int i = 0; i < 512; ++i)
		keys[i] = 0xffffffffu;
	u32 count = 0;
	u32* vout = vertices + size_t(c) * vertex_capacity;
	for (u64 j = begin; j < end; ++j)
	{
		u32 v = indices[j];
		u32 h = (v * 0x9E3779B1u) >> 23;
		for (;;)
		{
			if (keys[h] == v)
				break;
			if (keys[h] == 0xffffffffu)
			{
				keys[h] = v;
				vals[h] = u8(count);
				if (count < vertex_capacity)
					vout[count] = v;
				count++;
				break;
			}
			h = (h + 1) & 511;
		}
		triangles[j] = vals[h];
	}
	vertex_count[c] = count;
}

void local_indices(const u32* indices, const u64* cluster_index_offset, u32 cluster_count, u32 vertex_capacity, u32* vertices, u8* triangles, u32* vertex_count)
{
	LAUNCH(k_local_indices, cluster_count, indices, cluster_index_offset, cluster_count, vertex_capacity, vertices, triangles, vertex_count);
}

} // namespace clodb
