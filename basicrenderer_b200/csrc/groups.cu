// S5: group assembly (merged index lists) and boundary locks.
// Reference: runIterationTask's merge loop (clusterlod.h:708-711) and clod::lockBoundary (clusterlod.h:512-559).
#include "clodb.h"

namespace clodb
{

KERNEL k_group_cluster_counts(const u32* __restrict__ cluster_tri_offset, const u32* __restrict__ group_clusters, u32* counts, u32 K)
{
	size_t j = GTID;
	if (j >= K)
		return;
	u32 c = group_clusters[j];
	counts[j] = cluster_tri_offset[c + 1] - cluster_tri_offset[c];
}

KERNEL k_gather_triangles(const u32* __restrict__ tri, const u32* __restrict__ cluster_tri_offset, const u32* __restrict__ group_clusters, const u32* __restrict__ gc_tri_offset, u32 K, u32* out, u32 T)
{
	size_t t = GTID;
	if (t >= T)
		return;
	u32 lo = 0, hi = K; // slot j with gc_tri_offset[j] <= t < gc_tri_offset[j + 1]
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (gc_tri_offset[mid] <= u32(t))
			lo = mid;
		else
			hi = mid;
	}
	u32 src = cluster_tri_offset[group_clusters[lo]] + (u32(t) - gc_tri_offset[lo]);
	out[t * 3 + 0] = tri[size_t(src) * 3 + 0];
	out[t * 3 + 1] = tri[size_t(src) * 3 + 1];
	out[t * 3 + 2] = tri[size_t(src) * 3 + 2];
}

void gather_group_triangles(const u32* tri, const u32* cluster_tri_offset, const u32* group_clusters, u32 K, u32* gtri_out, u32* gc_tri_offset, Arena& temp)
{
	if (K == 0)
		return;
	ArenaScope scope(temp);
	u32* total = temp.alloc<u32>(1);
	LAUNCH(k_group_cluster_counts, K, cluster_tri_offset, group_clusters, gc_tri_offset, K);
	// K + 1 entries: the scan also produces the total in the last slot
	dev_memset(gc_tri_offset + K, 0, sizeof(u32));
	exclusive_scan_u32(gc_tri_offset, gc_tri_offset, size_t(K) + 1, total, temp);
	u32 T = dev_read(total);
	LAUNCH(k_gather_triangles, T, tri, cluster_tri_offset, group_clusters, gc_tri_offset, K, gtri_out, T);
}

static const u32 NO_GROUP = 0xffffffffu;

KERNEL k_lock_mark(const u32* __restrict__ gtri, const u32* __restrict__ group_tri_offset, u32 G, const u32* __restrict__ remap, u32* owner, u8* shared, size_t corners)
{
	size_t c = GTID;
	if (c >= corners)
		return;
	u32 t = u32(c / 3);
	u32 lo = 0, hi = G;
	while (hi - lo > 1)
	{
		u32 mid = (lo + hi) / 2;
		if (group_tri_offset[mid] <= t)
			lo = mid;
		else
			hi = mid;
	}
	u32 r = remap[gtri[c]];
	u32 prev = atomicCAS(&owner[r], NO_GROUP, lo);
	if (prev != NO_GROUP && prev != lo)
		shared[r] = 1;
}

KERNEL k_lock_finalize(const u32* __restrict__ remap, const u8* __restrict__ shared, const u8* __restrict__ vertex_lock, u8* locks, size_t vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	u8 l = (shared[remap[i]] ? 1 : 0) | (locks[i] & 2);
	if (vertex_lock)
		l |= vertex_lock[i];
	locks[i] = l;
}

void lock_boundary(const u32* gtri, const u32* group_tri_offset, u32 G, u32 T, const u32* remap, const u8* vertex_lock, size_t vertex_count, u8* locks, Arena& temp)
{
	ArenaScope scope(temp);
	u32* owner = temp.alloc<u32>(vertex_count);
	u8* shared = temp.alloc<u8>(vertex_count);
	dev_memset(owner, 0xff, vertex_count * sizeof(u32));
	dev_memset(shared, 0, vertex_count);
	LAUNCH(k_lock_mark, size_t(T) * 3, gtri, group_tri_offset, G, remap, owner, shared, size_t(T) * 3);
	LAUNCH(k_lock_finalize, vertex_count, remap, shared, vertex_lock, locks, vertex_count);
}

} // namespace clodb
