// S1: vertex dedup by position + protect bits.
//
// Reference semantics (bit-exact target):
//   meshopt_generatePositionRemap, ThirdParty/meshoptimizer/src/indexgenerator.cpp:442-465 with the position
//   hasher/equality at :90-128 : remap[i] = first (lowest) index whose x,y,z compare equal with IEEE == (so -0 == +0 and
//   a vertex with a NaN component only ever maps to itself). The reference gets "lowest" from in-order insertion into an
//   open-addressed hash set; here every vertex is inserted concurrently into a device hash table whose slots carry an
//   atomicMin of the member indices, which yields the same function of the input regardless of insertion order.
//   Protect pass: clusterlod.h:829-841.
#include "clodb.h"

namespace clodb
{

static const u32 EMPTY = 0xffffffffu;

DEVFN u32 hash_position(float x, float y, float z)
{
	// -0 and +0 must hash alike: x + 0.0f canonicalises the sign of zero
	u32 a = __float_as_uint(x + 0.0f), b = __float_as_uint(y + 0.0f), c = __float_as_uint(z + 0.0f);
	u32 h = a * 0x9E3779B1u;
	h = (h ^ (h >> 15)) + b * 0x85EBCA77u;
	h = (h ^ (h >> 13)) + c * 0xC2B2AE3Du;
	h ^= h >> 16;
	h *= 0x7FEB352Du;
	h ^= h >> 15;
	return h;
}

KERNEL k_remap_insert(const float* __restrict__ positions, size_t vertex_count, u32* table_rep, u32* table_min, u32 mask, u32* slot_out)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	float x = positions[i * 3 + 0], y = positions[i * 3 + 1], z = positions[i * 3 + 2];
	if (x != x || y != y || z != z)
	{
		slot_out[i] = EMPTY; // NaN never equals anything, itself included
		return;
	}
	u32 h = hash_position(x, y, z) & mask;
	for (;;)
	{
		u32 rep = atomicCAS(&table_rep[h], EMPTY, u32(i));
		if (rep == EMPTY || rep == u32(i))
			break;
		if (positions[size_t(rep) * 3 + 0] == x && positions[size_t(rep) * 3 + 1] == y && positions[size_t(rep) * 3 + 2] == z)
			break;
		h = (h + 1) & mask;
	}
	atomicMin(&table_min[h], u32(i));
	slot_out[i] = h;
}

KERNEL k_remap_resolve(u32* remap, const u32* __restrict__ table_min, size_t vertex_count)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	u32 slot = remap[i];
	remap[i] = slot == EMPTY ? u32(i) : table_min[slot];
}

void position_remap(const float* positions, size_t vertex_count, u32* remap, Arena& temp)
{
	if (vertex_count == 0)
		return;
	ArenaScope scope(temp);
	size_t table_size = 1;
	while (table_size < vertex_count * 2)
		table_size <<= 1;
	u32* table_rep = temp.alloc<u32>(table_size);
	u32* table_min = temp.alloc<u32>(table_size);
	dev_memset(table_rep, 0xff, table_size * sizeof(u32));
	dev_memset(table_min, 0xff, table_size * sizeof(u32));
	LAUNCH(k_remap_insert, vertex_count, positions, vertex_count, table_rep, table_min, u32(table_size - 1), remap);
	LAUNCH(k_remap_resolve, vertex_count, remap, table_min, vertex_count);
}

KERNEL k_protect_bits(const float* __restrict__ attributes, u32 attribute_stride, u32 protect_mask, const u32* __restrict__ remap, size_t vertex_count, u8* locks)
{
	size_t i = GTID;
	if (i >= vertex_count)
		return;
	u32 r = remap[i];
	if (r == u32(i))
		return;
	bool differs = false;
	for (u32 j = 0; j < attribute_stride; ++j)
		if ((protect_mask >> j) & 1u)
			differs |= attributes[i * attribute_stride + j] != attributes[size_t(r) * attribute_stride + j];
	if (differs)
		locks[i] |= 2; // meshopt_SimplifyVertex_Protect
}

void protect_bits(const float* attributes, u32 attribute_stride, u32 protect_mask, const u32* remap, size_t vertex_count, u8* locks)
{
	if (!protect_mask || !attributes)
		return;
	LAUNCH(k_protect_bits, vertex_count, attributes, attribute_stride, protect_mask, remap, vertex_count, locks);
}

} // namespace clodb
