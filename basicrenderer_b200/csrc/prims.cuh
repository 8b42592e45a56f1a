// clodb200 device-wide primitives (hand-written; no CUB/Thrust on the path):
//   exclusive_scan_u32   reduce-then-scan, 2048-element tiles, warp-shuffle block scans
//   radix_sort_pairs     stable LSD radix sort, 8-bit digits, warp match_any multi-split (keys u32/u64, values u32)
//   fill / iota helpers
// Under CLODB_EMU (development-only host emulation, see rt.cuh) each primitive has a serial stand-in with identical
// results, so stage logic built on top can be debugged without a GPU.
#pragma once

#include "rt.cuh"

namespace clodb
{

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;
typedef uint16_t u16;

// --------------------------------------------------------------------------------------------------- small kernels
template <typename T>
KERNEL k_fill(T* p, T v, size_t n)
{
	size_t i = GTID;
	if (i >= n)
		return;
	p[i] = v;
}

KERNEL k_iota(u32* p, size_t n)
{
	size_t i = GTID;
	if (i >= n)
		return;
	p[i] = u32(i);
}

template <typename T>
static inline void fill(T* p, T v, size_t n)
{
	LAUNCH(k_fill<T>, n, p, v, n);
}

static inline void iota(u32* p, size_t n)
{
	LAUNCH(k_iota, n, p, n);
}

struct OpAddU32
{
	HOSTDEVFN u32 identity()
	{
		return 0;
	}
	HOSTDEVFN u32 apply(u32 a, u32 b)
	{
		return a + b;
	}
	HOSTDEVFN bool prefix_independent(u32)
	{
		return false;
	}
};

struct OpMaxU64
{
	HOSTDEVFN u64 identity()
	{
		return 0;
	}
	HOSTDEVFN u64 apply(u64 a, u64 b)
	{
		return a > b ? a : b;
	}
	HOSTDEVFN bool prefix_independent(u64)
	{
		return false;
	}
};

#ifdef CLODB_EMU
// ------------------------------------------------------------------------------------------------------- emulation
template <typename T, typename Op>
static inline void exclusive_scan(const T* in, T* out, size_t n, T* total, Arena&)
{
	g_launches += 3;
	T sum = Op::identity();
	for (size_t i = 0; i < n; ++i)
	{
		T v = in[i];
		out[i] = sum;
		sum = Op::apply(sum, v);
	}
	if (total)
		*total = sum;
}

template <typename K>
static inline void radix_sort_pairs(K* keys, K* keys_tmp, u32* vals, u32* vals_tmp, size_t n, int bit_lo, int bit_hi, Arena&)
{
	if (n == 0 || bit_hi <= bit_lo)
		return;
	g_launches += 3 * ((bit_hi - bit_lo + 7) / 8);
	K mask = (bit_hi - bit_lo >= int(sizeof(K) * 8)) ? ~K(0) : ((K(1) << (bit_hi - bit_lo)) - 1);
	std::vector<u32> order(n);
	for (size_t i = 0; i < n; ++i)
		order[i] = u32(i);
	std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return ((keys[a] >> bit_lo) & mask) < ((keys[b] >> bit_lo) & mask); });
	for (size_t i = 0; i < n; ++i)
	{
		keys_tmp[i] = keys[order[i]];
		if (vals)
			vals_tmp[i] = vals[order[i]];
	}
	memcpy(keys, keys_tmp, n * sizeof(K));
	if (vals)
		memcpy(vals, vals_tmp, n * sizeof(u32));
}

#else
// ------------------------------------------------------------------------------------------------------------ CUDA
static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
DEVFN T shfl_up_any(T v, int d)
{
	return __shfl_up_sync(0xffffffffu, v, d);
}

template <typename T, typename Op>
DEVFN T warp_inclusive_scan(T v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		T t = shfl_up_any(v, d);
		if (lane >= d)
			v = Op::apply(t, v);
	}
	return v;
}

// exclusive scan of one value per thread across the CTA; smem must hold 34 values
template <typename T, typename Op>
DEVFN T block_exclusive_scan(T v, T* block_total, T* smem)
{
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int nwarps = blockDim.x >> 5;
	T inc = warp_inclusive_scan<T, Op>(v, lane);
	T ex = shfl_up_any(inc, 1);
	if (lane == 0)
		ex = Op::identity();
	if (lane == 31)
		smem[warp] = inc;
	__syncthreads();
	if (warp == 0)
	{
		T w = lane < nwarps ? smem[lane] : Op::identity();
		T winc = warp_inclusive_scan<T, Op>(w, lane);
		T wex = shfl_up_any(winc, 1);
		if (lane == 0)
			wex = Op::identity();
		if (lane < nwarps)
			smem[lane] = wex;
		if (lane == 31)
			smem[33] = winc;
	}
	__syncthreads();
	T result = Op::apply(smem[warp], ex);
	*block_total = smem[33];
	__syncthreads();
	return result;
}

// ---- single-pass chained scan ("decoupled look-back") ---------------------------------------------------------------
// Every tile publishes its aggregate, then its inclusive prefix, in a per-tile descriptor; a tile's exclusive prefix is
// assembled by warp 0 walking predecessor descriptors 32 at a time. One launch, each input element read once and written
// once. Descriptor flags carry the epoch of the scan call, so the descriptor arrays are never cleared between calls.
// Tiles are taken in blockIdx order (CTAs are dispatched in ascending blockIdx, so a predecessor is always resident or
// finished). The combine is applied in predecessor order, so non-commutative (segmented) operators are supported.
struct ScanChain
{
	// Packed {status, value} tile descriptors, one array per descriptor format. The formats must not share memory: a
	// stale 16-byte descriptor of an 8-byte scan read as the 8-byte descriptor of a 4-byte scan puts the high half of
	// an old *value* where the tag is expected, and values such as (group + 1) << 32 | bits do collide with small epochs.
	char* desc = nullptr;     // 4-byte scans: 8-byte {tag, value} words at a 16-byte stride
	char* desc64 = nullptr;   // 8-byte scans: 16-byte {value, tag} words
	char* box_desc = nullptr; // segmented box scans (clusterize.cu): two 16-byte {tag, xyz} words per tile
	size_t capacity_tiles = 0;
	u32 epoch = 0;
};
static const int SCAN_CHAIN_VALUE_BYTES = 32;
extern thread_local ScanChain g_scan_chain;
void scan_chain_reserve(size_t tiles);

// returns the next epoch (never 0 in the low 30 bits, so stale or zeroed descriptors never match)
static inline u32 scan_chain_next_epoch()
{
	g_scan_chain.epoch = (g_scan_chain.epoch + 1) & 0x3fffffffu;
	if (g_scan_chain.epoch == 0)
		g_scan_chain.epoch = 1;
	return g_scan_chain.epoch;
}

DEVFN u32 ld_volatile_u32(const u32* p)
{
	return *reinterpret_cast<const volatile u32*>(p);
}

template <typename T>
DEVFN T ld_cg_value(const char* base, size_t tile)
{
	// descriptor payloads are written by other CTAs: read through L2
	const T* p = reinterpret_cast<const T*>(base + tile * SCAN_CHAIN_VALUE_BYTES);
	T v;
	u32* dst = reinterpret_cast<u32*>(&v);
	const u32* src = reinterpret_cast<const u32*>(p);
#pragma unroll
	for (int k = 0; k < int(sizeof(T) / 4); ++k)
		dst[k] = __ldcg(src + k);
	return v;
}

template <typename T>
DEVFN void st_value(char* base, size_t tile, const T& v)
{
	*reinterpret_cast<T*>(base + tile * SCAN_CHAIN_VALUE_BYTES) = v;
}

template <typename T>
DEVFN T shfl_down_struct(const T& v, int d)
{
	T r;
	const u32* src = reinterpret_cast<const u32*>(&v);
	u32* dst = reinterpret_cast<u32*>(&r);
#pragma unroll
	for (int k = 0; k < int(sizeof(T) / 4); ++k)
		dst[k] = __shfl_down_sync(0xffffffffu, src[k], d);
	return r;
}

template <typename T>
DEVFN T shfl_idx_struct(const T& v, int lane)
{
	T r;
	const u32* src = reinterpret_cast<const u32*>(&v);
	u32* dst = reinterpret_cast<u32*>(&r);
#pragma unroll
	for (int k = 0; k < int(sizeof(T) / 4); ++k)
		dst[k] = __shfl_sync(0xffffffffu, src[k], lane);
	return r;
}

// ---- packed descriptors for 4- and 8-byte values ---------------------------------------------------------------------
// A look-back iteration costs one L2 round trip (~0.7 us on B200) and resolves one window of predecessors, so the chain
// advances at most window / round-trip tiles per second. With separate flag and value arrays (two dependent round trips and
// a fence per window of 32) that bound is ~20 tiles/us = 0.3 TB/s of 8 KB tiles. Here status and value share one word that
// is written and read with a single aligned access (no fence, one round trip), every lane inspects SCAN_LOOK descriptors per
// iteration (window = 32 * SCAN_LOOK tiles), and large inputs use 32 KB tiles: the bound moves above the HBM rate.
static const int SCAN_LOOK = 4;

template <typename T>
struct ScanDesc;

template <>
struct ScanDesc<u32>
{
	typedef unsigned long long Word;
	DEVFN Word pack(u32 tag, u32 v)
	{
		return (Word(tag) << 32) | Word(v);
	}
	DEVFN u32 tag(Word w)
	{
		return u32(w >> 32);
	}
	DEVFN u32 value(Word w)
	{
		return u32(w);
	}
	DEVFN Word load(const char* base, size_t tile)
	{
		return *reinterpret_cast<const volatile Word*>(base + tile * 16);
	}
	DEVFN void store(char* base, size_t tile, Word w)
	{
		*reinterpret_cast<volatile Word*>(base + tile * 16) = w;
	}
};

template <>
struct ScanDesc<u64>
{
	typedef ulonglong2 Word;
	DEVFN Word pack(u32 tag, u64 v)
	{
		Word w;
		w.x = v;
		w.y = tag;
		return w;
	}
	DEVFN u32 tag(Word w)
	{
		return u32(w.y);
	}
	DEVFN u64 value(Word w)
	{
		return w.x;
	}
	// one aligned 16-byte access: value and tag travel together
	DEVFN Word load(const char* base, size_t tile)
	{
		Word w;
		asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(base + tile * 16) : "memory");
		return w;
	}
	DEVFN void store(char* base, size_t tile, Word w)
	{
		asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(base + tile * 16), "l"(w.x), "l"(w.y) : "memory");
	}
};

// Called by all 32 lanes of warp 0 with the tile's aggregate; returns the tile's exclusive prefix (valid in every lane) and
// publishes the tile's inclusive prefix. desc: 16 bytes per tile. Combine(a, b): a precedes b.
template <typename T, typename Op>
DEVFN T scan_chain_lookback_packed(u32 tile, const T& tile_aggregate, char* desc, u32 epoch)
{
	typedef ScanDesc<T> D;
	const int lane = threadIdx.x & 31;
	const u32 tag_agg = (epoch << 2) | 1u, tag_inc = (epoch << 2) | 2u;
	if (tile == 0)
	{
		if (lane == 0)
			D::store(desc, 0, D::pack(tag_inc, tile_aggregate));
		return Op::identity();
	}
	const bool independent = Op::prefix_independent(tile_aggregate);
	if (lane == 0)
		D::store(desc, tile, D::pack(independent ? tag_inc : tag_agg, tile_aggregate));
	T prefix = Op::identity();
	int p = int(tile) - 1;
	for (;;)
	{
		// lane L looks at tiles p - L*SCAN_LOOK - k, k = 0..SCAN_LOOK-1 (k = 0 nearest)
		u32 tags[SCAN_LOOK];
		T vals[SCAN_LOOK];
		for (;;)
		{
			bool ready = true;
#pragma unroll
			for (int k = 0; k < SCAN_LOOK; ++k)
			{
				int idx = p - lane * SCAN_LOOK - k;
				if (idx >= 0)
				{
					typename D::Word w = D::load(desc, size_t(idx));
					tags[k] = D::tag(w);
					vals[k] = D::value(w);
				}
				else
				{
					tags[k] = tag_inc;
					vals[k] = Op::identity();
				}
				ready = ready && (tags[k] == tag_agg || tags[k] == tag_inc);
			}
			if (__all_sync(0xffffffffu, ready))
				break;
		}
		// nearest inclusive prefix inside this lane, then across lanes
		int local_inc = SCAN_LOOK;
#pragma unroll
		for (int k = SCAN_LOOK - 1; k >= 0; --k)
			if (tags[k] == tag_inc)
				local_inc = k;
		unsigned inc_mask = __ballot_sync(0xffffffffu, local_inc < SCAN_LOOK);
		int first_lane = inc_mask ? __ffs(inc_mask) - 1 : 32;
		// ordered combination of this lane's descriptors up to (and including) the cut: farther back = left operand
		T v = Op::identity();
		if (lane <= first_lane)
		{
			int last = lane == first_lane ? local_inc : SCAN_LOOK - 1;
#pragma unroll
			for (int k = 0; k < SCAN_LOOK; ++k)
				if (k <= last)
					v = Op::apply(vals[k], v);
		}
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			T t = shfl_down_struct(v, d);
			if (lane + d < 32)
				v = Op::apply(t, v);
		}
		v = shfl_idx_struct(v, 0);
		prefix = Op::apply(v, prefix);
		if (inc_mask)
			break;
		p -= 32 * SCAN_LOOK;
	}
	if (lane == 0 && !independent)
		D::store(desc, tile, D::pack(tag_inc, Op::apply(prefix, tile_aggregate)));
	return prefix;
}

template <typename T, typename Op, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_scan_chained(const T* __restrict__ in, T* __restrict__ out, size_t n, char* desc, u32 epoch, T* __restrict__ total_out)
{
	__shared__ T smem[34];
	__shared__ T s_prefix;
	const int TILE = THREADS * ITEMS;
	size_t base = size_t(blockIdx.x) * TILE + size_t(threadIdx.x) * ITEMS;
	T v[ITEMS];
	if (base + ITEMS <= n)
	{
		// ITEMS contiguous elements per thread: 16-byte vector loads
		const uint4* src = reinterpret_cast<const uint4*>(in + base);
		uint4* dst = reinterpret_cast<uint4*>(v);
#pragma unroll
		for (int k = 0; k < int(sizeof(T) * ITEMS / 16); ++k)
			dst[k] = src[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < ITEMS; ++k)
			v[k] = base + k < n ? in[base + k] : Op::identity();
	}
	T sum = Op::identity();
#pragma unroll
	for (int k = 0; k < ITEMS; ++k)
		sum = Op::apply(sum, v[k]);
	T total;
	T ex = block_exclusive_scan<T, Op>(sum, &total, smem);
	if (threadIdx.x < 32)
	{
		T prefix = scan_chain_lookback_packed<T, Op>(blockIdx.x, total, desc, epoch);
		if (threadIdx.x == 0)
		{
			s_prefix = prefix;
			if (total_out && blockIdx.x == gridDim.x - 1)
				*total_out = Op::apply(prefix, total);
		}
	}
	__syncthreads();
	T run = Op::apply(s_prefix, ex);
#pragma unroll
	for (int k = 0; k < ITEMS; ++k)
	{
		T t = v[k];
		v[k] = run;
		run = Op::apply(run, t);
	}
	if (base + ITEMS <= n)
	{
		uint4* dst = reinterpret_cast<uint4*>(out + base);
		const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
		for (int k = 0; k < int(sizeof(T) * ITEMS / 16); ++k)
			dst[k] = src[k];
	}
	else
	{
#pragma unroll
		for (int k = 0; k < ITEMS; ++k)
			if (base + k < n)
				out[base + k] = v[k];
	}
}

// out[i] = op(in[0..i)), in-place allowed; optional device-side total. `in`/`out` must be 16-byte aligned (arena
// allocations are 256-byte aligned).
static const size_t SCAN_LARGE_N = size_t(1) << 21;
static const int SCAN_LARGE_THREADS = 512;
static const int SCAN_LARGE_ITEMS = 16;
template <typename T, typename Op>
static inline void exclusive_scan(const T* in, T* out, size_t n, T* total, Arena&)
{
	if (n == 0)
	{
		if (total)
			dev_memset(total, 0, sizeof(T));
		return;
	}
	u32 epoch = scan_chain_next_epoch();
	const bool large = n >= SCAN_LARGE_N;
	const size_t tile = large ? size_t(SCAN_LARGE_THREADS) * SCAN_LARGE_ITEMS : size_t(SCAN_TILE);
	size_t nblocks = (n + tile - 1) / tile;
	scan_chain_reserve(nblocks); // may reallocate the descriptor arrays: take the pointer afterwards
#ifdef CLODB_DEBUG_SHARED_DESC
	// build variant for tests/test_prims.py::test_scan_descriptor_formats_do_not_alias: the pre-fix layout, to show the test bites
	char* const chain_desc = g_scan_chain.desc;
#else
	char* const chain_desc = sizeof(T) == 8 ? g_scan_chain.desc64 : g_scan_chain.desc;
#endif
	if (large)
		LAUNCH_GRID((k_scan_chained<T, Op, SCAN_LARGE_THREADS, SCAN_LARGE_ITEMS>), nblocks, SCAN_LARGE_THREADS, in, out, n, chain_desc, epoch, total);
	else
		LAUNCH_GRID((k_scan_chained<T, Op, SCAN_THREADS, SCAN_ITEMS>), nblocks, SCAN_THREADS, in, out, n, chain_desc, epoch, total);
}

// ---- radix sort ------------------------------------------------------------------------------------------------
static const int RS_THREADS = 256;
static const int RS_ITEMS = 8;
static const int RS_WARPS = RS_THREADS / 32;
static const int RS_TILE = RS_THREADS * RS_ITEMS;

template <typename K>
__global__ void k_rs_hist(const K* __restrict__ keys, u32* __restrict__ counts, size_t n, int shift, u32 nblocks)
{
	__shared__ u32 hist[256];
	hist[threadIdx.x] = 0;
	__syncthreads();
	size_t tile = size_t(blockIdx.x) * RS_TILE;
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r)
	{
		size_t i = tile + size_t(warp) * (32 * RS_ITEMS) + r * 32 + lane;
		if (i < n)
			atomicAdd(&hist[u32(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	counts[size_t(threadIdx.x) * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Scatter pass. Ranks come from warp match_any (in-order rank among the warp's items with the same digit) and running bases
// across the warps of the CTA, as before; what changed is where the keys go first: into shared memory at their position in the
// tile's digit-sorted order, and from there to global memory with consecutive threads writing consecutive addresses of a
// (digit, tile) run. The direct version issued one 4-byte store per key to 32 different sectors per warp instruction and ran at a
// fifth of the bandwidth its DRAM traffic needed (ncu: L2 transaction bound, not DRAM bound).
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K* __restrict__ keys, K* __restrict__ keys_out, const u32* __restrict__ vals, u32* __restrict__ vals_out, const u32* __restrict__ offsets, size_t n, int shift, u32 nblocks)
{
	__shared__ u32 hist[RS_WARPS][256];
	__shared__ K s_keys[RS_TILE];
	__shared__ u32 s_vals[RS_TILE];
	__shared__ u32 s_goff[256];  // global position of the run of digit d minus its first position in the tile order
	__shared__ u32 s_warp_total[RS_WARPS];
	for (int w = 0; w < RS_WARPS; ++w)
		hist[w][threadIdx.x] = 0;
	__syncthreads();

	const size_t tile = size_t(blockIdx.x) * RS_TILE;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 lt_mask = (1u << lane) - 1;

	K key[RS_ITEMS];
	u32 rank[RS_ITEMS];

#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r)
	{
		size_t i = tile + size_t(warp) * (32 * RS_ITEMS) + r * 32 + lane;
		bool valid = i < n;
		unsigned active = __ballot_sync(0xffffffffu, valid);
		rank[r] = 0;
		if (valid)
		{
			key[r] = keys[i];
			u32 digit = u32(key[r] >> shift) & 255u;
			unsigned peers = __match_any_sync(active, digit);
			u32 cnt = __popc(peers);
			bool leader = (peers & lt_mask) == 0;
			if (leader)
				hist[warp][digit] += cnt;
			__syncwarp(active);
			// in-order rank of this item among the warp's items with the same digit
			rank[r] = hist[warp][digit] - cnt + __popc(peers & lt_mask);
			__syncwarp(active);
		}
	}
	__syncthreads();

	// thread d: the digit's count in the tile, running bases across the warps, and (block-wide exclusive scan over the digits) the
	// first position of the digit in the tile's sorted order
	{
		const u32 d = threadIdx.x;
		u32 total = 0;
		for (int w = 0; w < RS_WARPS; ++w)
		{
			u32 c = hist[w][d];
			hist[w][d] = total;
			total += c;
		}
		u32 inc = total;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u32 t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o)
				inc += t;
		}
		if (lane == 31)
			s_warp_total[warp] = inc;
		__syncthreads();
		u32 before = 0;
		for (int w = 0; w < warp; ++w)
			before += s_warp_total[w];
		const u32 first = before + inc - total; // exclusive prefix over the digits
		s_goff[d] = offsets[size_t(d) * nblocks + blockIdx.x] - first;
		for (int w = 0; w < RS_WARPS; ++w)
			hist[w][d] += first;
	}
	__syncthreads();

#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r)
	{
		size_t i = tile + size_t(warp) * (32 * RS_ITEMS) + r * 32 + lane;
		if (i < n)
		{
			u32 digit = u32(key[r] >> shift) & 255u;
			u32 pos = hist[warp][digit] + rank[r];
			s_keys[pos] = key[r];
			if (vals)
				s_vals[pos] = vals[i];
		}
	}
	__syncthreads();

	const u32 count = n - tile < size_t(RS_TILE) ? u32(n - tile) : u32(RS_TILE);
#pragma unroll
	for (int q = 0; q < RS_ITEMS; ++q)
	{
		u32 pos = u32(q) * RS_THREADS + threadIdx.x;
		if (pos < count)
		{
			K k = s_keys[pos];
			u32 dst = s_goff[u32(k >> shift) & 255u] + pos;
			keys_out[dst] = k;
			if (vals)
				vals_out[dst] = s_vals[pos];
		}
	}
}

// Stable LSD radix sort of (key, value) pairs on key bits [bit_lo, bit_hi). Result ends in keys/vals; *_tmp are scratch of
// the same size. vals may be null (keys only).
template <typename K>
static inline void radix_sort_pairs(K* keys, K* keys_tmp, u32* vals, u32* vals_tmp, size_t n, int bit_lo, int bit_hi, Arena& arena)
{
	if (n == 0 || bit_hi <= bit_lo)
		return;
	ArenaScope scope(arena);
	u32 nblocks = u32((n + RS_TILE - 1) / RS_TILE);
	u32* counts = arena.alloc<u32>(size_t(256) * nblocks);
	K* src = keys;
	K* dst = keys_tmp;
	u32* vsrc = vals;
	u32* vdst = vals_tmp;
	for (int shift = bit_lo; shift < bit_hi; shift += 8)
	{
		LAUNCH_GRID(k_rs_hist<K>, nblocks, RS_THREADS, src, counts, n, shift, nblocks);
		exclusive_scan<u32, OpAddU32>(counts, counts, size_t(256) * nblocks, nullptr, arena);
		LAUNCH_GRID(k_rs_scatter<K>, nblocks, RS_THREADS, src, dst, vsrc, vdst, counts, n, shift, nblocks);
		K* t = src;
		src = dst;
		dst = t;
		u32* vt = vsrc;
		vsrc = vdst;
		vdst = vt;
	}
	if (src != keys)
	{
		dev_d2d(keys, src, n * sizeof(K));
		if (vals)
			dev_d2d(vals, vsrc, n * sizeof(u32));
	}
}
#endif

static inline void exclusive_scan_u32(const u32* in, u32* out, size_t n, u32* total, Arena& arena)
{
	exclusive_scan<u32, OpAddU32>(in, out, n, total, arena);
}

static inline int bits_for(u64 max_value)
{
	int b = 1;
	while (b < 64 && (max_value >> b) != 0)
		++b;
	return b;
}

} // namespace clodb
