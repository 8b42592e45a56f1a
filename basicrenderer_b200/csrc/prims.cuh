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

template <typename T, typename Op>
__global__ void k_scan_reduce(const T* __restrict__ in, T* __restrict__ block_sums, size_t n)
{
	__shared__ T smem[34];
	size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
	T sum = Op::identity();
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
		if (base + k < n)
			sum = Op::apply(sum, in[base + k]);
	T total;
	block_exclusive_scan<T, Op>(sum, &total, smem);
	if (threadIdx.x == 0)
		block_sums[blockIdx.x] = total;
}

// single-CTA scan of the per-tile sums (1024 threads, looping with a running carry)
template <typename T, typename Op>
__global__ void k_scan_blocksums(T* __restrict__ block_sums, size_t nblocks, T* __restrict__ total_out)
{
	__shared__ T smem[34];
	__shared__ T carry_s;
	if (threadIdx.x == 0)
		carry_s = Op::identity();
	__syncthreads();
	for (size_t base = 0; base < nblocks; base += blockDim.x)
	{
		size_t i = base + threadIdx.x;
		T v = i < nblocks ? block_sums[i] : Op::identity();
		T total;
		T ex = block_exclusive_scan<T, Op>(v, &total, smem);
		T carry = carry_s;
		if (i < nblocks)
			block_sums[i] = Op::apply(carry, ex);
		__syncthreads();
		if (threadIdx.x == 0)
			carry_s = Op::apply(carry, total);
		__syncthreads();
	}
	if (threadIdx.x == 0 && total_out)
		*total_out = carry_s;
}

template <typename T, typename Op>
__global__ void k_scan_apply(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ block_sums, size_t n)
{
	__shared__ T smem[34];
	size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
	T v[SCAN_ITEMS];
	T sum = Op::identity();
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		v[k] = base + k < n ? in[base + k] : Op::identity();
		sum = Op::apply(sum, v[k]);
	}
	T total;
	T run = Op::apply(block_sums[blockIdx.x], block_exclusive_scan<T, Op>(sum, &total, smem));
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
	{
		T t = v[k];
		v[k] = run;
		run = Op::apply(run, t);
	}
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
		if (base + k < n)
			out[base + k] = v[k];
}

// out[i] = op(in[0..i)), in-place allowed; optional device-side total
template <typename T, typename Op>
static inline void exclusive_scan(const T* in, T* out, size_t n, T* total, Arena& arena)
{
	if (n == 0)
	{
		if (total)
			dev_memset(total, 0, sizeof(T));
		return;
	}
	ArenaScope scope(arena);
	size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
	T* block_sums = arena.alloc<T>(nblocks);
	LAUNCH_GRID((k_scan_reduce<T, Op>), nblocks, SCAN_THREADS, in, block_sums, n);
	LAUNCH_GRID((k_scan_blocksums<T, Op>), 1, 1024, block_sums, nblocks, total);
	LAUNCH_GRID((k_scan_apply<T, Op>), nblocks, SCAN_THREADS, in, out, block_sums, n);
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

template <typename K>
__global__ void k_rs_scatter(const K* __restrict__ keys, K* __restrict__ keys_out, const u32* __restrict__ vals, u32* __restrict__ vals_out, const u32* __restrict__ offsets, size_t n, int shift, u32 nblocks)
{
	__shared__ u32 hist[RS_WARPS][256];
	for (int w = 0; w < RS_WARPS; ++w)
		hist[w][threadIdx.x] = 0;
	__syncthreads();

	size_t tile = size_t(blockIdx.x) * RS_TILE;
	int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u32 lt_mask = (1u << lane) - 1;

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

	// thread d: running base for digit d across the warps of this CTA (global offset of (digit, tile) first)
	{
		u32 d = threadIdx.x;
		u32 run = offsets[size_t(d) * nblocks + blockIdx.x];
		for (int w = 0; w < RS_WARPS; ++w)
		{
			u32 c = hist[w][d];
			hist[w][d] = run;
			run += c;
		}
	}
	__syncthreads();

#pragma unroll
	for (int r = 0; r < RS_ITEMS; ++r)
	{
		size_t i = tile + size_t(warp) * (32 * RS_ITEMS) + r * 32 + lane;
		if (i < n)
		{
			u32 digit = u32(key[r] >> shift) & 255u;
			size_t dst = size_t(hist[warp][digit]) + rank[r];
			keys_out[dst] = key[r];
			if (vals)
				vals_out[dst] = vals[i];
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
