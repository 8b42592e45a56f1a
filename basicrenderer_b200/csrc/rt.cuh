// clodb200 runtime layer: device memory arena, stream, kernel-launch macros.
//
// Product builds compile this with nvcc for sm_100a. The same kernel sources can also be compiled by g++ with
// -DCLODB_EMU into a *development-only* host emulation (tests/emu): every kernel body is then run as a serial loop over
// its thread index. That build exists so kernel logic can be debugged (gdb/asan) in a container without a GPU; it is
// never linked into libclodb200.so and the product library has no CPU code path.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#ifdef CLODB_EMU
#include <cmath>
#include <algorithm>
#else
#include <cuda_runtime.h>
#endif

namespace clodb
{

struct Error : std::runtime_error
{
	explicit Error(const std::string& what)
	    : std::runtime_error(what)
	{
	}
};

#ifdef CLODB_EMU
// ------------------------------------------------------------------------------------------------ host emulation
#define KERNEL static void
#define DEVFN static inline
#define HOSTDEVFN static inline
#define CONSTANT static const

extern thread_local size_t emu_tid;
extern int emu_reverse;
#define GTID (::clodb::emu_tid)

#define LAUNCH(kernel, n, ...)                                               \
	do                                                                       \
	{                                                                        \
		size_t _n = size_t(n);                                               \
		::clodb::g_launches++;                                               \
		if (::clodb::emu_reverse)                                            \
			for (size_t _i = _n; _i-- > 0;)                                  \
			{                                                                \
				::clodb::emu_tid = _i;                                       \
				kernel(__VA_ARGS__);                                         \
			}                                                                \
		else                                                                 \
			for (size_t _i = 0; _i < _n; ++_i)                               \
			{                                                                \
				::clodb::emu_tid = _i;                                       \
				kernel(__VA_ARGS__);                                         \
			}                                                                \
	} while (0)

template <typename T>
static inline T atomicAdd(T* p, T v)
{
	T old = *p;
	*p = old + v;
	return old;
}
template <typename T>
static inline T atomicMin(T* p, T v)
{
	T old = *p;
	if (v < old)
		*p = v;
	return old;
}
template <typename T>
static inline T atomicMax(T* p, T v)
{
	T old = *p;
	if (v > old)
		*p = v;
	return old;
}
template <typename T>
static inline T atomicOr(T* p, T v)
{
	T old = *p;
	*p = old | v;
	return old;
}
template <typename T>
static inline T atomicExch(T* p, T v)
{
	T old = *p;
	*p = v;
	return old;
}
template <typename T>
static inline T atomicCAS(T* p, T cmp, T v)
{
	T old = *p;
	if (old == cmp)
		*p = v;
	return old;
}
static inline unsigned int __float_as_uint(float f)
{
	unsigned int u;
	memcpy(&u, &f, 4);
	return u;
}
static inline float __uint_as_float(unsigned int u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}
static inline int __popc(unsigned int v)
{
	return __builtin_popcount(v);
}
static inline int __clz(unsigned int v)
{
	return v ? __builtin_clz(v) : 32;
}
template <typename T>
static inline T __ldg(const T* p)
{
	return *p;
}
typedef void* stream_t;

#else
// ------------------------------------------------------------------------------------------------ CUDA (sm_100a)
#define KERNEL static __global__ void
#define DEVFN static __device__ __forceinline__
#define HOSTDEVFN static __host__ __device__ __forceinline__
#define CONSTANT static __constant__ const
#define GTID (size_t(blockIdx.x) * blockDim.x + threadIdx.x)

typedef cudaStream_t stream_t;

#define CUDA_CHECK(expr)                                                                                                 \
	do                                                                                                                   \
	{                                                                                                                    \
		cudaError_t _e = (expr);                                                                                         \
		if (_e != cudaSuccess)                                                                                           \
			throw ::clodb::Error(std::string("CUDA error ") + cudaGetErrorString(_e) + " at " __FILE__ ":" + std::to_string(__LINE__) + " in " #expr); \
	} while (0)

// one thread per element, 256-thread CTAs; grid rounded up (callers guard with `if (i >= n) return`)
#define LAUNCH(kernel, n, ...)                                                                         \
	do                                                                                                 \
	{                                                                                                  \
		size_t _n = size_t(n);                                                                         \
		if (_n > 0)                                                                                    \
		{                                                                                              \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 0, _n);                                                 \
			kernel<<<(unsigned int)((_n + 255) / 256), 256, 0, ::clodb::g_stream>>>(__VA_ARGS__);       \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 1, 0);                                                     \
			::clodb::g_launches++;                                                                     \
			CUDA_CHECK(cudaGetLastError());                                                            \
			if (::clodb::g_sync_debug)                                                                 \
				CUDA_CHECK(cudaStreamSynchronize(::clodb::g_stream));                                  \
		}                                                                                              \
	} while (0)

// explicit grid/block launch for cooperative (block-level) kernels
#define LAUNCH_GRID(kernel, grid, block, ...)                                                          \
	do                                                                                                 \
	{                                                                                                  \
		if ((grid) > 0)                                                                                \
		{                                                                                              \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 0, size_t(grid) * size_t(block));                       \
			kernel<<<(unsigned int)(grid), (block), 0, ::clodb::g_stream>>>(__VA_ARGS__);              \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 1, 0);                                                     \
			::clodb::g_launches++;                                                                     \
			CUDA_CHECK(cudaGetLastError());                                                            \
			if (::clodb::g_sync_debug)                                                                 \
				CUDA_CHECK(cudaStreamSynchronize(::clodb::g_stream));                                  \
		}                                                                                              \
	} while (0)

// cooperative launch (grid-wide barriers inside the kernel); the caller sizes the grid to be co-resident
#define LAUNCH_COOP(kernel, grid, block, args_struct)                                                  \
	do                                                                                                 \
	{                                                                                                  \
		if ((grid) > 0)                                                                                \
		{                                                                                              \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 0, size_t(grid) * size_t(block));                       \
			void* _args[] = {(void*)&(args_struct)};                                                   \
			CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned int)(grid)), dim3(block), _args, 0, ::clodb::g_stream)); \
			if (::clodb::g_profile)                                                                    \
				::clodb::profile_mark(#kernel, 1, 0);                                                  \
			::clodb::g_launches++;                                                                     \
			if (::clodb::g_sync_debug)                                                                 \
				CUDA_CHECK(cudaStreamSynchronize(::clodb::g_stream));                                  \
		}                                                                                              \
	} while (0)
#endif

// Build context of the calling host thread: every thread that calls into the library gets its own stream, launch counter,
// workspace arenas, scan-chain descriptors and pinned staging (capi.cu creates them on first use), so independent meshes can
// be built concurrently from several threads of one process (scene batches of small meshes are launch-latency bound).
extern thread_local stream_t g_stream;
extern thread_local uint64_t g_launches;
extern int g_sync_debug;

// optional per-kernel timing with CUDA events on the launch stream (bench.py's roofline leg); off by default
extern thread_local int g_profile;
void profile_mark(const char* kernel_name, int end, size_t threads);
// persistent kernels: replace the thread count of the span just recorded by the number of work items it processed
void profile_set_last_work(size_t items);
// drains recorded events; returns "name,launches,total_ms,total_threads\n" lines sorted by time
std::string profile_report();

// Frees everything the runtime layer holds for the calling thread (build stream, copy stream and events, pinned upload ring,
// scan-chain descriptors, profiling events). Never throws: it runs from thread-exit destructors and from shutdown.
void rt_thread_release() noexcept;

// ---- raw device memory -------------------------------------------------------------------------------------------
void* dev_malloc(size_t bytes);
void dev_free(void* p);
void dev_memset(void* p, int value, size_t bytes);
void dev_h2d(void* dst, const void* src, size_t bytes);
void dev_d2h(void* dst, const void* src, size_t bytes);
void dev_d2d(void* dst, const void* src, size_t bytes);
void dev_sync();

// ---- pinned host staging + copy stream: level outputs are downloaded asynchronously while the build stream keeps working
struct HostStage
{
	char* base = nullptr;
	size_t capacity = 0;
	void reserve(size_t bytes); // grows (never shrinks); contents are not preserved
	void destroy();
};
// Enqueues dst_host <- src_dev on the copy stream once everything issued so far on the build stream has finished.
void dev_d2h_async(void* dst_host_pinned, const void* src, size_t bytes);
// Blocks the host until every dev_d2h_async issued so far has landed.
void dev_d2h_async_wait();

// ---- stack arena: all per-build temporaries come from one cudaMalloc'd slab (no allocator calls inside the loop) -----
struct Arena;
// Thrown by Arena::alloc_bytes. The slabs are sized from the triangle count; inputs whose adjacency is far denser than a
// surface's (triangle soups sharing each vertex among dozens of clusters) can need more, so the C-ABI build entry points
// grow the slab named here and run the (deterministic, side-effect free until it returns) build again.
struct ArenaExhausted : Error
{
	const Arena* arena;
	size_t need;
	ArenaExhausted(const Arena* a, size_t need_bytes, const std::string& what)
	    : Error(what), arena(a), need(need_bytes)
	{
	}
};

struct Arena
{
	char* base = nullptr;
	size_t capacity = 0;
	size_t offset = 0;
	size_t high_water = 0;

	void init(size_t bytes);
	void destroy();

	size_t mark() const
	{
		return offset;
	}
	void release(size_t m)
	{
		if (!debug_blocks.empty())
			debug_release(m);
		offset = m;
	}

	// CLODB200_DEBUG_MALLOC: every arena allocation becomes its own device allocation (freed when its scope is released), so
	// that compute-sanitizer sees overruns between neighbouring arrays of the slab. Development aid; off by default.
	std::vector<std::pair<size_t, void*>> debug_blocks;
	void* debug_alloc(size_t at, size_t bytes);
	void debug_release(size_t m);

	void* alloc_bytes(size_t bytes)
	{
		size_t aligned = (offset + 255) & ~size_t(255);
		static const bool debug_malloc = getenv("CLODB200_DEBUG_MALLOC") != nullptr;
		if (debug_malloc && aligned + bytes <= capacity)
		{
			offset = aligned + bytes;
			return debug_alloc(aligned, bytes);
		}
		if (aligned + bytes > capacity)
			throw ArenaExhausted(this, aligned + bytes, "clodb200: device arena exhausted (need " + std::to_string(aligned + bytes) + " of " + std::to_string(capacity) + " bytes)");
		offset = aligned + bytes;
		if (offset > high_water)
			high_water = offset;
		return base + aligned;
	}

	template <typename T>
	T* alloc(size_t count)
	{
		return static_cast<T*>(alloc_bytes(count * sizeof(T) + 16));
	}
};

struct ArenaScope
{
	Arena& arena;
	size_t m;
	explicit ArenaScope(Arena& a)
	    : arena(a), m(a.mark())
	{
	}
	~ArenaScope()
	{
		arena.release(m);
	}
};

// Host-side parallel loop for the replay stages that run between kernels (per-group table building, page accounting): splits
// [0, n) into contiguous chunks of at least `grain` items over up to 16 host threads; fn(begin, end, chunk_index) may only write
// to disjoint outputs. Runs inline when the range is small. CLODB200_HOST_THREADS caps the thread count (1 = serial).
unsigned host_thread_limit();
template <typename F>
static inline void host_parallel_for(size_t n, size_t grain, F&& fn)
{
	size_t chunks = grain ? n / grain : 1;
	unsigned limit = host_thread_limit();
	if (chunks > limit)
		chunks = limit;
	if (chunks <= 1)
	{
		if (n)
			fn(size_t(0), n, size_t(0));
		return;
	}
	std::vector<std::thread> threads;
	std::exception_ptr error;
	std::mutex error_mutex;
	auto run = [&](size_t c) {
		try
		{
			fn(n * c / chunks, n * (c + 1) / chunks, c);
		}
		catch (...)
		{
			std::lock_guard<std::mutex> lock(error_mutex);
			if (!error)
				error = std::current_exception();
		}
	};
	for (size_t c = 1; c < chunks; ++c)
		threads.emplace_back(run, c);
	run(0);
	for (std::thread& t : threads)
		t.join();
	if (error)
		std::rethrow_exception(error);
}
static inline size_t host_parallel_chunks(size_t n, size_t grain)
{
	size_t chunks = grain ? n / grain : 1;
	unsigned limit = host_thread_limit();
	return chunks > limit ? limit : (chunks < 1 ? 1 : chunks);
}

template <typename T>
static inline T dev_read(const T* p)
{
	T v;
	dev_d2h(&v, p, sizeof(T));
	return v;
}

template <typename T>
static inline std::vector<T> dev_download(const T* p, size_t count)
{
	std::vector<T> v(count);
	if (count)
		dev_d2h(v.data(), p, count * sizeof(T));
	return v;
}

} // namespace clodb
