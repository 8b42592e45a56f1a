// clodb200 runtime layer implementation (see rt.cuh).
#include "prims.cuh"

#include <algorithm>

namespace clodb
{

thread_local stream_t g_stream = 0;
thread_local uint64_t g_launches = 0;
int g_sync_debug = 0;
thread_local int g_profile = 0;

#ifdef CLODB_EMU
thread_local size_t emu_tid = 0;
int emu_reverse = 0;

void* dev_malloc(size_t bytes)
{
	void* p = malloc(bytes ? bytes : 1);
	if (!p)
		throw Error("clodb200(emu): out of memory");
	return p;
}
void dev_free(void* p)
{
	free(p);
}
void dev_memset(void* p, int value, size_t bytes)
{
	memset(p, value, bytes);
}
void dev_h2d(void* dst, const void* src, size_t bytes)
{
	memcpy(dst, src, bytes);
}
void dev_d2h(void* dst, const void* src, size_t bytes)
{
	memcpy(dst, src, bytes);
}
void dev_d2d(void* dst, const void* src, size_t bytes)
{
	memmove(dst, src, bytes);
}
void dev_sync()
{
}
void HostStage::reserve(size_t bytes)
{
	if (bytes <= capacity)
		return;
	free(base);
	base = static_cast<char*>(malloc(bytes));
	capacity = bytes;
}
void HostStage::destroy()
{
	free(base);
	base = nullptr;
	capacity = 0;
}
void dev_d2h_async(void* dst, const void* src, size_t bytes)
{
	memcpy(dst, src, bytes);
}
void dev_d2h_async_wait()
{
}
#else
namespace
{
thread_local cudaStream_t g_copy_stream = nullptr;
thread_local cudaEvent_t g_copy_ready = nullptr, g_copy_done = nullptr;
void ensure_copy_stream()
{
	if (g_copy_stream)
		return;
	CUDA_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
	CUDA_CHECK(cudaEventCreateWithFlags(&g_copy_ready, cudaEventDisableTiming));
	CUDA_CHECK(cudaEventCreateWithFlags(&g_copy_done, cudaEventDisableTiming));
}
} // namespace
void HostStage::reserve(size_t bytes)
{
	if (bytes <= capacity)
		return;
	dev_d2h_async_wait();
	if (base)
		cudaFreeHost(base);
	base = nullptr;
	size_t cap = bytes + bytes / 4;
	CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&base), cap, cudaHostAllocDefault));
	capacity = cap;
}
void HostStage::destroy()
{
	if (base)
		cudaFreeHost(base);
	base = nullptr;
	capacity = 0;
}
void dev_d2h_async(void* dst, const void* src, size_t bytes)
{
	if (!bytes)
		return;
	ensure_copy_stream();
	CUDA_CHECK(cudaEventRecord(g_copy_ready, g_stream));
	CUDA_CHECK(cudaStreamWaitEvent(g_copy_stream, g_copy_ready, 0));
	CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_copy_stream));
	CUDA_CHECK(cudaEventRecord(g_copy_done, g_copy_stream));
}
void dev_d2h_async_wait()
{
	if (g_copy_stream)
		CUDA_CHECK(cudaStreamSynchronize(g_copy_stream));
}

void* dev_malloc(size_t bytes)
{
	void* p = nullptr;
	CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
	return p;
}
void dev_free(void* p)
{
	if (p)
		cudaFree(p);
}
void dev_memset(void* p, int value, size_t bytes)
{
	if (bytes)
		CUDA_CHECK(cudaMemsetAsync(p, value, bytes, g_stream));
}
// Small control transfers (scalars, a few offsets) go through pinned memory: an upload is copied into a slot of a pinned
// ring and enqueued without waiting (the ring is drained once per wrap-around, so a slot is never overwritten while its
// copy is pending); a read-back lands in a pinned word and costs one stream synchronisation instead of a pageable-memory
// staging copy. Larger transfers keep the plain path.
static const size_t SMALL_XFER = 1024;
static const size_t RING_SLOTS = 512;
static thread_local char* g_pinned_ring = nullptr; // RING_SLOTS upload slots + one read-back slot
static thread_local size_t g_ring_next = 0;

static char* pinned_ring()
{
	if (!g_pinned_ring)
		CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&g_pinned_ring), (RING_SLOTS + 1) * SMALL_XFER));
	return g_pinned_ring;
}

void dev_h2d(void* dst, const void* src, size_t bytes)
{
	if (!bytes)
		return;
	if (bytes <= SMALL_XFER)
	{
		char* ring = pinned_ring();
		if (g_ring_next == RING_SLOTS)
		{
			CUDA_CHECK(cudaStreamSynchronize(g_stream));
			g_ring_next = 0;
		}
		char* slot = ring + g_ring_next++ * SMALL_XFER;
		memcpy(slot, src, bytes);
		CUDA_CHECK(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, g_stream));
		return;
	}
	CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
	CUDA_CHECK(cudaStreamSynchronize(g_stream));
}
void dev_d2h(void* dst, const void* src, size_t bytes)
{
	if (!bytes)
		return;
	if (bytes <= SMALL_XFER)
	{
		char* slot = pinned_ring() + RING_SLOTS * SMALL_XFER;
		CUDA_CHECK(cudaMemcpyAsync(slot, src, bytes, cudaMemcpyDeviceToHost, g_stream));
		CUDA_CHECK(cudaStreamSynchronize(g_stream));
		memcpy(dst, slot, bytes);
		g_ring_next = 0; // the stream is drained: every pending upload has landed
		return;
	}
	CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
	CUDA_CHECK(cudaStreamSynchronize(g_stream));
}
void dev_d2d(void* dst, const void* src, size_t bytes)
{
	if (bytes)
		CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
}
void dev_sync()
{
	CUDA_CHECK(cudaStreamSynchronize(g_stream));
}
#endif

#ifdef CLODB_EMU
void profile_mark(const char*, int, size_t)
{
}
void profile_set_last_work(size_t)
{
}
std::string profile_report()
{
	return std::string();
}
#else
namespace
{
struct ProfileSpan
{
	const char* name;
	cudaEvent_t start, stop;
	size_t threads;
};
// Heap objects behind trivially destructible thread_local pointers: rt_thread_release() may run from another thread_local's
// destructor at thread exit, when thread_local objects with destructors of their own could already be gone.
thread_local std::vector<ProfileSpan>* t_spans = nullptr;
thread_local std::vector<cudaEvent_t>* t_event_pool = nullptr;
std::vector<ProfileSpan>& spans_ref()
{
	if (!t_spans)
		t_spans = new std::vector<ProfileSpan>();
	return *t_spans;
}
std::vector<cudaEvent_t>& event_pool_ref()
{
	if (!t_event_pool)
		t_event_pool = new std::vector<cudaEvent_t>();
	return *t_event_pool;
}
#define g_spans (spans_ref())
#define g_event_pool (event_pool_ref())

cudaEvent_t take_event()
{
	if (!g_event_pool.empty())
	{
		cudaEvent_t e = g_event_pool.back();
		g_event_pool.pop_back();
		return e;
	}
	cudaEvent_t e;
	CUDA_CHECK(cudaEventCreate(&e));
	return e;
}
} // namespace

void profile_mark(const char* kernel_name, int end, size_t threads)
{
	if (!end)
	{
		ProfileSpan s;
		s.name = kernel_name;
		s.start = take_event();
		s.stop = take_event();
		s.threads = threads;
		CUDA_CHECK(cudaEventRecord(s.start, g_stream));
		g_spans.push_back(s);
	}
	else if (!g_spans.empty())
	{
		CUDA_CHECK(cudaEventRecord(g_spans.back().stop, g_stream));
	}
}

void profile_set_last_work(size_t items)
{
	if (!g_spans.empty())
		g_spans.back().threads = items;
}

std::string profile_report()
{
	CUDA_CHECK(cudaStreamSynchronize(g_stream));
	struct Acc
	{
		std::string name;
		double ms = 0;
		uint64_t count = 0;
		uint64_t threads = 0;
	};
	std::vector<Acc> accs;
	for (const ProfileSpan& s : g_spans)
	{
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, s.start, s.stop));
		Acc* a = nullptr;
		for (Acc& x : accs)
			if (x.name == s.name)
				a = &x;
		if (!a)
		{
			accs.push_back(Acc());
			a = &accs.back();
			a->name = s.name;
		}
		a->ms += ms;
		a->count++;
		a->threads += s.threads;
		g_event_pool.push_back(s.start);
		g_event_pool.push_back(s.stop);
	}
	g_spans.clear();
	std::sort(accs.begin(), accs.end(), [](const Acc& a, const Acc& b) { return a.ms > b.ms; });
	std::string out;
	for (const Acc& a : accs)
		out += a.name + "," + std::to_string(a.count) + "," + std::to_string(a.ms) + "," + std::to_string(a.threads) + "\n";
	return out;
}
#endif

#ifndef CLODB_EMU
thread_local ScanChain g_scan_chain;

void scan_chain_reserve(size_t tiles)
{
	if (tiles <= g_scan_chain.capacity_tiles)
		return;
	size_t cap = std::max<size_t>(tiles * 2, size_t(1) << 16);
	CUDA_CHECK(cudaStreamSynchronize(g_stream));
	dev_free(g_scan_chain.desc);
	dev_free(g_scan_chain.desc64);
	dev_free(g_scan_chain.box_desc);
	g_scan_chain.desc = static_cast<char*>(dev_malloc(cap * 16));
	g_scan_chain.desc64 = static_cast<char*>(dev_malloc(cap * 16));
	g_scan_chain.box_desc = static_cast<char*>(dev_malloc(cap * SCAN_CHAIN_VALUE_BYTES));
	CUDA_CHECK(cudaMemsetAsync(g_scan_chain.desc, 0, cap * 16, g_stream));
	CUDA_CHECK(cudaMemsetAsync(g_scan_chain.desc64, 0, cap * 16, g_stream));
	CUDA_CHECK(cudaMemsetAsync(g_scan_chain.box_desc, 0, cap * SCAN_CHAIN_VALUE_BYTES, g_stream));
	g_scan_chain.capacity_tiles = cap;
}
#endif

unsigned host_thread_limit()
{
	static unsigned limit = 0;
	if (!limit)
	{
		unsigned hw = std::thread::hardware_concurrency();
		limit = hw ? (hw < 16 ? hw : 16u) : 1u;
		if (const char* e = getenv("CLODB200_HOST_THREADS"))
			limit = unsigned(std::max(1, atoi(e)));
	}
	return limit;
}

void rt_thread_release() noexcept
{
#ifndef CLODB_EMU
	// errors are ignored on purpose: at process exit the CUDA runtime may already be unloading
	if (g_stream)
		cudaStreamSynchronize(g_stream);
	if (g_copy_stream)
	{
		cudaStreamSynchronize(g_copy_stream);
		cudaStreamDestroy(g_copy_stream);
		cudaEventDestroy(g_copy_ready);
		cudaEventDestroy(g_copy_done);
		g_copy_stream = nullptr;
		g_copy_ready = g_copy_done = nullptr;
	}
	if (g_pinned_ring)
		cudaFreeHost(g_pinned_ring);
	g_pinned_ring = nullptr;
	g_ring_next = 0;
	cudaFree(g_scan_chain.desc);
	cudaFree(g_scan_chain.desc64);
	cudaFree(g_scan_chain.box_desc);
	g_scan_chain = ScanChain();
	if (t_spans)
	{
		for (const ProfileSpan& s : *t_spans)
		{
			cudaEventDestroy(s.start);
			cudaEventDestroy(s.stop);
		}
		delete t_spans;
		t_spans = nullptr;
	}
	if (t_event_pool)
	{
		for (cudaEvent_t e : *t_event_pool)
			cudaEventDestroy(e);
		delete t_event_pool;
		t_event_pool = nullptr;
	}
	if (g_stream)
		cudaStreamDestroy(g_stream);
	g_stream = 0;
	cudaGetLastError();
#endif
}

void Arena::init(size_t bytes)
{
	destroy();
	base = static_cast<char*>(dev_malloc(bytes));
	capacity = bytes;
	offset = 0;
	high_water = 0;
}

void* Arena::debug_alloc(size_t at, size_t bytes)
{
	void* p = dev_malloc(bytes);
	debug_blocks.push_back(std::make_pair(at, p));
	return p;
}

void Arena::debug_release(size_t m)
{
	dev_sync();
	while (!debug_blocks.empty() && debug_blocks.back().first >= m)
	{
		dev_free(debug_blocks.back().second);
		debug_blocks.pop_back();
	}
}

void Arena::destroy()
{
	if (!debug_blocks.empty())
		debug_release(0);
	if (base)
		dev_free(base);
	base = nullptr;
	capacity = offset = 0;
}

} // namespace clodb
