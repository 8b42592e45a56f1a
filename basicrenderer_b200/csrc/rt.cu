// clodb200 runtime layer implementation (see rt.cuh).
#include "rt.cuh"

namespace clodb
{

stream_t g_stream = 0;
uint64_t g_launches = 0;
int g_sync_debug = 0;

#ifdef CLODB_EMU
size_t emu_tid = 0;
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
#else
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
void dev_h2d(void* dst, const void* src, size_t bytes)
{
	if (bytes)
	{
		CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
		CUDA_CHECK(cudaStreamSynchronize(g_stream));
	}
}
void dev_d2h(void* dst, const void* src, size_t bytes)
{
	if (bytes)
	{
		CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
		CUDA_CHECK(cudaStreamSynchronize(g_stream));
	}
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

void Arena::init(size_t bytes)
{
	destroy();
	base = static_cast<char*>(dev_malloc(bytes));
	capacity = bytes;
	offset = 0;
	high_water = 0;
}

void Arena::destroy()
{
	if (base)
		dev_free(base);
	base = nullptr;
	capacity = offset = 0;
}

} // namespace clodb
