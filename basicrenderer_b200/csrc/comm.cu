// Scene batches across the GPUs of one box: meshes are independent units sharded by mesh, so the build has
// no data-path collective. The path's only exchange is the gather of the per-mesh cache metadata blobs (the bytes
// CLodCache::SerializeMetadata writes, CLodCache.cpp:169-207) so that one rank can write the scene-level cache index:
// one NCCL all-gather of byte counts and one padded all-gather of the payloads over NVLink, on a communication stream of their own,
// so a rank keeps building its next mesh while the previous batch's blobs travel.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy PyTorch already loaded), so the
// library itself loads and builds single-GPU without it.
#include "../../include/clodb200.h"
#include "clodb.h"

#include <mutex>
#include <string>
#include <vector>

#ifndef CLODB_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

using namespace clodb;

namespace clodb
{
void capi_set_error(const std::string& message); // capi.cu: stores the calling thread's last error
bool capi_initialized();
} // namespace clodb

struct clodb200_gather
{
	int world = 1, rank = 0;
	size_t stride = 0;
	std::vector<unsigned long long> sizes;
	std::vector<unsigned char> loopback; // world size 1: the payload itself
#ifndef CLODB_EMU
	unsigned char* host = nullptr; // pinned, world * stride
	unsigned char *d_send = nullptr, *d_recv = nullptr;
	cudaEvent_t done = nullptr;
#endif
};

namespace
{
std::mutex g_comm_mutex;
int g_world = 1, g_rank = 0;
#ifndef CLODB_EMU
struct NcclApi
{
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;
ncclComm_t g_comm = nullptr;
cudaStream_t g_comm_stream = nullptr;
unsigned long long *g_d_sizes = nullptr, *g_h_sizes = nullptr; // world + 1 entries each (own size at [world])

bool load_nccl(std::string& why)
{
	if (g_nccl.handle)
		return true;
	const char* names[] = {getenv("CLODB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char* n : names)
	{
		if (!n || !*n)
			continue;
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle)
			break;
	}
	if (!g_nccl.handle)
	{
		why = std::string("clodb200: NCCL not found (") + (dlerror() ? dlerror() : "dlopen failed") + ")";
		return false;
	}
	g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.handle, "ncclGetUniqueId"));
	g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.handle, "ncclCommInitRank"));
	g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.handle, "ncclCommDestroy"));
	g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(g_nccl.handle, "ncclAllGather"));
	g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.handle, "ncclGetErrorString"));
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather)
	{
		why = "clodb200: libnccl lacks the expected entry points";
		return false;
	}
	return true;
}

int nccl_fail(ncclResult_t r, const char* where)
{
	capi_set_error(std::string("clodb200: ") + where + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
	return CLODB200_ERR_RUNTIME;
}

int cuda_fail(cudaError_t e, const char* where)
{
	capi_set_error(std::string("clodb200: ") + where + " failed: " + cudaGetErrorString(e));
	return CLODB200_ERR_RUNTIME;
}
#define COMM_CUDA(expr)                  \
	do                                   \
	{                                    \
		cudaError_t _e = (expr);         \
		if (_e != cudaSuccess)           \
			return cuda_fail(_e, #expr); \
	} while (0)
#endif
} // namespace

extern "C"
{

int clodb200_commGetUniqueId(void* out_id)
{
	if (!out_id)
		return CLODB200_ERR_INVALID;
#ifdef CLODB_EMU
	memset(out_id, 0, CLODB200_COMM_ID_BYTES);
	return CLODB200_OK;
#else
	std::lock_guard<std::mutex> lock(g_comm_mutex);
	std::string why;
	if (!load_nccl(why))
	{
		capi_set_error(why);
		return CLODB200_ERR_RUNTIME;
	}
	static_assert(sizeof(ncclUniqueId) == CLODB200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r != ncclSuccess)
		return nccl_fail(r, "ncclGetUniqueId");
	memcpy(out_id, &id, sizeof(id));
	return CLODB200_OK;
#endif
}

int clodb200_commInit(const void* id, int world_size, int rank)
{
	if (world_size < 1 || rank < 0 || rank >= world_size || (world_size > 1 && !id))
		return CLODB200_ERR_INVALID;
	if (!capi_initialized())
	{
		capi_set_error("clodb200: clodb200_commInit needs clodb200_init first (it selects this rank's device)");
		return CLODB200_ERR_NO_DEVICE;
	}
	std::lock_guard<std::mutex> lock(g_comm_mutex);
#ifdef CLODB_EMU
	if (world_size > 1)
	{
		capi_set_error("clodb200: the development emulation has no NCCL; world size must be 1");
		return CLODB200_ERR_RUNTIME;
	}
#else
	if (g_comm)
	{
		capi_set_error("clodb200: communicator already initialised");
		return CLODB200_ERR_INVALID;
	}
	if (world_size > 1)
	{
		std::string why;
		if (!load_nccl(why))
		{
			capi_set_error(why);
			return CLODB200_ERR_RUNTIME;
		}
		ncclUniqueId nid;
		memcpy(&nid, id, sizeof(nid));
		ncclResult_t r = g_nccl.CommInitRank(&g_comm, world_size, nid, rank);
		if (r != ncclSuccess)
			return nccl_fail(r, "ncclCommInitRank");
		COMM_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
		COMM_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_d_sizes), sizeof(unsigned long long) * (size_t(world_size) + 1)));
		COMM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g_h_sizes), sizeof(unsigned long long) * (size_t(world_size) + 1)));
	}
#endif
	g_world = world_size;
	g_rank = rank;
	return CLODB200_OK;
}

int clodb200_commWorldSize(void)
{
	return g_world;
}

int clodb200_commRank(void)
{
	return g_rank;
}

void clodb200_commDestroy(void)
{
	std::lock_guard<std::mutex> lock(g_comm_mutex);
#ifndef CLODB_EMU
	if (g_comm)
	{
		cudaStreamSynchronize(g_comm_stream);
		g_nccl.CommDestroy(g_comm);
		g_comm = nullptr;
		cudaStreamDestroy(g_comm_stream);
		g_comm_stream = nullptr;
		cudaFree(g_d_sizes);
		cudaFreeHost(g_h_sizes);
		g_d_sizes = g_h_sizes = nullptr;
	}
#endif
	g_world = 1;
	g_rank = 0;
}

clodb200_gather* clodb200_commGatherBegin(const void* payload, size_t bytes)
{
	if (bytes && !payload)
	{
		capi_set_error("clodb200_commGatherBegin: null payload");
		return nullptr;
	}
	clodb200_gather* g = new clodb200_gather();
	g->world = g_world;
	g->rank = g_rank;
	if (g_world == 1)
	{
		g->sizes.assign(1, bytes);
		g->loopback.assign(static_cast<const unsigned char*>(payload), static_cast<const unsigned char*>(payload) + bytes);
		return g;
	}
#ifdef CLODB_EMU
	delete g;
	return nullptr;
#else
	std::lock_guard<std::mutex> lock(g_comm_mutex);
	auto fail_with = [&](const std::string& what) -> clodb200_gather* {
		capi_set_error(what);
		clodb200_commGatherFree(g);
		return nullptr;
	};
	if (!g_comm)
		return fail_with("clodb200_commGatherBegin: communicator not initialised");
	// 1) byte counts (a few microseconds over NVLink; the padded exchange below needs the maximum)
	g_h_sizes[g_world] = bytes;
	cudaError_t e = cudaMemcpyAsync(g_d_sizes + g_world, g_h_sizes + g_world, sizeof(unsigned long long), cudaMemcpyHostToDevice, g_comm_stream);
	if (e != cudaSuccess)
		return fail_with(std::string("clodb200: size upload failed: ") + cudaGetErrorString(e));
	ncclResult_t r = g_nccl.AllGather(g_d_sizes + g_world, g_d_sizes, 1, ncclUint64, g_comm, g_comm_stream);
	if (r != ncclSuccess)
		return fail_with("clodb200: ncclAllGather (sizes) failed");
	cudaMemcpyAsync(g_h_sizes, g_d_sizes, sizeof(unsigned long long) * size_t(g_world), cudaMemcpyDeviceToHost, g_comm_stream);
	e = cudaStreamSynchronize(g_comm_stream);
	if (e != cudaSuccess)
		return fail_with(std::string("clodb200: size exchange failed: ") + cudaGetErrorString(e));
	g->sizes.assign(g_h_sizes, g_h_sizes + g_world);
	size_t stride = 16;
	for (unsigned long long s : g->sizes)
		stride = std::max<size_t>(stride, (size_t(s) + 15) & ~size_t(15));
	g->stride = stride;
	// 2) payloads, padded to the largest; the result lands in pinned host memory; nothing here blocks the caller
	if (cudaMalloc(reinterpret_cast<void**>(&g->d_send), stride) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&g->d_recv), stride * size_t(g_world)) != cudaSuccess ||
	    cudaMallocHost(reinterpret_cast<void**>(&g->host), stride * size_t(g_world)) != cudaSuccess || cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming) != cudaSuccess)
		return fail_with("clodb200: gather buffers could not be allocated");
	unsigned char* mine = g->host + stride * size_t(g_rank); // own slot of the pinned block doubles as the upload staging
	memcpy(mine, payload, bytes);
	cudaMemcpyAsync(g->d_send, mine, stride, cudaMemcpyHostToDevice, g_comm_stream);
	r = g_nccl.AllGather(g->d_send, g->d_recv, stride, ncclUint8, g_comm, g_comm_stream);
	if (r != ncclSuccess)
		return fail_with("clodb200: ncclAllGather (payloads) failed");
	cudaMemcpyAsync(g->host, g->d_recv, stride * size_t(g_world), cudaMemcpyDeviceToHost, g_comm_stream);
	cudaEventRecord(g->done, g_comm_stream);
	return g;
#endif
}

int clodb200_commGatherWait(clodb200_gather* gather)
{
	if (!gather)
		return CLODB200_ERR_INVALID;
#ifndef CLODB_EMU
	if (gather->done)
	{
		cudaError_t e = cudaEventSynchronize(gather->done);
		if (e != cudaSuccess)
			return cuda_fail(e, "gather wait");
	}
#endif
	return CLODB200_OK;
}

const void* clodb200_commGatherGet(const clodb200_gather* gather, int rank, size_t* out_bytes)
{
	if (!gather || rank < 0 || rank >= gather->world)
		return nullptr;
	if (out_bytes)
		*out_bytes = size_t(gather->sizes[size_t(rank)]);
	if (gather->world == 1)
		return gather->loopback.data();
#ifndef CLODB_EMU
	return gather->host + gather->stride * size_t(rank);
#else
	return nullptr;
#endif
}

void clodb200_commGatherFree(clodb200_gather* gather)
{
	if (!gather)
		return;
#ifndef CLODB_EMU
	if (gather->done)
	{
		cudaEventSynchronize(gather->done);
		cudaEventDestroy(gather->done);
	}
	cudaFree(gather->d_send);
	cudaFree(gather->d_recv);
	if (gather->host)
		cudaFreeHost(gather->host);
#endif
	delete gather;
}

} // extern "C"
