// Cache naming: the renderer only finds a cache through CLodCache::BuildCacheFileName /
// BuildSceneCacheSubdirectory / ComputeBuildConfigHash (BasicRenderer/src/Import/CLodCache.cpp:62-80, 586-633), which are
// boost::hash_combine chains over three strings, fourteen literal constants and eleven environment variables.
//
// PARITY UNPINNED. Boost is not vendored in the reference tree and is absent from this image, and no reference test holds a
// known name. What is restated here is the published algorithm of Boost.ContainerHash as shipped since Boost 1.82 (the vcpkg
// baseline pinned by the reference's vcpkg.json, edffab1b, is a 2025 snapshot): hash_combine(seed, v) = hash_mix(seed +
// 0x9e3779b9 + hash<T>(v)) with the 64-bit hash_mix (multiplier 0xe9846af9b1a615d), hash<integral> = value, and
// hash<std::string> = hash_range(0, chars) = mulxp1_hash (8 bytes per step, 128-bit multiply folded by xor). A maintainer with
// Boost at hand can pin it with tests/cpp/boost_names_check.cpp (prints the same values from <boost/container_hash/hash.hpp>).
#include "../../include/clodb200.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace
{
typedef unsigned long long u64;

u64 hash_mix(u64 x)
{
	const u64 m = 0xe9846af9b1a615dULL;
	x ^= x >> 32;
	x *= m;
	x ^= x >> 32;
	x *= m;
	x ^= x >> 28;
	return x;
}

u64 mulx(u64 x, u64 y)
{
	unsigned __int128 r = (unsigned __int128)x * y;
	return u64(r) ^ u64(r >> 64);
}

u64 read32le(const unsigned char* p)
{
	return u64(p[0]) | u64(p[1]) << 8 | u64(p[2]) << 16 | u64(p[3]) << 24;
}

u64 read64le(const unsigned char* p)
{
	return read32le(p) | read32le(p + 4) << 32;
}

// boost::hash_range for ranges of char on 64-bit targets (mulxp1_hash)
u64 hash_chars(u64 seed, const unsigned char* p, size_t n)
{
	const u64 q = 0x9e3779b97f4a7c15ULL;
	const u64 k = q * q;
	u64 w = mulx(seed + q, k);
	u64 h = w ^ u64(n);
	while (n >= 8)
	{
		u64 v1 = read64le(p);
		w += q;
		h ^= mulx(v1 + w, k);
		p += 8;
		n -= 8;
	}
	{
		u64 v1 = 0;
		if (n >= 4)
			v1 = read32le(p + n - 4) << ((n - 4) * 8) | read32le(p);
		else if (n >= 1)
		{
			const size_t x1 = (n - 1) & 2; // 1: 0, 2: 0, 3: 2
			const size_t x2 = n >> 1;      // 1: 0, 2: 1, 3: 1
			v1 = u64(p[x1]) << (x1 * 8) | u64(p[x2]) << (x2 * 8) | u64(p[0]);
		}
		w += q;
		h ^= mulx(v1 + w, k);
	}
	return mulx(h + w, k);
}

void combine_value(u64& seed, u64 hashed)
{
	seed = hash_mix(seed + 0x9e3779b9ULL + hashed);
}

void combine_string(u64& seed, const std::string& s)
{
	combine_value(seed, hash_chars(0, reinterpret_cast<const unsigned char*>(s.data()), s.size()));
}

size_t emit(const std::string& s, char* out, size_t capacity)
{
	if (out && capacity)
	{
		size_t n = s.size() < capacity - 1 ? s.size() : capacity - 1;
		memcpy(out, s.data(), n);
		out[n] = 0;
	}
	return s.size() + 1;
}

std::string hex(u64 v)
{
	char buf[32];
	snprintf(buf, sizeof(buf), "%llx", v);
	return buf;
}
} // namespace

extern "C"
{

uint64_t clodb200_cacheBuildConfigHash(void)
{
	// CLodCache.cpp:586-619: schema version, meshlet size and twelve literal layout/heuristic versions, then the voxel environment
	u64 seed = 0;
	const unsigned int constants[14] = {47u /* kSchemaVersion */, 128u /* MS_MESHLET_SIZE */, 32u, 4u, 4u, 1u, 1u, 7u, 1u, 1u, 1u, 7u, 27u, 3u};
	for (unsigned int c : constants)
		combine_value(seed, c);
	const char* env[11] = {"BASICRENDERER_CLOD_VOXEL_MODE", "BASICRENDERER_CLOD_VOXEL_GRID", "BASICRENDERER_CLOD_VOXEL_MIN_RES", "BASICRENDERER_CLOD_VOXEL_RAYS", "BASICRENDERER_CLOD_VOXEL_SCALE",
	    "BASICRENDERER_CLOD_VOXEL_RETRIES", "BASICRENDERER_CLOD_VOXEL_GROWTH", "BASICRENDERER_CLOD_VOXEL_ACCEPTANCE_BIAS", "BASICRENDERER_CLOD_VOXEL_OPACITY_THRESHOLD",
	    "BASICRENDERER_CLOD_VOXEL_CARRY_ZERO_COVERAGE", "BASICRENDERER_CLOD_VOXEL_PRUNING"};
	for (const char* name : env)
	{
		const char* value = getenv(name);
		combine_string(seed, value ? std::string(value) : std::string());
	}
	return seed;
}

size_t clodb200_cacheFileName(const char* source_identifier, const char* prim_path, const char* subset_name, uint64_t build_config_hash, char* out, size_t capacity)
{
	// CLodCache.cpp:621-633
	u64 seed = 0;
	combine_string(seed, source_identifier ? source_identifier : "");
	combine_string(seed, prim_path ? prim_path : "");
	combine_string(seed, subset_name ? subset_name : "");
	combine_value(seed, build_config_hash);
	return emit("clod_" + hex(seed) + ".usdc", out, capacity);
}

size_t clodb200_cacheSubdirectory(const char* source_identifier, char* out, size_t capacity)
{
	// CLodCache.cpp:38-80: "clod\<sanitised stem of the source path>_<hex hash of the whole identifier>"
	const std::string source = source_identifier ? source_identifier : "";
	std::string stem = "scene";
	if (!source.empty())
	{
		size_t slash = source.find_last_of("/\\");
		std::string file = slash == std::string::npos ? source : source.substr(slash + 1);
		size_t dot = file.find_last_of('.');
		std::string s = (dot == std::string::npos || dot == 0 || file == "..") ? file : file.substr(0, dot);
		if (!s.empty())
			stem = s;
	}
	std::string clean;
	for (size_t i = 0; i < stem.size();)
	{
		unsigned char c = static_cast<unsigned char>(stem[i]);
		if (c < 0x80)
		{
			clean.push_back((isalnum(c) || c == '_' || c == '-') ? char(c) : '_');
			++i;
		}
		else
		{
			// one wide character of the reference's std::wstring: not alphanumeric in the "C" locale
			size_t len = c >= 0xf0 ? 4 : (c >= 0xe0 ? 3 : (c >= 0xc0 ? 2 : 1));
			clean.push_back('_');
			i += len;
		}
	}
	if (clean.empty())
		clean = "scene";
	u64 seed = 0;
	combine_string(seed, source);
	return emit("clod/" + clean + "_" + hex(seed), out, capacity);
}

} // extern "C"
