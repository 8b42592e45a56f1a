// For a machine that HAS Boost (this image does not): prints the values the reference's naming code computes, from the real
// <boost/container_hash/hash.hpp>, for the inputs tests/test_cache_names.py uses. Compare with the library:
//   g++ -std=c++17 -I<boost> tests/cpp/boost_names_check.cpp -o boost_names_check && ./boost_names_check
//   python -c "from basicrenderer_b200 import cache; print(hex(cache.build_config_hash({})), cache.cache_file_name('models/zorah.usd', '/World/Mesh_0', '', cache.build_config_hash({})))"
// Mirrors CLodCache.cpp:586-633 (ComputeBuildConfigHash with an empty environment, BuildCacheFileName).
#include <boost/container_hash/hash.hpp>
#include <cstdint>
#include <cstdio>
#include <string>

int main()
{
	size_t seed = 0;
	const uint32_t constants[14] = {47, 128, 32, 4, 4, 1, 1, 7, 1, 1, 1, 7, 27, 3};
	for (uint32_t c : constants)
		boost::hash_combine(seed, c);
	for (int i = 0; i < 11; ++i)
		boost::hash_combine(seed, std::string());
	printf("config hash (empty environment): %zx\n", seed);
	size_t name = 0;
	boost::hash_combine(name, std::string("models/zorah.usd"));
	boost::hash_combine(name, std::string("/World/Mesh_0"));
	boost::hash_combine(name, std::string());
	boost::hash_combine(name, uint64_t(seed));
	printf("clod_%zx.usdc\n", name);
	for (const char* s : {"", "a", "ab", "abc", "abcd", "abcdefg", "abcdefgh", "abcdefghi", "models/zorah.usd"})
		printf("hash(\"%s\") = %zx\n", s, boost::hash<std::string>()(std::string(s)));
	return 0;
}
