// ORACLE BUILD SHIM (test infrastructure, not product code).
// Minimal POD stand-ins for the four DirectXMath value types the reference's
// headless CLod builder uses (ClusterLODTypes.h:11, ClusterLODShaderTypes.h:8).
// No XM math functions are used on that path, so none are provided.
#pragma once
#include <cstdint>
namespace DirectX {
struct XMFLOAT2 { float x, y; XMFLOAT2() = default; constexpr XMFLOAT2(float x_, float y_) : x(x_), y(y_) {} };
struct XMFLOAT3 { float x, y, z; XMFLOAT3() = default; constexpr XMFLOAT3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };
struct XMFLOAT4 { float x, y, z, w; XMFLOAT4() = default; constexpr XMFLOAT4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {} };
struct XMUINT4 { uint32_t x, y, z, w; XMUINT4() = default; constexpr XMUINT4(uint32_t x_, uint32_t y_, uint32_t z_, uint32_t w_) : x(x_), y(y_), z(z_), w(w_) {} };
}
