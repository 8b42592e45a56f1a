// ORACLE BUILD SHIM: no-op logging so the reference builder compiles without spdlog.
#pragma once
namespace spdlog {
template <class... A> inline void info(A&&...) {}
template <class... A> inline void warn(A&&...) {}
template <class... A> inline void error(A&&...) {}
template <class... A> inline void debug(A&&...) {}
template <class... A> inline void trace(A&&...) {}
template <class... A> inline void critical(A&&...) {}
}
