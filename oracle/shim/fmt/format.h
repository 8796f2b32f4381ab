// oracle/shim/fmt/format.h — TEST INFRASTRUCTURE: fmt::print / fmt::format swallowed (debug output only)
#pragma once
#include <string>
namespace fmt {
template <typename... A>
inline void print(const A&...) {}
template <typename... A>
inline std::string format(const A&...) { return std::string(); }
}  // namespace fmt
