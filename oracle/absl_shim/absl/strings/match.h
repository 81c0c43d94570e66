// Test-harness shim (NOT product code).
#pragma once
#include <string_view>
namespace absl {
inline bool StartsWith(std::string_view s, std::string_view p) {
  return s.substr(0, p.size()) == p;
}
inline bool EndsWith(std::string_view s, std::string_view p) {
  return s.size() >= p.size() && s.substr(s.size() - p.size()) == p;
}
inline bool StrContains(std::string_view s, std::string_view p) {
  return s.find(p) != std::string_view::npos;
}
}  // namespace absl
