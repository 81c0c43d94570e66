// Test-harness shim (NOT product code).
#pragma once
#include <cstdio>
#include <string>
#include <string_view>
#include <type_traits>
#include "absl/strings/has_absl_stringify.h"
namespace absl {
namespace shim_internal {
inline void AppendPiece(std::string& s, std::string_view v) { s.append(v); }
inline void AppendPiece(std::string& s, const char* v) { s.append(v ? v : ""); }
inline void AppendPiece(std::string& s, const std::string& v) { s.append(v); }
inline void AppendPiece(std::string& s, char v) { s.push_back(v); }
inline void AppendPiece(std::string& s, bool v) { s.append(v ? "1" : "0"); }
template <class T>
std::enable_if_t<std::is_integral_v<T> && !std::is_same_v<T, bool> &&
                 !std::is_same_v<T, char>>
AppendPiece(std::string& s, T v) {
  s.append(std::to_string(v));
}
template <class T>
std::enable_if_t<std::is_floating_point_v<T>> AppendPiece(std::string& s, T v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%g", static_cast<double>(v));
  s.append(buf);
}
template <class T>
std::enable_if_t<HasAbslStringify<T>::value> AppendPiece(std::string& s,
                                                        const T& v) {
  StringSink sink;
  AbslStringify(sink, v);
  s.append(sink.out);
}
}  // namespace shim_internal
template <class... Args>
std::string StrCat(const Args&... args) {
  std::string s;
  (shim_internal::AppendPiece(s, args), ...);
  return s;
}
template <class... Args>
void StrAppend(std::string* s, const Args&... args) {
  (shim_internal::AppendPiece(*s, args), ...);
}
}  // namespace absl
