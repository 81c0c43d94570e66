// Test-harness shim (NOT product code).
#pragma once
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
namespace absl {
namespace shim_internal {
struct StringSink {
  std::string out;
  void Append(size_t n, char c) { out.append(n, c); }
  void Append(std::string_view v) { out.append(v); }
};
}  // namespace shim_internal
template <class T, class = void>
struct HasAbslStringify : std::false_type {};
template <class T>
struct HasAbslStringify<
    T, std::void_t<decltype(AbslStringify(
           std::declval<shim_internal::StringSink&>(), std::declval<const T&>()))>>
    : std::true_type {};
template <class... Args>
void Format(shim_internal::StringSink* sink, const char* fmt, Args... args);
}  // namespace absl
