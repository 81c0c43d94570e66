// Test-harness shim (NOT product code).
#pragma once
#include <algorithm>
#include <iterator>
namespace absl {
template <class C, class T>
bool c_linear_search(const C& c, const T& v) {
  return std::find(std::begin(c), std::end(c), v) != std::end(c);
}
template <class C, class T>
auto c_lower_bound(C& c, const T& v) {
  return std::lower_bound(std::begin(c), std::end(c), v);
}
template <class C, class T>
auto c_upper_bound(C& c, const T& v) {
  return std::upper_bound(std::begin(c), std::end(c), v);
}
template <class C, class T, class Cmp>
auto c_lower_bound(C& c, const T& v, Cmp cmp) {
  return std::lower_bound(std::begin(c), std::end(c), v, cmp);
}
template <class C, class T, class Cmp>
auto c_upper_bound(C& c, const T& v, Cmp cmp) {
  return std::upper_bound(std::begin(c), std::end(c), v, cmp);
}
template <class C>
auto c_adjacent_find(C& c) {
  return std::adjacent_find(std::begin(c), std::end(c));
}
template <class C, class P>
auto c_adjacent_find(C& c, P p) {
  return std::adjacent_find(std::begin(c), std::end(c), p);
}
}  // namespace absl
