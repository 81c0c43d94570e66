// Test-harness shim (NOT product code).
#pragma once
#include <memory>
namespace absl {
template <class T>
std::unique_ptr<T> WrapUnique(T* p) {
  return std::unique_ptr<T>(p);
}
}  // namespace absl
